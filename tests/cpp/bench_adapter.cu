// Side-by-side timing of the b200 policies and the reference's own CUDA back-end (cuda_exec / cuda_reduce /
// CUB scan and sort: "the kernels to beat") through the SAME RAJA front-end calls with
// the SAME lambdas, on one GPU.  CUDA events on the resource's stream, 3 warm-ups, best of 10.
//   bench_adapter [log2n=28]
#include "RAJA_b200/RAJA.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

template <class F> float best_ms(cudaStream_t s, F f, int reps = 10, int warm = 3)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < warm; ++i) f();
  CK(cudaStreamSynchronize(s));
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(e0, s);
    f();
    cudaEventRecord(e1, s);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  return best;
}

__global__ void fill_f64(double* p, long n, unsigned seed) { for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) { unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; p[i] = 0.5 + (h >> 8) * (1.0 / 16777216.0); } }
__global__ void fill_u32(unsigned* p, long n, unsigned seed, unsigned mask) { for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) { unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16; p[i] = h & mask; } }
__global__ void iota_u32(unsigned* p, long n) { for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = (unsigned)i; }

void row(const char* what, const char* pol, double bytes, float ms) { printf("%-34s %-28s %9.4f ms %9.1f GB/s\n", what, pol, ms, bytes / ms / 1e6); fflush(stdout); }

template <class EX, class AP>
void run_pi_atomic(const char* pol, RAJA::resources::Cuda res, cudaStream_t s, RAJA::TypedRangeSegment<long> seg, double* pi, double dx, long N)
{
  float ms = best_ms(s, [&] { cudaMemsetAsync(pi, 0, 8, s); RAJA::forall<EX>(res, seg, [=] RAJA_DEVICE(long i) { double x = (double(i) + 0.5) * dx; RAJA::atomicAdd<AP>(pi, dx / (1.0 + x * x)); }); }, 2, 1);
  double h = 0;
  cudaMemcpy(&h, pi, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %-36s %9.4f ms %9.2f Gupdates/s   pi = %.12f\n", "pi, atomicAdd into one address", pol, ms, N / ms / 1e6, 4.0 * h);
}

int main(int argc, char** argv)
{
  const int log2n = argc > 1 ? atoi(argv[1]) : 28;
  const long N = 1L << log2n;
  RAJA::resources::Cuda res;
  cudaStream_t s = res.get_stream();
  using B = RAJA::b200_exec_async<256>;
  using C = RAJA::cuda_exec_async<256>;
  using CO = RAJA::cuda_exec_occ_calc_async<256>;
  RAJA::TypedRangeSegment<long> seg(0, N);

  double *a, *b, *c;
  CK(cudaMalloc(&a, N * 8)); CK(cudaMalloc(&b, N * 8)); CK(cudaMalloc(&c, N * 8));
  fill_f64<<<1184, 256, 0, s>>>(a, N, 1); fill_f64<<<1184, 256, 0, s>>>(b, N, 2); fill_f64<<<1184, 256, 0, s>>>(c, N, 3);
  printf("# N = 2^%d, fp64 arrays %.2f GiB each\n", log2n, N * 8.0 / (1 << 30));

  // ---- triad a = b + s*c (24 B/elem) and daxpy a += s*b (24 B/elem)
  const double sc = 3.0;
  row("triad f64", "b200_exec<256>", 24.0 * N, best_ms(s, [&] { RAJA::forall<B>(res, seg, [=] RAJA_DEVICE(long i) { a[i] = b[i] + sc * c[i]; }); }));
  row("triad f64", "cuda_exec<256>", 24.0 * N, best_ms(s, [&] { RAJA::forall<C>(res, seg, [=] RAJA_DEVICE(long i) { a[i] = b[i] + sc * c[i]; }); }));
  row("triad f64", "cuda_exec_occ_calc<256>", 24.0 * N, best_ms(s, [&] { RAJA::forall<CO>(res, seg, [=] RAJA_DEVICE(long i) { a[i] = b[i] + sc * c[i]; }); }));
  row("daxpy f64", "b200_exec<256>", 24.0 * N, best_ms(s, [&] { RAJA::forall<B>(res, seg, [=] RAJA_DEVICE(long i) { a[i] += sc * b[i]; }); }));
  row("daxpy f64", "cuda_exec<256>", 24.0 * N, best_ms(s, [&] { RAJA::forall<C>(res, seg, [=] RAJA_DEVICE(long i) { a[i] += sc * b[i]; }); }));

  // ---- dot with expt::Reduce (16 B/elem); the call returns after the result is on the host
  double dot = 0;
  row("dot f64, expt::Reduce<plus>", "b200_exec<256>", 16.0 * N, best_ms(s, [&] { dot = 0; RAJA::forall<B>(res, seg, RAJA::expt::Reduce<RAJA::operators::plus>(&dot), [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<double, RAJA::operators::plus>& d) { d += b[i] * c[i]; }); }));
  const double dot_b200 = dot;
  row("dot f64, expt::Reduce<plus>", "cuda_exec<256>", 16.0 * N, best_ms(s, [&] { dot = 0; RAJA::forall<C>(res, seg, RAJA::expt::Reduce<RAJA::operators::plus>(&dot), [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<double, RAJA::operators::plus>& d) { d += b[i] * c[i]; }); }));
  printf("#   dot: b200 %.17g  cuda %.17g  rel diff %.3g\n", dot_b200, dot, (dot_b200 - dot) / dot);
  // ---- classic reducer objects, sum + min + maxloc in one kernel (8 B/elem)
  {
    double s1 = 0, s2 = 0;
    row("sum+min+maxloc f64, classic", "b200_exec + b200_reduce", 8.0 * N, best_ms(s, [&] {
          RAJA::ReduceSum<RAJA::b200_reduce, double> rs(0.0); RAJA::ReduceMin<RAJA::b200_reduce, double> rm(1e300); RAJA::ReduceMaxLoc<RAJA::b200_reduce, double, long> rl(-1e300, -1);
          RAJA::forall<B>(res, seg, [=] RAJA_DEVICE(long i) { rs += b[i]; rm.min(b[i]); rl.maxloc(b[i], i); }); s1 = rs.get() + rm.get() + (double)rl.getLoc(); }));
    row("sum+min+maxloc f64, classic", "cuda_exec + cuda_reduce", 8.0 * N, best_ms(s, [&] {
          RAJA::ReduceSum<RAJA::cuda_reduce, double> rs(0.0); RAJA::ReduceMin<RAJA::cuda_reduce, double> rm(1e300); RAJA::ReduceMaxLoc<RAJA::cuda_reduce, double, long> rl(-1e300, -1);
          RAJA::forall<C>(res, seg, [=] RAJA_DEVICE(long i) { rs += b[i]; rm.min(b[i]); rl.maxloc(b[i], i); }); s2 = rs.get() + rm.get() + (double)rl.getLoc(); }));
    printf("#   sum+min+loc: b200 %.17g  cuda %.17g\n", s1, s2);
  }

  // ---- scans (2*sizeof(T) B/elem)
  {
    int* in = (int*)a; int* out = (int*)b;
    fill_u32<<<1184, 256, 0, s>>>((unsigned*)in, N, 5, 7);
    row("inclusive_scan int32", "b200_exec<256>", 8.0 * N, best_ms(s, [&] { RAJA::inclusive_scan<B>(res, RAJA::make_span(in, N), RAJA::make_span(out, N), RAJA::operators::plus<int> {}); }));
    row("inclusive_scan int32", "cuda_exec<256> (CUB)", 8.0 * N, best_ms(s, [&] { RAJA::inclusive_scan<C>(res, RAJA::make_span(in, N), RAJA::make_span(out, N), RAJA::operators::plus<int> {}); }));
    row("exclusive_scan_inplace int32", "b200_exec<256>", 8.0 * N, best_ms(s, [&] { RAJA::exclusive_scan_inplace<B>(res, RAJA::make_span(out, N), RAJA::operators::plus<int> {}, 0); }));
    row("exclusive_scan_inplace int32", "cuda_exec<256> (CUB)", 8.0 * N, best_ms(s, [&] { RAJA::exclusive_scan_inplace<C>(res, RAJA::make_span(out, N), RAJA::operators::plus<int> {}, 0); }));
    double* din = c; double* dout = b;
    row("inclusive_scan f64", "b200_exec<256>", 16.0 * N, best_ms(s, [&] { RAJA::inclusive_scan<B>(res, RAJA::make_span(din, N), RAJA::make_span(dout, N), RAJA::operators::plus<double> {}); }));
    row("inclusive_scan f64", "cuda_exec<256> (CUB)", 16.0 * N, best_ms(s, [&] { RAJA::inclusive_scan<C>(res, RAJA::make_span(din, N), RAJA::make_span(dout, N), RAJA::operators::plus<double> {}); }));
  }

  // ---- sort_pairs u32 + u32 (onesweep model 68 B/elem); each run sorts a fresh copy (copy not timed separately:
  //      both policies pay the same 16 B/elem device copy inside the timed region)
  {
    unsigned *k0 = (unsigned*)a, *v0 = k0 + N, *k = (unsigned*)b, *v = k + N;
    fill_u32<<<1184, 256, 0, s>>>(k0, N, 42, 0xffffffffu);
    iota_u32<<<1184, 256, 0, s>>>(v0, N);
    float copy_ms = best_ms(s, [&] { cudaMemcpyAsync(k, k0, N * 8, cudaMemcpyDeviceToDevice, s); });
    float t1 = best_ms(s, [&] { cudaMemcpyAsync(k, k0, N * 8, cudaMemcpyDeviceToDevice, s); RAJA::sort_pairs<B>(res, RAJA::make_span(k, N), RAJA::make_span(v, N)); }, 5, 2) - copy_ms;
    float t2 = best_ms(s, [&] { cudaMemcpyAsync(k, k0, N * 8, cudaMemcpyDeviceToDevice, s); RAJA::sort_pairs<C>(res, RAJA::make_span(k, N), RAJA::make_span(v, N)); }, 5, 2) - copy_ms;
    row("sort_pairs u32+u32 (68 B model)", "b200_exec<256>", 68.0 * N, t1);
    row("sort_pairs u32+u32 (68 B model)", "cuda_exec<256> (CUB)", 68.0 * N, t2);
    double* dk = b; double* dk0 = a;
    fill_f64<<<1184, 256, 0, s>>>(dk0, N, 9);
    copy_ms = best_ms(s, [&] { cudaMemcpyAsync(dk, dk0, N * 8, cudaMemcpyDeviceToDevice, s); });
    t1 = best_ms(s, [&] { cudaMemcpyAsync(dk, dk0, N * 8, cudaMemcpyDeviceToDevice, s); RAJA::stable_sort<B>(res, RAJA::make_span(dk, N)); }, 3, 1) - copy_ms;
    t2 = best_ms(s, [&] { cudaMemcpyAsync(dk, dk0, N * 8, cudaMemcpyDeviceToDevice, s); RAJA::stable_sort<C>(res, RAJA::make_span(dk, N)); }, 3, 1) - copy_ms;
    row("stable_sort f64 (136 B model)", "b200_exec<256>", 136.0 * N, t1);
    row("stable_sort f64 (136 B model)", "cuda_exec<256> (CUB)", 136.0 * N, t2);
  }

  // ---- atomic histogram, 64 K and 256 bins (4 B/elem)
  for (int bins : {65536, 256}) {
    int* keys = (int*)a; int* hist = (int*)c;
    fill_u32<<<1184, 256, 0, s>>>((unsigned*)keys, N, 7, (unsigned)bins - 1);
    cudaMemsetAsync(hist, 0, bins * 4, s);
    char name[64]; snprintf(name, sizeof name, "atomicAdd histogram, %d bins", bins);
    const int reps = bins == 256 ? 2 : 5;
    row(name, "b200_exec + b200_atomic", 4.0 * N, best_ms(s, [&] { RAJA::forall<B>(res, seg, [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::b200_atomic>(&hist[keys[i]], 1); }); }, reps, 1));
    row(name, "cuda_exec + cuda_atomic", 4.0 * N, best_ms(s, [&] { RAJA::forall<C>(res, seg, [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::cuda_atomic>(&hist[keys[i]], 1); }); }, reps, 1));
    // warp-aggregated adds (one atomic per distinct address per warp): pays on contended bins, costs on spread ones
    row(name, "b200_exec + b200_atomic_aggregated", 4.0 * N, best_ms(s, [&] { RAJA::forall<B>(res, seg, [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::b200_atomic_aggregated>(&hist[keys[i]], 1); }); }, reps, 1));
    // sensitivity of an atomic-bearing body to the block size / grid of the persistent engine
    row(name, "b200_exec<128> + b200_atomic", 4.0 * N, best_ms(s, [&] { RAJA::forall<RAJA::b200_exec_async<128>>(res, seg, [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::b200_atomic>(&hist[keys[i]], 1); }); }, reps, 1));
    row(name, "b200_exec<512> + b200_atomic", 4.0 * N, best_ms(s, [&] { RAJA::forall<RAJA::b200_exec_async<512>>(res, seg, [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::b200_atomic>(&hist[keys[i]], 1); }); }, reps, 1));
    row(name, "b200_exec<1024> + b200_atomic", 4.0 * N, best_ms(s, [&] { RAJA::forall<RAJA::b200_exec_async<1024>>(res, seg, [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::b200_atomic>(&hist[keys[i]], 1); }); }, reps, 1));
    row(name, "b200_exec_grid<256,296> + b200_atomic", 4.0 * N, best_ms(s, [&] { RAJA::forall<RAJA::b200_exec_grid_async<256, 296>>(res, seg, [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::b200_atomic>(&hist[keys[i]], 1); }); }, reps, 1));
    // the same histogram as a MultiReduceSum (the framework-visible form: bins privatised per CTA in shared memory)
    snprintf(name, sizeof name, "MultiReduceSum histogram, %d bins", bins);
    {
      RAJA::MultiReduceSum<RAJA::b200_multi_reduce, int> mr((size_t)bins);
      row(name, "b200_exec + b200_multi_reduce", 4.0 * N, best_ms(s, [&] { RAJA::forall<B>(res, seg, [=] RAJA_DEVICE(long i) { mr[keys[i]] += 1; }); }, 5, 1));
      long long tot = 0;
      for (int b = 0; b < bins; ++b) tot += mr.get(b);
      if (tot != 6 * (long long)N) printf("MultiReduceSum(b200) total %lld != %lld\n", tot, 6 * (long long)N);
    }
    {
      RAJA::MultiReduceSum<RAJA::cuda_multi_reduce_atomic, int> mr((size_t)bins);
      row(name, "cuda_exec + cuda_multi_reduce_atomic", 4.0 * N, best_ms(s, [&] { RAJA::forall<C>(res, seg, [=] RAJA_DEVICE(long i) { mr[keys[i]] += 1; }); }, 5, 1));
    }
  }
  // ---- pi by atomics (examples/pi-reduce_vs_atomic.cpp: every iteration adds into ONE address) and by a reducer
  {
    const long bins = N;
    const double dx = 1.0 / double(bins);
    double* pi = c;
    run_pi_atomic<B, RAJA::b200_atomic>("b200_exec + b200_atomic", res, s, seg, pi, dx, N);
    run_pi_atomic<B, RAJA::b200_atomic_aggregated>("b200_exec + b200_atomic_aggregated", res, s, seg, pi, dx, N);
    run_pi_atomic<C, RAJA::cuda_atomic>("cuda_exec + cuda_atomic", res, s, seg, pi, dx, N);
    double pr = 0;
    float ms = best_ms(s, [&] { pr = 0; RAJA::forall<B>(res, seg, RAJA::expt::Reduce<RAJA::operators::plus>(&pr), [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<double, RAJA::operators::plus>& p) { double x = (double(i) + 0.5) * dx; p += dx / (1.0 + x * x); }); }, 3, 1);
    printf("%-34s %-36s %9.4f ms %9.2f Gupdates/s   pi = %.12f\n", "pi, expt::Reduce<plus>", "b200_exec", ms, N / ms / 1e6, 4.0 * pr);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
