// Tuning / comparison harness for b200/scan.cuh (NOT part of the product library).
// Sweeps kernel shapes at N = 2^log2n, checks each against cub::DeviceScan (the kernel the
// reference's cuda_exec scan delegates to, policy/cuda/scan.hpp:52-235) and prints GB/s.
//   nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -I raja_b200/csrc tools/tune/tune_scan.cu -o tools/tune/tune_scan
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include "b200/scan.cuh"

using namespace b200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <class T> __global__ void fill_kernel(T* p, long long n, unsigned seed)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = (T)(h & 7);
  }
}
template <class T> __global__ void diff_kernel(const T* a, const T* b, long long n, unsigned long long* bad)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (a[i] != b[i]) atomicAdd(bad, 1ull);
}

template <class F> float time_ms(F f, int reps)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

static int g_clear = 0, g_set = 0;
template <class T, class S> void run_shape(const char* tname, const T* in, T* out, const T* ref, long long n, void* ws, unsigned long long* bad)
{
  CK(cudaMemset(out, 0xff, n * sizeof(T)));
  float ms = time_ms([&] { CK((scan::launch<T, Plus<T>, true, S>(0, in, out, n, T(0), ws, g_clear, g_set))); }, 5);
  CK(cudaMemset(bad, 0, 8));
  diff_kernel<<<1184, 256>>>(out, ref, n, bad);
  unsigned long long h = 0;
  CK(cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost));
  int occ = 0;
  scan::configure<T, Plus<T>, true, S>(&occ);
  printf("%-4s BLOCK=%4d VPT=%2d ROWS=%2d SLOTS=%d LAG=%d SERIAL=%d tile=%6.1fKB smem=%6.1fKB occ=%d : %8.4f ms %8.1f GB/s %s\n", tname, S::BLOCK, S::VPT, S::ROWS, S::SLOTS, S::LAG, (int)S::SERIAL,
         scan::Layout<T, S>::TILE_BYTES / 1024.0, scan::Layout<T, S>::SMEM_BYTES / 1024.0, occ, ms, 2.0 * n * sizeof(T) / ms / 1e6, h ? "MISMATCH" : "ok");
  fflush(stdout);
}

template <class T> void run_type(const char* tname, long long n)
{
  T *in, *out, *ref;
  CK(cudaMalloc(&in, n * sizeof(T))); CK(cudaMalloc(&out, n * sizeof(T))); CK(cudaMalloc(&ref, n * sizeof(T)));
  unsigned long long* bad; CK(cudaMalloc(&bad, 8));
  fill_kernel<<<1184, 256>>>(in, n, 12345u);
  void* ws; CK(cudaMalloc(&ws, 64 << 20));
  // CUB: the existing kernel
  size_t tb = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tb, in, ref, n);
  void* tmp; CK(cudaMalloc(&tmp, tb));
  float ms = time_ms([&] { cub::DeviceScan::InclusiveSum(tmp, tb, in, ref, n); }, 5);
  printf("%-4s cub::DeviceScan::InclusiveSum (CCCL %d) : %8.4f ms %8.1f GB/s\n", tname, CUB_VERSION, ms, 2.0 * n * sizeof(T) / ms / 1e6);
  // plain copy for the roofline denominator of this box
  ms = time_ms([&] { cudaMemcpyAsync(out, in, n * sizeof(T), cudaMemcpyDeviceToDevice, 0); }, 5);
  printf("%-4s cudaMemcpy D2D : %8.4f ms %8.1f GB/s\n", tname, ms, 2.0 * n * sizeof(T) / ms / 1e6);

  run_shape<T, scan::Shape<384, 7, 10>>(tname, in, out, ref, n, ws, bad);        // 168 KB
  run_shape<T, scan::Shape<384, 7, 10, 5, 0>>(tname, in, out, ref, n, ws, bad);  // 210 KB
  run_shape<T, scan::Shape<384, 7, 10, 5, 1>>(tname, in, out, ref, n, ws, bad);
  run_shape<T, scan::Shape<384, 9, 10>>(tname, in, out, ref, n, ws, bad);        // 216 KB
  run_shape<T, scan::Shape<384, 9, 10, 4, 0, true>>(tname, in, out, ref, n, ws, bad);  // + thread totals: 226 KB
  run_shape<T, scan::Shape<384, 7, 10, 5, 0, true>>(tname, in, out, ref, n, ws, bad);  // 220 KB
  run_shape<T, scan::Shape<512, 5, 10>>(tname, in, out, ref, n, ws, bad);        // 160 KB
  run_shape<T, scan::Shape<512, 5, 10, 5, 0>>(tname, in, out, ref, n, ws, bad);  // 200 KB
  run_shape<T, scan::Shape<512, 5, 10, 5, 1>>(tname, in, out, ref, n, ws, bad);
  run_shape<T, scan::Shape<512, 7, 10>>(tname, in, out, ref, n, ws, bad);        // 224 KB
  run_shape<T, scan::Shape<640, 5, 10>>(tname, in, out, ref, n, ws, bad);        // 200 KB
  run_shape<T, scan::Shape<768, 3, 10, 5, 1>>(tname, in, out, ref, n, ws, bad);  // 180 KB
  run_shape<T, scan::Shape<768, 3, 10, 6, 1>>(tname, in, out, ref, n, ws, bad);  // 216 KB
  run_shape<T, scan::Shape<256, 9, 10, 5, 1>>(tname, in, out, ref, n, ws, bad);  // 180 KB
  run_shape<T, scan::Shape<256, 9, 10, 6, 1>>(tname, in, out, ref, n, ws, bad);  // 216 KB
  run_shape<T, scan::Shape<256, 5, 19, 5, 1>>(tname, in, out, ref, n, ws, bad);  // 100 KB occ 2
  run_shape<T, scan::Shape<128, 9, 28, 4, 0>>(tname, in, out, ref, n, ws, bad);  // 72 KB occ 3
  {  // look-back statistics of the default shape
    using S = scan::Shape<384, 7, 10>;
    const long long nt = scan::num_tiles<T, S>(n);
    unsigned* tr; CK(cudaMalloc(&tr, nt * 4)); CK(cudaMemset(tr, 0, nt * 4));
    CK((scan::launch<T, Plus<T>, true, S>(0, in, out, n, T(0), ws, 0, 0, tr)));
    std::vector<unsigned> h(nt);
    CK(cudaMemcpy(h.data(), tr, nt * 4, cudaMemcpyDeviceToHost));
    unsigned long long sum = 0, mx = 0, one = 0;
    for (auto v : h) { sum += v; if (v > mx) mx = v; if (v <= 1) ++one; }
    printf("%-4s look-back polls/tile: mean %.2f max %llu, resolved by the first poll: %.1f%%\n", tname, (double)sum / nt, mx, 100.0 * one / nt);
    cudaFree(tr);
  }
  cudaFree(in); cudaFree(out); cudaFree(ref); cudaFree(ws); cudaFree(tmp); cudaFree(bad);
}

// diagnostic operator: double-typed scan whose combine is an INTEGER add on the bit patterns -- same loads,
// stores, shuffles and tile-state traffic as Plus<double>, but no FP64 instruction (results are meaningless)
struct BitAdd {
  using value_type = double;
  __host__ __device__ static double identity() { return 0.0; }
  __device__ double operator()(double a, double b) const { return __longlong_as_double(__double_as_longlong(a) + __double_as_longlong(b)); }
};
void run_bitadd(long long n)
{
  double *in, *out;
  CK(cudaMalloc(&in, n * 8)); CK(cudaMalloc(&out, n * 8));
  fill_kernel<<<1184, 256>>>(in, n, 12345u);
  void* ws; CK(cudaMalloc(&ws, 64 << 20));
  using S = scan::Shape<384, 7, 10, 5, 0>;
  float ms = time_ms([&] { CK((scan::launch<double, BitAdd, true, S>(0, in, out, n, 0.0, ws, g_clear, g_set))); }, 5);
  printf("f64  BitAdd (integer add on the bits) BLOCK=384 VPT=7 : %8.4f ms %8.1f GB/s\n", ms, 16.0 * n / ms / 1e6);
  ms = time_ms([&] { CK((scan::launch<double, Plus<double>, true, S>(0, in, out, n, 0.0, ws, g_clear, g_set))); }, 5);
  printf("f64  Plus<double>                     BLOCK=384 VPT=7 : %8.4f ms %8.1f GB/s\n", ms, 16.0 * n / ms / 1e6);
  cudaFree(in); cudaFree(out); cudaFree(ws);
}

int main(int argc, char** argv)
{
  int log2n = argc > 1 ? atoi(argv[1]) : 28;
  long long n = 1ll << log2n;
  if (argc > 2) { run_bitadd(n); return 0; }
  run_type<int>("i32", n);
  run_type<long long>("i64", n);
  run_type<double>("f64", n);  // integer-valued doubles: every association gives the same bits, so CUB still checks it
  return 0;
}
