// WorkGroup timing: the b200 runners next to the reference's cuda_work runners through the SAME WorkPool /
// WorkGroup calls with the SAME loop bodies, on one GPU.  CUDA events on the
// resource's stream around group.run(), 3 warm-ups, best of 20.  Two groups:
//   halo  : the pack step of examples/tut_halo-exchange.cpp on a 128^3 grid with 3 variables: 78 loops
//           (18 faces of 16384, 36 edges of 128, 24 corners of 1)
//   mixed : 1000 loops with log-uniform lengths in [1, 200000]
#include "RAJA_b200/RAJA.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

template <class T> struct PinnedAllocator {
  using value_type = T;
  PinnedAllocator() = default;
  template <class U> PinnedAllocator(PinnedAllocator<U> const&) noexcept {}
  T* allocate(size_t n) { void* p = nullptr; if (cudaMallocHost(&p, n * sizeof(T)) != cudaSuccess) throw std::bad_alloc(); return static_cast<T*>(p); }
  void deallocate(T* p, size_t) noexcept { cudaFreeHost(p); }
  template <class U> friend bool operator==(PinnedAllocator const&, PinnedAllocator<U> const&) { return true; }
  template <class U> friend bool operator!=(PinnedAllocator const&, PinnedAllocator<U> const&) { return false; }
};

// managed storage: the work structs migrate to the device on first touch and are then read from HBM / L2, not over PCIe
template <class T> struct ManagedAllocator {
  using value_type = T;
  ManagedAllocator() = default;
  template <class U> ManagedAllocator(ManagedAllocator<U> const&) noexcept {}
  T* allocate(size_t n) { void* p = nullptr; if (cudaMallocManaged(&p, n * sizeof(T)) != cudaSuccess) throw std::bad_alloc(); return static_cast<T*>(p); }
  void deallocate(T* p, size_t) noexcept { cudaFree(p); }
  template <class U> friend bool operator==(ManagedAllocator const&, ManagedAllocator<U> const&) { return true; }
  template <class U> friend bool operator!=(ManagedAllocator const&, ManagedAllocator<U> const&) { return false; }
};

struct Loop { long len; long src_off; long stride; long dst_off; };

template <class Exec, class Order, class Alloc = PinnedAllocator<char>> float time_group(const char* what, const char* pol, const std::vector<Loop>& loops, const double* src, double* dst, long total)
{
  using Pol = RAJA::WorkGroupPolicy<Exec, Order, RAJA::constant_stride_array_of_objects, RAJA::indirect_function_call_dispatch>;
  RAJA::WorkPool<Pol, long, RAJA::xargs<>, Alloc> pool(Alloc {});
  for (const Loop& l : loops) {
    const long so = l.src_off, st = l.stride, d0 = l.dst_off;
    pool.enqueue(RAJA::TypedRangeSegment<long>(0, l.len), [=] RAJA_DEVICE(long i) { dst[d0 + i] = src[so + i * st]; });
  }
  auto group = pool.instantiate();
  RAJA::resources::Cuda res = RAJA::resources::Cuda::get_default();
  cudaStream_t s = res.get_stream();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) auto site = group.run(res);
  CK(cudaStreamSynchronize(s));
  float best = 1e30f;
  for (int i = 0; i < 20; ++i) {
    cudaEventRecord(e0, s);
    auto site = group.run(res);
    cudaEventRecord(e1, s);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  printf("%-8s %-70s %5zu loops %9ld iterations %9.4f ms %8.1f GB/s\n", what, pol, loops.size(), total, best, 16.0 * total / best / 1e6);
  fflush(stdout);
  return best;
}

template <class F> void all_policies(const char* what, const std::vector<Loop>& loops, const double* src, double* dst, long total, F check)
{
  time_group<RAJA::b200_work_async<256>, RAJA::unordered_b200_fused>(what, "b200_work_async<256> + unordered_b200_fused", loops, src, dst, total);
  check("b200 fused");
  time_group<RAJA::cuda_work_async<1024>, RAJA::unordered_cuda_loop_y_block_iter_x_threadblock_average>(what, "cuda_work_async<1024> + unordered_cuda_loop_y_block_iter_x_threadblock_average", loops, src, dst, total);
  check("cuda fused");
  time_group<RAJA::b200_work_async<256>, RAJA::unordered_b200_fused, ManagedAllocator<char>>(what, "b200_work_async<256> + unordered_b200_fused, managed storage", loops, src, dst, total);
  check("b200 fused managed");
  time_group<RAJA::cuda_work_async<1024>, RAJA::unordered_cuda_loop_y_block_iter_x_threadblock_average, ManagedAllocator<char>>(what, "cuda_work_async<1024> + unordered_cuda_loop_y_block..., managed storage", loops, src, dst, total);
  time_group<RAJA::cuda_work_async<256>, RAJA::unordered_cuda_loop_y_block_iter_x_threadblock_average>(what, "cuda_work_async<256> + unordered_cuda_loop_y_block_iter_x_threadblock_average", loops, src, dst, total);
  time_group<RAJA::b200_work_async<256>, RAJA::ordered>(what, "b200_work_async<256> + ordered (one forall per loop)", loops, src, dst, total);
  check("b200 ordered");
  time_group<RAJA::cuda_work_async<256>, RAJA::ordered>(what, "cuda_work_async<256> + ordered (one forall per loop)", loops, src, dst, total);
}

int main()
{
  const long G = 128, NV = 3, GRID = G * G * G;
  double* src; double* dst;
  CK(cudaMalloc(&src, NV * GRID * 8));
  std::vector<double> h(NV * GRID);
  for (long i = 0; i < NV * GRID; ++i) h[i] = (double)(i % 1000003);
  CK(cudaMemcpy(src, h.data(), NV * GRID * 8, cudaMemcpyHostToDevice));

  // halo pack: faces (contiguous or strided), edges, corners of every variable
  std::vector<Loop> halo;
  long total = 0;
  for (long v = 0; v < NV; ++v) {
    const long base = v * GRID;
    const long face_off[6] = {0, (G - 1) * G * G, 1, 2, 0, G - 1};
    const long face_stride[6] = {1, 1, G, G, G, G};  // two contiguous faces, four strided ones
    for (int f = 0; f < 6; ++f) { halo.push_back({G * G, base + face_off[f], face_stride[f], total}); total += G * G; }
    for (int e = 0; e < 12; ++e) { halo.push_back({G, base + e * G * G, (e % 3 == 0) ? 1 : (e % 3 == 1 ? G : G * G / 4), total}); total += G; }
    for (int c = 0; c < 8; ++c) { halo.push_back({1, base + c * (GRID / 8), 1, total}); total += 1; }
  }
  CK(cudaMalloc(&dst, std::max(total, 60000000L) * 8));
  auto check_halo = [&](const char* tag) {
    std::vector<double> out(total);
    CK(cudaMemcpy(out.data(), dst, total * 8, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (const Loop& l : halo)
      for (long i = 0; i < l.len; ++i) bad += out[l.dst_off + i] != h[l.src_off + i * l.stride];
    if (bad) printf("MISMATCH (%s, halo): %ld elements\n", tag, bad);
    CK(cudaMemset(dst, 0, total * 8));
  };
  all_policies("halo", halo, src, dst, total, check_halo);

  // mixed: 1000 loops, log-uniform lengths
  std::mt19937 rng(7);
  std::vector<Loop> mixed;
  long mtotal = 0;
  for (int j = 0; j < 1000; ++j) {
    const double u = std::uniform_real_distribution<double>(0.0, 1.0)(rng);
    const long len = std::max(1L, (long)std::exp(u * std::log(200000.0)));
    const long so = (long)(rng() % (unsigned long)(NV * GRID - len));
    mixed.push_back({len, so, 1, mtotal});
    mtotal += len;
  }
  auto check_mixed = [&](const char* tag) {
    std::vector<double> out(mtotal);
    CK(cudaMemcpy(out.data(), dst, mtotal * 8, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (const Loop& l : mixed)
      for (long i = 0; i < l.len; ++i) bad += out[l.dst_off + i] != h[l.src_off + i];
    if (bad) printf("MISMATCH (%s, mixed): %ld elements\n", tag, bad);
    CK(cudaMemset(dst, 0, mtotal * 8));
  };
  all_policies("mixed", mixed, src, dst, mtotal, check_mixed);
  CK(cudaDeviceSynchronize());
  return 0;
}
