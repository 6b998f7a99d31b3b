// Tuning / comparison harness for b200/radix_sort.cuh (NOT part of the product library).
// Sweeps sweep-kernel shapes on uint32 keys + uint32 values at N = 2^log2n, checks each against
// cub::DeviceRadixSort::SortPairs (what the reference's cuda_exec sort_pairs delegates to,
// policy/cuda/sort.hpp:391-544; both are stable, so outputs must be identical) and prints timings.
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include "b200/radix_sort.cuh"

using namespace b200;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_keys(uint32_t* k, uint32_t* v, long long n, unsigned seed, unsigned mask)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    k[i] = h & mask;
    v[i] = (uint32_t)i;
  }
}
__global__ void diff_kernel(const uint32_t* a, const uint32_t* b, long long n, unsigned long long* bad)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (a[i] != b[i]) atomicAdd(bad, 1ull);
}

struct Bufs {
  uint32_t *k0, *v0, *k, *v, *ka, *va, *rk, *rv;
  unsigned long long* bad;
  void* ws;
  long long n;
};

template <class F> float time_ms(const Bufs& b, F f, int reps)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {
    CK(cudaMemcpyAsync(b.k, b.k0, b.n * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpyAsync(b.v, b.v0, b.n * 4, cudaMemcpyDeviceToDevice));
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

template <class S> void run_shape(const Bufs& b)
{
  float ms = time_ms(b, [&] { CK((rsort::sort<uint32_t, uint32_t, S>(0, b.k, b.ka, b.v, b.va, b.n, rsort::KeyOrder{rsort::KEY_UNSIGNED, 0}, b.ws))); }, 3);
  CK(cudaMemset(b.bad, 0, 8));
  diff_kernel<<<1184, 256>>>(b.k, b.rk, b.n, b.bad);
  diff_kernel<<<1184, 256>>>(b.v, b.rv, b.n, b.bad);
  unsigned long long h = 0;
  CK(cudaMemcpy(&h, b.bad, 8, cudaMemcpyDeviceToHost));
  int occ = 0;
  using K = rsort::SweepKernel<uint32_t, uint32_t, false, uint32_t, S>;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, K::kernel(), S::THREADS, sizeof(typename K::SM));
  printf("BLOCK=%4d ITEMS=%2d MIN_CTAS=%d MATCH=%d PACK=%d tile=%5d smem=%6.1fKB occ=%d : %8.3f ms  %6.2f Gpairs/s  model68 %7.1f GB/s %s\n", S::BLOCK, S::ITEMS, S::MIN_CTAS,
         S::MATCH, (int)S::PACK + 10 * S::LB_WARPS + 1000 * S::PREFETCH_TILES, S::TILE, sizeof(typename K::SM) / 1024.0, occ, ms, b.n / ms / 1e6, 68.0 * b.n / ms / 1e6, h ? "MISMATCH" : "ok");
  fflush(stdout);
}

int main(int argc, char** argv)
{
  int log2n = argc > 1 ? atoi(argv[1]) : 28;
  unsigned mask = argc > 2 ? (unsigned)strtoul(argv[2], 0, 0) : 0xffffffffu;
  Bufs b;
  b.n = 1ll << log2n;
  size_t bytes = b.n * 4;
  CK(cudaMalloc(&b.k0, bytes)); CK(cudaMalloc(&b.v0, bytes)); CK(cudaMalloc(&b.k, bytes)); CK(cudaMalloc(&b.v, bytes));
  CK(cudaMalloc(&b.ka, bytes)); CK(cudaMalloc(&b.va, bytes)); CK(cudaMalloc(&b.rk, bytes)); CK(cudaMalloc(&b.rv, bytes));
  CK(cudaMalloc(&b.bad, 8));
  CK(cudaMalloc(&b.ws, 256 << 20));
  fill_keys<<<1184, 256>>>(b.k0, b.v0, b.n, 42u, mask);
  // CUB reference result + timing
  {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, b.k0, b.rk, b.v0, b.rv, b.n);
    void* tmp; CK(cudaMalloc(&tmp, tb));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(e0);
      cub::DeviceRadixSort::SortPairs(tmp, tb, b.k0, b.rk, b.v0, b.rv, b.n);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0 && ms < best) best = ms;
    }
    printf("cub::DeviceRadixSort::SortPairs u32/u32 N=2^%d mask=%#x (CCCL %d): %8.3f ms  %6.2f Gpairs/s  model68 %7.1f GB/s\n", log2n, mask, CUB_VERSION, best, b.n / best / 1e6,
           68.0 * b.n / best / 1e6);
    cudaFree(tmp);
  }
  run_shape<rsort::SweepShape<384, 24, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 30, 3, 3, false, 444>>(b);
  run_shape<rsort::SweepShape<256, 30, 3, 3, false, 300>>(b);
  run_shape<rsort::SweepShape<256, 30, 3, 3, false, 600>>(b);
  run_shape<rsort::SweepShape<256, 30, 3, 3, false, 444, 1>>(b);
  run_shape<rsort::SweepShape<256, 32, 3, 3, false, 444, 1>>(b);
  run_shape<rsort::SweepShape<256, 32, 3, 3, true, 444, 1>>(b);
  run_shape<rsort::SweepShape<256, 30, 3, 3, true, 444, 1>>(b);
  run_shape<rsort::SweepShape<384, 24, 2, 3, false, 296>>(b);
  run_shape<rsort::SweepShape<384, 28, 2, 3, false, 296, 1>>(b);
  run_shape<rsort::SweepShape<512, 20, 2, 3, false, 296, 1>>(b);
  return 0;
}
