// Tuning / comparison harness for b200/radix_sort.cuh (NOT part of the product library).
// Sweeps sweep-kernel shapes at N = 2^log2n for one (key, value) type pair chosen at compile time
// (-DTK=0 u32+u32 [default], 1 u32 keys only, 2 u64 keys only (fp64-style), 3 u64+u32, 4 u64+u64,
// 5 u32+u64), checks each against cub::DeviceRadixSort (what the reference's cuda_exec sorts delegate
// to, policy/cuda/sort.hpp:101-227, 391-544; both are stable, so outputs must be identical) and
// prints timings.   tune_sort [log2n] [key mask]
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include "b200/radix_sort.cuh"

using namespace b200;
#ifndef TK
#define TK 0
#endif
#if TK == 0
using K = uint32_t; using V = uint32_t;
#elif TK == 1
using K = uint32_t; using V = rsort::NoValue;
#elif TK == 2
using K = unsigned long long; using V = rsort::NoValue;
#elif TK == 3
using K = unsigned long long; using V = uint32_t;
#elif TK == 4
using K = unsigned long long; using V = unsigned long long;
#else
using K = uint32_t; using V = unsigned long long;
#endif
constexpr bool HASV = !std::is_same<V, rsort::NoValue>::value;
using VS = std::conditional<HASV, V, uint32_t>::type;  // storage type of the value buffers
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_keys(K* k, VS* v, long long n, unsigned seed, unsigned long long mask)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    unsigned g = h * 2654435761u + 12345u;
    g ^= g >> 15; g *= 2246822519u; g ^= g >> 13;
    unsigned long long kk = sizeof(K) == 8 ? (((unsigned long long)g << 32) | h) : h;
    k[i] = (K)(kk & mask);
    v[i] = (VS)i;
  }
}
template <class T> __global__ void diff_kernel(const T* a, const T* b, long long n, unsigned long long* bad)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (a[i] != b[i]) atomicAdd(bad, 1ull);
}

struct Bufs {
  K *k0, *k, *ka, *rk;
  VS *v0, *v, *va, *rv;
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
    CK(cudaMemcpyAsync(b.k, b.k0, b.n * sizeof(K), cudaMemcpyDeviceToDevice));
    CK(cudaMemcpyAsync(b.v, b.v0, b.n * sizeof(VS), cudaMemcpyDeviceToDevice));
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

constexpr double MODEL_BYTES = sizeof(K) + sizeof(K) * 2.0 * (sizeof(K) + (HASV ? sizeof(V) : 0));

template <class S> void run_shape(const Bufs& b)
{
  float ms = time_ms(b, [&] { CK((rsort::sort<K, V, S>(0, b.k, b.ka, (V*)b.v, (V*)b.va, b.n, rsort::KeyOrder{rsort::KEY_UNSIGNED, 0}, b.ws))); }, 3);
  CK(cudaMemset(b.bad, 0, 8));
  diff_kernel<<<1184, 256>>>(b.k, b.rk, b.n, b.bad);
  if (HASV) diff_kernel<<<1184, 256>>>(b.v, b.rv, b.n, b.bad);
  unsigned long long h = 0;
  CK(cudaMemcpy(&h, b.bad, 8, cudaMemcpyDeviceToHost));
  int occ = 0;
  using KK = rsort::SweepKernel<K, V, false, uint32_t, S>;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, KK::kernel(), S::THREADS, sizeof(typename KK::SM));
  printf("BLOCK=%4d ITEMS=%2d MIN_CTAS=%d MATCH=%d PACK=%d PREFETCH=%3d LBW=%d LBB=%d tile=%5d smem=%6.1fKB occ=%d : %8.3f ms  %6.2f Gkeys/s  model%.0f %7.1f GB/s %s\n", S::BLOCK, S::ITEMS,
         S::MIN_CTAS, S::MATCH, (int)S::PACK, S::PREFETCH_TILES, S::LB_WARPS, S::LB_BATCH, S::TILE, sizeof(typename KK::SM) / 1024.0, occ, ms, b.n / ms / 1e6, MODEL_BYTES, MODEL_BYTES * b.n / ms / 1e6,
         h ? "MISMATCH" : "ok");
  fflush(stdout);
}

int main(int argc, char** argv)
{
  int log2n = argc > 1 ? atoi(argv[1]) : 28;
  unsigned long long mask = argc > 2 ? strtoull(argv[2], 0, 0) : ~0ull;
  Bufs b;
  b.n = 1ll << log2n;
  size_t kb = b.n * sizeof(K), vb = b.n * sizeof(VS);
  CK(cudaMalloc(&b.k0, kb)); CK(cudaMalloc(&b.k, kb)); CK(cudaMalloc(&b.ka, kb)); CK(cudaMalloc(&b.rk, kb));
  CK(cudaMalloc(&b.v0, vb)); CK(cudaMalloc(&b.v, vb)); CK(cudaMalloc(&b.va, vb)); CK(cudaMalloc(&b.rv, vb));
  CK(cudaMalloc(&b.bad, 8));
  CK(cudaMalloc(&b.ws, 512 << 20));
  fill_keys<<<1184, 256>>>(b.k0, b.v0, b.n, 42u, mask);
  // CUB reference result + timing
  {
    size_t tb = 0;
    if (HASV) cub::DeviceRadixSort::SortPairs(nullptr, tb, b.k0, b.rk, b.v0, b.rv, b.n);
    else cub::DeviceRadixSort::SortKeys(nullptr, tb, b.k0, b.rk, b.n);
    void* tmp; CK(cudaMalloc(&tmp, tb));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(e0);
      if (HASV) cub::DeviceRadixSort::SortPairs(tmp, tb, b.k0, b.rk, b.v0, b.rv, b.n);
      else cub::DeviceRadixSort::SortKeys(tmp, tb, b.k0, b.rk, b.n);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0 && ms < best) best = ms;
    }
    printf("cub::DeviceRadixSort key %zu B value %zu B N=2^%d mask=%#llx (CCCL %d): %8.3f ms  %6.2f Gkeys/s  model%.0f %7.1f GB/s\n", sizeof(K), HASV ? sizeof(V) : (size_t)0, log2n, mask,
           CUB_VERSION, best, b.n / best / 1e6, MODEL_BYTES, MODEL_BYTES * b.n / best / 1e6);
    cudaFree(tmp);
  }
#if TK == 0 && defined(SMALL_CTAS)
  // r02: more, smaller CTAs per SM (the sweep is bound by neither pipe alone: LSU 76 %, ALU 61 %, issue 54 % -- do more phases in flight help?)
  run_shape<rsort::SweepShape<256, 32, 3, 3, false, 300, 2, 2>>(b);
  run_shape<rsort::SweepShape<192, 32, 4, 3, false, 400, 2, 2>>(b);
  run_shape<rsort::SweepShape<192, 28, 4, 3, false, 400, 2, 2>>(b);
  run_shape<rsort::SweepShape<192, 32, 4, 3, false, 600, 2, 2>>(b);
  run_shape<rsort::SweepShape<192, 32, 4, 3, false, 400, 2, 2, 400, 12>>(b);
  run_shape<rsort::SweepShape<192, 30, 4, 3, false, 400, 2, 2>>(b);
  run_shape<rsort::SweepShape<192, 24, 5, 3, false, 500, 2, 2>>(b);
  run_shape<rsort::SweepShape<224, 32, 3, 3, false, 300, 2, 2>>(b);
#elif TK == 0
  // r02: larger tiles at 2 CTAs per SM (what paid for the 64-bit keys)
  run_shape<rsort::SweepShape<256, 32, 3, 3, false, 300, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 40, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 48, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 28, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 32, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<512, 24, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<512, 20, 2, 3, false, 200, 2, 2>>(b);
#elif TK == 1
  run_shape<rsort::SweepShape<384, 24, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 30, 3, 3, false, 300>>(b);
  run_shape<rsort::SweepShape<256, 30, 4, 3, false, 300>>(b);
  run_shape<rsort::SweepShape<256, 24, 4, 3, false, 400>>(b);
#elif TK == 2
  // r02: 64-bit keys move twice the bytes per key, so the r01 shapes (4608-key tiles) pay the per-tile costs (look-back,
  // barriers) twice as often per byte as the u32 pairs; larger tiles at 2 CTAs per SM
  run_shape<rsort::SweepShape<256, 18, 3, 3, false, 300, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 24, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 28, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 32, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 20, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 22, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 24, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 26, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 36, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<512, 16, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<512, 18, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 28, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<448, 22, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<448, 24, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 26, 2, 3, false, 200, 2, 2, 400, 12>>(b);
  run_shape<rsort::SweepShape<384, 26, 2, 3, false, 200, 2, 4>>(b);
#elif TK == 3
  run_shape<rsort::SweepShape<256, 16, 3, 3, false, 300, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 24, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 16, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 18, 2, 3, false, 200, 2, 2>>(b);
#elif TK == 4
  run_shape<rsort::SweepShape<256, 14, 3, 3, false, 300, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 20, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 14, 2, 3, false, 200, 2, 2>>(b);
  run_shape<rsort::SweepShape<384, 16, 2, 3, false, 200, 2, 2>>(b);
#else
  run_shape<rsort::SweepShape<384, 16, 2, 2>>(b);
  run_shape<rsort::SweepShape<256, 20, 3, 3, false, 300>>(b);
  run_shape<rsort::SweepShape<256, 18, 3, 3, false, 300>>(b);
  run_shape<rsort::SweepShape<384, 16, 2, 3, false, 200>>(b);
#endif
  return 0;
}
