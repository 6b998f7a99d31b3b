// How does an atomic-bearing loop body (histogram: hist[keys[i]] += 1) react to the shape of the persistent forall engine?
// Sweeps grid size and UNROLL of b200::launch_forall with the body a user lambda compiles to (one load, one RED per index).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iraja_b200/csrc tools/micro/atomic_grid_sweep.cu -o atomic_grid_sweep
#include <cstdio>
#include <cstdlib>
#include "b200/forall.cuh"
using namespace b200;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
struct HistBody {
  int* hist;
  const int* keys;
  __device__ void operator()(long i) const { atomicAdd(&hist[keys[i]], 1); }
};
__global__ void fill(int* k, long n, unsigned mask) { for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) { unsigned h = (unsigned)i * 2654435761u + 7u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16; k[i] = (int)(h & mask); } }
// the reference's mapping: one index per thread, one block per 256 indices
__global__ void direct_kernel(HistBody b, long n) { long i = blockIdx.x * (long)blockDim.x + threadIdx.x; if (i < n) b(i); }

struct TriadBody {
  double* a;
  const double* b;
  const double* c;
  double s;
  __device__ void operator()(long i) const { a[i] = b[i] + s * c[i]; }
};
template <int BLOCK, int UNROLL, class Body> float run(Body b, long n, int grid)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    CK((launch_forall<BLOCK, UNROLL, long>(0, b, n, NoFinalize {}, grid)));
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  return best;
}
int main(int argc, char** argv)
{
  const long n = 1L << (argc > 1 ? atoi(argv[1]) : 28);
  const int bins = argc > 2 ? atoi(argv[2]) : 65536;
  int *keys, *hist;
  CK(cudaMalloc(&keys, n * 4)); CK(cudaMalloc(&hist, bins * 4)); CK(cudaMemset(hist, 0, bins * 4));
  fill<<<1184, 256>>>(keys, n, bins - 1);
  HistBody b {hist, keys};
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); direct_kernel<<<(unsigned)((n + 255) / 256), 256>>>(b, n); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    printf("direct (1 index per thread, %ld blocks of 256)            : %8.4f ms\n", (n + 255) / 256, best);
  }
  const int sms = device_info().sm_count;
  for (int per_sm : {1, 2, 4, 8, 16, 32, 64, 128, 256, 512}) {
    const int g = sms * per_sm;
    printf("grid %5d (%2d/SM)  BLOCK 256:  UNROLL 1 %8.4f  2 %8.4f  4 %8.4f  8 %8.4f ms   | BLOCK 1024 UNROLL 1 %8.4f  4 %8.4f ms\n", g, per_sm, run<256, 1>(b, n, g), run<256, 2>(b, n, g),
           run<256, 4>(b, n, g), run<256, 8>(b, n, g), run<1024, 1>(b, n, g / 4 > 0 ? g / 4 : 1), run<1024, 4>(b, n, g / 4 > 0 ? g / 4 : 1));
    fflush(stdout);
  }
  // the same question for a streaming body (triad, 24 B per index): what does oversubscribing the grid cost?
  double *a, *bb, *c;
  CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&bb, n * 8)); CK(cudaMalloc(&c, n * 8));
  CK(cudaMemset(bb, 0, n * 8)); CK(cudaMemset(c, 0, n * 8));
  TriadBody t {a, bb, c, 3.0};
  for (int per_sm : {2, 4, 8, 16, 32, 64}) {
    const int g = sms * per_sm;
    const float m1 = run<256, 4>(t, n, g), m2 = run<256, 2>(t, n, g), m3 = run<256, 1>(t, n, g);
    printf("triad f64  grid %5d (%2d/SM) BLOCK 256: UNROLL 4 %8.4f ms %7.1f GB/s | UNROLL 2 %8.4f ms %7.1f GB/s | UNROLL 1 %8.4f ms %7.1f GB/s\n", g, per_sm, m1, 24.0 * n / m1 / 1e6, m2,
           24.0 * n / m2 / 1e6, m3, 24.0 * n / m3 / 1e6);
  }
  return 0;
}
