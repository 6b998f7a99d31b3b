// How fast can SM stores push data into a peer GPU's memory over NVLink, as a function of the store width per thread?
// (the fused partition + all-to-all sweep of the multi-GPU sort writes 4 bytes per thread, one 128-byte line per warp)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/peer_store_bw.cu -o tools/micro/peer_store_bw ; needs 2 GPUs
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
template <class V> __global__ void store_kernel(V* __restrict__ dst, const V* __restrict__ src, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// runs of `run` elements whose destinations start at arbitrary (4-byte aligned) places: the write-out pattern of the sweep
template <class V> float bench(V* dst, const V* src, size_t bytes, int grid, int block)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    store_kernel<V><<<grid, block>>>(dst, src, bytes / sizeof(V));
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  return best;
}
int main()
{
  int ndev = 0; CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t bytes = 1ull << 30;
  void *src, *local, *peer;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, bytes));
  CK(cudaSetDevice(0)); CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&local, bytes)); CK(cudaMemset(src, 1, bytes));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  for (int per_sm : {2, 4, 8, 16}) {
    const int grid = 148 * per_sm;
    printf("grid %5d x 256:  local  4B %7.1f  16B %7.1f GB/s | peer  4B %7.1f   8B %7.1f  16B %7.1f GB/s\n", grid,
           bytes / bench<uint32_t>((uint32_t*)local, (const uint32_t*)src, bytes, grid, 256) / 1e6,
           bytes / bench<uint4>((uint4*)local, (const uint4*)src, bytes, grid, 256) / 1e6,
           bytes / bench<uint32_t>((uint32_t*)peer, (const uint32_t*)src, bytes, grid, 256) / 1e6,
           bytes / bench<uint2>((uint2*)peer, (const uint2*)src, bytes, grid, 256) / 1e6,
           bytes / bench<uint4>((uint4*)peer, (const uint4*)src, bytes, grid, 256) / 1e6);
    fflush(stdout);
  }
  // copy engine for comparison
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); CK(cudaMemcpyPeerAsync(peer, 1, src, 0, bytes)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  cudaEventRecord(e0); CK(cudaMemcpyPeerAsync(peer, 1, src, 0, bytes)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemcpyPeerAsync: %7.1f GB/s\n", bytes / ms / 1e6);
  return 0;
}
