// Micro-benchmark: DADD / IADD64 latency (1 dependent chain, 1 warp) and throughput (8 chains, full SM) on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <class T, int CH> __global__ void chains(T* out, T inc, int iters)
{
  T a[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) a[c] = (T)(threadIdx.x + c);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) a[c] = a[c] + inc;
  }
  long long t1 = clock64();
  T s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[1 << 20] = t1 - t0;
}
template <class T, int CH> void run(const char* name, int blocks, int threads)
{
  T* out; cudaMalloc(&out, (1 << 20) * 8 + 64);
  const int iters = 4096;
  chains<T, CH><<<blocks, threads>>>(out, (T)1, iters);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  chains<T, CH><<<blocks, threads>>>(out, (T)1, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, ((long long*)out) + (1 << 20), 8, cudaMemcpyDeviceToHost);
  double ops = (double)blocks * threads * CH * iters;
  printf("%-6s chains=%d blocks=%4d threads=%4d : %.1f cycles per dependent op (block 0), %.1f Gop/s, %.1f ops/clk/SM at 1.9 GHz x 148\n", name, CH, blocks, threads,
         (double)cyc / iters / (CH > 1 ? 1 : 1), ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(out);
}
int main()
{
  run<double, 1>("DADD", 1, 32);
  run<long long, 1>("IADD64", 1, 32);
  run<float, 1>("FADD", 1, 32);
  run<double, 8>("DADD", 148 * 4, 512);
  run<long long, 8>("IADD64", 148 * 4, 512);
  run<float, 8>("FADD", 148 * 4, 512);
  run<double, 2>("DADD", 148, 384);
  run<double, 1>("DADD", 148, 384);
  return 0;
}
