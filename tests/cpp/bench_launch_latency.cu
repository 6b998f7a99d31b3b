// Per-launch cost of a small forall: 2000 async foralls of 1000 iterations back to back on one stream, b200_exec next to
// the reference's cuda_exec.  "host" = wall time of the enqueue loop per launch, "stream" = CUDA-event time per launch.
#include "RAJA_b200/RAJA.hpp"
#include <chrono>
#include <cstdio>
template <class Pol, class... P> void run(const char* name, RAJA::resources::Cuda res, double* a, long len)
{
  cudaStream_t s = res.get_stream();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int n = 2000;
  double best_host = 1e30, best_dev = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaStreamSynchronize(s);
    cudaEventRecord(e0, s);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; ++i) RAJA::forall<Pol>(res, RAJA::TypedRangeSegment<long>(0, len), [=] RAJA_DEVICE(long j) { a[j] += 1.0; });
    auto t1 = std::chrono::steady_clock::now();
    cudaEventRecord(e1, s); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best_host = std::min(best_host, std::chrono::duration<double, std::micro>(t1 - t0).count() / n);
    best_dev = std::min(best_dev, ms * 1e3 / n);
  }
  printf("%-28s len %7ld : host %6.2f us / launch, stream %6.2f us / launch\n", name, len, best_host, best_dev);
}
int main()
{
  RAJA::resources::Cuda res;
  double* a; cudaMalloc(&a, 8 << 20); cudaMemset(a, 0, 8 << 20);
  for (long len : {1L, 1000L, 100000L, 1000000L}) {
    run<RAJA::b200_exec_async<256>>("b200_exec_async<256>", res, a, len);
    run<RAJA::cuda_exec_async<256>>("cuda_exec_async<256>", res, a, len);
  }
  return 0;
}
