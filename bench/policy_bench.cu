// bench/policy_bench.cu -- the headline step through the PUBLIC RAJA front-ends, for bench.py.
//
// extern "C" doors onto RAJA::forall / RAJA::inclusive_scan / RAJA::sort_pairs instantiated in THIS translation
// unit with user-style lambdas, once with the b200 policies (the product: include/RAJA_b200/RAJA.hpp) and once with
// the reference's own CUDA back-end (cuda_exec + cuda_reduce + cub::DeviceScan / cub::DeviceRadixSort through
// policy/cuda/scan.hpp:149-184 and sort.hpp:391-459) -- "the kernels to beat" of SURVEY.md section 8d, timed by
// bench.py in the same process on the same buffers.  Built in the build container against the reference headers
// (bench/Makefile, called by __graft_entry__.build()); the .so travels to the GPU box.
#include "RAJA_b200/RAJA.hpp"
#include <cstdint>
#include <exception>
#include <string>

namespace
{
thread_local std::string g_err;
template <class F> int guarded(F&& f)
{
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
using Seg = RAJA::TypedRangeSegment<long>;
inline RAJA::resources::Cuda res_of(void* stream) { return RAJA::resources::Cuda::CudaFromStream((cudaStream_t)stream); }

using B   = RAJA::b200_exec_async<256>;
using C   = RAJA::cuda_exec_async<256>;
}  // namespace

extern "C" {

const char* pb_last_error(void) { return g_err.c_str(); }

// ---- forall + ReduceSum: dot product, fp64.  The call returns when the result is on the host (what RAJA guarantees).
int pb_dot_expt_b200(const double* x, const double* y, long n, double* out, void* stream)
{
  return guarded([&] {
    double dot = 0.0;
    RAJA::forall<B>(res_of(stream), Seg(0, n), RAJA::expt::Reduce<RAJA::operators::plus>(&dot),
                    [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<double, RAJA::operators::plus>& d) { d += x[i] * y[i]; });
    *out = dot;
  });
}
int pb_dot_classic_b200(const double* x, const double* y, long n, double* out, void* stream)
{
  return guarded([&] {
    RAJA::ReduceSum<RAJA::b200_reduce, double> dot(0.0);
    RAJA::forall<B>(res_of(stream), Seg(0, n), [=] RAJA_DEVICE(long i) { dot += x[i] * y[i]; });
    *out = dot.get();
  });
}
int pb_dot_expt_cuda(const double* x, const double* y, long n, double* out, void* stream)
{
  return guarded([&] {
    double dot = 0.0;
    RAJA::forall<C>(res_of(stream), Seg(0, n), RAJA::expt::Reduce<RAJA::operators::plus>(&dot),
                    [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<double, RAJA::operators::plus>& d) { d += x[i] * y[i]; });
    *out = dot;
  });
}
int pb_dot_classic_cuda(const double* x, const double* y, long n, double* out, void* stream)
{
  return guarded([&] {
    RAJA::ReduceSum<RAJA::cuda_reduce, double> dot(0.0);
    RAJA::forall<C>(res_of(stream), Seg(0, n), [=] RAJA_DEVICE(long i) { dot += x[i] * y[i]; });
    *out = dot.get();
  });
}

// ---- inclusive_scan, int32, operators::plus
int pb_inclusive_scan_b200(const int* in, int* out, long n, void* stream)
{
  return guarded([&] { RAJA::inclusive_scan<B>(res_of(stream), RAJA::make_span(in, n), RAJA::make_span(out, n), RAJA::operators::plus<int> {}); });
}
int pb_inclusive_scan_cuda(const int* in, int* out, long n, void* stream)
{
  return guarded([&] { RAJA::inclusive_scan<C>(res_of(stream), RAJA::make_span(in, n), RAJA::make_span(out, n), RAJA::operators::plus<int> {}); });
}

// ---- sort_pairs, uint32 keys + uint32 values, ascending
int pb_sort_pairs_b200(unsigned* keys, unsigned* vals, long n, void* stream)
{
  return guarded([&] { RAJA::sort_pairs<B>(res_of(stream), RAJA::make_span(keys, n), RAJA::make_span(vals, n)); });
}
int pb_sort_pairs_cuda(unsigned* keys, unsigned* vals, long n, void* stream)
{
  return guarded([&] { RAJA::sort_pairs<C>(res_of(stream), RAJA::make_span(keys, n), RAJA::make_span(vals, n)); });
}

// ---- the other BASELINE configurations through the policy layer (bench.py --paths)
int pb_triad_b200(double* a, const double* b, const double* c, double s, long n, void* stream)
{
  return guarded([&] { RAJA::forall<B>(res_of(stream), Seg(0, n), [=] RAJA_DEVICE(long i) { a[i] = b[i] + s * c[i]; }); });
}
int pb_triad_cuda(double* a, const double* b, const double* c, double s, long n, void* stream)
{
  return guarded([&] { RAJA::forall<C>(res_of(stream), Seg(0, n), [=] RAJA_DEVICE(long i) { a[i] = b[i] + s * c[i]; }); });
}
int pb_stable_sort_f64_b200(double* keys, long n, void* stream)
{
  return guarded([&] { RAJA::stable_sort<B>(res_of(stream), RAJA::make_span(keys, n)); });
}
int pb_stable_sort_f64_cuda(double* keys, long n, void* stream)
{
  return guarded([&] { RAJA::stable_sort<C>(res_of(stream), RAJA::make_span(keys, n)); });
}
int pb_histogram_b200(int* hist, const int* keys, long n, void* stream)
{
  return guarded([&] { RAJA::forall<B>(res_of(stream), Seg(0, n), [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::b200_atomic>(&hist[keys[i]], 1); }); });
}
int pb_histogram_aggregated_b200(int* hist, const int* keys, long n, void* stream)
{
  return guarded([&] { RAJA::forall<B>(res_of(stream), Seg(0, n), [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::b200_atomic_aggregated>(&hist[keys[i]], 1); }); });
}
int pb_histogram_cuda(int* hist, const int* keys, long n, void* stream)
{
  return guarded([&] { RAJA::forall<C>(res_of(stream), Seg(0, n), [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::cuda_atomic>(&hist[keys[i]], 1); }); });
}

}  // extern "C"
