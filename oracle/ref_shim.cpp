// ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" doors onto the UNMODIFIED reference: every function below
// instantiates RAJA's own templates (seq_exec / omp_parallel_for_exec and their
// reducers, scans, sorts, atomics) from the headers where they lie under
// /root/reference.  Built by oracle/Makefile into oracle/_ref/libraja_ref.so
// (git-ignored).  Used to (a) pin oracle/oracle.c, (b) generate tests/golden/,
// (c) serve as the timed CPU baseline (`cpu_baseline.kind = "reference"`).
//
// pol: 0 = RAJA::seq_exec (+seq_reduce, seq_atomic), 1 = RAJA::omp_parallel_for_exec
// (+omp_reduce, omp_atomic).
#include "RAJA/RAJA.hpp"
#include <cstdint>
#include <omp.h>

using Idx = RAJA::Index_type;
using Range = RAJA::TypedRangeSegment<Idx>;

template <class T, template <class...> class OpT = RAJA::operators::plus> struct dummy {};

#define FOR_ALL_TYPES(X) X(i32, int32_t) X(u32, uint32_t) X(i64, int64_t) X(u64, uint64_t) X(f32, float) X(f64, double)

template <class Pol, class RPol, class T> struct Ops {
  static void triad(T* a, const T* b, const T* c, T s, Idx n)
  {
    RAJA::forall<Pol>(Range(0, n), [=](Idx i) { a[i] = b[i] + s * c[i]; });
  }
  static void daxpy(T* a, const T* b, T s, Idx n)
  {
    RAJA::forall<Pol>(Range(0, n), [=](Idx i) { a[i] += b[i] * s; });
  }
  static T sum(const T* x, Idx n)
  {
    RAJA::ReduceSum<RPol, T> r(T(0));
    RAJA::forall<Pol>(Range(0, n), [=](Idx i) { r += x[i]; });
    return r.get();
  }
  static T dot(const T* x, const T* y, Idx n)
  {
    RAJA::ReduceSum<RPol, T> r(T(0));
    RAJA::forall<Pol>(Range(0, n), [=](Idx i) { r += x[i] * y[i]; });
    return r.get();
  }
  static T min(const T* x, Idx n)
  {
    RAJA::ReduceMin<RPol, T> r(RAJA::operators::limits<T>::max());
    RAJA::forall<Pol>(Range(0, n), [=](Idx i) { r.min(x[i]); });
    return r.get();
  }
  static T max(const T* x, Idx n)
  {
    RAJA::ReduceMax<RPol, T> r(RAJA::operators::limits<T>::min());
    RAJA::forall<Pol>(Range(0, n), [=](Idx i) { r.max(x[i]); });
    return r.get();
  }
  static T minloc(const T* x, Idx n, int64_t* loc)
  {
    RAJA::ReduceMinLoc<RPol, T, Idx> r(RAJA::operators::limits<T>::max(), -1);
    RAJA::forall<Pol>(Range(0, n), [=](Idx i) { r.minloc(x[i], i); });
    *loc = r.getLoc();
    return r.get();
  }
  static T maxloc(const T* x, Idx n, int64_t* loc)
  {
    RAJA::ReduceMaxLoc<RPol, T, Idx> r(RAJA::operators::limits<T>::min(), -1);
    RAJA::forall<Pol>(Range(0, n), [=](Idx i) { r.maxloc(x[i], i); });
    *loc = r.getLoc();
    return r.get();
  }
  // examples/forall-param-reductions.cpp:137-160: five expt::Reduce params in one forall
  static void multi(const T* x, Idx n, T* sum_, T* min_, T* max_, T* minv, int64_t* minl, T* maxv, int64_t* maxl)
  {
    using VL = RAJA::expt::ValLoc<T, Idx>;
    T s = 0, mn = RAJA::operators::limits<T>::max(), mx = RAJA::operators::limits<T>::min();
    VL lo(RAJA::operators::limits<T>::max(), -1), hi(RAJA::operators::limits<T>::min(), -1);
    RAJA::forall<Pol>(Range(0, n), RAJA::expt::Reduce<RAJA::operators::plus>(&s),
                      RAJA::expt::Reduce<RAJA::operators::minimum>(&mn), RAJA::expt::Reduce<RAJA::operators::maximum>(&mx),
                      RAJA::expt::Reduce<RAJA::operators::minimum>(&lo), RAJA::expt::Reduce<RAJA::operators::maximum>(&hi),
                      [=](Idx i, RAJA::expt::ValOp<T, RAJA::operators::plus>& s_, RAJA::expt::ValOp<T, RAJA::operators::minimum>& mn_,
                          RAJA::expt::ValOp<T, RAJA::operators::maximum>& mx_,
                          RAJA::expt::ValLocOp<T, Idx, RAJA::operators::minimum>& lo_,
                          RAJA::expt::ValLocOp<T, Idx, RAJA::operators::maximum>& hi_) {
                        s_ += x[i];
                        mn_.min(x[i]);
                        mx_.max(x[i]);
                        lo_.minloc(x[i], i);
                        hi_.maxloc(x[i], i);
                      });
    *sum_ = s; *min_ = mn; *max_ = mx;
    *minv = lo.getVal(); *minl = lo.getLoc(); *maxv = hi.getVal(); *maxl = hi.getLoc();
  }
  template <class Op> static void iscan(const T* in, T* out, Idx n)
  {
    if (in == out) RAJA::inclusive_scan_inplace<Pol>(RAJA::make_span(out, n), Op{});
    else RAJA::inclusive_scan<Pol>(RAJA::make_span(in, n), RAJA::make_span(out, n), Op{});
  }
  template <class Op> static void escan(const T* in, T* out, Idx n, T init)
  {
    if (in == out) RAJA::exclusive_scan_inplace<Pol>(RAJA::make_span(out, n), Op{}, init);
    else RAJA::exclusive_scan<Pol>(RAJA::make_span(in, n), RAJA::make_span(out, n), Op{}, init);
  }
  static void inclusive_scan(int op, const T* in, T* out, Idx n)
  {
    if (op == 0) iscan<RAJA::operators::plus<T>>(in, out, n);
    else if (op == 1) iscan<RAJA::operators::minimum<T>>(in, out, n);
    else iscan<RAJA::operators::maximum<T>>(in, out, n);
  }
  static void exclusive_scan(int op, const T* in, T* out, Idx n, T init)
  {
    if (op == 0) escan<RAJA::operators::plus<T>>(in, out, n, init);
    else if (op == 1) escan<RAJA::operators::minimum<T>>(in, out, n, init);
    else escan<RAJA::operators::maximum<T>>(in, out, n, init);
  }
  template <class V> static void sort_pairs(int stable, T* k, V* v, Idx n, int desc)
  {
    auto ks = RAJA::make_span(k, n);
    auto vs = RAJA::make_span(v, n);
    if (stable) {
      if (desc) RAJA::stable_sort_pairs<Pol>(ks, vs, RAJA::operators::greater<T>{});
      else RAJA::stable_sort_pairs<Pol>(ks, vs, RAJA::operators::less<T>{});
    } else {
      if (desc) RAJA::sort_pairs<Pol>(ks, vs, RAJA::operators::greater<T>{});
      else RAJA::sort_pairs<Pol>(ks, vs, RAJA::operators::less<T>{});
    }
  }
  static void sort(int stable, T* k, Idx n, int desc)
  {
    auto ks = RAJA::make_span(k, n);
    if (stable) {
      if (desc) RAJA::stable_sort<Pol>(ks, RAJA::operators::greater<T>{});
      else RAJA::stable_sort<Pol>(ks, RAJA::operators::less<T>{});
    } else {
      if (desc) RAJA::sort<Pol>(ks, RAJA::operators::greater<T>{});
      else RAJA::sort<Pol>(ks, RAJA::operators::less<T>{});
    }
  }
};

template <class T> using Seq = Ops<RAJA::seq_exec, RAJA::seq_reduce, T>;
template <class T> using Omp = Ops<RAJA::omp_parallel_for_exec, RAJA::omp_reduce, T>;

#define PICK(T, CALL) (pol == 0 ? Seq<T>::CALL : Omp<T>::CALL)

extern "C" {

int ref_omp_max_threads(void) { return omp_get_max_threads(); }
// launchers such as torchrun export OMP_NUM_THREADS=1: the timed baseline sets its thread count explicitly
void ref_omp_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
const char* ref_version(void) { return "RAJA " RAJA_STRINGIFY(RAJA_VERSION_MAJOR) "." RAJA_STRINGIFY(RAJA_VERSION_MINOR) "." RAJA_STRINGIFY(RAJA_VERSION_PATCHLEVEL); }

#define DEF(S, T)                                                                                                        \
  void ref_triad_##S(int pol, T* a, const T* b, const T* c, T s, int64_t n) { PICK(T, triad(a, b, c, s, n)); }            \
  void ref_daxpy_##S(int pol, T* a, const T* b, T s, int64_t n) { PICK(T, daxpy(a, b, s, n)); }                           \
  T ref_sum_##S(int pol, const T* x, int64_t n) { return PICK(T, sum(x, n)); }                                            \
  T ref_dot_##S(int pol, const T* x, const T* y, int64_t n) { return PICK(T, dot(x, y, n)); }                             \
  T ref_min_##S(int pol, const T* x, int64_t n) { return PICK(T, min(x, n)); }                                            \
  T ref_max_##S(int pol, const T* x, int64_t n) { return PICK(T, max(x, n)); }                                            \
  T ref_minloc_##S(int pol, const T* x, int64_t n, int64_t* loc) { return PICK(T, minloc(x, n, loc)); }                   \
  T ref_maxloc_##S(int pol, const T* x, int64_t n, int64_t* loc) { return PICK(T, maxloc(x, n, loc)); }                   \
  void ref_multi_##S(int pol, const T* x, int64_t n, T* s, T* mn, T* mx, T* lv, int64_t* ll, T* hv, int64_t* hl)          \
  { PICK(T, multi(x, n, s, mn, mx, lv, ll, hv, hl)); }                                                                    \
  void ref_inclusive_scan_##S(int pol, int op, const T* in, T* out, int64_t n) { PICK(T, inclusive_scan(op, in, out, n)); } \
  void ref_exclusive_scan_##S(int pol, int op, const T* in, T* out, int64_t n, T init)                                    \
  { PICK(T, exclusive_scan(op, in, out, n, init)); }                                                                      \
  void ref_sort_pairs_##S(int pol, int stable, T* k, uint64_t* v, int64_t n, int desc)                                    \
  { PICK(T, template sort_pairs<uint64_t>(stable, k, v, n, desc)); }                                                      \
  void ref_sort_pairs32_##S(int pol, int stable, T* k, uint32_t* v, int64_t n, int desc)                                  \
  { PICK(T, template sort_pairs<uint32_t>(stable, k, v, n, desc)); }                                                      \
  void ref_sort_##S(int pol, int stable, T* k, int64_t n, int desc) { PICK(T, sort(stable, k, n, desc)); }
FOR_ALL_TYPES(DEF)

// exercises/atomic-histogram_solution.cpp:131-134 / :152-155
void ref_histogram_i32(int pol, int32_t* hist, const int32_t* keys, int64_t n)
{
  if (pol == 0) RAJA::forall<RAJA::seq_exec>(Range(0, n), [=](Idx i) { RAJA::atomicAdd<RAJA::seq_atomic>(&hist[keys[i]], 1); });
  else RAJA::forall<RAJA::omp_parallel_for_exec>(Range(0, n), [=](Idx i) { RAJA::atomicAdd<RAJA::omp_atomic>(&hist[keys[i]], 1); });
}
// examples/pi-reduce_vs_atomic.cpp:105-109 / :143-147
double ref_pi_atomic(int pol, int64_t num_bins)
{
  const double dx = 1.0 / double(num_bins);
  double pi = 0.0;
  double* p = &pi;
  if (pol == 0)
    RAJA::forall<RAJA::seq_exec>(Range(0, num_bins), [=](Idx i) {
      double x = (double(i) + 0.5) * dx;
      RAJA::atomicAdd<RAJA::seq_atomic>(p, dx / (1.0 + x * x));
    });
  else
    RAJA::forall<RAJA::omp_parallel_for_exec>(Range(0, num_bins), [=](Idx i) {
      double x = (double(i) + 0.5) * dx;
      RAJA::atomicAdd<RAJA::omp_atomic>(p, dx / (1.0 + x * x));
    });
  return pi;
}

}  // extern "C"
