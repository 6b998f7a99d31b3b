/* oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of what the reference (LLNL/RAJA,
 * /root/reference, v2025.09) computes on its `seq_exec` path for the hot path of
 * this repository.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product
 * (raja_b200/, include/) never does.
 *
 * PARITY PINNING: this restatement is checked against the reference itself --
 * oracle/_ref/libraja_ref.so, built by oracle/Makefile from the reference's own
 * headers where they lie under /root/reference (ref_shim.cpp instantiates
 * RAJA::forall/scan/sort<seq_exec>) -- by tests/test_oracle_vs_reference.py, and
 * against tests/golden/ fixtures generated from that same library
 * (tests/golden/generate.py).  The reference's tests hold no golden files
 * (SURVEY.md section 4): their oracles are inline host loops, restated in
 * tests/test_reference_cases.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (no FMA contraction: the
 * reference's default host build, g++ -O2 without -march, rounds a*b before +).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference/include/RAJA/).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FOR_ALL_TYPES(X) X(i32, int32_t) X(u32, uint32_t) X(i64, int64_t) X(u64, uint64_t) X(f32, float) X(f64, double)

/* util/Operators.hpp:644-656 limits<T>: min() of a floating type is -max, not -inf */
static int32_t lim_min_i32(void) { return INT32_MIN; }  static int32_t lim_max_i32(void) { return INT32_MAX; }
static uint32_t lim_min_u32(void) { return 0; }         static uint32_t lim_max_u32(void) { return UINT32_MAX; }
static int64_t lim_min_i64(void) { return INT64_MIN; }  static int64_t lim_max_i64(void) { return INT64_MAX; }
static uint64_t lim_min_u64(void) { return 0; }         static uint64_t lim_max_u64(void) { return UINT64_MAX; }
static float lim_min_f32(void) { return -3.402823466e+38f; }  static float lim_max_f32(void) { return 3.402823466e+38f; }
static double lim_min_f64(void) { return -1.7976931348623157e+308; } static double lim_max_f64(void) { return 1.7976931348623157e+308; }

/* ---- forall, no reducer: policy/sequential/forall.hpp:58-102 (plain for loop over the
 *      segment) with the bodies of the BASELINE configs ------------------------------ */
#define DEF_FORALL(S, T)                                                                    \
  void orc_triad_##S(T* a, const T* b, const T* c, T s, int64_t n)                          \
  { for (int64_t i = 0; i < n; ++i) a[i] = b[i] + s * c[i]; }                               \
  /* examples/tut_daxpy.cpp:102-104: a[i] += b[i] * c */                                    \
  void orc_daxpy_##S(T* a, const T* b, T s, int64_t n)                                      \
  { for (int64_t i = 0; i < n; ++i) a[i] += b[i] * s; }
FOR_ALL_TYPES(DEF_FORALL)

/* ---- reducers: policy/sequential/reduce.hpp:41-56 (seq_reduce = left fold in iteration
 *      order), combiners pattern/detail/reduce.hpp:67-101 (op_adapter: val = Op(val, v)),
 *      operators util/Operators.hpp:555-594 (minimum: rhs < lhs ? rhs : lhs -- ties keep lhs),
 *      MinLoc/MaxLoc pattern/detail/reduce.hpp:127-179 (compares val only -> first extremum
 *      wins), default loc -1. ---------------------------------------------------------- */
#define DEF_REDUCE(S, T)                                                                    \
  T orc_sum_##S(const T* x, int64_t n)                                                      \
  { T v = 0; for (int64_t i = 0; i < n; ++i) v += x[i]; return v; }                          \
  T orc_min_##S(const T* x, int64_t n)                                                      \
  { T v = lim_max_##S(); for (int64_t i = 0; i < n; ++i) if (x[i] < v) v = x[i]; return v; } \
  T orc_max_##S(const T* x, int64_t n)                                                      \
  { T v = lim_min_##S(); for (int64_t i = 0; i < n; ++i) if (v < x[i]) v = x[i]; return v; } \
  T orc_minloc_##S(const T* x, int64_t n, int64_t* loc)                                     \
  { T v = lim_max_##S(); int64_t l = -1;                                                    \
    for (int64_t i = 0; i < n; ++i) if (x[i] < v) { v = x[i]; l = i; }                       \
    *loc = l; return v; }                                                                   \
  T orc_maxloc_##S(const T* x, int64_t n, int64_t* loc)                                     \
  { T v = lim_min_##S(); int64_t l = -1;                                                    \
    for (int64_t i = 0; i < n; ++i) if (v < x[i]) { v = x[i]; l = i; }                       \
    *loc = l; return v; }                                                                   \
  /* exercises/dot-product_solution.cpp:87-89 */                                            \
  T orc_dot_##S(const T* x, const T* y, int64_t n)                                          \
  { T v = 0; for (int64_t i = 0; i < n; ++i) v += x[i] * y[i]; return v; }
FOR_ALL_TYPES(DEF_REDUCE)

/* ---- scans: policy/sequential/scan.hpp:41-152.  op: 0 plus, 1 minimum, 2 maximum.
 *      in == out allowed (the *_inplace forms, :41-96). --------------------------------- */
#define OP_APPLY(op, a, b) ((op) == 0 ? (a) + (b) : (op) == 1 ? ((b) < (a) ? (b) : (a)) : ((a) < (b) ? (b) : (a)))
#define DEF_SCAN(S, T)                                                                      \
  void orc_inclusive_scan_##S(int op, const T* in, T* out, int64_t n)                       \
  { if (n <= 0) return; T agg = in[0]; out[0] = agg;                                        \
    for (int64_t i = 1; i < n; ++i) { agg = OP_APPLY(op, agg, in[i]); out[i] = agg; } }      \
  void orc_exclusive_scan_##S(int op, const T* in, T* out, int64_t n, T init)               \
  { T agg = init;                                                                           \
    for (int64_t i = 0; i < n; ++i) { T t = in[i]; out[i] = agg; agg = OP_APPLY(op, agg, t); } }
FOR_ALL_TYPES(DEF_SCAN)

/* ---- stable sort of (key, value) pairs: util/sort.hpp:560-690 merge_sort (bottom-up:
 *      insertion sort on 16-element chunks, then pairwise merges doubling from 16,
 *      ping-ponging between the array and one copy), called by
 *      policy/sequential/sort.hpp:95-107 (keys) and :137-157 (pairs, comparing keys only).
 *      Values are opaque 8-byte payloads here (4-byte values are widened by the caller). */
#define DEF_SORT(S, T)                                                                      \
  typedef struct { T k; uint64_t v; } pair_##S;                                             \
  static int less_##S(T a, T b, int desc) { return desc ? (a > b) : (a < b); }              \
  static void insertion_##S(pair_##S* p, int64_t n, int desc)                               \
  { for (int64_t i = 1; i < n; ++i) { pair_##S t = p[i]; int64_t j = i;                      \
      while (j > 0 && less_##S(t.k, p[j - 1].k, desc)) { p[j] = p[j - 1]; --j; }             \
      p[j] = t; } }                                                                         \
  static void merge_##S(const pair_##S* a, const pair_##S* ae, const pair_##S* b,           \
                        const pair_##S* be, pair_##S* d, int desc)                          \
  { while (a != ae && b != be) { if (less_##S(b->k, a->k, desc)) *d++ = *b++; else *d++ = *a++; } \
    while (a != ae) *d++ = *a++;                                                            \
    while (b != be) *d++ = *b++; }                                                          \
  int orc_stable_sort_pairs_##S(T* keys, uint64_t* vals, int64_t n, int desc)               \
  { if (n <= 1) return 0;                                                                   \
    pair_##S* x = (pair_##S*)malloc(sizeof(pair_##S) * (size_t)n);                           \
    pair_##S* y = (pair_##S*)malloc(sizeof(pair_##S) * (size_t)n);                           \
    if (!x || !y) { free(x); free(y); return -1; }                                          \
    for (int64_t i = 0; i < n; ++i) { x[i].k = keys[i]; x[i].v = vals ? vals[i] : 0; }        \
    for (int64_t s = 0; s < n; s += 16) insertion_##S(x + s, (n - s < 16) ? n - s : 16, desc); \
    pair_##S *src = x, *dst = y;                                                            \
    for (int64_t mid = 16; mid < n; mid *= 2) {                                             \
      for (int64_t s = 0; s < n; s += 2 * mid) {                                            \
        int64_t m = s + mid < n ? s + mid : n, f = s + 2 * mid < n ? s + 2 * mid : n;        \
        merge_##S(src + s, src + m, src + m, src + f, dst + s, desc); }                     \
      pair_##S* t = src; src = dst; dst = t; }                                              \
    for (int64_t i = 0; i < n; ++i) { keys[i] = src[i].k; if (vals) vals[i] = src[i].v; }     \
    free(x); free(y); return 0; }
FOR_ALL_TYPES(DEF_SORT)

/* ---- atomics: policy/sequential/atomic.hpp:43-49 (seq_atomic atomicAdd = plain += ,
 *      returns the old value), bodies of exercises/atomic-histogram_solution.cpp:131-134
 *      and examples/pi-reduce_vs_atomic.cpp:105-109. ------------------------------------- */
void orc_histogram_i32(int32_t* hist, const int32_t* keys, int64_t n)
{ for (int64_t i = 0; i < n; ++i) hist[keys[i]] += 1; }
void orc_histogram_u64(uint64_t* hist, const int32_t* keys, int64_t n)
{ for (int64_t i = 0; i < n; ++i) hist[keys[i]] += 1; }
void orc_histogram_f64(double* hist, const int32_t* keys, int64_t n)
{ for (int64_t i = 0; i < n; ++i) hist[keys[i]] += 1.0; }
double orc_pi_atomic(int64_t num_bins)
{ const double dx = 1.0 / (double)num_bins; double pi = 0.0;
  for (int64_t i = 0; i < num_bins; ++i) { double x = ((double)i + 0.5) * dx; pi += dx / (1.0 + x * x); }
  return pi; }
