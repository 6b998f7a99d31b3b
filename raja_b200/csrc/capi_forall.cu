// raja_b200 C ABI: forall workloads, reducers and atomics.
//
// Each entry point is the b200 forall engine (b200/forall.cuh) instantiated with
// a fixed loop body taken from the reference's examples/exercises (file:line in
// include/raja_b200.h).  The RAJA policy adapter (include/RAJA_b200/) instantiates
// the SAME engine with the user's lambda.
#include "runtime.hpp"
#include "b200/forall.cuh"
#include "b200/atomic.cuh"

using namespace b200;
using namespace rb200;

namespace {

constexpr int BLOCK = 256;
using Index = long long;

// mul/add that ptxas may not contract into an FMA: the reference's host path
// (g++ -O2, no -march) rounds the product, so bit-exact parity needs two roundings
B200_D double mul_add(double b, double s, double c) { return __dadd_rn(b, __dmul_rn(s, c)); }
B200_D float mul_add(float b, float s, float c) { return __fadd_rn(b, __fmul_rn(s, c)); }
template <class T> B200_D T mul_add(T b, T s, T c) { return (T)(b + s * c); }
B200_D double mul_nofma(double a, double b) { return __dmul_rn(a, b); }
B200_D float mul_nofma(float a, float b) { return __fmul_rn(a, b); }
template <class T> B200_D T mul_nofma(T a, T b) { return (T)(a * b); }
B200_D double add_nofma(double a, double b) { return __dadd_rn(a, b); }
B200_D float add_nofma(float a, float b) { return __fadd_rn(a, b); }
template <class T> B200_D T add_nofma(T a, T b) { return (T)(a + b); }

// A V-wide body walks nv = n / V vector indices through operator() / tile<U>();
// the n % V leftover elements go through the engine's tail hook.
template <class T, int V> struct VecSpace {
  Index nv, n;
  VecSpace(Index n_) : nv(n_ / V), n(n_) {}
};

// ---- forall bodies ---------------------------------------------------------------
template <class T, int V> struct TriadBody {
  T* a;
  const T* b;
  const T* c;
  T s;
  Index nv, n;
  B200_D void operator()(Index i) const { tile<1>(i, 0); }
  template <int U> B200_D void tile(Index base, Index stride) const
  {
    using VT = Vec<T, V>;
    VT vb[U], vc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      vb[u] = ld_stream(reinterpret_cast<const VT*>(b) + base + u * stride);
      vc[u] = ld_stream(reinterpret_cast<const VT*>(c) + base + u * stride);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      VT r;
#pragma unroll
      for (int j = 0; j < V; ++j) r.e[j] = mul_add(vb[u].e[j], s, vc[u].e[j]);
      st_stream(reinterpret_cast<VT*>(a) + base + u * stride, r);
    }
  }
  B200_D Index tail_len() const { return n - nv * V; }
  B200_D void tail(Index j) const
  {
    const Index e = nv * V + j;
    a[e] = mul_add(b[e], s, c[e]);
  }
};

template <class T, int V> struct DaxpyBody {
  T* a;
  const T* b;
  T s;
  Index nv, n;
  B200_D void operator()(Index i) const { tile<1>(i, 0); }
  template <int U> B200_D void tile(Index base, Index stride) const
  {
    using VT = Vec<T, V>;
    VT va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      va[u] = ld_stream_coherent(reinterpret_cast<const VT*>(a) + base + u * stride);
      vb[u] = ld_stream(reinterpret_cast<const VT*>(b) + base + u * stride);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      VT r;
#pragma unroll
      for (int j = 0; j < V; ++j) r.e[j] = mul_add(va[u].e[j], s, vb[u].e[j]);  // a + s*b, as a[i] += b[i]*c
      st_stream(reinterpret_cast<VT*>(a) + base + u * stride, r);
    }
  }
  B200_D Index tail_len() const { return n - nv * V; }
  B200_D void tail(Index j) const
  {
    const Index e = nv * V + j;
    a[e] = mul_add(a[e], s, b[e]);
  }
};

// ---- reducer bodies: thread-private accumulator in `acc` ------------------------------
template <class T, class Op, int V> struct ReduceBody {
  const T* x;
  Index nv, n;
  T acc;
  B200_D void operator()(Index i) { tile<1>(i, 0); }
  template <int U> B200_D void tile(Index base, Index stride)
  {
    using VT = Vec<T, V>;
    Op op;
    VT v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_stream(reinterpret_cast<const VT*>(x) + base + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc = op(acc, v[u].e[j]);
    }
  }
  B200_D Index tail_len() const { return n - nv * V; }
  B200_D void tail(Index j) { acc = Op{}(acc, x[nv * V + j]); }
};

template <class T, class Op, int V> struct ReduceLocBody {  // Op = MinLoc / MaxLoc
  const T* x;
  Index nv, n;
  typename Op::value_type acc;
  B200_D void operator()(Index i) { tile<1>(i, 0); }
  template <int U> B200_D void tile(Index base, Index stride)
  {
    using VT = Vec<T, V>;
    Op op;
    VT v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_stream(reinterpret_cast<const VT*>(x) + base + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc = op(acc, {v[u].e[j], (base + u * stride) * V + j});
    }
  }
  B200_D Index tail_len() const { return n - nv * V; }
  B200_D void tail(Index j)
  {
    const Index e = nv * V + j;
    acc = Op{}(acc, {x[e], e});
  }
};

template <class T, int V> struct DotBody {
  const T* x;
  const T* y;
  Index nv, n;
  T acc;
  B200_D void operator()(Index i) { tile<1>(i, 0); }
  template <int U> B200_D void tile(Index base, Index stride)
  {
    using VT = Vec<T, V>;
    VT a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      a[u] = ld_stream(reinterpret_cast<const VT*>(x) + base + u * stride);
      b[u] = ld_stream(reinterpret_cast<const VT*>(y) + base + u * stride);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc = add_nofma(acc, mul_nofma(a[u].e[j], b[u].e[j]));
    }
  }
  B200_D Index tail_len() const { return n - nv * V; }
  B200_D void tail(Index j)
  {
    const Index e = nv * V + j;
    acc = add_nofma(acc, mul_nofma(x[e], y[e]));
  }
};

// sum, min, max, minloc, maxloc in one pass
template <class T> struct MultiVal {
  T sum, mn, mx, mnv, mxv;
  long long mnl, mxl;
};
template <class T> struct MultiOp {
  using value_type = MultiVal<T>;
  B200_HD static value_type identity()
  {
    return {T(0), limits<T>::max(), limits<T>::min(), limits<T>::max(), limits<T>::min(), -1, -1};
  }
  B200_HD value_type operator()(const value_type& a, const value_type& b) const
  {
    value_type r;
    r.sum = a.sum + b.sum;
    r.mn = Minimum<T>{}(a.mn, b.mn);
    r.mx = Maximum<T>{}(a.mx, b.mx);
    ValLoc<T, long long> lo = MinLoc<T, long long>{}({a.mnv, a.mnl}, {b.mnv, b.mnl});
    ValLoc<T, long long> hi = MaxLoc<T, long long>{}({a.mxv, a.mxl}, {b.mxv, b.mxl});
    r.mnv = lo.val;
    r.mnl = lo.loc;
    r.mxv = hi.val;
    r.mxl = hi.loc;
    return r;
  }
};
template <class T, int V> struct MultiBody {
  const T* x;
  Index nv, n;
  MultiVal<T> acc;
  B200_D void one(T v, Index e)
  {
    acc.sum += v;
    if (v < acc.mn) acc.mn = v;
    if (acc.mx < v) acc.mx = v;
    if (v < acc.mnv) { acc.mnv = v; acc.mnl = e; }   // first extremum within a thread: indices ascend
    if (acc.mxv < v) { acc.mxv = v; acc.mxl = e; }
  }
  B200_D void operator()(Index i) { tile<1>(i, 0); }
  template <int U> B200_D void tile(Index base, Index stride)
  {
    using VT = Vec<T, V>;
    VT v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_stream(reinterpret_cast<const VT*>(x) + base + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < V; ++j) one(v[u].e[j], (base + u * stride) * V + j);
    }
  }
  B200_D Index tail_len() const { return n - nv * V; }
  B200_D void tail(Index j) { one(x[nv * V + j], nv * V + j); }
};

// ---- atomic bodies -------------------------------------------------------------------
template <class H, bool AGG> struct HistogramBody {
  H* hist;
  const int32_t* keys;
  B200_D void operator()(Index i) const
  {
    const int32_t k = ld_stream(keys + i);
    if (AGG) (void)atomic_add_aggregated(hist + k, H(1));
    else (void)atomic_add(hist + k, H(1));
  }
};
template <bool AGG> struct PiBody {
  double* pi;
  double dx;
  B200_D void operator()(Index i) const
  {
    const double x = (double(i) + 0.5) * dx;
    const double v = dx / (1.0 + x * x);
    if (AGG) (void)atomic_add_aggregated(pi, v);
    else (void)atomic_add(pi, v);
  }
};

template <class T> __global__ void store_value_kernel(T* p, T v) { *p = v; }

template <class T> bool aligned16(const T* p) { return ((uintptr_t)p & 15u) == 0; }

// Runs `Body` through the engine with a grid reduction on body.acc -> *out.
template <int UNROLL, class Op, class Body>
int run_reduce(cudaStream_t stream, Body body, Index len, typename Op::value_type* out, bool ordered)
{
  using T = typename Op::value_type;
  if (ordered) {
    using Fin = ReduceFinalize<BLOCK, Op, true>;
    const int grid = forall_grid<BLOCK, UNROLL, Index, Body, Fin>(len);
    Scratch partials(POOL_DEVICE, sizeof(T) * (size_t)grid, stream);
    Scratch counter(POOL_DEVICE_ZEROED, sizeof(unsigned), stream);
    if (!partials.p || !counter.p) return RB200_ENOMEM;
    Fin fin{(T*)partials.p, (unsigned*)counter.p, out};
    RB200_CUDA((launch_forall<BLOCK, UNROLL, Index, Body, Fin>(stream, body, len, fin, grid, "forall_reduce_kernel")));
  } else {
    // only instantiated for ops with hardware atomics (see dispatch below)
    return RB200_ENOTSUP;
  }
  return 0;
}

template <int UNROLL, class Op, class Body>
int run_reduce_atomic(cudaStream_t stream, Body body, Index len, typename Op::value_type* out)
{
  using T = typename Op::value_type;
  using Fin = ReduceFinalize<BLOCK, Op, false>;
  const int grid = forall_grid<BLOCK, UNROLL, Index, Body, Fin>(len);
  { LaunchScope sc_("store_value_kernel", stream); store_value_kernel<T><<<1, 1, 0, stream>>>(out, Op::identity()); }
  Fin fin{nullptr, nullptr, out};
  RB200_CUDA((launch_forall<BLOCK, UNROLL, Index, Body, Fin>(stream, body, len, fin, grid, "forall_reduce_kernel")));
  return 0;
}

// reducer-free streaming bodies: oversubscribed grid, as the policy layer launches user lambdas (b200/forall.cuh: persistent_grid)
constexpr int STREAM_WAVES = 32;
template <class T> int triad_t(T* a, const T* b, const T* c, double s, Index n, cudaStream_t st)
{
  constexpr int V = 16 / sizeof(T);
  if (aligned16(a) && aligned16(b) && aligned16(c)) {
    VecSpace<T, V> sp(n);
    TriadBody<T, V> body{a, b, c, (T)s, sp.nv, n};
    const int grid = forall_grid<BLOCK, 4, Index, TriadBody<T, V>, NoFinalize>(sp.nv, 0, STREAM_WAVES);
    RB200_CUDA((launch_forall<BLOCK, 4, Index>(st, body, sp.nv, NoFinalize{}, grid)));
  } else {
    TriadBody<T, 1> body{a, b, c, (T)s, n, n};
    const int grid = forall_grid<BLOCK, 8, Index, TriadBody<T, 1>, NoFinalize>(n, 0, STREAM_WAVES);
    RB200_CUDA((launch_forall<BLOCK, 8, Index>(st, body, n, NoFinalize{}, grid)));
  }
  return 0;
}

template <class T> int daxpy_t(T* a, const T* b, double s, Index n, cudaStream_t st)
{
  constexpr int V = 16 / sizeof(T);
  if (aligned16(a) && aligned16(b)) {
    VecSpace<T, V> sp(n);
    DaxpyBody<T, V> body{a, b, (T)s, sp.nv, n};
    const int grid = forall_grid<BLOCK, 4, Index, DaxpyBody<T, V>, NoFinalize>(sp.nv, 0, STREAM_WAVES);
    RB200_CUDA((launch_forall<BLOCK, 4, Index>(st, body, sp.nv, NoFinalize{}, grid)));
  } else {
    DaxpyBody<T, 1> body{a, b, (T)s, n, n};
    const int grid = forall_grid<BLOCK, 8, Index, DaxpyBody<T, 1>, NoFinalize>(n, 0, STREAM_WAVES);
    RB200_CUDA((launch_forall<BLOCK, 8, Index>(st, body, n, NoFinalize{}, grid)));
  }
  return 0;
}

template <class T, class Op, bool HasAtomic> int reduce_val_t(const T* x, Index n, T* out, int deterministic, cudaStream_t st)
{
  constexpr int V = 16 / sizeof(T);
  if (n == 0) {
    { LaunchScope sc_("store_value_kernel", st); store_value_kernel<T><<<1, 1, 0, st>>>(out, Op::identity()); }
    RB200_CUDA(cudaGetLastError());
    return 0;
  }
  if (!deterministic) {
    if constexpr (HasAtomic) {
      if (aligned16(x)) {
        VecSpace<T, V> sp(n);
        return run_reduce_atomic<4, Op>(st, ReduceBody<T, Op, V>{x, sp.nv, n, Op::identity()}, sp.nv, out);
      }
      return run_reduce_atomic<8, Op>(st, ReduceBody<T, Op, 1>{x, n, n, Op::identity()}, n, out);
    } else {
      return fail(RB200_ENOTSUP, "rb200_reduce: no hardware atomic for this op/type; use deterministic=1");
    }
  }
  if (aligned16(x)) {
    VecSpace<T, V> sp(n);
    return run_reduce<4, Op>(st, ReduceBody<T, Op, V>{x, sp.nv, n, Op::identity()}, sp.nv, out, true);
  }
  return run_reduce<8, Op>(st, ReduceBody<T, Op, 1>{x, n, n, Op::identity()}, n, out, true);
}

template <class T> struct LocOut {
  T* val;
  long long* loc;
};
template <class T> __global__ void split_valloc_kernel(const ValLoc<T, long long>* in, T* val, long long* loc)
{
  *val = in->val;
  if (loc) *loc = in->loc;
}

template <class T, class Op> int reduce_loc_t(const T* x, Index n, T* out_val, long long* out_loc, cudaStream_t st)
{
  constexpr int V = 16 / sizeof(T);
  using VL = typename Op::value_type;
  Scratch tmp(POOL_DEVICE, sizeof(VL), st);
  if (!tmp.p) return RB200_ENOMEM;
  if (n == 0) {
    { LaunchScope sc_("store_value_kernel", st); store_value_kernel<VL><<<1, 1, 0, st>>>((VL*)tmp.p, Op::identity()); }
  } else if (aligned16(x)) {
    VecSpace<T, V> sp(n);
    int rc = run_reduce<4, Op>(st, ReduceLocBody<T, Op, V>{x, sp.nv, n, Op::identity()}, sp.nv, (VL*)tmp.p, true);
    if (rc) return rc;
  } else {
    int rc = run_reduce<8, Op>(st, ReduceLocBody<T, Op, 1>{x, n, n, Op::identity()}, n, (VL*)tmp.p, true);
    if (rc) return rc;
  }
  { LaunchScope sc_("split_valloc_kernel", st); split_valloc_kernel<T><<<1, 1, 0, st>>>((const VL*)tmp.p, out_val, out_loc); }
  RB200_CUDA(cudaGetLastError());
  return 0;
}

template <class T> int reduce_t(int kind, const T* x, Index n, T* out, long long* loc, int det, cudaStream_t st)
{
  constexpr bool fp_or_64 = true;
  (void)fp_or_64;
  switch (kind) {
    case RB200_REDUCE_SUM: return reduce_val_t<T, Plus<T>, true>(x, n, out, det, st);
    case RB200_REDUCE_MIN: return reduce_val_t<T, Minimum<T>, std::is_integral<T>::value>(x, n, out, det, st);
    case RB200_REDUCE_MAX: return reduce_val_t<T, Maximum<T>, std::is_integral<T>::value>(x, n, out, det, st);
    case RB200_REDUCE_MINLOC: return reduce_loc_t<T, MinLoc<T, long long>>(x, n, out, loc, st);
    case RB200_REDUCE_MAXLOC: return reduce_loc_t<T, MaxLoc<T, long long>>(x, n, out, loc, st);
    default: return fail(RB200_EINVAL, "rb200_reduce: bad kind");
  }
}

template <class T> int dot_t(const T* x, const T* y, Index n, T* out, cudaStream_t st)
{
  constexpr int V = 16 / sizeof(T);
  if (n == 0) {
    { LaunchScope sc_("store_value_kernel", st); store_value_kernel<T><<<1, 1, 0, st>>>(out, T(0)); }
    RB200_CUDA(cudaGetLastError());
    return 0;
  }
  if (aligned16(x) && aligned16(y)) {
    VecSpace<T, V> sp(n);
    return run_reduce<4, Plus<T>>(st, DotBody<T, V>{x, y, sp.nv, n, T(0)}, sp.nv, out, true);
  }
  return run_reduce<8, Plus<T>>(st, DotBody<T, 1>{x, y, n, n, T(0)}, n, out, true);
}

template <class T> int multi_t(const T* x, Index n, void* out, cudaStream_t st)
{
  constexpr int V = 16 / sizeof(T);
  using Op = MultiOp<T>;
  using MV = MultiVal<T>;
  if (n == 0) {
    { LaunchScope sc_("store_value_kernel", st); store_value_kernel<MV><<<1, 1, 0, st>>>((MV*)out, Op::identity()); }
    RB200_CUDA(cudaGetLastError());
    return 0;
  }
  if (aligned16(x)) {
    VecSpace<T, V> sp(n);
    return run_reduce<4, Op>(st, MultiBody<T, V>{x, sp.nv, n, Op::identity()}, sp.nv, (MV*)out, true);
  }
  return run_reduce<8, Op>(st, MultiBody<T, 1>{x, n, n, Op::identity()}, n, (MV*)out, true);
}

template <class H> int hist_t(H* hist, const int32_t* keys, Index n, int agg, cudaStream_t st)
{
  if (n <= 0) return 0;
  if (agg) {
    using B = HistogramBody<H, true>;
    const int grid = forall_grid<BLOCK, 8, Index, B, NoFinalize>(n, 0, 32);  // atomic-bound: 32 waves (b200/forall.cuh)
    RB200_CUDA((launch_forall<BLOCK, 8, Index>(st, B{hist, keys}, n, NoFinalize{}, grid)));
  } else {
    using B = HistogramBody<H, false>;
    const int grid = forall_grid<BLOCK, 8, Index, B, NoFinalize>(n, 0, 32);  // atomic-bound: 32 waves (b200/forall.cuh)
    RB200_CUDA((launch_forall<BLOCK, 8, Index>(st, B{hist, keys}, n, NoFinalize{}, grid)));
  }
  return 0;
}

}  // namespace

#define CHECK_ARGS(cond, msg) \
  if (!(cond)) return fail(RB200_EINVAL, msg)

#define DISPATCH_T(t, CALL)                                              \
  switch (t) {                                                           \
    case RB200_I32: { using T = int32_t; return CALL; }                  \
    case RB200_U32: { using T = uint32_t; return CALL; }                 \
    case RB200_I64: { using T = int64_t; return CALL; }                  \
    case RB200_U64: { using T = uint64_t; return CALL; }                 \
    case RB200_F32: { using T = float; return CALL; }                    \
    case RB200_F64: { using T = double; return CALL; }                   \
    default: return fail(RB200_EINVAL, "bad dtype");                     \
  }

extern "C" {

int rb200_forall_triad(rb200_dtype t, void* a, const void* b, const void* c, double s, int64_t n, rb200_stream_t stream)
{
  CHECK_ARGS(n >= 0 && (n == 0 || (a && b && c)), "rb200_forall_triad: bad arguments");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(t, (triad_t<T>((T*)a, (const T*)b, (const T*)c, s, n, st)));
}

int rb200_forall_daxpy(rb200_dtype t, void* a, const void* b, double s, int64_t n, rb200_stream_t stream)
{
  CHECK_ARGS(n >= 0 && (n == 0 || (a && b)), "rb200_forall_daxpy: bad arguments");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(t, (daxpy_t<T>((T*)a, (const T*)b, s, n, st)));
}

int rb200_reduce(rb200_reduce_kind kind, rb200_dtype t, const void* x, int64_t n, void* result_val, int64_t* result_loc,
                 int deterministic, rb200_stream_t stream)
{
  CHECK_ARGS(n >= 0 && result_val && (n == 0 || x), "rb200_reduce: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(t, (reduce_t<T>((int)kind, (const T*)x, n, (T*)result_val, (long long*)result_loc, deterministic, st)));
}

int rb200_reduce_dot(rb200_dtype t, const void* x, const void* y, int64_t n, void* result, rb200_stream_t stream)
{
  CHECK_ARGS(n >= 0 && result && (n == 0 || (x && y)), "rb200_reduce_dot: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(t, (dot_t<T>((const T*)x, (const T*)y, n, (T*)result, st)));
}

int rb200_reduce_multi(rb200_dtype t, const void* x, int64_t n, void* result, rb200_stream_t stream)
{
  CHECK_ARGS(n >= 0 && result && (n == 0 || x), "rb200_reduce_multi: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(t, (multi_t<T>((const T*)x, n, result, st)));
}

size_t rb200_reduce_multi_result_bytes(rb200_dtype t)
{
  switch (t) {
    case RB200_I32: case RB200_U32: case RB200_F32: return sizeof(MultiVal<float>);
    case RB200_I64: case RB200_U64: case RB200_F64: return sizeof(MultiVal<double>);
    default: return 0;
  }
}
size_t rb200_reduce_multi_loc_offset(rb200_dtype t)
{
  switch (t) {
    case RB200_I32: case RB200_U32: case RB200_F32: return offsetof(MultiVal<float>, mnl);
    case RB200_I64: case RB200_U64: case RB200_F64: return offsetof(MultiVal<double>, mnl);
    default: return 0;
  }
}

int rb200_atomic_histogram(rb200_dtype hist_t_, void* hist, const int32_t* keys, int64_t n, int aggregate, rb200_stream_t stream)
{
  CHECK_ARGS(n >= 0 && (n == 0 || (hist && keys)), "rb200_atomic_histogram: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  switch (hist_t_) {
    case RB200_I32: return hist_t<int32_t>((int32_t*)hist, keys, n, aggregate, st);
    case RB200_U32: return hist_t<uint32_t>((uint32_t*)hist, keys, n, aggregate, st);
    case RB200_U64: return hist_t<unsigned long long>((unsigned long long*)hist, keys, n, aggregate, st);
    case RB200_F32: return hist_t<float>((float*)hist, keys, n, aggregate, st);
    case RB200_F64: return hist_t<double>((double*)hist, keys, n, aggregate, st);
    default: return fail(RB200_ENOTSUP, "rb200_atomic_histogram: hist dtype must be I32/U32/U64/F32/F64");
  }
}

int rb200_atomic_pi(double* pi, int64_t num_bins, int aggregate, rb200_stream_t stream)
{
  CHECK_ARGS(num_bins >= 0 && pi, "rb200_atomic_pi: bad arguments");
  if (num_bins == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const double dx = 1.0 / double(num_bins);
  if (aggregate) {
    using B = PiBody<true>;
    const int grid = forall_grid<BLOCK, 4, Index, B, NoFinalize>((Index)num_bins, 0, 4);
    RB200_CUDA((launch_forall<BLOCK, 4, Index>(st, B{pi, dx}, (Index)num_bins, NoFinalize{}, grid)));
  } else {
    using B = PiBody<false>;
    const int grid = forall_grid<BLOCK, 4, Index, B, NoFinalize>((Index)num_bins, 0, 4);
    RB200_CUDA((launch_forall<BLOCK, 4, Index>(st, B{pi, dx}, (Index)num_bins, NoFinalize{}, grid)));
  }
  return 0;
}

}  // extern "C"
