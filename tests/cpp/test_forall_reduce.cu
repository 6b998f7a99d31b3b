// Parity of the RAJA::b200_exec policy family against the reference's own seq_exec on the same inputs
// and the same lambdas: forall (plain, daxpy/triad), expt::Reduce params, classic reducer objects.
// Patterns follow the reference's tests: test/functional/forall/reduce-basic/tests/*.hpp (data rand()%100,
// several segment shapes), exercises/dot-product_solution.cpp, examples/forall-param-reductions.cpp:137-145.
#include "RAJA_b200/RAJA.hpp"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <random>

static int g_fail = 0;
#define EXPECT(cond)                                                         \
  do {                                                                       \
    if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
  } while (0)

template <class T> T* dev_alloc(size_t n) { T* p; cudaMalloc(&p, n * sizeof(T) + 16); return p; }
template <class T> T* to_dev(const std::vector<T>& h) { T* p = dev_alloc<T>(h.size()); cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice); return p; }
template <class T> std::vector<T> to_host(const T* d, size_t n) { std::vector<T> h(n); cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost); return h; }

using EXEC = RAJA::b200_exec<256>;
using EXEC_ASYNC = RAJA::b200_exec_async<128>;

template <class T> void test_elementwise(long N)
{
  std::mt19937_64 gen(1);
  std::uniform_real_distribution<double> u(0, 1);
  std::vector<T> b(N), c(N), a_ref(N), a0(N);
  for (long i = 0; i < N; ++i) { b[i] = (T)u(gen); c[i] = (T)u(gen); a0[i] = (T)u(gen); }
  const T s = (T)3.0;
  T *db = to_dev(b), *dc = to_dev(c), *da = to_dev(a0);
  // triad
  {
    T* pa = a_ref.data(); const T *pb = b.data(), *pc = c.data();
    RAJA::forall<RAJA::seq_exec>(RAJA::TypedRangeSegment<long>(0, N), [=](long i) { pa[i] = pb[i] + s * pc[i]; });
    RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, N), [=] RAJA_DEVICE(long i) { da[i] = db[i] + s * dc[i]; });
    auto h = to_host(da, N);
    // same expression, contraction pinned OFF on both sides (tests/cpp/Makefile: nvcc --fmad=false, g++ -ffp-contract=off):
    // elementwise floating point is bit-exact (SURVEY.md hard part 2)
    long bad = 0;
    for (long i = 0; i < N; ++i) bad += (memcmp(&h[i], &a_ref[i], sizeof(T)) != 0);
    EXPECT(bad == 0);
  }
  // daxpy through a resource + async policy + RangeStrideSegment + negative offsets
  {
    std::vector<T> a1 = a0;
    T* pa = a1.data(); const T* pb = b.data();
    cudaMemcpy(da, a0.data(), N * sizeof(T), cudaMemcpyHostToDevice);
    RAJA::forall<RAJA::seq_exec>(RAJA::TypedRangeStrideSegment<long>(N - 1, -1, -2), [=](long i) { pa[i] += s * pb[i]; });
    RAJA::resources::Cuda res;
    auto ev = RAJA::forall<EXEC_ASYNC>(res, RAJA::TypedRangeStrideSegment<long>(N - 1, -1, -2), [=] RAJA_DEVICE(long i) { da[i] += s * db[i]; });
    RAJA::resources::Event e = ev;
    e.wait();
    auto h = to_host(da, N);
    long bad = 0;
    for (long i = 0; i < N; ++i) bad += (memcmp(&h[i], &a1[i], sizeof(T)) != 0);
    EXPECT(bad == 0);
  }
  cudaFree(da); cudaFree(db); cudaFree(dc);
}

void test_int_reductions(long N)
{
  // test-forall-basic-ReduceSum/Min/Max/MinLoc: data rand()%100 - 50, exact
  std::vector<int> x(N);
  srand(4793);
  for (auto& v : x) v = rand() % 100 - 50;
  if (N > 10) { x[N / 3] = -77; x[N / 2] = -77; x[N - 2] = 91; x[5] = 91; }  // duplicate extrema: first index wins
  int* dx = to_dev(x);
  const int* hx = x.data();
  RAJA::TypedRangeSegment<int> seg(0, (int)N);

  // expt::Reduce params, five at once (examples/forall-param-reductions.cpp:137-145)
  using VL = RAJA::expt::ValLoc<int, int>;
  int s0 = 0, mn0 = 1000, mx0 = -1000; VL lo0(1000, -1), hi0(-1000, -1);
  int s1 = 0, mn1 = 1000, mx1 = -1000; VL lo1(1000, -1), hi1(-1000, -1);
  auto body_host = [=](int i, RAJA::expt::ValOp<int, RAJA::operators::plus>& s, RAJA::expt::ValOp<int, RAJA::operators::minimum>& mn,
                       RAJA::expt::ValOp<int, RAJA::operators::maximum>& mx, RAJA::expt::ValLocOp<int, int, RAJA::operators::minimum>& lo,
                       RAJA::expt::ValLocOp<int, int, RAJA::operators::maximum>& hi) {
    s += hx[i]; mn.min(hx[i]); mx.max(hx[i]); lo.minloc(hx[i], i); hi.maxloc(hx[i], i);
  };
  RAJA::forall<RAJA::seq_exec>(seg, RAJA::expt::Reduce<RAJA::operators::plus>(&s0), RAJA::expt::Reduce<RAJA::operators::minimum>(&mn0),
                               RAJA::expt::Reduce<RAJA::operators::maximum>(&mx0), RAJA::expt::Reduce<RAJA::operators::minimum>(&lo0),
                               RAJA::expt::Reduce<RAJA::operators::maximum>(&hi0), body_host);
  RAJA::forall<EXEC>(seg, RAJA::expt::Reduce<RAJA::operators::plus>(&s1), RAJA::expt::Reduce<RAJA::operators::minimum>(&mn1),
                     RAJA::expt::Reduce<RAJA::operators::maximum>(&mx1), RAJA::expt::Reduce<RAJA::operators::minimum>(&lo1),
                     RAJA::expt::Reduce<RAJA::operators::maximum>(&hi1), RAJA::Name("b200_multi_reduce"),
                     [=] RAJA_DEVICE(int i, RAJA::expt::ValOp<int, RAJA::operators::plus>& s, RAJA::expt::ValOp<int, RAJA::operators::minimum>& mn,
                                     RAJA::expt::ValOp<int, RAJA::operators::maximum>& mx, RAJA::expt::ValLocOp<int, int, RAJA::operators::minimum>& lo,
                                     RAJA::expt::ValLocOp<int, int, RAJA::operators::maximum>& hi) {
                       s += dx[i]; mn.min(dx[i]); mx.max(dx[i]); lo.minloc(dx[i], i); hi.maxloc(dx[i], i);
                     });
  EXPECT(s0 == s1); EXPECT(mn0 == mn1); EXPECT(mx0 == mx1);
  EXPECT(lo0.getVal() == lo1.getVal()); EXPECT(lo0.getLoc() == lo1.getLoc());
  EXPECT(hi0.getVal() == hi1.getVal()); EXPECT(hi0.getLoc() == hi1.getLoc());

  // ReduceLoc with separate value / index targets
  int v0 = 1000, i0 = -1, v1 = 1000, i1 = -1;
  RAJA::forall<RAJA::seq_exec>(seg, RAJA::expt::ReduceLoc<RAJA::operators::minimum>(&v0, &i0),
                               [=](int i, RAJA::expt::ValLocOp<int, int, RAJA::operators::minimum>& lo) { lo.minloc(hx[i], i); });
  RAJA::forall<EXEC>(seg, RAJA::expt::ReduceLoc<RAJA::operators::minimum>(&v1, &i1),
                     [=] RAJA_DEVICE(int i, RAJA::expt::ValLocOp<int, int, RAJA::operators::minimum>& lo) { lo.minloc(dx[i], i); });
  EXPECT(v0 == v1); EXPECT(i0 == i1);

  // classic reducer objects, accumulating over two launches and a reset
  {
    RAJA::ReduceSum<RAJA::seq_reduce, long> hs(5);
    RAJA::ReduceMin<RAJA::seq_reduce, int> hmin(1000);
    RAJA::ReduceMax<RAJA::seq_reduce, int> hmax(-1000);
    RAJA::ReduceMinLoc<RAJA::seq_reduce, int, int> hlo(1000, -1);
    RAJA::ReduceMaxLoc<RAJA::seq_reduce, int, int> hhi(-1000, -1);
    RAJA::ReduceSum<RAJA::b200_reduce, long> ds(5);
    RAJA::ReduceMin<RAJA::b200_reduce, int> dmin(1000);
    RAJA::ReduceMax<RAJA::b200_reduce, int> dmax(-1000);
    RAJA::ReduceMinLoc<RAJA::b200_reduce, int, int> dlo(1000, -1);
    RAJA::ReduceMaxLoc<RAJA::b200_reduce, int, int> dhi(-1000, -1);
    for (int rep = 0; rep < 2; ++rep)
    {
      RAJA::forall<RAJA::seq_exec>(seg, [=](int i) { hs += hx[i]; hmin.min(hx[i]); hmax.max(hx[i]); hlo.minloc(hx[i], i); hhi.maxloc(hx[i], i); });
      RAJA::forall<EXEC>(seg, [=] RAJA_DEVICE(int i) { ds += dx[i]; dmin.min(dx[i]); dmax.max(dx[i]); dlo.minloc(dx[i], i); dhi.maxloc(dx[i], i); });
    }
    EXPECT(hs.get() == ds.get()); EXPECT(hmin.get() == dmin.get()); EXPECT(hmax.get() == dmax.get());
    EXPECT(hlo.get() == dlo.get()); EXPECT(hlo.getLoc() == dlo.getLoc());
    EXPECT(hhi.get() == dhi.get()); EXPECT(hhi.getLoc() == dhi.getLoc());
    ds.reset(0); hs.reset(0);
    RAJA::forall<RAJA::seq_exec>(RAJA::TypedRangeSegment<int>(0, (int)(N / 2)), [=](int i) { hs += hx[i]; });
    RAJA::forall<EXEC>(RAJA::TypedRangeSegment<int>(0, (int)(N / 2)), [=] RAJA_DEVICE(int i) { ds += dx[i]; });
    EXPECT(hs.get() == ds.get());
  }
  cudaFree(dx);
}

void test_dot(long N)
{
  // exercises/dot-product_solution.cpp: a = 1.0, b = 1.0 -> exact; randomized variant within n*eps*sum|x|
  std::vector<double> a(N, 1.0), b(N, 1.0);
  double *da = to_dev(a), *db = to_dev(b);
  double dot = 0.0;
  RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, N), RAJA::expt::Reduce<RAJA::operators::plus>(&dot),
                     [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<double, RAJA::operators::plus>& d) { d += da[i] * db[i]; });
  EXPECT(dot == (double)N);
  RAJA::ReduceSum<RAJA::b200_reduce, double> cdot(0.0);
  RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, N), [=] RAJA_DEVICE(long i) { cdot += da[i] * db[i]; });
  EXPECT(cdot.get() == (double)N);

  std::mt19937_64 gen(12345);
  std::uniform_real_distribution<double> u(0.5, 1.5);
  double abs_sum = 0;
  for (long i = 0; i < N; ++i) { a[i] = u(gen); b[i] = u(gen); abs_sum += a[i] * b[i]; }
  cudaMemcpy(da, a.data(), N * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), N * 8, cudaMemcpyHostToDevice);
  const double *ha = a.data(), *hb = b.data();
  double ref = 0.0;
  RAJA::forall<RAJA::seq_exec>(RAJA::TypedRangeSegment<long>(0, N), RAJA::expt::Reduce<RAJA::operators::plus>(&ref),
                               [=](long i, RAJA::expt::ValOp<double, RAJA::operators::plus>& d) { d += ha[i] * hb[i]; });
  dot = 0.0;
  RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, N), RAJA::expt::Reduce<RAJA::operators::plus>(&dot),
                     [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<double, RAJA::operators::plus>& d) { d += da[i] * db[i]; });
  EXPECT(std::fabs(dot - ref) <= N * std::numeric_limits<double>::epsilon() * abs_sum);
  cudaFree(da); cudaFree(db);
}

void test_list_segment_and_empty()
{
  // ListSegment (gather indices) and an empty range (no launch, targets untouched)
  const int N = 5000;
  std::vector<int> idx;
  for (int i = 0; i < N; i += 3) idx.push_back(i);
  std::vector<int> x(N);
  for (int i = 0; i < N; ++i) x[i] = i % 17;
  int* dx = to_dev(x);
  RAJA::resources::Cuda res;
  RAJA::TypedListSegment<int> lseg(idx.data(), idx.size(), res);
  long s = 0, ref = 0;
  for (int i : idx) ref += x[i];
  RAJA::forall<EXEC>(res, lseg, RAJA::expt::Reduce<RAJA::operators::plus>(&s), [=] RAJA_DEVICE(int i, RAJA::expt::ValOp<long, RAJA::operators::plus>& v) { v += dx[i]; });
  EXPECT(s == ref);
  long e = 42;
  RAJA::forall<EXEC>(RAJA::TypedRangeSegment<int>(3, 3), RAJA::expt::Reduce<RAJA::operators::plus>(&e), [=] RAJA_DEVICE(int i, RAJA::expt::ValOp<long, RAJA::operators::plus>& v) { v += dx[i]; });
  EXPECT(e == 42);
  RAJA::synchronize<RAJA::b200_synchronize>();
  cudaFree(dx);
}

void test_index_set()
{
  // TypedIndexSet of range and list segments, one b200 launch per segment (seq_segit), classic reducers
  // (test/functional/forall/indexset-view, reduce-basic: same bodies under seq_exec and b200_exec)
  using ISet = RAJA::TypedIndexSet<RAJA::TypedRangeSegment<int>, RAJA::TypedListSegment<int>>;
  const int N = 6000;
  std::vector<int> x(N);
  for (int i = 0; i < N; ++i) x[i] = (i * 7) % 23 - 11;
  std::vector<int> idx;
  for (int i = 1000; i < 2000; i += 3) idx.push_back(i);
  RAJA::resources::Cuda res;
  RAJA::resources::Host hres;
  ISet iset, hset;
  iset.push_back(RAJA::TypedRangeSegment<int>(0, 1000));
  iset.push_back(RAJA::TypedListSegment<int>(idx.data(), idx.size(), res));
  iset.push_back(RAJA::TypedRangeSegment<int>(2000, N));
  hset.push_back(RAJA::TypedRangeSegment<int>(0, 1000));
  hset.push_back(RAJA::TypedListSegment<int>(idx.data(), idx.size(), hres));
  hset.push_back(RAJA::TypedRangeSegment<int>(2000, N));
  int* dx = to_dev(x);
  std::vector<int> y(N, -1), yref(N, -1);
  int* dy = to_dev(y);
  RAJA::ReduceSum<RAJA::b200_reduce, long> s(0);
  RAJA::ReduceMin<RAJA::b200_reduce, int> mn(1 << 30);
  RAJA::ReduceSum<RAJA::seq_reduce, long> rs(0);
  RAJA::ReduceMin<RAJA::seq_reduce, int> rmn(1 << 30);
  RAJA::forall<RAJA::ExecPolicy<RAJA::seq_segit, EXEC>>(iset, [=] RAJA_DEVICE(int i) { s += dx[i]; mn.min(dx[i] * 3); dy[i] = 2 * dx[i]; });
  RAJA::forall<RAJA::ExecPolicy<RAJA::seq_segit, RAJA::seq_exec>>(hset, [=, &x, &yref](int i) { rs += x[i]; rmn.min(x[i] * 3); yref[i] = 2 * x[i]; });
  EXPECT(s.get() == rs.get());
  EXPECT(mn.get() == rmn.get());
  EXPECT(to_host(dy, N) == yref);
  cudaFree(dx); cudaFree(dy);
}

int main()
{
  for (long N : {1L, 37L, 1000L, 100003L, 3000017L}) {
    test_elementwise<double>(N);
    test_elementwise<float>(N);
    test_int_reductions(N);
    test_dot(N);
  }
  test_list_segment_and_empty();
  test_index_set();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  printf(g_fail ? "test_forall_reduce: %d FAILED\n" : "test_forall_reduce: OK\n", g_fail);
  return g_fail ? 1 : 0;
}
