// Parity of RAJA::MultiReduceSum / Min / Max / BitOr / BitAnd under RAJA::b200_multi_reduce against the
// reference's own seq_multi_reduce on the same inputs and bodies.
// Patterns follow test/functional/forall/multi-reduce-basic/tests/test-forall-basic-MultiReduce.hpp (random
// bins and values, get / get_all / reset, several launches into one reducer) and exercises/atomic-histogram.
#include "RAJA_b200/RAJA.hpp"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>

static int g_fail = 0;
#define EXPECT(cond)                                                         \
  do {                                                                       \
    if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
  } while (0)

template <class T> T* to_dev(const std::vector<T>& h) { T* p; cudaMalloc(&p, h.size() * sizeof(T) + 16); cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice); return p; }

using EXEC = RAJA::b200_exec<256>;
using MR   = RAJA::b200_multi_reduce;
using SEQ  = RAJA::seq_multi_reduce;

template <class T> bool same(T a, T b, double tol) { return tol == 0 ? a == b : std::fabs((double)a - (double)b) <= tol; }

// sum / min / max over random (bin, value) pairs, twice into the same reducers, then reset
template <class T> void test_sum_min_max(long N, int bins, bool exact)
{
  std::mt19937 gen(11 + (unsigned)N + bins);
  std::vector<int> hb(N);
  std::vector<T> hv(N);
  double mag = 0;
  for (long i = 0; i < N; ++i) {
    hb[i] = (int)(gen() % bins);
    hv[i] = exact ? (T)((int)(gen() % 201) - 100) : (T)((gen() % 2000) / 1000.0 - 1.0);
    mag += std::fabs((double)hv[i]);
  }
  int* db = to_dev(hb);
  T* dv = to_dev(hv);
  std::vector<T> init(bins);
  for (int b = 0; b < bins; ++b) init[b] = (T)(b % 5);
  RAJA::MultiReduceSum<MR, T> s(init);
  RAJA::MultiReduceMin<MR, T> mn(bins);
  RAJA::MultiReduceMax<MR, T> mx(bins);
  RAJA::MultiReduceSum<SEQ, T> rs(init);
  RAJA::MultiReduceMin<SEQ, T> rmn(bins);
  RAJA::MultiReduceMax<SEQ, T> rmx(bins);
  for (int rep = 0; rep < 2; ++rep) {
    RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, N), [=] RAJA_DEVICE(long i) {
      s[db[i]] += dv[i];
      mn[db[i]].min(dv[i]);
      mx[db[i]].max(dv[i]);
    });
    RAJA::forall<RAJA::seq_exec>(RAJA::TypedRangeSegment<long>(0, N), [=, &hb, &hv](long i) {
      rs[hb[i]] += hv[i];
      rmn[hb[i]].min(hv[i]);
      rmx[hb[i]].max(hv[i]);
    });
  }
  const double tol = exact ? 0.0 : 2.0 * N * (double)std::numeric_limits<T>::epsilon() * (mag + 1);
  EXPECT((long)s.size() == bins);
  std::vector<T> all(bins);
  s.get_all(all);
  for (int b = 0; b < bins; ++b) {
    EXPECT(same(s.get(b), rs.get(b), tol));
    EXPECT(same(all[b], rs.get(b), tol));
    EXPECT(mn.get(b) == rmn.get(b));
    EXPECT(mx[b].get() == rmx.get(b));
  }
  // reset, one more launch
  s.reset((size_t)bins);
  rs.reset((size_t)bins);
  RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, N), [=] RAJA_DEVICE(long i) { s[db[i]] += dv[i]; });
  RAJA::forall<RAJA::seq_exec>(RAJA::TypedRangeSegment<long>(0, N), [=, &hb, &hv](long i) { rs[hb[i]] += hv[i]; });
  for (int b = 0; b < bins; ++b) EXPECT(same(s.get(b), rs.get(b), tol));
  cudaFree(db); cudaFree(dv);
}

void test_bits(long N, int bins)
{
  std::mt19937 gen(5);
  std::vector<int> hb(N);
  std::vector<unsigned> hv(N);
  for (long i = 0; i < N; ++i) { hb[i] = (int)(gen() % bins); hv[i] = gen() | (gen() & 0xff00ff00u); }
  int* db = to_dev(hb);
  unsigned* dv = to_dev(hv);
  RAJA::MultiReduceBitOr<MR, unsigned> o(bins);
  RAJA::MultiReduceBitAnd<MR, unsigned> a(bins);
  RAJA::MultiReduceBitOr<SEQ, unsigned> ro(bins);
  RAJA::MultiReduceBitAnd<SEQ, unsigned> ra(bins);
  RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, N), [=] RAJA_DEVICE(long i) { o[db[i]] |= dv[i] & 0x0f0f0f0fu; a[db[i]] &= dv[i] | 0x0000ffffu; });
  RAJA::forall<RAJA::seq_exec>(RAJA::TypedRangeSegment<long>(0, N), [=, &hb, &hv](long i) { ro[hb[i]] |= hv[i] & 0x0f0f0f0fu; ra[hb[i]] &= hv[i] | 0x0000ffffu; });
  for (int b = 0; b < bins; ++b) { EXPECT(o.get(b) == ro.get(b)); EXPECT(a.get(b) == ra.get(b)); }
  cudaFree(db); cudaFree(dv);
}

// the histogram of exercises/atomic-histogram with a MultiReduceSum next to a classic reducer and an expt::Reduce
void test_histogram_mixed(long N, int bins)
{
  std::mt19937 gen(99);
  std::vector<int> hk(N);
  std::vector<long long> ref(bins, 0);
  long long total = 0;
  for (long i = 0; i < N; ++i) { hk[i] = (int)(gen() % bins); ++ref[hk[i]]; total += hk[i]; }
  int* dk = to_dev(hk);
  RAJA::MultiReduceSum<MR, long long> hist(bins);
  RAJA::ReduceSum<RAJA::b200_reduce, long long> classic(0);
  long long param = 0;
  RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, N), RAJA::expt::Reduce<RAJA::operators::plus>(&param),
                     [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<long long, RAJA::operators::plus>& p) {
                       hist[dk[i]] += 1;
                       classic += dk[i];
                       p += dk[i];
                     });
  for (int b = 0; b < bins; ++b) EXPECT(hist.get(b) == ref[b]);
  EXPECT(classic.get() == total);
  EXPECT(param == total);
  cudaFree(dk);
}

int main()
{
  for (long N : {1L, 1000L, 300007L}) {
    test_sum_min_max<int>(N, 256, true);
    test_sum_min_max<long long>(N, 10, true);
    test_sum_min_max<unsigned long long>(N, 1, true);
    test_sum_min_max<double>(N, 100, false);
    test_sum_min_max<float>(N, 7, false);
  }
  test_sum_min_max<int>(1 << 20, 65536, true);  // bins do not fit shared memory: global tally
  test_sum_min_max<double>(1 << 20, 20000, false);
  test_bits(100000, 7);
  test_histogram_mixed(1 << 20, 256);
  test_histogram_mixed(1 << 20, 100000);
  {
    RAJA::MultiReduceSum<MR, int> empty;  // no bins: usable, nothing to do
    EXPECT(empty.size() == 0);
    RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, 100), [=] RAJA_DEVICE(long) { (void)empty; });
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  printf(g_fail ? "test_multi_reduce: %d FAILED\n" : "test_multi_reduce: OK\n", g_fail);
  return g_fail ? 1 : 0;
}
