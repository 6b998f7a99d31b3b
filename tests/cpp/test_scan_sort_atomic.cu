// Parity of RAJA::inclusive_scan / exclusive_scan[_inplace], RAJA::sort / stable_sort[_pairs] and the
// RAJA::atomic* functions under the b200 policies against the reference's own seq_exec on the same inputs.
// Patterns follow the reference's tests: test/functional/scan/tests/*.hpp (iota, N in {0, 357, 32000},
// plus/minimum/maximum), test/unit/algorithm/tests/test-algorithm-sort-utils.hpp:610-705 (mt19937, [-N, N]),
// test/functional/forall/atomic-basic/tests/test-forall-atomic-basic.hpp:97-128 (known answers),
// exercises/atomic-histogram_solution.cpp.
#include "RAJA_b200/RAJA.hpp"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <numeric>

static int g_fail = 0;
#define EXPECT(cond)                                                         \
  do {                                                                       \
    if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
  } while (0)

template <class T> T* to_dev(const std::vector<T>& h) { T* p; cudaMalloc(&p, h.size() * sizeof(T) + 16); cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice); return p; }
template <class T> std::vector<T> to_host(const T* d, size_t n) { std::vector<T> h(n); cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost); return h; }

using EXEC = RAJA::b200_exec<256>;

template <class T, class Op> void test_scan(long N, bool exact)
{
  std::vector<T> in(N), ref(N);
  std::mt19937 gen(7 + (unsigned)N);
  for (long i = 0; i < N; ++i) in[i] = exact ? (T)((int)(gen() % 19) - 9) : (T)((gen() % 1000) / 1000.0);
  T* din = to_dev(in);
  T* dout; cudaMalloc(&dout, N * sizeof(T) + 16);
  auto close = [&](const std::vector<T>& a, const std::vector<T>& b) {
    if (exact) return a == b;
    double s = 0;
    for (long i = 0; i < N; ++i) s += std::fabs((double)in[i]);
    for (long i = 0; i < N; ++i) if (std::fabs((double)a[i] - (double)b[i]) > N * (double)std::numeric_limits<T>::epsilon() * (s + 1)) return false;
    return true;
  };
  RAJA::inclusive_scan<RAJA::seq_exec>(RAJA::make_span(in.data(), N), RAJA::make_span(ref.data(), N), Op {});
  RAJA::inclusive_scan<EXEC>(RAJA::make_span(din, N), RAJA::make_span(dout, N), Op {});
  EXPECT(close(to_host(dout, N), ref));
  RAJA::exclusive_scan<RAJA::seq_exec>(RAJA::make_span(in.data(), N), RAJA::make_span(ref.data(), N), Op {}, (T)2);
  RAJA::exclusive_scan<EXEC>(RAJA::make_span(din, N), RAJA::make_span(dout, N), Op {}, (T)2);
  EXPECT(close(to_host(dout, N), ref));
  // in place
  std::vector<T> tmp = in;
  RAJA::inclusive_scan_inplace<RAJA::seq_exec>(RAJA::make_span(tmp.data(), N), Op {});
  cudaMemcpy(dout, in.data(), N * sizeof(T), cudaMemcpyHostToDevice);
  RAJA::inclusive_scan_inplace<EXEC>(RAJA::make_span(dout, N), Op {});
  EXPECT(close(to_host(dout, N), tmp));
  tmp = in;
  RAJA::exclusive_scan_inplace<RAJA::seq_exec>(RAJA::make_span(tmp.data(), N), Op {}, (T)5);
  cudaMemcpy(dout, in.data(), N * sizeof(T), cudaMemcpyHostToDevice);
  RAJA::exclusive_scan_inplace<EXEC>(RAJA::make_span(dout, N), Op {}, (T)5);
  EXPECT(close(to_host(dout, N), tmp));
  cudaFree(din); cudaFree(dout);
}

// an element type the TMA scan is not written for: 16 bytes, with a functor that is associative but NOT commutative
// (composition of the affine maps x -> a x + b over the integers mod 2^64)
struct Affine { unsigned long a, b; bool operator==(const Affine& o) const { return a == o.a && b == o.b; } };
struct Compose { RAJA_HOST_DEVICE Affine operator()(const Affine& f, const Affine& g) const { return Affine {f.a * g.a, f.b * g.a + g.b}; } };
void test_scan_struct(long N)
{
  std::vector<Affine> in(N), ref(N);
  std::mt19937 gen(11 + (unsigned)N);
  for (auto& f : in) f = Affine {(unsigned long)(gen() | 1u), (unsigned long)gen()};
  Affine* din = to_dev(in);
  Affine* dout; cudaMalloc(&dout, N * sizeof(Affine) + 16);
  const Affine seed {3, 7};
  RAJA::inclusive_scan<RAJA::seq_exec>(RAJA::make_span(in.data(), N), RAJA::make_span(ref.data(), N), Compose {});
  RAJA::inclusive_scan<EXEC>(RAJA::make_span(din, N), RAJA::make_span(dout, N), Compose {});
  EXPECT(to_host(dout, N) == ref);
  RAJA::exclusive_scan<RAJA::seq_exec>(RAJA::make_span(in.data(), N), RAJA::make_span(ref.data(), N), Compose {}, seed);
  RAJA::exclusive_scan<EXEC>(RAJA::make_span(din, N), RAJA::make_span(dout, N), Compose {}, seed);
  EXPECT(to_host(dout, N) == ref);
  std::vector<Affine> tmp = in;
  RAJA::inclusive_scan_inplace<RAJA::seq_exec>(RAJA::make_span(tmp.data(), N), Compose {});
  RAJA::inclusive_scan_inplace<EXEC>(RAJA::make_span(din, N), Compose {});
  EXPECT(to_host(din, N) == tmp);
  tmp = in;
  cudaMemcpy(din, in.data(), N * sizeof(Affine), cudaMemcpyHostToDevice);
  RAJA::exclusive_scan_inplace<RAJA::seq_exec>(RAJA::make_span(tmp.data(), N), Compose {}, seed);
  RAJA::exclusive_scan_inplace<EXEC>(RAJA::make_span(din, N), Compose {}, seed);
  EXPECT(to_host(din, N) == tmp);
  cudaFree(din); cudaFree(dout);
}

template <class V> V make_val(long i) { return (V)i; }
struct Val16 { long a; long b; bool operator==(const Val16& o) const { return a == o.a && b == o.b; } };
template <> Val16 make_val<Val16>(long i) { return Val16 {i, ~i}; }
struct Val12 { int a, b, c; bool operator==(const Val12& o) const { return a == o.a && b == o.b && c == o.c; } };  // 4-byte aligned, not a power of two
template <> Val12 make_val<Val12>(long i) { return Val12 {(int)i, (int)(i * 3), (int)-i}; }
struct Val8 { float a, b; bool operator==(const Val8& o) const { return a == o.a && b == o.b; } };  // 8 bytes, 4-byte aligned
template <> Val8 make_val<Val8>(long i) { return Val8 {(float)i, (float)(i & 1023)}; }

template <class K, class V = int> void test_sort(long N)
{
  std::mt19937 gen(42 + (unsigned)N);
  std::vector<K> keys(N);
  for (auto& k : keys) {
    if (std::is_floating_point<K>::value) k = (K)(((double)gen() / 4294967296.0 - 0.5) * 2.0 * N);
    else k = (K)((long long)(gen() % (2 * N + 1)) - (std::is_signed<K>::value ? N : 0));
  }
  if (std::is_floating_point<K>::value && N > 4) { keys[1] = (K)0.0; keys[3] = (K)-0.0; }
  std::vector<V> vals(N);
  for (long i = 0; i < N; ++i) vals[i] = make_val<V>(i);
  // keys only, ascending and descending
  {
    std::vector<K> ref = keys;
    RAJA::stable_sort<RAJA::seq_exec>(RAJA::make_span(ref.data(), N), RAJA::operators::less<K> {});
    K* dk = to_dev(keys);
    RAJA::sort<EXEC>(RAJA::make_span(dk, N), RAJA::operators::less<K> {});
    auto h = to_host(dk, N);
    EXPECT(memcmp(h.data(), ref.data(), N * sizeof(K)) == 0 || std::is_floating_point<K>::value);  // +-0.0 order is checked through pairs
    bool sorted = true;
    for (long i = 1; i < N; ++i) sorted = sorted && !(h[i] < h[i - 1]);
    EXPECT(sorted);
    ref = keys;
    RAJA::stable_sort<RAJA::seq_exec>(RAJA::make_span(ref.data(), N), RAJA::operators::greater<K> {});
    cudaMemcpy(dk, keys.data(), N * sizeof(K), cudaMemcpyHostToDevice);
    RAJA::stable_sort<EXEC>(RAJA::make_span(dk, N), RAJA::operators::greater<K> {});
    h = to_host(dk, N);
    sorted = true;
    for (long i = 1; i < N; ++i) sorted = sorted && !(h[i] > h[i - 1]);
    EXPECT(sorted);
    cudaFree(dk);
  }
  // pairs: stable, bit-exact keys AND values against stable_sort_pairs<seq_exec>
  for (int desc = 0; desc < 2; ++desc) {
    std::vector<K> rk = keys;
    std::vector<V> rv = vals;
    K* dk = to_dev(keys);
    V* dv = to_dev(vals);
    if (desc) {
      RAJA::stable_sort_pairs<RAJA::seq_exec>(RAJA::make_span(rk.data(), N), RAJA::make_span(rv.data(), N), RAJA::operators::greater<K> {});
      RAJA::stable_sort_pairs<EXEC>(RAJA::make_span(dk, N), RAJA::make_span(dv, N), RAJA::operators::greater<K> {});
    } else {
      RAJA::stable_sort_pairs<RAJA::seq_exec>(RAJA::make_span(rk.data(), N), RAJA::make_span(rv.data(), N), RAJA::operators::less<K> {});
      RAJA::sort_pairs<EXEC>(RAJA::make_span(dk, N), RAJA::make_span(dv, N), RAJA::operators::less<K> {});
    }
    auto hk = to_host(dk, N);
    auto hv = to_host(dv, N);
    EXPECT(memcmp(hk.data(), rk.data(), N * sizeof(K)) == 0);
    EXPECT(hv == rv);
    cudaFree(dk); cudaFree(dv);
  }
}

template <class T> void test_atomics(int seglimit)
{
  std::vector<T> init = {(T)0, (T)seglimit, (T)seglimit, (T)0, (T)0, (T)(seglimit + 1), (T)seglimit, (T)0, (T)0, (T)0, (T)0, (T)0};
  T* w = to_dev(init);
  using AT = RAJA::b200_atomic;
  RAJA::forall<EXEC>(RAJA::TypedRangeSegment<int>(0, seglimit), [=] RAJA_HOST_DEVICE(int i) {
    RAJA::atomicAdd<AT>(w + 0, (T)1);
    RAJA::atomicSub<AT>(w + 1, (T)1);
    RAJA::atomicMin<AT>(w + 2, (T)i);
    RAJA::atomicMax<AT>(w + 3, (T)i);
    RAJA::atomicInc<AT>(w + 4);
    RAJA::atomicDec<AT>(w + 5);
    RAJA::atomicExchange<AT>(w + 6, (T)i);
    RAJA::atomicCAS<AT>(w + 7, (T)i, (T)(i + 1));
    RAJA::atomicLoad<AT>(w + 8);
    RAJA::atomicStore<AT>(w + 9, (T)1);
    RAJA::atomicInc<AT>(w + 10, (T)16);
    RAJA::atomicDec<AT>(w + 11, (T)16);
  });
  auto c = to_host(w, init.size());
  EXPECT(c[0] == (T)seglimit); EXPECT(c[1] == (T)0); EXPECT(c[2] == (T)0); EXPECT(c[3] == (T)(seglimit - 1));
  EXPECT(c[4] == (T)seglimit); EXPECT(c[5] == (T)1); EXPECT(c[6] >= (T)0 && c[6] < (T)seglimit);
  EXPECT(c[7] > (T)0 && c[7] <= (T)seglimit); EXPECT(c[8] == (T)0); EXPECT(c[9] == (T)1);
  EXPECT(c[10] == (T)(seglimit % 17)); EXPECT(c[11] == (T)((17 - seglimit % 17) % 17 == 0 ? 0 : 17 - seglimit % 17));
  cudaFree(w);
}

void test_histogram(long N, int bins)
{
  std::vector<int> keys(N), ref(bins, 0);
  std::mt19937 gen(7);
  for (auto& k : keys) { k = (int)(gen() % bins); ++ref[k]; }
  int* dk = to_dev(keys);
  std::vector<int> zero(bins, 0);
  int* dh = to_dev(zero);
  RAJA::forall<EXEC>(RAJA::TypedRangeSegment<long>(0, N), [=] RAJA_DEVICE(long i) { RAJA::atomicAdd<RAJA::b200_atomic>(&dh[dk[i]], 1); });
  EXPECT(to_host(dh, bins) == ref);
  cudaFree(dk); cudaFree(dh);
}

int main()
{
  for (long N : {1L, 357L, 32000L, 1000003L}) {
    test_scan<int, RAJA::operators::plus<int>>(N, true);
    test_scan<int, RAJA::operators::minimum<int>>(N, true);
    test_scan<long, RAJA::operators::maximum<long>>(N, true);
    test_scan<long long, RAJA::operators::plus<long long>>(N, true);
    test_scan<double, RAJA::operators::plus<double>>(N, false);
    test_scan<float, RAJA::operators::maximum<float>>(N, true);
    // element sizes outside 4 / 8 bytes (three-kernel scan of RAJA_b200/scan.hpp)
    test_scan<short, RAJA::operators::plus<short>>(N, true);
    test_scan<signed char, RAJA::operators::maximum<signed char>>(N, true);
    test_scan<unsigned short, RAJA::operators::minimum<unsigned short>>(N, true);
    test_scan_struct(N);
  }
  for (long N : {2L, 33L, 1000L, 9217L, 300007L}) {
    test_sort<int>(N);
    test_sort<unsigned>(N);
    test_sort<long long>(N);
    test_sort<float>(N);
    test_sort<double>(N);
    // 1- / 2-byte keys (widened) and values of other sizes (sorted by position, gathered)
    test_sort<short>(N);
    test_sort<unsigned short, short>(N);
    test_sort<signed char, Val16>(N);
    test_sort<unsigned char, unsigned char>(N);
    test_sort<int, Val16>(N);
    test_sort<double, Val12>(N);
    test_sort<unsigned, Val8>(N);
    test_sort<long long, short>(N);
  }
  test_atomics<int>(1000);
  test_atomics<unsigned>(1000);
  test_atomics<long long>(777);
  test_atomics<unsigned long long>(50);
  test_atomics<float>(1000);
  test_atomics<double>(1000);
  test_histogram(1 << 20, 256);
  test_histogram(1 << 20, 65536);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  printf(g_fail ? "test_scan_sort_atomic: %d FAILED\n" : "test_scan_sort_atomic: OK\n", g_fail);
  return g_fail ? 1 : 0;
}
