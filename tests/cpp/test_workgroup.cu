// RAJA::WorkPool / WorkGroup / WorkSite under RAJA::b200_work: the same enqueued loops
// run under the reference's seq_work and under b200_work (ordered, reverse_ordered, unordered_b200_fused) and the
// arrays they update are compared bit for bit.  Patterns follow test/functional/workgroup/tests/
// test-workgroup-{Ordered,Unordered}-{Single,MultipleReuse}.hpp: three callable types over three element types,
// random sub-ranges, pool and group reuse, every storage policy, indirect-function-call and direct dispatch,
// extra run-time arguments (xargs); plus what the fused runner must balance: a mix of 3-element and
// 300 000-element loops, list segments, an empty group.
#include "RAJA_b200/RAJA.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>

static int g_fail = 0;
#define EXPECT(cond)                                                         \
  do {                                                                       \
    if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
  } while (0)

// pinned storage for the work structs (what test/include/RAJA_test-workgroup.hpp's ResourceAllocator does)
template <class T> struct PinnedAllocator {
  using value_type = T;
  PinnedAllocator() = default;
  template <class U> PinnedAllocator(PinnedAllocator<U> const&) noexcept {}
  T* allocate(size_t n)
  {
    void* p = nullptr;
    if (cudaMallocHost(&p, n * sizeof(T)) != cudaSuccess) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, size_t) noexcept { cudaFreeHost(p); }
  template <class U> friend bool operator==(PinnedAllocator const&, PinnedAllocator<U> const&) { return true; }
  template <class U> friend bool operator!=(PinnedAllocator const&, PinnedAllocator<U> const&) { return false; }
};
using HostAllocator = std::allocator<char>;

template <class I, class T> struct callable1 {
  T* p; T v;
  RAJA_HOST_DEVICE void operator()(I i) const { p[i] += T(i) + v; }
};
template <class I, class T> struct callable2 {
  T* p; T v;
  RAJA_HOST_DEVICE void operator()(I i) const { p[i] = p[i] * T(3) + T(i) + v; }
};
template <class I, class T> struct callable3 {
  T* p; T v;
  RAJA_HOST_DEVICE void operator()(I i) const { p[i] += T(i) * T(0.5) + v; }
};

struct indirect_typer { template <class...> using type = RAJA::indirect_function_call_dispatch; };
struct direct_typer { template <class... Ts> using type = RAJA::direct_dispatch<Ts...>; };

template <class T> struct Mirror {  // device array + host array updated by the seq_work run
  std::vector<T> h;
  T* d = nullptr;
  explicit Mirror(size_t n) : h(n) {
    for (size_t i = 0; i < n; ++i) h[i] = T(i % 17);
    cudaMalloc(&d, n * sizeof(T) + 16);
    cudaMemcpy(d, h.data(), n * sizeof(T), cudaMemcpyHostToDevice);
  }
  ~Mirror() { cudaFree(d); }
  bool same() const {
    std::vector<T> c(h.size());
    cudaMemcpy(c.data(), d, h.size() * sizeof(T), cudaMemcpyDeviceToHost);
    return c == h;
  }
};

// num loops of each of the three callables over random sub-ranges of [lo, hi), pool and group reused
template <class Order, class Storage, class Typer, class I> void test_multiple(std::mt19937& rng, I lo, I hi, int num, int pool_reuse, int group_reuse)
{
  using Seg = RAJA::TypedRangeSegment<I>;
  using C1 = callable1<I, I>; using C2 = callable2<I, size_t>; using C3 = callable3<I, double>;
  using Dispatch = typename Typer::template type<camp::list<Seg, C1>, camp::list<Seg, C2>, camp::list<Seg, C3>>;
  using DevPol = RAJA::WorkGroupPolicy<RAJA::b200_work_async<256>, Order, Storage, Dispatch>;
  using HostOrder = typename std::conditional<std::is_same<Order, RAJA::reverse_ordered>::value, RAJA::reverse_ordered, RAJA::ordered>::type;
  using HostPol = RAJA::WorkGroupPolicy<RAJA::seq_work, HostOrder, Storage, Dispatch>;
  using DevPool = RAJA::WorkPool<DevPol, I, RAJA::xargs<>, PinnedAllocator<char>>;
  using DevGroup = RAJA::WorkGroup<DevPol, I, RAJA::xargs<>, PinnedAllocator<char>>;
  using DevSite = RAJA::WorkSite<DevPol, I, RAJA::xargs<>, PinnedAllocator<char>>;
  using HostPool = RAJA::WorkPool<HostPol, I, RAJA::xargs<>, HostAllocator>;
  using HostGroup = RAJA::WorkGroup<HostPol, I, RAJA::xargs<>, HostAllocator>;
  using HostSite = RAJA::WorkSite<HostPol, I, RAJA::xargs<>, HostAllocator>;

  const size_t N = (size_t)(hi + lo);
  // every loop owns its own slice of the arrays, so unordered execution is race-free
  Mirror<I> a1(N * num); Mirror<size_t> a2(N * num); Mirror<double> a3(N * num);
  DevPool pool(PinnedAllocator<char> {});
  HostPool hpool(HostAllocator {});
  std::uniform_int_distribution<I> db(lo, hi - 1);
  for (int pr = 0; pr < pool_reuse; ++pr) {
    for (int j = 0; j < num; ++j) {
      for (int k = 0; k < 3; ++k) {
        const I b = db(rng);
        const I e = std::uniform_int_distribution<I>(b, hi)(rng);  // may be empty
        const size_t off = (size_t)j * N;
        if (k == 0) { pool.enqueue(Seg {b, e}, C1 {a1.d + off, I(5)}); hpool.enqueue(Seg {b, e}, C1 {a1.h.data() + off, I(5)}); }
        if (k == 1) { pool.enqueue(Seg {b, e}, C2 {a2.d + off, size_t(7)}); hpool.enqueue(Seg {b, e}, C2 {a2.h.data() + off, size_t(7)}); }
        if (k == 2) { pool.enqueue(Seg {b, e}, C3 {a3.d + off, 11.0}); hpool.enqueue(Seg {b, e}, C3 {a3.h.data() + off, 11.0}); }
      }
    }
    EXPECT(pool.num_loops() == hpool.num_loops());
    DevGroup group = pool.instantiate();
    HostGroup hgroup = hpool.instantiate();
    EXPECT(pool.num_loops() == 0);
    for (int gr = 0; gr < group_reuse; ++gr) {
      DevSite site = group.run();
      HostSite hsite = hgroup.run();
      site.get_resource().get_event().wait();
      EXPECT(a1.same());
      EXPECT(a2.same());
      EXPECT(a3.same());
      site.clear();
      hsite.clear();
    }
    group.clear();
    hgroup.clear();
  }
}

// extra run-time arguments, list segments, wildly different loop lengths, synchronous policy
template <class Order> void test_xargs_mixed_lengths()
{
  using I = long;
  using Dispatch = RAJA::indirect_function_call_dispatch;
  using DevPol = RAJA::WorkGroupPolicy<RAJA::b200_work<128>, Order, RAJA::ragged_array_of_objects, Dispatch>;
  using HostPol = RAJA::WorkGroupPolicy<RAJA::seq_work, RAJA::ordered, RAJA::ragged_array_of_objects, Dispatch>;
  using DevPool = RAJA::WorkPool<DevPol, I, RAJA::xargs<double*, long>, PinnedAllocator<char>>;
  using HostPool = RAJA::WorkPool<HostPol, I, RAJA::xargs<double*, long>, HostAllocator>;
  const long lens[] = {3, 300000, 1, 512, 513, 77777, 2, 1024, 100000, 5};
  long total = 0;
  for (long l : lens) total += l;
  Mirror<double> out(total + 300003);
  std::vector<long> idx;  // a list segment over a slice of its own (unordered loops must not overlap): every third index, descending
  for (long i = 299999; i >= 0; i -= 3) idx.push_back(total + 3 + i);
  RAJA::resources::Cuda res;
  RAJA::resources::Host hres;
  DevPool pool(PinnedAllocator<char> {});
  HostPool hpool(HostAllocator {});
  long off = 0;
  for (long l : lens) {
    auto body = [=] RAJA_HOST_DEVICE(long i, double* dst, long scale) { dst[i] = dst[i] * 2.0 + double(i * scale); };
    pool.enqueue(RAJA::TypedRangeSegment<long>(off, off + l), body);
    hpool.enqueue(RAJA::TypedRangeSegment<long>(off, off + l), body);
    off += l;
  }
  auto body2 = [=] RAJA_HOST_DEVICE(long i, double* dst, long scale) { dst[i] += 0.25 * double(scale); };
  // the pool stores non-owning copies of a ListSegment (index/ListSegment.hpp:184-190): keep the owners alive
  RAJA::TypedListSegment<long> dseg(idx.data(), idx.size(), res), hseg(idx.data(), idx.size(), hres);
  pool.enqueue(dseg, body2);
  hpool.enqueue(hseg, body2);
  auto group = pool.instantiate();
  auto hgroup = hpool.instantiate();
  auto site = group.run(out.d, 3L);  // synchronous policy: complete on return
  auto hsite = hgroup.run(out.h.data(), 3L);
  EXPECT(out.same());
  auto site2 = group.run(out.d, 5L);  // run the same group again with other arguments
  auto hsite2 = hgroup.run(out.h.data(), 5L);
  EXPECT(out.same());
  // an empty group runs nothing and does not fail
  pool.clear();
  auto empty = pool.instantiate();
  auto esite = empty.run(out.d, 1L);
  EXPECT(cudaDeviceSynchronize() == cudaSuccess);
}

template <class Order> void all_for_order(std::mt19937& rng)
{
  test_multiple<Order, RAJA::array_of_pointers, indirect_typer, int>(rng, 0, 16, 1, 1, 1);       // Single-style
  test_multiple<Order, RAJA::ragged_array_of_objects, indirect_typer, int>(rng, 1, 1024, 3, 2, 2);
  test_multiple<Order, RAJA::constant_stride_array_of_objects, indirect_typer, long>(rng, 5, 5000, 7, 2, 3);
  test_multiple<Order, RAJA::ragged_array_of_objects, direct_typer, long>(rng, 0, 3000, 4, 2, 2);
  test_multiple<Order, RAJA::array_of_pointers, indirect_typer, long>(rng, 0, 200000, 3, 1, 2);  // many tiles per loop
  test_xargs_mixed_lengths<Order>();
}

int main()
{
  std::mt19937 rng(20250117);
  all_for_order<RAJA::ordered>(rng);
  all_for_order<RAJA::reverse_ordered>(rng);
  all_for_order<RAJA::unordered_b200_fused>(rng);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
  printf("test_workgroup: %s (%d failures)\n", g_fail ? "FAILED" : "OK", g_fail);
  return g_fail ? 1 : 0;
}
