// RAJA::WorkPool / WorkGroup / WorkSite for the b200 back-end: many small loops
// (halo pack / unpack, examples/tut_halo-exchange.cpp) enqueued once and run as ONE launch.
// Replaces policy/cuda/WorkGroup/Dispatcher.hpp:36-123 (device function pointers for the dispatcher) and
// policy/cuda/WorkGroup/WorkRunner.hpp:44-411 (ordered / reverse_ordered runners over forall, and the fused
// unordered_cuda_loop_y_block_iter_x_threadblock_average runner).
//
//   RAJA::b200_work<BLOCK, Async = false>          workgroup exec policy
//   RAJA::ordered / RAJA::reverse_ordered          one async b200 forall per loop, one wait at the end
//   RAJA::unordered_b200_fused                     ONE persistent kernel for the whole group
//
// The reference's fused runner launches a (ceil(average length / BLOCK), loops) grid: every loop gets the
// same number of blocks, so short loops idle most of their blocks and long loops stride.  Here the group
// is flattened into TILES of BLOCK x 4 iterations; the host keeps a running tile count per loop at enqueue
// time, `run` ships that prefix array (4 bytes per loop) and launches min(tiles, SMs x resident CTAs) CTAs
// that walk the tiles round-robin: a CTA finds its tile's loop by binary search in the prefix array and makes
// ONE dispatcher call per tile (the holder runs the tile's iterations with four independent accesses in
// flight per thread) -- balanced whatever the mix of loop lengths, no idle blocks, no per-loop launch.
#ifndef RAJA_B200_WORKGROUP_HPP
#define RAJA_B200_WORKGROUP_HPP

#include <mutex>
#include <vector>

namespace RAJA
{
namespace policy
{
namespace b200
{
template<size_t BLOCK_SIZE, bool Async = false>
struct b200_work
    : public RAJA::make_policy_pattern_launch_platform_t<policy_id,
                                                         RAJA::Pattern::workgroup_exec,
                                                         (Async ? RAJA::Launch::async : RAJA::Launch::sync),
                                                         RAJA::Platform::cuda>
{
  static_assert(BLOCK_SIZE % 32 == 0 && BLOCK_SIZE >= 32 && BLOCK_SIZE <= 1024, "BLOCK_SIZE must be a multiple of 32, at most 1024");
};
template<size_t BLOCK_SIZE>
using b200_work_async = b200_work<BLOCK_SIZE, true>;

struct unordered_b200_fused
    : public RAJA::make_policy_pattern_platform_t<policy_id, RAJA::Pattern::workgroup_order, RAJA::Platform::cuda>
{};
}  // namespace b200
}  // namespace policy
using policy::b200::b200_work;
using policy::b200::b200_work_async;
using policy::b200::unordered_b200_fused;

namespace resources
{
template<size_t BLOCK_SIZE, bool Async>
struct get_resource<::RAJA::policy::b200::b200_work<BLOCK_SIZE, Async>>
{
  using type = camp::resources::Cuda;
};
}  // namespace resources

namespace b200_detail
{
// ---- device-side values for the dispatcher (function pointers / vtable objects must be made on the device) ----
template<typename Factory>
__global__ void workgroup_get_value_kernel(typename Factory::value_type* ptr, Factory factory)
{
  *ptr = factory();
}

template<typename Factory>
inline auto workgroup_get_value(Factory&& factory)
{
  using F          = std::decay_t<Factory>;
  using value_type = typename F::value_type;
  static std::mutex mtx;
  const std::lock_guard<std::mutex> lock(mtx);
  value_type* ptr = nullptr;
  check(cudaMallocHost(&ptr, sizeof(value_type)), "workgroup dispatcher value");
  workgroup_get_value_kernel<F><<<1, 1>>>(ptr, factory);
  check(cudaGetLastError(), "workgroup dispatcher value");
  check(cudaDeviceSynchronize(), "workgroup dispatcher value");
  value_type v = *ptr;
  cudaFreeHost(ptr);
  return v;
}

template<typename Factory>
inline auto workgroup_cached_value(Factory&& factory)
{
  static auto value = workgroup_get_value(std::forward<Factory>(factory));  // once per Factory type
  return value;
}

template<typename T, typename Dispatcher_T>
inline const Dispatcher_T* workgroup_device_dispatcher()
{
  static Dispatcher_T dispatcher {
      Dispatcher_T::template makeDispatcher<T>([](auto&& factory) { return workgroup_cached_value(std::forward<decltype(factory)>(factory)); })};
  return &dispatcher;
}

constexpr int WORK_UNROLL = 4;

// ---- what the fused runner stores per loop: segment + body; called once per TILE on the device ----
template<int BLOCK, typename Segment_type, typename LoopBody, typename index_type, typename... Args>
struct HoldB200TileLoop
{
  template<typename segment_in, typename body_in>
  HoldB200TileLoop(segment_in&& segment, body_in&& body) : m_segment(std::forward<segment_in>(segment)), m_body(std::forward<body_in>(body))
  {}

  // tile = which BLOCK x WORK_UNROLL chunk of this loop's iteration space
  RAJA_DEVICE RAJA_INLINE void operator()(unsigned tile, Args... args) const
  {
    const auto begin = m_segment.begin();
    const index_type len(m_segment.end() - begin);
    const index_type base = index_type(tile) * index_type(BLOCK * WORK_UNROLL) + index_type(threadIdx.x);
    if (base + index_type((WORK_UNROLL - 1) * BLOCK) < len)
    {
#pragma unroll
      for (int u = 0; u < WORK_UNROLL; ++u) m_body(begin[base + index_type(u * BLOCK)], args...);
    }
    else
    {
      for (int u = 0; u < WORK_UNROLL; ++u)
      {
        const index_type i = base + index_type(u * BLOCK);
        if (i < len) m_body(begin[i], args...);
      }
    }
  }

private:
  Segment_type m_segment;
  LoopBody m_body;
};

template<int BLOCK, typename StorageIter, typename value_type, typename... Args>
__global__ void __launch_bounds__(BLOCK)
workgroup_fused_kernel(const __grid_constant__ StorageIter iter, const unsigned* __restrict__ tile_end,  // [loops]: tiles through loop l
                       const unsigned num_loops, const unsigned total_tiles, Args... args)
{
  for (unsigned t = blockIdx.x; t < total_tiles; t += gridDim.x)
  {
    // first loop whose running tile count exceeds t (uniform across the CTA: broadcast loads)
    unsigned lo = 0, hi = num_loops - 1;
    while (lo < hi)
    {
      const unsigned mid = (lo + hi) >> 1;
      if (__ldg(tile_end + mid) > t) hi = mid;
      else lo = mid + 1;
    }
    const unsigned first = lo == 0 ? 0u : __ldg(tile_end + lo - 1);
    value_type::device_call(&iter[lo], t - first, args...);
  }
}

// ---- what the ordered runners store per loop: an async b200 forall of the loop ----
template<size_t BLOCK, typename Segment_type, typename LoopBody, typename index_type, typename... Args>
struct HoldB200Forall
{
  using resource_type = resources::Cuda;
  using BodyArgs      = RAJA::detail::HoldBodyArgs_device<LoopBody, index_type, Args...>;

  template<typename segment_in, typename body_in>
  HoldB200Forall(segment_in&& segment, body_in&& body) : m_segment(std::forward<segment_in>(segment)), m_body(std::forward<body_in>(body))
  {}

  RAJA_INLINE void operator()(resource_type r, Args... args) const
  {
    wrap::forall(r, ::RAJA::policy::b200::b200_exec<BLOCK, true>(), m_segment, BodyArgs {m_body, std::forward<Args>(args)...});
  }

private:
  Segment_type m_segment;
  LoopBody m_body;
};
}  // namespace b200_detail

namespace detail
{
// ---- ordered / reverse_ordered: one async forall per loop in (reverse) enqueue order, one wait ----
template<size_t BLOCK_SIZE, bool Async, typename ORDER_POLICY_T, typename DISPATCH_POLICY_T, typename ALLOCATOR_T, typename INDEX_T, typename... Args>
struct B200WorkRunnerForall
{
  using exec_policy     = RAJA::b200_work<BLOCK_SIZE, Async>;
  using order_policy    = ORDER_POLICY_T;
  using dispatch_policy = DISPATCH_POLICY_T;
  using Allocator       = ALLOCATOR_T;
  using index_type      = INDEX_T;
  using resource_type   = resources::Cuda;

  struct holder_type
  {
    template<typename T>
    using type = b200_detail::HoldB200Forall<BLOCK_SIZE,
                                             typename camp::at<T, camp::num<0>>::type,  // segment
                                             typename camp::at<T, camp::num<1>>::type,  // loop body
                                             index_type,
                                             Args...>;
  };
  template<typename T>
  using holder_type_t = typename holder_type::template type<T>;

  // the stored loops are called on the host, each one enqueues a kernel
  using dispatcher_exec_policy   = RAJA::seq_work;
  using dispatcher_holder_policy = dispatcher_transform_types_t<dispatch_policy, holder_type>;
  using dispatcher_type          = Dispatcher<Platform::host, dispatcher_holder_policy, void, resource_type, Args...>;

  B200WorkRunnerForall()                                       = default;
  B200WorkRunnerForall(B200WorkRunnerForall const&)            = delete;
  B200WorkRunnerForall& operator=(B200WorkRunnerForall const&) = delete;
  B200WorkRunnerForall(B200WorkRunnerForall&&)                 = default;
  B200WorkRunnerForall& operator=(B200WorkRunnerForall&&)      = default;

  template<typename WorkContainer, typename segment_T, typename loop_T>
  inline void enqueue(WorkContainer& storage, segment_T&& seg, loop_T&& loop)
  {
    using holder = holder_type_t<camp::list<camp::decay<segment_T>, camp::decay<loop_T>>>;
    storage.template emplace<holder>(RAJA::detail::get_Dispatcher<holder, dispatcher_type>(dispatcher_exec_policy {}), std::forward<segment_T>(seg),
                                     std::forward<loop_T>(loop));
  }

  void clear() {}
  using per_run_storage = int;

  template<typename WorkContainer>
  per_run_storage run(WorkContainer const& storage, resource_type r, Args... args) const
  {
    using value_type = typename WorkContainer::value_type;
    auto begin = storage.begin();
    auto end   = storage.end();
    if (std::is_same<order_policy, RAJA::reverse_ordered>::value)
    {
      for (auto it = end; it != begin; --it) value_type::host_call(&*(it - 1), r, args...);
    }
    else
    {
      for (auto it = begin; it != end; ++it) value_type::host_call(&*it, r, args...);
    }
    if (begin != end && !Async) r.wait();
    return per_run_storage {};
  }
};

template<size_t BLOCK_SIZE, bool Async, typename DISPATCH_POLICY_T, typename ALLOCATOR_T, typename INDEX_T, typename... Args>
struct WorkRunner<RAJA::b200_work<BLOCK_SIZE, Async>, RAJA::ordered, DISPATCH_POLICY_T, ALLOCATOR_T, INDEX_T, Args...>
    : B200WorkRunnerForall<BLOCK_SIZE, Async, RAJA::ordered, DISPATCH_POLICY_T, ALLOCATOR_T, INDEX_T, Args...>
{};
template<size_t BLOCK_SIZE, bool Async, typename DISPATCH_POLICY_T, typename ALLOCATOR_T, typename INDEX_T, typename... Args>
struct WorkRunner<RAJA::b200_work<BLOCK_SIZE, Async>, RAJA::reverse_ordered, DISPATCH_POLICY_T, ALLOCATOR_T, INDEX_T, Args...>
    : B200WorkRunnerForall<BLOCK_SIZE, Async, RAJA::reverse_ordered, DISPATCH_POLICY_T, ALLOCATOR_T, INDEX_T, Args...>
{};

// ---- unordered_b200_fused: one persistent kernel over the tiles of every loop ----
template<size_t BLOCK_SIZE, bool Async, typename DISPATCH_POLICY_T, typename ALLOCATOR_T, typename INDEX_T, typename... Args>
struct WorkRunner<RAJA::b200_work<BLOCK_SIZE, Async>, RAJA::unordered_b200_fused, DISPATCH_POLICY_T, ALLOCATOR_T, INDEX_T, Args...>
{
  using exec_policy     = RAJA::b200_work<BLOCK_SIZE, Async>;
  using order_policy    = RAJA::unordered_b200_fused;
  using dispatch_policy = DISPATCH_POLICY_T;
  using Allocator       = ALLOCATOR_T;
  using index_type      = INDEX_T;
  using resource_type   = resources::Cuda;

  struct holder_type
  {
    template<typename T>
    using type = b200_detail::HoldB200TileLoop<(int)BLOCK_SIZE,
                                               typename camp::at<T, camp::num<0>>::type,  // iterable
                                               typename camp::at<T, camp::num<1>>::type,  // loop body
                                               index_type,
                                               Args...>;
  };
  template<typename T>
  using holder_type_t = typename holder_type::template type<T>;

  // the stored loops are called on the device, once per tile: (tile within the loop, user args...)
  using dispatcher_exec_policy   = exec_policy;
  using dispatcher_holder_policy = dispatcher_transform_types_t<dispatch_policy, holder_type>;
  using dispatcher_type          = Dispatcher<Platform::cuda, dispatcher_holder_policy, RAJA::b200_work<BLOCK_SIZE, true>, unsigned, Args...>;

  WorkRunner()                             = default;
  WorkRunner(WorkRunner const&)            = delete;
  WorkRunner& operator=(WorkRunner const&) = delete;
  WorkRunner(WorkRunner&& o) : m_tile_end(std::move(o.m_tile_end)) { o.m_tile_end.clear(); }
  WorkRunner& operator=(WorkRunner&& o)
  {
    m_tile_end = std::move(o.m_tile_end);
    o.m_tile_end.clear();
    return *this;
  }

  template<typename WorkContainer, typename Iterable, typename LoopBody>
  inline void enqueue(WorkContainer& storage, Iterable&& iter, LoopBody&& loop_body)
  {
    using holder    = holder_type_t<camp::list<camp::decay<Iterable>, camp::decay<LoopBody>>>;
    const auto len  = std::distance(std::begin(iter), std::end(iter));
    if (len > 0)  // zero-length loops store nothing (policy/cuda/WorkGroup/WorkRunner.hpp:322)
    {
      constexpr unsigned long long TILE = (unsigned long long)BLOCK_SIZE * b200_detail::WORK_UNROLL;
      const unsigned long long tiles    = ((unsigned long long)len + TILE - 1) / TILE;
      const unsigned long long through  = (m_tile_end.empty() ? 0ull : (unsigned long long)m_tile_end.back()) + tiles;
      if (through > 0xffffffffull) RAJA_ABORT_OR_THROW("raja_b200: WorkGroup holds more than 2^32 tiles");
      m_tile_end.push_back((unsigned)through);
      storage.template emplace<holder>(b200_detail::workgroup_device_dispatcher<holder, dispatcher_type>(), std::forward<Iterable>(iter),
                                       std::forward<LoopBody>(loop_body));
    }
  }

  using per_run_storage = int;  // the prefix array comes from the stream-ordered pool: returned right after the launch

  template<typename WorkContainer>
  per_run_storage run(WorkContainer const& storage, resource_type r, Args... args) const
  {
    using Iterator   = camp::decay<decltype(std::begin(storage))>;
    using value_type = typename WorkContainer::value_type;
    Iterator begin   = std::begin(storage);
    const size_t num_loops = (size_t)std::distance(begin, std::end(storage));
    if (num_loops > 0)
    {
      if (num_loops != m_tile_end.size()) RAJA_ABORT_OR_THROW("raja_b200: WorkGroup storage and runner disagree on the number of loops");
      cudaStream_t stream = r.get_stream();
      const unsigned total_tiles = m_tile_end.back();
      RAJA::b200_detail::Scratch prefix(num_loops * sizeof(unsigned), stream);
      // pageable source: the driver stages it before returning, so the vector may change afterwards
      RAJA::b200_detail::check(cudaMemcpyAsync(prefix.p, m_tile_end.data(), num_loops * sizeof(unsigned), cudaMemcpyHostToDevice, stream), "workgroup prefix");
      auto kern = &b200_detail::workgroup_fused_kernel<(int)BLOCK_SIZE, Iterator, value_type, Args...>;
      static thread_local ::b200::PerDevice occ_cache;  // per instantiation, host thread and device
      int& occ = occ_cache.current();
      if (occ == 0)
      {
        RAJA::b200_detail::check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)BLOCK_SIZE, 0), "workgroup occupancy");
        if (occ < 1) occ = 1;
      }
      const int grid = ::b200::persistent_grid(occ, (long long)total_tiles);
      kern<<<grid, (unsigned)BLOCK_SIZE, 0, stream>>>(begin, static_cast<const unsigned*>(prefix.p), (unsigned)num_loops, total_tiles, args...);
      RAJA::b200_detail::check(cudaGetLastError(), "workgroup launch");
      if (!Async) r.wait();
    }
    return per_run_storage {};
  }

  void clear() { m_tile_end.clear(); }

private:
  std::vector<unsigned> m_tile_end;  // tiles through loop l, in enqueue order
};

}  // namespace detail
}  // namespace RAJA

#endif
