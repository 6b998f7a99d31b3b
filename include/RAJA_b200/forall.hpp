// forall_impl for b200_exec: replaces policy/cuda/forall.hpp:565-655 (host launch) and :383-542
// (the four forallp_cuda_kernel variants) with the persistent engine of b200/forall.cuh.
// Found by ADL from pattern/forall.hpp:192-195 and CallForall (:662-673).
#ifndef RAJA_B200_FORALL_HPP
#define RAJA_B200_FORALL_HPP

namespace RAJA
{
namespace policy
{
namespace b200
{

// what param_init / param_resolve receive (the reference hands its cudaInfo: MemUtils_CUDA.hpp:380-394).
// Besides the launch shape it collects the device addresses of the reducer results of ONE launch, so that
// param_resolve fetches them all with one batch of copies into one pinned block and ONE stream wait -- the
// reference waits once per reducer (policy/cuda/params/reduce.hpp:46-60).
struct LaunchInfo
{
  int grid;
  int block;
  cudaStream_t stream;
  RAJA::resources::Cuda res;

  // non-null for a b200_multi_exec launch: every reducer result is all-gathered over the ranks before it is fetched
  rb200_comm_t comm = RAJA::b200_detail::tl_multi_launch_comm();
  int world         = comm ? RAJA::b200_detail::tl_multi().world : 1;

  static constexpr int max_results = 16;
  struct Result
  {
    const void* dev;
    unsigned bytes, offset;
  };
  Result results[max_results] = {};
  int n_results    = 0;
  unsigned staged  = 0;        // bytes of the pinned block in use
  void* staging    = nullptr;  // pinned block, filled at the first fetch
  bool fetched     = false;

  // param_init: a result of `bytes` bytes will appear at device address `dev`
  void expect(const void* dev, size_t bytes)
  {
    if (n_results < max_results)
    {
      results[n_results++] = Result {dev, (unsigned)bytes, staged};
      staged += (unsigned)(((bytes + 15) & ~size_t(15)) * (size_t)world);  // multi launch: one copy per rank, rank order
    }
  }
  // param_resolve: host copy of the result at `dev` (valid until this LaunchInfo is destroyed)
  const void* fetch(const void* dev, size_t bytes)
  {
    if (!fetched && n_results > 0)
    {
      staging = rb200_pinned_malloc(staged);
      if (!staging) throw std::runtime_error(std::string("raja_b200: reducer staging: ") + rb200_last_error());
      if (comm)
      {
        // all ranks' results of every reducer, gathered through peer memory in rank order (one small kernel each)
        gathered = rb200_device_malloc(staged, stream);
        if (!gathered) throw std::runtime_error(std::string("raja_b200: reducer staging: ") + rb200_last_error());
        for (int i = 0; i < n_results; ++i)
          RAJA::b200_detail::check_rc(rb200_comm_allgather_small(comm, results[i].dev, (int)stride(i), static_cast<char*>(gathered) + results[i].offset, stream),
                                      "reducer all-gather");
        RAJA::b200_detail::check(cudaMemcpyAsync(staging, gathered, staged, cudaMemcpyDeviceToHost, stream), "reducer readback");
      }
      else
      {
        for (int i = 0; i < n_results; ++i)
          RAJA::b200_detail::check(cudaMemcpyAsync(static_cast<char*>(staging) + results[i].offset, results[i].dev, results[i].bytes,
                                                   cudaMemcpyDeviceToHost, stream),
                                   "reducer readback");
      }
      RAJA::b200_detail::check(cudaStreamSynchronize(stream), "reducer readback");
      fetched = true;
    }
    for (int i = 0; i < n_results; ++i)
      if (results[i].dev == dev) return static_cast<char*>(staging) + results[i].offset;
    if (comm) throw std::runtime_error("raja_b200: a b200_multi_exec launch takes at most 16 reducer parameters");
    // more reducers than slots: fetch this one on its own
    if (!overflow) overflow = rb200_pinned_malloc(64);
    RAJA::b200_detail::check(cudaMemcpyAsync(overflow, dev, bytes, cudaMemcpyDeviceToHost, stream), "reducer readback");
    RAJA::b200_detail::check(cudaStreamSynchronize(stream), "reducer readback");
    return overflow;
  }
  // bytes between the copies of two ranks' results of reducer i (multi launch)
  size_t stride(int i) const { return (results[i].bytes + 15) & ~size_t(15); }
  size_t stride_of(const void* dev) const
  {
    for (int i = 0; i < n_results; ++i)
      if (results[i].dev == dev) return stride(i);
    return 0;
  }
  void* gathered = nullptr;
  ~LaunchInfo()
  {
    if (gathered) rb200_device_free(gathered, stream);
    if (staging) rb200_pinned_free(staging);
    if (overflow) rb200_pinned_free(overflow);
  }
  void* overflow = nullptr;
};

namespace detail
{
// The engine's per-thread body: the user's loop body, the iterator and the parameter pack travel
// together by value; each thread works on its own copy (what thread_privatize does in the reference,
// pattern/detail/privatizer.hpp:45-99), so ValOp accumulators inside the pack are thread-private.
template<typename LoopBody, typename Iterator, typename ForallParam>
struct Body
{
  LoopBody body;
  Iterator begin;
  ForallParam params;
  template<typename Index>
  __device__ __forceinline__ void operator()(Index i)
  {
    if constexpr (RAJA::expt::type_traits::is_ForallParamPack_empty<ForallParam>::value) body(begin[i]);
    else RAJA::expt::invoke_body(params, body, begin[i]);
  }
};

template<typename ExecPol, typename ForallParam>
struct Finalize
{
  template<typename B>
  __device__ __forceinline__ void operator()(B& b) const
  {
    if constexpr (!RAJA::expt::type_traits::is_ForallParamPack_empty<ForallParam>::value)
      RAJA::expt::ParamMultiplexer::parampack_combine(ExecPol {}, b.params);
  }
};
}  // namespace detail

template<typename Iterable, typename LoopBody, size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename ForallParam>
RAJA_INLINE concepts::enable_if_t<resources::EventProxy<resources::Cuda>, RAJA::expt::type_traits::is_ForallParamPack<ForallParam>>
forall_impl(resources::Cuda res, b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const& pol, Iterable&& iter, LoopBody&& loop_body, ForallParam f_params)
{
  using Iterator  = camp::decay<decltype(std::begin(iter))>;
  using LOOP_BODY = camp::decay<LoopBody>;
  using IndexType = camp::decay<decltype(std::distance(std::begin(iter), std::end(iter)))>;
  using EXEC_POL  = b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async>;
  using BodyT     = detail::Body<LOOP_BODY, Iterator, camp::decay<ForallParam>>;
  using FinT      = detail::Finalize<EXEC_POL, camp::decay<ForallParam>>;
  constexpr int UNROLL = 4;
  constexpr bool empty = RAJA::expt::type_traits::is_ForallParamPack_empty<ForallParam>::value;

  Iterator begin      = std::begin(iter);
  Iterator end        = std::end(iter);
  const IndexType len = std::distance(begin, end);

  // only launch if there is something to iterate over (policy/cuda/forall.hpp:598) -- except that a rank with an empty shard
  // still takes part in the exchange of a multi-GPU launch's reducer results (it contributes the identity)
  if (len > 0 || (!empty && RAJA::b200_detail::tl_multi_launch_comm() != nullptr))
  {
    cudaStream_t stream = res.get_stream();
    // b200_exec_grid<BLOCK, GRID>: the caller's grid (never more blocks than units of work); otherwise SMs x resident CTAs
    // User lambdas are opaque.  Without reducers: 32 waves of CTAs (256 CTAs per SM at BLOCK 256), which lets the block scheduler
    // balance bodies of uneven cost (atomics: -12 %, level with one block per 256 indices) and is free for streaming ones
    // (b200/forall.cuh: persistent_grid).  With reducers -- parameters of
    // the pack, or classic / multi reducer objects that announce themselves while the body is copied below -- one wave: every
    // CTA pays the block + grid combine (and a multi-reducer its flush), so fewer CTAs are cheaper.
    constexpr int waves = empty ? 32 : 1;
    int grid = ::b200::forall_grid<(int)BLOCK_SIZE, UNROLL, IndexType, BodyT, FinT>(len, (long long)GRID_SIZE, waves);
    LaunchInfo info {grid, (int)BLOCK_SIZE, stream, res};
    if constexpr (!empty) RAJA::expt::ParamMultiplexer::parampack_init(pol, f_params, info);
    {
      // classic reducer objects captured by the body set themselves up for this launch when the body
      // is copied -- exactly once, inside the scope (RAJA_b200/reduce_classic.hpp); later copies (the
      // by-value kernel argument) just alias that set-up
      size_t dyn_smem = 0;  // what multi-reducers in the body reserved (RAJA_b200/multi_reduce.hpp)
      int captured    = 0;  // reducer objects captured by the body
      auto make_body = [&]() {
        RAJA::b200_detail::LaunchScope scope(info.grid, stream);
        BodyT b {loop_body, begin, f_params};
        dyn_smem = RAJA::b200_detail::tl_launch().dyn_smem;
        captured = RAJA::b200_detail::tl_launch().reducers;
        return b;
      };
      BodyT body = make_body();
      // the captured reducers sized their scratch for info.grid; a smaller launch just leaves slots unused
      if (captured > 0 && waves > 1 && GRID_SIZE == 0) grid = ::b200::forall_grid<(int)BLOCK_SIZE, UNROLL, IndexType, BodyT, FinT>(len, 0, 1);
      RAJA::b200_detail::check(::b200::launch_forall<(int)BLOCK_SIZE, UNROLL, IndexType>(stream, body, len, FinT {}, grid, "forall_kernel", dyn_smem),
                               "forall launch");
    }
    if constexpr (!empty) RAJA::expt::ParamMultiplexer::parampack_resolve(pol, f_params, info);
    if (!Async) res.wait();
  }
  return resources::EventProxy<resources::Cuda>(res);
}

// b200_multi_exec: the same launch on this rank's shard; the reducer parameters of the launch are combined over the ranks
// in param_resolve (LaunchInfo::comm is picked up from the guard below)
template<typename Iterable, typename LoopBody, size_t BLOCK_SIZE, bool Async, typename ForallParam>
RAJA_INLINE concepts::enable_if_t<resources::EventProxy<resources::Cuda>, RAJA::expt::type_traits::is_ForallParamPack<ForallParam>>
forall_impl(resources::Cuda res, b200_multi_exec<BLOCK_SIZE, Async> const&, Iterable&& iter, LoopBody&& loop_body, ForallParam f_params)
{
  const RAJA::b200_detail::MultiState& m = RAJA::b200_detail::multi_state_or_throw("forall<b200_multi_exec>");
  struct Guard
  {
    explicit Guard(rb200_comm_t c) { RAJA::b200_detail::tl_multi_launch_comm() = c; }
    ~Guard() { RAJA::b200_detail::tl_multi_launch_comm() = nullptr; }
  } guard(m.world > 1 ? m.comm : nullptr);
  return forall_impl(res, b200_exec_explicit<BLOCK_SIZE, 0, Async> {}, std::forward<Iterable>(iter), std::forward<LoopBody>(loop_body), f_params);
}

// RAJA::synchronize<b200_synchronize>() (policy/cuda/synchronize.hpp:39-42)
// waits for the streams the back-end has launched on since they were last waited for, not for the whole device
// (RAJA::cuda::synchronize(), policy/cuda/MemUtils_CUDA.hpp:207-231)
RAJA_INLINE void synchronize_impl(const b200_synchronize&)
{
  if (rb200_synchronize_streams() < 0) throw std::runtime_error(std::string("raja_b200: synchronize: ") + rb200_last_error());
}

}  // namespace b200
}  // namespace policy

namespace resources
{
// TypedIndexSet foralls (ExecPolicy<seq_segit, b200_exec<...>>) run on a Cuda resource; the per-segment
// dispatch is the generic one (pattern/forall.hpp:279-299 -> CallForall -> ADL forall_impl above).
// Mirrors util/resource.hpp:95-113.
template<typename ISetIter, size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async>
struct get_resource<ExecPolicy<ISetIter, ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async>>>
{
  using type = camp::resources::Cuda;
};
}  // namespace resources
}  // namespace RAJA

#endif
