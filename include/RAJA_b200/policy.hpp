// Policy tags of the b200 back-end (mirrors policy/cuda/policy.hpp:347-361, 410-445, 1257-1620).
#ifndef RAJA_B200_POLICY_HPP
#define RAJA_B200_POLICY_HPP

namespace RAJA
{
namespace policy
{
namespace b200
{

// The reference's Policy enum is closed (policy/PolicyBase.hpp:29-39).  Reusing Policy::cuda would make
// the reference's is_cuda_policy-constrained param_* hooks ambiguous with ours, so the family gets a
// value of its own; Platform::cuda keeps plugins and get_platform (internal/get_platform.hpp:73-81) honest.
constexpr ::RAJA::Policy policy_id = static_cast<::RAJA::Policy>(64);

// One mapping: a persistent grid (SMs x resident CTAs) walking BLOCK*UNROLL tiles.  The reference's
// Direct / StridedLoop / occupancy-calculator variants (policy/cuda/policy.hpp:1257-1322) collapse to it.
// GRID_SIZE == 0: the grid is sized from the occupancy of the instantiated kernel (SMs x resident CTAs, never more
// blocks than tiles); GRID_SIZE > 0 fixes it (cuda_exec_grid, policy/cuda/policy.hpp:1276-1290) -- the persistent
// kernel walks its tiles round-robin with whatever grid it is given.
template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async = false>
struct b200_exec_explicit
    : public RAJA::make_policy_pattern_launch_platform_t<policy_id,
                                                         RAJA::Pattern::forall,
                                                         (Async ? RAJA::Launch::async : RAJA::Launch::sync),
                                                         RAJA::Platform::cuda>
{
  static_assert(BLOCK_SIZE % 32 == 0 && BLOCK_SIZE >= 32 && BLOCK_SIZE <= 1024, "BLOCK_SIZE must be a multiple of 32, at most 1024");
  static constexpr size_t block_size = BLOCK_SIZE;
  static constexpr size_t grid_size  = GRID_SIZE;
  static constexpr bool async        = Async;
};
template<size_t BLOCK_SIZE, bool Async = false>
using b200_exec = b200_exec_explicit<BLOCK_SIZE, 0, Async>;

// Multi-GPU range policy: the calling rank -- one process, or one host thread per device -- owns a
// contiguous shard of the global iteration space; kernels run on the shard, and only the paths with a real exchange step
// talk to the other GPUs, through the peer-memory communicator of libraja_b200.so (rb200_comm_*): reducer results are
// all-gathered and folded in rank order, a scan gets the fold of the lower ranks' totals as its carry, a sort is a sample
// sort with a fused partition + all-to-all.  The rank's communicator is the one made current by RAJA::b200::MultiResource
// (RAJA_b200/multi.hpp) on this host thread.  Everything a b200_exec does, it does the same way (it IS one).
template<size_t BLOCK_SIZE, bool Async = false>
struct b200_multi_exec : public b200_exec_explicit<BLOCK_SIZE, 0, Async>
{};

// classic reducer objects (policy/cuda/policy.hpp:1550-1620).  Ordered: last-block-done over the grid's partials in fixed
// slot order (run-to-run deterministic).  Unordered ("atomic"): one device atomic per block on a self-resetting slot;
// value types / operators the ISA has no atomic for fall back to the ordered combine, as the reference's atomic
// tunings do (policy/cuda/reduce.hpp:1001-1030).
template<bool Ordered>
struct b200_reduce_policy
    : public RAJA::make_policy_pattern_launch_platform_t<policy_id, RAJA::Pattern::reduce, RAJA::Launch::undefined, RAJA::Platform::cuda>
{
  static constexpr bool ordered = Ordered;
};
using b200_reduce        = b200_reduce_policy<true>;
using b200_reduce_atomic = b200_reduce_policy<false>;
template<bool with_atomic>
using b200_reduce_base = std::conditional_t<with_atomic, b200_reduce_atomic, b200_reduce>;

// atomics: device code uses hardware atomics, host code falls through to host_policy
// (policy/cuda/policy.hpp:437-445, policy/cuda/atomic.hpp:690-700)
template<typename host_policy>
struct b200_atomic_explicit
{};
using b200_atomic = b200_atomic_explicit<RAJA::seq_atomic>;

struct b200_synchronize
    : public RAJA::make_policy_pattern_launch_t<policy_id, RAJA::Pattern::synchronize, RAJA::Launch::sync>
{};

}  // namespace b200
}  // namespace policy

using policy::b200::b200_atomic;
using policy::b200::b200_atomic_explicit;
using policy::b200::b200_exec;
using policy::b200::b200_exec_explicit;
using policy::b200::b200_multi_exec;
using policy::b200::b200_reduce;
using policy::b200::b200_reduce_atomic;
using policy::b200::b200_reduce_base;
using policy::b200::b200_reduce_policy;
using policy::b200::b200_synchronize;
template<size_t BLOCK_SIZE>
using b200_exec_async = policy::b200::b200_exec_explicit<BLOCK_SIZE, 0, true>;
template<size_t BLOCK_SIZE, size_t GRID_SIZE>
using b200_exec_grid = policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, false>;
template<size_t BLOCK_SIZE, size_t GRID_SIZE>
using b200_exec_grid_async = policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, true>;

namespace resources
{
// primary template: util/resource.hpp:44-48; CUDA specialisations: 72-113
template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async>
struct get_resource<::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async>>
{
  using type = camp::resources::Cuda;
};
template<size_t BLOCK_SIZE, bool Async>
struct get_resource<::RAJA::policy::b200::b200_multi_exec<BLOCK_SIZE, Async>>
{
  using type = camp::resources::Cuda;
};
}  // namespace resources

namespace b200_detail
{
// the rank this host thread acts as (set by RAJA::b200::MultiResource::make_current)
struct MultiState
{
  rb200_comm_t comm = nullptr;
  int rank          = 0;
  int world         = 1;
  size_t ws_bytes   = 0;
  char* ws          = nullptr;  // this rank's own workspace (receive buffers of the sort live there)
};
inline MultiState& tl_multi()
{
  static thread_local MultiState s;
  return s;
}
// non-null while a b200_multi_exec launch is being set up on this thread: the launch's reducer results are combined across ranks
inline rb200_comm_t& tl_multi_launch_comm()
{
  static thread_local rb200_comm_t c = nullptr;
  return c;
}
inline const MultiState& multi_state_or_throw(const char* what)
{
  const MultiState& m = tl_multi();
  if (m.comm == nullptr) throw std::runtime_error(std::string("raja_b200: ") + what + ": no RAJA::b200::MultiResource is current on this thread");
  return m;
}

inline void check(cudaError_t e, const char* what)
{
  if (e != cudaSuccess) throw std::runtime_error(std::string("raja_b200: ") + what + ": " + cudaGetErrorString(e));
}
inline void check_rc(int rc, const char* what)
{
  if (rc != 0) throw std::runtime_error(std::string("raja_b200: ") + what + ": " + rb200_last_error());
}

// RAJA operator -> kernel-side operator.  The arithmetic plus/minimum/maximum map onto the b200 ones
// (same identities: util/Operators.hpp:644-656), which unlocks redux.sync for 32-bit integers; anything
// else is wrapped as is (needs a static identity()).
template<class Fn, class T>
struct WrappedOp
{
  using value_type = T;
  RAJA_HOST_DEVICE static T identity() { return Fn::identity(); }
  RAJA_HOST_DEVICE T operator()(T a, T b) const { return Fn {}(a, b); }
};
template<class Fn, class T, class = void>
struct kernel_op
{
  using type = WrappedOp<Fn, T>;
};
template<class T>
struct kernel_op<RAJA::operators::plus<T, T, T>, T, std::enable_if_t<std::is_arithmetic<T>::value>>
{
  using type = ::b200::Plus<T>;
};
template<class T>
struct kernel_op<RAJA::operators::minimum<T, T, T>, T, std::enable_if_t<std::is_arithmetic<T>::value>>
{
  using type = ::b200::Minimum<T>;
};
template<class T>
struct kernel_op<RAJA::operators::maximum<T, T, T>, T, std::enable_if_t<std::is_arithmetic<T>::value>>
{
  using type = ::b200::Maximum<T>;
};

// The launch a forall_impl is setting up on this host thread.  Classic reducer objects captured in the
// loop body read it in their copy constructor (the role of tl_status + setup_reducers in the reference,
// policy/cuda/MemUtils_CUDA.hpp:192-195, 380-394).
struct LaunchRecord
{
  bool active         = false;
  int grid            = 0;
  cudaStream_t stream = 0;
  size_t dyn_smem     = 0;  // dynamic shared memory reserved so far by objects captured in the body (multi-reducers)
  int reducers        = 0;  // reducer objects (classic, multi) that set themselves up for this launch while the body was copied
};
inline LaunchRecord& tl_launch()
{
  static thread_local LaunchRecord r;
  return r;
}
struct LaunchScope
{
  LaunchScope(int grid, cudaStream_t s)
  {
    LaunchRecord& r = tl_launch();
    r.active   = true;
    r.grid     = grid;
    r.stream   = s;
    r.dyn_smem = 0;
    r.reducers = 0;
  }
  ~LaunchScope() { tl_launch().active = false; }
};

// The header-only engine brackets every launch with b200::launch_hook() (b200/common.cuh).  Kernels instantiated in the
// user's translation unit report to the library through its exported entry points, so they are counted, marked on their
// stream (RAJA::synchronize<b200_synchronize>) and profiled like the library's own -- whatever the symbol binding between
// the executable and the shared library is.
inline void hook_begin(const char* name, cudaStream_t s) { rb200_launch_begin(name, s); }
inline void hook_end(const char* name, cudaStream_t s) { rb200_launch_end(name, s); }
struct HookInstaller
{
  HookInstaller()
  {
    ::b200::launch_hook().begin = hook_begin;
    ::b200::launch_hook().end   = hook_end;
  }
};
static const HookInstaller hook_installer_instance;  // one per translation unit; idempotent

// scratch from the stream-ordered pool, returned on scope exit (safe right after enqueue)
struct Scratch
{
  void* p;
  cudaStream_t s;
  Scratch(size_t bytes, cudaStream_t st) : p(rb200_device_malloc(bytes, st)), s(st)
  {
    if (!p) throw std::runtime_error(std::string("raja_b200: device pool: ") + rb200_last_error());
  }
  ~Scratch() { rb200_device_free(p, s); }
  Scratch(const Scratch&)            = delete;
  Scratch& operator=(const Scratch&) = delete;
};
}  // namespace b200_detail
}  // namespace RAJA

#endif
