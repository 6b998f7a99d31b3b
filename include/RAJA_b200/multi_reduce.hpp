// MultiReduceSum / Min / Max / BitOr / BitAnd <RAJA::b200_multi_reduce, T>: an array of reducers
// indexed at run time (RAJA::MultiReduce*, pattern/multi_reduce.hpp) on the b200 back-end.
// Replaces policy/cuda/multi_reduce.hpp:80-853 (block-then-grid atomic algorithm and its tally data).
//
// This is the framework-visible histogram: a raw RAJA::atomicAdd in a loop body is opaque to the
// back-end (2^28 updates into 256 bins run at the L2 atomic rate: 60 ms), a MultiReduce is not --
// every CTA of the persistent forall grid accumulates into its own copy of the bins in SHARED
// memory and flushes that copy once, at the end of its life, with one global atomic per touched bin.
//
// Protocol (same copy-constructor hooks as the classic reducers, RAJA_b200/reduce_classic.hpp):
//   ROOT    the object the user declared: owns the control block (host values, device tally).
//   LAUNCH  the copy made while forall_impl copies the loop body (launch record active): reserves
//           bins x replication elements of the launch's dynamic shared memory; if the bins do not
//           fit, the launch combines straight into the replicated global tally.
//   ALIAS   further host copies of a LAUNCH copy (the kernel argument): plain pointers.
//   device  the engine's per-thread copy (b200/forall.cuh): the first-level copies of a block
//           initialise the block's shared bins together in their constructor and flush them
//           together in their destructor (both run uniformly for all threads of the block).
// get(bin) waits for the streams used, folds the tally's replications into the host value.
// Replication: shared bins are replicated up to 32x (one column per lane: no same-address conflicts
// inside a warp for few bins), the global tally TALLY_REPL x by block index.
#ifndef RAJA_B200_MULTI_REDUCE_HPP
#define RAJA_B200_MULTI_REDUCE_HPP

#include <set>
#include <vector>

#include "RAJA/pattern/detail/multi_reduce.hpp"
#include "RAJA/pattern/multi_reduce.hpp"

namespace RAJA
{
namespace policy
{
namespace b200
{
struct b200_multi_reduce_tuning_default
{};
template<typename tuning>
struct b200_multi_reduce_policy
    : public RAJA::make_policy_pattern_launch_platform_t<policy_id, RAJA::Pattern::multi_reduce, RAJA::Launch::undefined, RAJA::Platform::cuda>
{};
}  // namespace b200
}  // namespace policy
using b200_multi_reduce = policy::b200::b200_multi_reduce_policy<policy::b200::b200_multi_reduce_tuning_default>;

namespace b200
{
namespace detail
{

// device-side atomic combine for the reference's reduce operators (pattern/detail/reduce.hpp:84-101);
// works on global and shared addresses alike
template<class Op>
struct AtomicCombine;
template<class T>
struct AtomicCombine<RAJA::reduce::sum<T>>
{
  __device__ static void apply(T* p, T v) { RAJA::atomicAdd(RAJA::b200_atomic {}, p, v); }
};
template<class T>
struct AtomicCombine<RAJA::reduce::min<T>>
{
  __device__ static void apply(T* p, T v) { RAJA::atomicMin(RAJA::b200_atomic {}, p, v); }
};
template<class T>
struct AtomicCombine<RAJA::reduce::max<T>>
{
  __device__ static void apply(T* p, T v) { RAJA::atomicMax(RAJA::b200_atomic {}, p, v); }
};
template<class T>
struct AtomicCombine<RAJA::reduce::or_bit<T>>
{
  __device__ static void apply(T* p, T v) { RAJA::atomicOr(RAJA::b200_atomic {}, p, v); }
};
template<class T>
struct AtomicCombine<RAJA::reduce::and_bit<T>>
{
  __device__ static void apply(T* p, T v) { RAJA::atomicAnd(RAJA::b200_atomic {}, p, v); }
};

constexpr int MR_TALLY_REPL      = 8;          // copies of every bin in the global tally (by block index)
constexpr size_t MR_SMEM_BUDGET  = 40 * 1024;  // dynamic shared memory one launch may spend on multi-reducers

template<typename T>
__global__ void mr_fill_kernel(T* p, size_t n, T v)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

template<typename T, typename t_MultiReduceOp, typename tuning>
struct MultiReduceDataB200
{
  using value_type    = T;
  using MultiReduceOp = t_MultiReduceOp;

  struct Control
  {
    std::vector<T> host;          // per bin: initial value combined with everything folded in so far
    T* tally        = nullptr;    // device [num_bins][MR_TALLY_REPL], holds the identity between folds
    bool dirty      = false;      // launches since the last fold
    std::set<cudaStream_t> streams;
  };

  MultiReduceDataB200() = delete;

  template<typename Container, std::enable_if_t<!std::is_same<Container, MultiReduceDataB200>::value>* = nullptr>
  MultiReduceDataB200(Container const& container, T identity)
      : m_ctl(new Control), m_tally(nullptr), m_identity(identity), m_num_bins(0), m_smem_offset(-1), m_smem_repl(0), m_parent(this), m_kind(ROOT),
        m_first_level(false)
  {
    assign(container, identity);
  }

  RAJA_HOST_DEVICE MultiReduceDataB200(MultiReduceDataB200 const& other)
      : m_ctl(other.m_ctl), m_tally(other.m_tally), m_identity(other.m_identity), m_num_bins(other.m_num_bins), m_smem_offset(other.m_smem_offset),
        m_smem_repl(other.m_smem_repl), m_parent(other.m_parent), m_kind(HOST), m_first_level(false)
  {
#if !defined(__CUDA_ARCH__)
    if (other.m_kind == LAUNCH || other.m_kind == ALIAS)
    {
      m_kind = ALIAS;
    }
    else
    {
      ::RAJA::b200_detail::LaunchRecord& lr = ::RAJA::b200_detail::tl_launch();
      if (lr.active && m_num_bins > 0)
      {
        m_kind   = LAUNCH;
        m_parent = nullptr;  // marks the kernel argument: its device copies are the first-level ones
        // shared bins: as many copies as fit the budget, a power of two up to one per lane
        const size_t bin_bytes = sizeof(T) * m_num_bins;
        const size_t left      = lr.dyn_smem < MR_SMEM_BUDGET ? MR_SMEM_BUDGET - lr.dyn_smem : 0;
        int repl               = 0;
        for (int r = 32; r >= 1; r >>= 1)
          if (bin_bytes * r <= left)
          {
            repl = r;
            break;
          }
        if (repl > 0)
        {
          m_smem_offset = (int)lr.dyn_smem;
          m_smem_repl   = repl;
          lr.dyn_smem += (bin_bytes * repl + 15) & ~size_t(15);
          ++lr.reducers;  // every CTA flushes its shared bins once: one wave of CTAs (RAJA_b200/forall.hpp)
        }
        else
        {
          m_smem_offset = -1;  // too many bins: combine into the global tally directly
          m_smem_repl   = 0;
        }
        m_ctl->streams.insert(lr.stream);
        m_ctl->dirty = true;
      }
    }
#else
    m_kind = DEVICE;
    if (other.m_parent == nullptr)
    {
      // first-level device copy (the engine's per-thread copy of the kernel argument): the block's
      // threads initialise the block's shared bins together
      m_parent      = &other;
      m_first_level = true;
      if (m_smem_offset >= 0)
      {
        T* sm = shared_bins();
        for (size_t i = threadIdx.x; i < m_num_bins * (size_t)m_smem_repl; i += blockDim.x) sm[i] = m_identity;
        __syncthreads();
      }
    }
#endif
  }

  MultiReduceDataB200(MultiReduceDataB200&&)                 = delete;
  MultiReduceDataB200& operator=(MultiReduceDataB200 const&) = delete;
  MultiReduceDataB200& operator=(MultiReduceDataB200&&)      = delete;

  RAJA_HOST_DEVICE ~MultiReduceDataB200()
  {
#if !defined(__CUDA_ARCH__)
    if (m_kind == ROOT)
    {
      if (m_ctl)
      {
        for (cudaStream_t s : m_ctl->streams) cudaStreamSynchronize(s);
        if (m_ctl->tally) rb200_device_free(m_ctl->tally, nullptr);
        delete m_ctl;
      }
    }
#else
    if (m_kind == DEVICE && m_first_level && m_smem_offset >= 0)
    {
      // flush the block's shared bins: fold the replications, one global atomic per touched bin
      __syncthreads();
      T* sm = shared_bins();
      for (size_t bin = threadIdx.x; bin < m_num_bins; bin += blockDim.x)
      {
        T v = m_identity;
        for (int r = 0; r < m_smem_repl; ++r) MultiReduceOp {}(v, sm[bin * m_smem_repl + r]);
        if (v != m_identity) AtomicCombine<MultiReduceOp>::apply(m_tally + bin * MR_TALLY_REPL + (blockIdx.x % MR_TALLY_REPL), v);
      }
    }
#endif
  }

  template<typename Container>
  void reset(Container const& container, T identity)
  {
    fold();
    assign(container, identity);
  }

  RAJA_HOST_DEVICE size_t num_bins() const { return m_num_bins; }

  RAJA_HOST_DEVICE T identity() const { return m_identity; }

  RAJA_HOST_DEVICE void combine(size_t bin, T const& val)
  {
#if !defined(__CUDA_ARCH__)
    MultiReduceOp {}(m_ctl->host[bin], val);  // used in a host loop
#else
    if (m_smem_offset >= 0) AtomicCombine<MultiReduceOp>::apply(shared_bins() + bin * m_smem_repl + (threadIdx.x & (m_smem_repl - 1)), val);
    else AtomicCombine<MultiReduceOp>::apply(m_tally + bin * MR_TALLY_REPL + (blockIdx.x % MR_TALLY_REPL), val);
#endif
  }

  T get(size_t bin)
  {
    fold();
    return m_ctl->host[bin];
  }

private:
  enum Kind : int
  {
    ROOT,
    HOST,
    LAUNCH,
    ALIAS,
    DEVICE
  };

  __device__ T* shared_bins() const
  {
    extern __shared__ __align__(16) unsigned char raja_b200_dyn_smem[];
    return reinterpret_cast<T*>(raja_b200_dyn_smem + m_smem_offset);
  }

  template<typename Container>
  void assign(Container const& container, T identity)
  {
    m_identity = identity;
    m_ctl->host.clear();
    for (auto const& v : container) m_ctl->host.push_back(v);
    const size_t nb = m_ctl->host.size();
    if (nb != m_num_bins || !m_ctl->tally)
    {
      if (m_ctl->tally) rb200_device_free(m_ctl->tally, nullptr);
      m_ctl->tally = nullptr;
      m_num_bins   = nb;
      if (nb > 0)
      {
        m_ctl->tally = static_cast<T*>(rb200_device_malloc(sizeof(T) * nb * MR_TALLY_REPL, nullptr));
        if (!m_ctl->tally) throw std::runtime_error(std::string("raja_b200: multi-reducer tally: ") + rb200_last_error());
      }
    }
    m_tally = m_ctl->tally;
    fill_tally();
  }

  void fill_tally()
  {
    if (m_num_bins == 0) return;
    const size_t n = m_num_bins * MR_TALLY_REPL;
    mr_fill_kernel<T><<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, 0>>>(m_tally, n, m_identity);
    ::RAJA::b200_detail::check(cudaStreamSynchronize(0), "multi-reducer tally init");
    m_ctl->dirty = false;
  }

  // wait for the launches, fold the tally into the host values, reset the tally to the identity
  void fold()
  {
    if (!m_ctl->dirty || m_num_bins == 0) return;
    for (cudaStream_t s : m_ctl->streams) ::RAJA::b200_detail::check(cudaStreamSynchronize(s), "multi-reducer get");
    m_ctl->streams.clear();
    std::vector<T> t(m_num_bins * MR_TALLY_REPL);
    ::RAJA::b200_detail::check(cudaMemcpy(t.data(), m_tally, sizeof(T) * t.size(), cudaMemcpyDeviceToHost), "multi-reducer tally read");
    for (size_t bin = 0; bin < m_num_bins; ++bin)
      for (int r = 0; r < MR_TALLY_REPL; ++r) MultiReduceOp {}(m_ctl->host[bin], t[bin * MR_TALLY_REPL + r]);
    fill_tally();
  }

  Control* m_ctl;
  T* m_tally;
  T m_identity;
  size_t m_num_bins;
  int m_smem_offset;  // bytes into the launch's dynamic shared memory, -1 = no shared bins
  int m_smem_repl;    // power of two
  const MultiReduceDataB200* m_parent;
  Kind m_kind;
  bool m_first_level;  // device: this thread's copy of the kernel argument (initialises / flushes the shared bins)
};

}  // namespace detail
}  // namespace b200

RAJA_DECLARE_ALL_MULTI_REDUCERS(policy::b200::b200_multi_reduce_policy, b200::detail::MultiReduceDataB200)

}  // namespace RAJA

#endif
