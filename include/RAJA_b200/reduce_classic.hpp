// Classic reducer objects for the b200 back-end: ReduceSum / ReduceMin / ReduceMax / ReduceMinLoc /
// ReduceMaxLoc / ReduceBitOr / ReduceBitAnd <RAJA::b200_reduce, T[, IndexType]>.
// Replaces policy/cuda/reduce.hpp:968-1378 (cuda::Reduce, its data classes :678-965 and PinnedTally :474-661).
//
// Protocol (same idea as the reference, much less machinery):
//   ROOT     the object the user declared.  Owns a control block with one pinned, device-visible tally.
//   LAUNCH   the copy made while forall_impl copies the loop body for a launch (the launch record of
//            RAJA_b200/policy.hpp is active): takes per-launch scratch -- one partial per block of the
//            persistent grid and a self-resetting counter -- from the stream-ordered pools, and returns
//            it when it is destroyed (safe right after enqueue).
//   ALIAS    any further host copy of a LAUNCH copy (the by-value kernel argument): just carries pointers.
//   HOST     any other host copy: starts at the identity, folds into its parent on destruction.
//   device   every thread's private copy (made by the engine, b200/forall.cuh): starts at the identity;
//            its destructor runs the block reduce and the last-block-done grid combine, and the last
//            block folds the launch's result into the pinned tally.
// get() waits for the streams that were used, folds the tally into the host value and returns it.
// One reducer may be used in launches on several streams only if those launches are ordered
// (the tally is one slot, not the reference's per-stream list).
#ifndef RAJA_B200_REDUCE_CLASSIC_HPP
#define RAJA_B200_REDUCE_CLASSIC_HPP

#include <vector>

namespace RAJA
{
namespace b200
{
namespace detail
{

// block reduce for a block size only known at run time (a destructor has no policy to read it from)
template<class Op>
__device__ __forceinline__ typename Op::value_type block_reduce_rt(typename Op::value_type v, Op op, typename Op::value_type* smem)
{
  using T = typename Op::value_type;
  v = ::b200::warp_allreduce(v, op);
  const int nw = (blockDim.x + 31) >> 5;
  if (nw == 1) return v;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) smem[w] = v;
  __syncthreads();
  if (w == 0)
  {
    v = (l < nw) ? smem[l] : Op::identity();
    v = ::b200::warp_allreduce(v, op);
  }
  return v;
}

template<class Op>
class Reduce
{
public:
  using T = typename Op::value_type;

  Reduce() : Reduce(T(), Op::identity()) {}

  explicit Reduce(T init_val, T identity_ = Op::identity())
      : m_parent(this), m_ctl(new Control), m_value(init_val), m_identity(identity_), m_partials(nullptr), m_counter(nullptr), m_kind(ROOT)
  {
    m_tally = static_cast<T*>(rb200_pinned_malloc(sizeof(T)));  // mapped pinned memory: one address for host and device
    if (!m_tally) throw std::runtime_error(std::string("raja_b200: reducer tally: ") + rb200_last_error());
    *m_tally = m_identity;
  }

  void reset(T in_val, T identity_ = Op::identity())
  {
    (void)get();  // syncs outstanding launches
    m_value    = in_val;
    m_identity = identity_;
    *m_tally   = m_identity;
  }

  RAJA_HOST_DEVICE Reduce(const Reduce& other)
      : m_parent(other.m_parent), m_ctl(other.m_ctl), m_value(other.m_identity), m_identity(other.m_identity), m_partials(other.m_partials),
        m_counter(other.m_counter), m_kind(HOST), m_stream(other.m_stream), m_tally(other.m_tally)
  {
#if !defined(__CUDA_ARCH__)
    if (other.m_kind == LAUNCH || other.m_kind == ALIAS)
    {
      m_kind = ALIAS;
    }
    else
    {
      ::RAJA::b200_detail::LaunchRecord& lr = ::RAJA::b200_detail::tl_launch();
      if (lr.active)
      {
        m_kind     = LAUNCH;
        m_stream   = lr.stream;
        m_partials = static_cast<T*>(rb200_device_malloc(sizeof(T) * (size_t)lr.grid, lr.stream));
        m_counter  = static_cast<unsigned*>(rb200_device_zeroed_malloc(sizeof(unsigned), lr.stream));
        if (!m_partials || !m_counter) throw std::runtime_error(std::string("raja_b200: reducer scratch: ") + rb200_last_error());
        m_ctl->note_stream(lr.stream);
      }
      else if (other.m_kind == ROOT || other.m_kind == HOST)
      {
        m_parent = &other;  // fold into the object this was copied from
      }
    }
#else
    m_kind = DEVICE;
#endif
  }

  Reduce& operator=(const Reduce&) = delete;

  RAJA_HOST_DEVICE ~Reduce()
  {
#if !defined(__CUDA_ARCH__)
    switch (m_kind)
    {
      case ROOT:
        if (m_tally) rb200_pinned_free(m_tally);
        delete m_ctl;
        break;
      case LAUNCH:
        rb200_device_free(m_partials, m_stream);
        rb200_device_zeroed_free(m_counter, m_stream);
        break;
      case HOST:
        if (m_parent && m_parent != this) m_parent->combine(m_value);
        break;
      default: break;
    }
#else
    // every thread of the block gets here (the engine's private copy goes out of scope once per thread)
    __shared__ alignas(16) unsigned char raw[32 * sizeof(T)];
    T* smem = reinterpret_cast<T*>(raw);
    Op op;
    T v = block_reduce_rt(m_value, op, smem);
    __shared__ bool s_last;
    if (threadIdx.x == 0)
    {
      ::b200::st_cg(m_partials + blockIdx.x, v);
      __threadfence();
      const unsigned ticket = ::atomicInc(m_counter, gridDim.x - 1);  // wraps back to 0: scratch is returned zeroed
      s_last                = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last)
    {
      __threadfence();
      T acc = Op::identity();
      for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) acc = op(acc, ::b200::ld_cg(m_partials + i));
      __syncthreads();
      acc = block_reduce_rt(acc, op, smem);
      if (threadIdx.x == 0)
      {
        *m_tally = op(*m_tally, acc);  // launches that share a reducer are stream-ordered
      }
    }
    __syncthreads();
#endif
  }

  //! synchronises the streams this reducer was used on, folds the device tally in, returns the value
  T get()
  {
    Control* c = m_ctl;
    if (!c->streams.empty())
    {
      for (cudaStream_t s : c->streams) ::RAJA::b200_detail::check(cudaStreamSynchronize(s), "reducer get");
      c->streams.clear();
      Op op;
      m_value  = op(m_value, *m_tally);
      *m_tally = m_identity;
    }
    return m_value;
  }
  operator T() { return get(); }

  RAJA_HOST_DEVICE void combine(T other) const { m_value = Op {}(m_value, other); }
  T& local() const { return m_value; }
  T get_combined() const { return m_value; }

private:
  enum Kind : int { ROOT, HOST, LAUNCH, ALIAS, DEVICE };
  struct Control  // shared by the root and its copies: which streams have launches in flight
  {
    std::vector<cudaStream_t> streams;
    void note_stream(cudaStream_t s)
    {
      for (cudaStream_t t : streams)
        if (t == s) return;
      streams.push_back(s);
    }
  };
  const Reduce* m_parent;
  Control* m_ctl;
  mutable T m_value;
  T m_identity;
  T* m_partials;
  unsigned* m_counter;
  Kind m_kind;
  cudaStream_t m_stream = 0;
  T* m_tally            = nullptr;  // pinned, device-visible running result of the launches so far
};

}  // namespace detail
}  // namespace b200

namespace b200_detail
{
// (value, location) pairs of the classic MinLoc / MaxLoc objects (pattern/detail/reduce.hpp:127-179):
// ties go to the smaller location when the location is an integer (first extremum, as seq_exec)
template<class T, class I, bool doing_min>
struct ValueLocOp
{
  using value_type = RAJA::reduce::detail::ValueLoc<T, I, doing_min>;
  RAJA_HOST_DEVICE static value_type identity()
  {
    return value_type(doing_min ? RAJA::operators::limits<T>::max() : RAJA::operators::limits<T>::min(), RAJA::reduce::detail::DefaultLoc<I>().value());
  }
  RAJA_HOST_DEVICE value_type operator()(value_type a, value_type b) const
  {
    const T av = a, bv = b;
    const bool better = doing_min ? (bv < av) : (av < bv);
    const bool worse  = doing_min ? (av < bv) : (bv < av);
    bool take_b       = better;
    if constexpr (std::is_integral<I>::value)
    {
      const I none = RAJA::reduce::detail::DefaultLoc<I>().value();
      if (!better && !worse) take_b = (b.getLoc() != none) && (a.getLoc() == none || b.getLoc() < a.getLoc());
    }
    return take_b ? b : a;
  }
};
template<class T> struct BitOrOp
{
  using value_type = T;
  RAJA_HOST_DEVICE static T identity() { return T(0); }
  RAJA_HOST_DEVICE T operator()(T a, T b) const { return a | b; }
};
template<class T> struct BitAndOp
{
  using value_type = T;
  RAJA_HOST_DEVICE static T identity() { return ~T(0); }
  RAJA_HOST_DEVICE T operator()(T a, T b) const { return a & b; }
};
}  // namespace b200_detail

// ---- the specialisations user code names (primary templates: pattern/reduce.hpp:80-233) -----------------
template<typename T>
class ReduceSum<b200_reduce, T> : public b200::detail::Reduce<::b200::Plus<T>>
{
public:
  using Base = b200::detail::Reduce<::b200::Plus<T>>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceSum& operator+=(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<typename T>
class ReduceMin<b200_reduce, T> : public b200::detail::Reduce<::b200::Minimum<T>>
{
public:
  using Base = b200::detail::Reduce<::b200::Minimum<T>>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceMin& min(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<typename T>
class ReduceMax<b200_reduce, T> : public b200::detail::Reduce<::b200::Maximum<T>>
{
public:
  using Base = b200::detail::Reduce<::b200::Maximum<T>>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceMax& max(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<typename T>
class ReduceBitOr<b200_reduce, T> : public b200::detail::Reduce<b200_detail::BitOrOp<T>>
{
public:
  using Base = b200::detail::Reduce<b200_detail::BitOrOp<T>>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceBitOr& operator|=(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<typename T>
class ReduceBitAnd<b200_reduce, T> : public b200::detail::Reduce<b200_detail::BitAndOp<T>>
{
public:
  using Base = b200::detail::Reduce<b200_detail::BitAndOp<T>>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceBitAnd& operator&=(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<typename T, typename IndexType>
class ReduceMinLoc<b200_reduce, T, IndexType> : public b200::detail::Reduce<b200_detail::ValueLocOp<T, IndexType, true>>
{
public:
  using Op         = b200_detail::ValueLocOp<T, IndexType, true>;
  using value_type = typename Op::value_type;
  using Base       = b200::detail::Reduce<Op>;
  ReduceMinLoc(T init_val, IndexType init_idx, T identity_val = RAJA::operators::limits<T>::max(),
               IndexType identity_idx = RAJA::reduce::detail::DefaultLoc<IndexType>().value())
      : Base(value_type(init_val, init_idx), value_type(identity_val, identity_idx))
  {}
  void reset(T init_val, IndexType init_idx, T identity_val = RAJA::operators::limits<T>::max(),
             IndexType identity_idx = RAJA::reduce::detail::DefaultLoc<IndexType>().value())
  {
    Base::reset(value_type(init_val, init_idx), value_type(identity_val, identity_idx));
  }
  RAJA_HOST_DEVICE const ReduceMinLoc& minloc(T rhs, IndexType loc) const
  {
    this->combine(value_type(rhs, loc));
    return *this;
  }
  IndexType getLoc() { return Base::get().getLoc(); }
  operator T() { return Base::get(); }
  T get() { return Base::get(); }
};
template<typename T, typename IndexType>
class ReduceMaxLoc<b200_reduce, T, IndexType> : public b200::detail::Reduce<b200_detail::ValueLocOp<T, IndexType, false>>
{
public:
  using Op         = b200_detail::ValueLocOp<T, IndexType, false>;
  using value_type = typename Op::value_type;
  using Base       = b200::detail::Reduce<Op>;
  ReduceMaxLoc(T init_val, IndexType init_idx, T identity_val = RAJA::operators::limits<T>::min(),
               IndexType identity_idx = RAJA::reduce::detail::DefaultLoc<IndexType>().value())
      : Base(value_type(init_val, init_idx), value_type(identity_val, identity_idx))
  {}
  void reset(T init_val, IndexType init_idx, T identity_val = RAJA::operators::limits<T>::min(),
             IndexType identity_idx = RAJA::reduce::detail::DefaultLoc<IndexType>().value())
  {
    Base::reset(value_type(init_val, init_idx), value_type(identity_val, identity_idx));
  }
  RAJA_HOST_DEVICE const ReduceMaxLoc& maxloc(T rhs, IndexType loc) const
  {
    this->combine(value_type(rhs, loc));
    return *this;
  }
  IndexType getLoc() { return Base::get().getLoc(); }
  operator T() { return Base::get(); }
  T get() { return Base::get(); }
};

}  // namespace RAJA

#endif
