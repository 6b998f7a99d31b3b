// Classic reducer objects for the b200 back-end: ReduceSum / ReduceMin / ReduceMax / ReduceMinLoc /
// ReduceMaxLoc / ReduceBitOr / ReduceBitAnd <RAJA::b200_reduce | RAJA::b200_reduce_atomic, T[, IndexType]>.
// Replaces policy/cuda/reduce.hpp:968-1378 (cuda::Reduce, its data classes :678-965 and PinnedTally :474-661).
//
// Protocol (same object model as the reference, much less machinery):
//   ROOT     the object the user declared.  Owns a control block: the streams that have launches of this reducer
//            in flight and a growing list of pinned, device-visible result slots, ONE PER LAUNCH (the role of
//            PinnedTally's per-stream value nodes) -- launches on different, unordered streams never touch the
//            same slot, so sharing one reducer between streams is safe.
//   LAUNCH   the copy made while a launch copies its loop body (the launch record of RAJA_b200/policy.hpp is
//            active): takes a result slot and per-launch scratch from the stream-ordered pools -- ordered policy:
//            one partial per block + a self-resetting counter; atomic policy: 16 zeroed bytes (counter + a
//            self-resetting atomic slot, b200/reduce.cuh) -- and returns the scratch when it is destroyed (safe
//            right after enqueue).
//   ALIAS    any further host copy of a LAUNCH copy (the by-value kernel argument): just carries pointers.
//   HOST     any other host copy: starts at the identity, folds into the object it was copied from on destruction.
//   DEVICE   copies made in device code.  A copy OF THE KERNEL ARGUMENT (every thread's private copy, made by the
//            engine) runs the block + grid combine in its destructor and the last block stores the launch's result
//            into the slot; a copy of a device copy (a by-value re-capture inside the loop body) folds into the copy
//            it was made from (policy/cuda/reduce.hpp:1109-1118: `if (!parent->parent) grid_reduce else combine`).
// get() waits for the streams that were used, folds the slots into the host value in launch order and returns it.
// Block shapes are read at run time (1-3 D, partial last warp allowed), so the same objects work inside
// RAJA::forall, RAJA::launch and RAJA::kernel launches of this back-end.
#ifndef RAJA_B200_REDUCE_CLASSIC_HPP
#define RAJA_B200_REDUCE_CLASSIC_HPP

#include <mutex>
#include <new>
#include <vector>

namespace RAJA
{
namespace b200_detail
{
// host-side fold of (initial value, launch 1, launch 2, ...): operators that distinguish "in order" from "in a tree"
// (the value-location pairs below) provide Op::fold_in_order; everything else folds with the operator itself
template<class Op, class V>
inline auto fold_in_order(Op, V a, V b, int) -> decltype(Op::fold_in_order(a, b))
{
  return Op::fold_in_order(a, b);
}
template<class Op, class V>
inline V fold_in_order(Op op, V a, V b, long)
{
  return op(a, b);
}
}  // namespace b200_detail

namespace b200
{
namespace detail
{

template<class Op, bool Ordered = true>
class Reduce
{
public:
  using T = typename Op::value_type;
  // operators / value types without a hardware atomic fall back to the ordered combine
  static constexpr bool use_atomic = !Ordered && ::b200::AtomicSlot<Op>::available;

  Reduce() : Reduce(T(), Op::identity()) {}

  explicit Reduce(T init_val, T identity_ = Op::identity())
      : m_parent(this), m_ctl(new Control), m_value(init_val), m_identity(identity_), m_scratch(nullptr), m_counter(nullptr), m_out(nullptr), m_kind(ROOT)
  {}

  void reset(T in_val, T identity_ = Op::identity())
  {
    (void)get();  // syncs outstanding launches
    m_value    = in_val;
    m_identity = identity_;
  }

  RAJA_HOST_DEVICE Reduce(const Reduce& other)
      : m_parent(&other), m_ctl(other.m_ctl), m_value(other.m_identity), m_identity(other.m_identity), m_scratch(other.m_scratch),
        m_counter(other.m_counter), m_out(other.m_out), m_kind(HOST), m_stream(other.m_stream)
  {
#if !defined(__CUDA_ARCH__)
    if (other.m_kind == LAUNCH || other.m_kind == ALIAS)
    {
      m_kind   = ALIAS;
      m_parent = nullptr;
    }
    else
    {
      ::RAJA::b200_detail::LaunchRecord& lr = ::RAJA::b200_detail::tl_launch();
      if (lr.active)
      {
        m_kind   = LAUNCH;
        m_parent = nullptr;
        m_stream = lr.stream;
        ++lr.reducers;
        if (use_atomic)
        {
          m_counter = static_cast<unsigned*>(rb200_device_zeroed_malloc(16, lr.stream));
          m_scratch = nullptr;
        }
        else
        {
          m_scratch = static_cast<T*>(rb200_device_malloc(sizeof(T) * (size_t)lr.grid, lr.stream));
          m_counter = static_cast<unsigned*>(rb200_device_zeroed_malloc(sizeof(unsigned), lr.stream));
        }
        if (!m_counter || (!use_atomic && !m_scratch)) throw std::runtime_error(std::string("raja_b200: reducer scratch: ") + rb200_last_error());
        m_out = m_ctl->new_slot(lr.stream, m_identity);
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
      case ROOT: delete m_ctl; break;
      case LAUNCH:
        if (m_scratch) rb200_device_free(m_scratch, m_stream);
        rb200_device_zeroed_free(m_counter, m_stream);
        break;
      case HOST:
        if (m_parent && m_parent != this)
        {
          std::lock_guard<std::mutex> lock(m_ctl->mu);
          m_parent->combine(m_value);
        }
        break;
      default: break;
    }
#else
    if (m_kind != DEVICE || m_parent == nullptr) return;
    const Kind from = m_parent->m_kind;
    if (from == DEVICE)
    {
      m_parent->combine(m_value);  // a re-capture inside the loop body: same thread, no synchronisation
    }
    else if ((from == LAUNCH || from == ALIAS) && m_counter != nullptr)
    {
      // every thread of the block gets here (the engine's private copy goes out of scope once per thread)
      T r;
      bool mine;
      if constexpr (use_atomic) mine = ::RAJA::b200_detail::shaped_grid_reduce_atomic<Op>(m_value, m_counter, r);
      else mine = ::RAJA::b200_detail::shaped_grid_reduce<Op>(m_value, m_scratch, m_counter, r);
      if (mine) *m_out = r;  // this launch's own pinned slot
    }
    // else: copied into a kernel without a launch record of this back-end -- nothing to combine into
#endif
  }

  //! synchronises the streams this reducer was used on, folds the launches' results in, returns the value
  T get()
  {
    Control* c = m_ctl;
    std::lock_guard<std::mutex> lock(c->mu);
    if (c->used > 0)
    {
      for (cudaStream_t s : c->streams) ::RAJA::b200_detail::check(cudaStreamSynchronize(s), "reducer get");
      c->streams.clear();
      Op op;
      for (size_t i = 0; i < c->used; ++i) m_value = ::RAJA::b200_detail::fold_in_order(op, m_value, *c->slot(i), 0);
      c->used = 0;
    }
    return m_value;
  }
  operator T() { return get(); }

  RAJA_HOST_DEVICE void combine(T other) const { m_value = Op {}(m_value, other); }
  T& local() const { return m_value; }
  T get_combined() const { return m_value; }

private:
  enum Kind : int { ROOT, HOST, LAUNCH, ALIAS, DEVICE };
  struct Control  // shared by the root and its host copies
  {
    static constexpr size_t chunk = 64;
    std::mutex mu;
    std::vector<cudaStream_t> streams;  // streams with launches in flight since the last get()
    std::vector<T*> chunks;             // pinned, device-visible slots, `chunk` per allocation
    size_t used = 0;                    // slots handed out since the last get()
    T* slot(size_t i) { return chunks[i / chunk] + (i % chunk); }
    T* new_slot(cudaStream_t s, const T& identity)
    {
      std::lock_guard<std::mutex> lock(mu);
      bool seen = false;
      for (cudaStream_t t : streams) seen = seen || (t == s);
      if (!seen) streams.push_back(s);
      if (used / chunk >= chunks.size())
      {
        T* p = static_cast<T*>(rb200_pinned_malloc(sizeof(T) * chunk));  // mapped pinned memory: one address for host and device
        if (!p) throw std::runtime_error(std::string("raja_b200: reducer tally: ") + rb200_last_error());
        chunks.push_back(p);
      }
      T* p = slot(used++);
      new (p) T(identity);  // a launch whose grid never runs (error) contributes nothing
      return p;
    }
    ~Control()
    {
      for (T* p : chunks) rb200_pinned_free(p);
    }
  };
  const Reduce* m_parent;
  Control* m_ctl;
  mutable T m_value;
  T m_identity;
  T* m_scratch;         // ordered: one partial per block
  unsigned* m_counter;  // zeroed, self-resetting ticket (atomic policy: 16 bytes, ticket + slot)
  T* m_out;             // pinned result slot of this launch
  Kind m_kind;
  cudaStream_t m_stream = 0;
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
      // ties inside the parallel tree: the smaller location wins and "no location" (the identity's, DefaultLoc) counts as
      // the smallest -- in the sequential fold an accumulator that starts at the identity keeps it on a tie
      const I none = RAJA::reduce::detail::DefaultLoc<I>().value();
      if (!better && !worse) take_b = (a.getLoc() != none) && (b.getLoc() == none || b.getLoc() < a.getLoc());
    }
    return take_b ? b : a;
  }
  // host-side fold of (initial value, launch 1, launch 2, ...): the order is known, so a tie keeps the left operand exactly
  // as the sequential loop does (util/Operators.hpp:555-594) -- e.g. an initial value that equals the minimum keeps its location
  RAJA_HOST_DEVICE static value_type fold_in_order(value_type a, value_type b)
  {
    const T av = a, bv = b;
    return (doing_min ? (bv < av) : (av < bv)) ? b : a;
  }
};

}  // namespace b200_detail

// ---- the specialisations user code names (primary templates: pattern/reduce.hpp:80-233) -----------------
template<bool Ordered, typename T>
class ReduceSum<b200_reduce_policy<Ordered>, T> : public b200::detail::Reduce<::b200::Plus<T>, Ordered>
{
public:
  using Base = b200::detail::Reduce<::b200::Plus<T>, Ordered>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceSum& operator+=(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<bool Ordered, typename T>
class ReduceMin<b200_reduce_policy<Ordered>, T> : public b200::detail::Reduce<::b200::Minimum<T>, Ordered>
{
public:
  using Base = b200::detail::Reduce<::b200::Minimum<T>, Ordered>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceMin& min(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<bool Ordered, typename T>
class ReduceMax<b200_reduce_policy<Ordered>, T> : public b200::detail::Reduce<::b200::Maximum<T>, Ordered>
{
public:
  using Base = b200::detail::Reduce<::b200::Maximum<T>, Ordered>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceMax& max(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<bool Ordered, typename T>
class ReduceBitOr<b200_reduce_policy<Ordered>, T> : public b200::detail::Reduce<::b200::BitOr<T>, Ordered>
{
public:
  using Base = b200::detail::Reduce<::b200::BitOr<T>, Ordered>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceBitOr& operator|=(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<bool Ordered, typename T>
class ReduceBitAnd<b200_reduce_policy<Ordered>, T> : public b200::detail::Reduce<::b200::BitAnd<T>, Ordered>
{
public:
  using Base = b200::detail::Reduce<::b200::BitAnd<T>, Ordered>;
  using Base::Base;
  RAJA_HOST_DEVICE const ReduceBitAnd& operator&=(T rhs) const
  {
    this->combine(rhs);
    return *this;
  }
};
template<bool Ordered, typename T, typename IndexType>
class ReduceMinLoc<b200_reduce_policy<Ordered>, T, IndexType> : public b200::detail::Reduce<b200_detail::ValueLocOp<T, IndexType, true>, Ordered>
{
public:
  using Op         = b200_detail::ValueLocOp<T, IndexType, true>;
  using value_type = typename Op::value_type;
  using Base       = b200::detail::Reduce<Op, Ordered>;
  // (T(), IndexType()), as pattern/detail/reduce.hpp:403
  ReduceMinLoc() : Base(value_type(T(), IndexType()), Op::identity()) {}
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
template<bool Ordered, typename T, typename IndexType>
class ReduceMaxLoc<b200_reduce_policy<Ordered>, T, IndexType> : public b200::detail::Reduce<b200_detail::ValueLocOp<T, IndexType, false>, Ordered>
{
public:
  using Op         = b200_detail::ValueLocOp<T, IndexType, false>;
  using value_type = typename Op::value_type;
  using Base       = b200::detail::Reduce<Op, Ordered>;
  ReduceMaxLoc() : Base(value_type(T(), IndexType()), Op::identity()) {}
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
