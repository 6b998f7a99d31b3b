// Reductions for b200_exec.
//   * RAJA::expt::Reduce / ReduceLoc parameters: param_init / param_combine / param_resolve hooks
//     (replaces policy/cuda/params/reduce.hpp:21-62 and the grid_reduce of policy/cuda/reduce.hpp:311-366)
//   * RAJA::detail::Name parameter hooks (policy/cuda/params/kernel_name.hpp:28-71): an NVTX range per launch
//   * classic reducer objects ReduceSum/Min/Max/MinLoc/MaxLoc/BitOr/BitAnd<b200_reduce | b200_reduce_atomic, T[, I]>
//     (replaces policy/cuda/reduce.hpp:968-1378), in RAJA_b200/reduce_classic.hpp
// Grid combine: warp shuffle / redux -> one shared-memory hop -> last-block-done over the persistent
// grid's <= ~600 partials in fixed slot order (b200/reduce.cuh), i.e. deterministic for a given (N, grid).
#ifndef RAJA_B200_REDUCE_HPP
#define RAJA_B200_REDUCE_HPP

namespace RAJA
{
namespace b200_detail
{
// ---- min/max over (value, location) pairs: ties go to the smaller location, which makes the parallel
//      tree return what the sequential loop returns -- the FIRST extremum (util/Operators.hpp:555-594:
//      ties keep the left operand).  A location of -1 marks "no element yet" (expt::ValLoc's default) and counts as the
//      smallest: a sequential accumulator that starts at the identity keeps it when an element ties with it.
template<class T, class I>
struct kernel_op<RAJA::operators::minimum<RAJA::expt::ValLoc<T, I>, RAJA::expt::ValLoc<T, I>, RAJA::expt::ValLoc<T, I>>, RAJA::expt::ValLoc<T, I>, void>
{
  struct type
  {
    using value_type = RAJA::expt::ValLoc<T, I>;
    RAJA_HOST_DEVICE static value_type identity() { return RAJA::operators::minimum<value_type, value_type, value_type>::identity(); }
    RAJA_HOST_DEVICE value_type operator()(value_type a, value_type b) const
    {
      bool take_b = (b.val < a.val);
      if constexpr (std::is_integral<I>::value) take_b = take_b || (!(a.val < b.val) && a.loc != I(-1) && (b.loc == I(-1) || b.loc < a.loc));
      return take_b ? b : a;  // location types without an order (user structs): a tie keeps the left operand
    }
  };
};
template<class T, class I>
struct kernel_op<RAJA::operators::maximum<RAJA::expt::ValLoc<T, I>, RAJA::expt::ValLoc<T, I>, RAJA::expt::ValLoc<T, I>>, RAJA::expt::ValLoc<T, I>, void>
{
  struct type
  {
    using value_type = RAJA::expt::ValLoc<T, I>;
    RAJA_HOST_DEVICE static value_type identity() { return RAJA::operators::maximum<value_type, value_type, value_type>::identity(); }
    RAJA_HOST_DEVICE value_type operator()(value_type a, value_type b) const
    {
      bool take_b = (a.val < b.val);
      if constexpr (std::is_integral<I>::value) take_b = take_b || (!(b.val < a.val) && a.loc != I(-1) && (b.loc == I(-1) || b.loc < a.loc));
      return take_b ? b : a;
    }
  };
};
// bitwise reductions map onto the kernel-side operators (redux.sync for 32-bit words)
template<class T>
struct kernel_op<RAJA::operators::bit_or<T, T, T>, T, std::enable_if_t<std::is_integral<T>::value>>
{
  using type = ::b200::BitOr<T>;
};
template<class T>
struct kernel_op<RAJA::operators::bit_and<T, T, T>, T, std::enable_if_t<std::is_integral<T>::value>>
{
  using type = ::b200::BitAnd<T>;
};

// block + grid combine of one value per thread; `slot[0]` receives the result, slot[1 ..] are the
// per-block partials, `counter` is a zero-initialised, self-resetting ticket.
template<int BLOCK, class Op>
__device__ __forceinline__ void grid_reduce_to_slot(typename Op::value_type v, typename Op::value_type* slot, unsigned* counter)
{
  using T = typename Op::value_type;
  __shared__ alignas(16) unsigned char raw[32 * sizeof(T)];  // T may have a non-trivial default constructor
  T* smem = reinterpret_cast<T*>(raw);
  Op op;
  v = ::b200::block_reduce<BLOCK>(v, op, smem);
  T r = v;
  if (::b200::grid_combine_last_block<BLOCK>(v, op, slot + 1, counter, smem, r))
  {
    if (threadIdx.x == 0) ::b200::st_cg(slot, r);
  }
  __syncthreads();  // smem is reused by the next reducer of the pack
}

// ---- the same for a launch of ANY shape (RAJA::launch teams, classic reducer destructors, RAJA::kernel):
//      1-3 D blocks and grids, team sizes that are not a multiple of 32.  Fold order is fixed: lanes ascending
//      inside a warp, warps ascending inside a block, blocks ascending.
struct BlockShape
{
  unsigned nthreads, tid, nblocks, bid, lane, w, nw, in_warp;
  __device__ __forceinline__ BlockShape()
  {
    nthreads = blockDim.x * blockDim.y * blockDim.z;
    tid      = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    nblocks  = gridDim.x * gridDim.y * gridDim.z;
    bid      = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    lane     = tid & 31u;
    w        = tid >> 5;
    nw       = (nthreads + 31u) >> 5;
    in_warp  = (nthreads - w * 32u) < 32u ? (nthreads - w * 32u) : 32u;  // the last warp of a block may be partial
  }
};

// ordered tree over the `count` lanes that exist in this warp; valid in lane 0
template<class Op>
__device__ __forceinline__ typename Op::value_type shaped_warp_fold(typename Op::value_type x, unsigned count, unsigned lane, Op op)
{
  using T             = typename Op::value_type;
  const unsigned mask = count >= 32u ? 0xffffffffu : ((1u << count) - 1u);
#pragma unroll
  for (unsigned m = 1; m < 32u; m <<= 1)
  {
    T o = ::b200::shfl_idx_mask(mask, x, (int)((lane + m) < count ? lane + m : lane));
    if (lane + m < count && (lane & (2u * m - 1u)) == 0u) x = op(x, o);
  }
  return x;
}

// block fold; valid in flattened thread 0.  smem: 32 values; may be reused after the next __syncthreads().
template<class Op>
__device__ __forceinline__ typename Op::value_type shaped_block_fold(typename Op::value_type v, typename Op::value_type* smem, const BlockShape& s, Op op)
{
  v = shaped_warp_fold(v, s.in_warp, s.lane, op);
  if (s.nw == 1) return v;
  if (s.lane == 0) smem[s.w] = v;
  __syncthreads();
  if (s.w == 0)  // with more than one warp, warp 0 is a full warp
  {
    v = (s.lane < s.nw) ? smem[s.lane] : Op::identity();
    v = shaped_warp_fold(v, 32u, s.lane, op);
  }
  return v;
}

// Ordered grid combine.  Every thread of every block calls it with its value; returns true in flattened thread 0 of
// the block that arrived last, with `result` = fold of all blocks in block order.  `counter`: zeroed, self-resetting.
template<class Op>
__device__ __forceinline__ bool shaped_grid_reduce(typename Op::value_type v, typename Op::value_type* partials, unsigned* counter,
                                                   typename Op::value_type& result)
{
  using T = typename Op::value_type;
  __shared__ alignas(16) unsigned char raw[32 * sizeof(T)];
  __shared__ bool s_last;
  T* smem = reinterpret_cast<T*>(raw);
  Op op;
  const BlockShape s;
  v = shaped_block_fold(v, smem, s, op);
  if (s.tid == 0)
  {
    ::b200::st_cg(partials + s.bid, v);
    __threadfence();
    const unsigned ticket = ::atomicInc(counter, s.nblocks - 1u);  // wraps back to 0: scratch is returned zeroed
    s_last                = (ticket == s.nblocks - 1u);
  }
  __syncthreads();
  bool mine = false;
  if (s_last)
  {
    __threadfence();
    // every thread folds a CONTIGUOUS range of partials, so the final order is block order
    const unsigned per = (s.nblocks + s.nthreads - 1u) / s.nthreads;
    const unsigned lo = s.tid * per, hi = (lo + per < s.nblocks) ? lo + per : s.nblocks;
    T acc = Op::identity();
    for (unsigned i = lo; i < hi; ++i) acc = op(acc, ::b200::ld_cg(partials + i));
    __syncthreads();  // smem reuse
    acc    = shaped_block_fold(acc, smem, s, op);
    result = acc;
    mine   = (s.tid == 0);
  }
  __syncthreads();  // smem and s_last are reused by the next reducer
  return mine;
}

// Unordered grid combine: one hardware atomic per block on a self-resetting slot (b200/reduce.cuh: AtomicSlot).
// `block` points at 16 zeroed bytes: [counter u32][pad][slot 8 B]; both are zero again when the kernel ends.
template<class Op>
__device__ __forceinline__ bool shaped_grid_reduce_atomic(typename Op::value_type v, void* block, typename Op::value_type& result)
{
  using T = typename Op::value_type;
  __shared__ alignas(16) unsigned char raw[32 * sizeof(T)];
  T* smem = reinterpret_cast<T*>(raw);
  Op op;
  const BlockShape s;
  v         = shaped_block_fold(v, smem, s, op);
  bool mine = false;
  if (s.tid == 0)
  {
    unsigned* counter = static_cast<unsigned*>(block);
    void* slot        = static_cast<char*>(block) + 8;
    ::b200::AtomicSlot<Op>::combine(slot, v);
    __threadfence();
    const unsigned ticket = ::atomicInc(counter, s.nblocks - 1u);
    if (ticket == s.nblocks - 1u)
    {
      __threadfence();
      result = ::b200::AtomicSlot<Op>::take(slot);
      mine   = true;
    }
  }
  __syncthreads();  // smem is reused by the next reducer
  return mine;
}
}  // namespace b200_detail

namespace expt
{
namespace detail
{
// ---- expt::Reduce / ReduceLoc ----------------------------------------------------------------------
template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename OP, typename T, typename VOp>
inline void param_init(::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, Reducer<OP, T, VOp>& red, ::RAJA::policy::b200::LaunchInfo& li)
{
  // one device block: [result][partial per block]; one zeroed, self-resetting counter
  red.devicetarget = static_cast<T*>(rb200_device_malloc(sizeof(T) * (size_t)(li.grid + 1), li.stream));
  red.device_count = static_cast<unsigned int*>(rb200_device_zeroed_malloc(sizeof(unsigned int), li.stream));
  if (!red.devicetarget || !red.device_count) throw std::runtime_error(std::string("raja_b200: reducer scratch: ") + rb200_last_error());
  li.expect(red.devicetarget, sizeof(T));
}

template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename OP, typename T, typename VOp>
RAJA_HOST_DEVICE inline void param_combine(::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, Reducer<OP, T, VOp>& red)
{
#if defined(__CUDA_ARCH__)
  using Op = typename ::RAJA::b200_detail::kernel_op<OP, T>::type;
  ::RAJA::b200_detail::grid_reduce_to_slot<(int)BLOCK_SIZE, Op>(red.getVal(), red.devicetarget, red.device_count);
#else
  RAJA_UNUSED_VAR(red);
#endif
}

namespace b200_resolve
{
// the result is folded into the user's target on the host, as the reference does after res.wait()
// (policy/cuda/params/reduce.hpp:46-60); all reducers of the launch arrive with ONE wait (LaunchInfo::fetch)
template<typename Red>
inline void resolve(Red& red, ::RAJA::policy::b200::LaunchInfo& li)
{
  using T = typename Red::value_type;
  alignas(alignof(T)) unsigned char raw[sizeof(T)];
  const char* host = static_cast<const char*>(li.fetch(red.devicetarget, sizeof(T)));
  // one value per rank for a b200_multi_exec launch, folded into the target in rank order (deterministic; ranks own
  // ascending index ranges, so a location reducer keeps the first extremum)
  const size_t stride = li.comm ? li.stride_of(red.devicetarget) : 0;
  for (int r = 0; r < (li.comm ? li.world : 1); ++r)
  {
    memcpy(raw, host + (size_t)r * stride, sizeof(T));
    red.combineTarget(*reinterpret_cast<T*>(raw));
  }
  rb200_device_zeroed_free(red.device_count, li.stream);
  red.device_count = nullptr;
  rb200_device_free(red.devicetarget, li.stream);
  red.devicetarget = nullptr;
}
}  // namespace b200_resolve

template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename OP, typename T, typename VOp>
inline void param_resolve(::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, Reducer<OP, T, VOp>& red, ::RAJA::policy::b200::LaunchInfo& li)
{
  b200_resolve::resolve(red, li);
}

// ---- RAJA::Name("..."): an NVTX range around the launch (policy/cuda/params/kernel_name.hpp:28-71) -----------
template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async>
inline void param_init(::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, RAJA::detail::Name& kn, const ::RAJA::policy::b200::LaunchInfo&)
{
  if (kn.name != nullptr) rb200_range_push(kn.name);
}
template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async>
RAJA_HOST_DEVICE inline void param_combine(::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, RAJA::detail::Name&)
{}
template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async>
inline void param_resolve(::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, RAJA::detail::Name& kn, const ::RAJA::policy::b200::LaunchInfo&)
{
  if (kn.name != nullptr) rb200_range_pop();
}

}  // namespace detail
}  // namespace expt
}  // namespace RAJA

#include "RAJA_b200/reduce_classic.hpp"

#endif
