// Reductions for b200_exec.
//   * RAJA::expt::Reduce / ReduceLoc parameters: param_init / param_combine / param_resolve hooks
//     (replaces policy/cuda/params/reduce.hpp:21-62 and the grid_reduce of policy/cuda/reduce.hpp:311-366)
//   * RAJA::detail::Name parameter hooks (policy/cuda/params/kernel_name.hpp)
//   * classic reducer objects ReduceSum/Min/Max/MinLoc/MaxLoc<b200_reduce, T[, I]>
//     (replaces policy/cuda/reduce.hpp:968-1378)
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
//      ties keep the left operand).  A location of -1 marks "no element yet" (expt::ValLoc's default).
template<class T, class I>
struct kernel_op<RAJA::operators::minimum<RAJA::expt::ValLoc<T, I>, RAJA::expt::ValLoc<T, I>, RAJA::expt::ValLoc<T, I>>, RAJA::expt::ValLoc<T, I>, void>
{
  struct type
  {
    using value_type = RAJA::expt::ValLoc<T, I>;
    RAJA_HOST_DEVICE static value_type identity() { return RAJA::operators::minimum<value_type, value_type, value_type>::identity(); }
    RAJA_HOST_DEVICE value_type operator()(value_type a, value_type b) const
    {
      const bool take_b = (b.val < a.val) || (!(a.val < b.val) && b.loc != I(-1) && (a.loc == I(-1) || b.loc < a.loc));
      return take_b ? b : a;
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
      const bool take_b = (a.val < b.val) || (!(b.val < a.val) && b.loc != I(-1) && (a.loc == I(-1) || b.loc < a.loc));
      return take_b ? b : a;
    }
  };
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
}  // namespace b200_detail

namespace expt
{
namespace detail
{
// ---- expt::Reduce / ReduceLoc ----------------------------------------------------------------------
template<size_t BLOCK_SIZE, bool Async, typename OP, typename T, typename VOp>
inline void param_init(::RAJA::policy::b200::b200_exec<BLOCK_SIZE, Async> const&, Reducer<OP, T, VOp>& red, ::RAJA::policy::b200::LaunchInfo& li)
{
  // one device block: [result][partial per block]; one zeroed, self-resetting counter
  red.devicetarget = static_cast<T*>(rb200_device_malloc(sizeof(T) * (size_t)(li.grid + 1), li.stream));
  red.device_count = static_cast<unsigned int*>(rb200_device_zeroed_malloc(sizeof(unsigned int), li.stream));
  if (!red.devicetarget || !red.device_count) throw std::runtime_error(std::string("raja_b200: reducer scratch: ") + rb200_last_error());
}

template<size_t BLOCK_SIZE, bool Async, typename OP, typename T, typename VOp>
RAJA_HOST_DEVICE inline void param_combine(::RAJA::policy::b200::b200_exec<BLOCK_SIZE, Async> const&, Reducer<OP, T, VOp>& red)
{
#if defined(__CUDA_ARCH__)
  using Op = typename ::RAJA::b200_detail::kernel_op<OP, T>::type;
  ::RAJA::b200_detail::grid_reduce_to_slot<(int)BLOCK_SIZE, Op>(red.getVal(), red.devicetarget, red.device_count);
#else
  RAJA_UNUSED_VAR(red);
#endif
}

template<size_t BLOCK_SIZE, bool Async, typename OP, typename T, typename VOp>
inline void param_resolve(::RAJA::policy::b200::b200_exec<BLOCK_SIZE, Async> const&, Reducer<OP, T, VOp>& red, ::RAJA::policy::b200::LaunchInfo& li)
{
  // the result is folded into the user's target on the host, as the reference does after res.wait()
  // (policy/cuda/params/reduce.hpp:46-60)
  alignas(alignof(T)) unsigned char raw[sizeof(T)];
  ::RAJA::b200_detail::check(cudaMemcpyAsync(raw, red.devicetarget, sizeof(T), cudaMemcpyDeviceToHost, li.stream), "reducer readback");
  ::RAJA::b200_detail::check(cudaStreamSynchronize(li.stream), "reducer readback");
  T value = *reinterpret_cast<T*>(raw);
  red.combineTarget(value);
  rb200_device_zeroed_free(red.device_count, li.stream);
  red.device_count = nullptr;
  rb200_device_free(red.devicetarget, li.stream);
  red.devicetarget = nullptr;
}

// ---- RAJA::Name("...") ---------------------------------------------------------------------------------
template<size_t BLOCK_SIZE, bool Async>
inline void param_init(::RAJA::policy::b200::b200_exec<BLOCK_SIZE, Async> const&, RAJA::detail::Name&, const ::RAJA::policy::b200::LaunchInfo&)
{}
template<size_t BLOCK_SIZE, bool Async>
RAJA_HOST_DEVICE inline void param_combine(::RAJA::policy::b200::b200_exec<BLOCK_SIZE, Async> const&, RAJA::detail::Name&)
{}
template<size_t BLOCK_SIZE, bool Async>
inline void param_resolve(::RAJA::policy::b200::b200_exec<BLOCK_SIZE, Async> const&, RAJA::detail::Name&, const ::RAJA::policy::b200::LaunchInfo&)
{}

}  // namespace detail
}  // namespace expt
}  // namespace RAJA

#include "RAJA_b200/reduce_classic.hpp"

#endif
