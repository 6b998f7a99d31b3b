// RAJA::impl::scan::* for b200_exec: replaces policy/cuda/scan.hpp:52-235 (cub::DeviceScan) with the
// single-pass TMA-staged scan of b200/scan.cuh.  Must be seen before RAJA/pattern/scan.hpp.
#ifndef RAJA_B200_SCAN_HPP
#define RAJA_B200_SCAN_HPP

namespace RAJA
{
namespace impl
{
namespace scan
{
namespace b200_detail
{
template<bool Inclusive, size_t BLOCK_SIZE, bool Async, typename InIter, typename OutIter, typename Function, typename T>
inline resources::EventProxy<resources::Cuda> run(resources::Cuda res, InIter begin, InIter end, OutIter out, Function, T init)
{
  using Op = typename ::RAJA::b200_detail::kernel_op<Function, T>::type;
  static_assert(std::is_same<std::remove_cv_t<std::remove_reference_t<decltype(*out)>>, T>::value, "scan: input and output element types must match");
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "b200 scan supports 4- and 8-byte element types");
  const long long n = (long long)(end - begin);  // 64-bit (the reference truncates to int: policy/cuda/scan.hpp:65)
  if (n > 0)
  {
    cudaStream_t stream = res.get_stream();
    const T* in_p       = &*begin;  // iterators over contiguous storage (pointers, RAJA::Span)
    T* out_p            = &*out;
    using Shape         = typename ::b200::scan::Tuning<T>::shape;
    ::RAJA::b200_detail::Scratch ws(::b200::scan::workspace_bytes<T, Shape>(n), stream);
    ::RAJA::b200_detail::check(::b200::scan::launch<T, Op, Inclusive>(stream, in_p, out_p, n, init, ws.p), "scan launch");
    if (!Async) res.wait();
  }
  return resources::EventProxy<resources::Cuda>(res);
}

// ---- b200_multi_exec: this rank scans its shard seeded with the fold of the lower ranks' totals.  Three kernels on the stream:
//      shard total (grid reduce) -> exclusive prefix over the ranks through peer memory (rb200_comm_allreduce_small) -> the scan
//      with its carry read from device memory.  3 N sizeof(T) bytes of traffic per GPU, no host synchronisation in between.
template<typename T> struct multi_dtype;
template<> struct multi_dtype<int> { static constexpr rb200_dtype value = RB200_I32; };
template<> struct multi_dtype<unsigned int> { static constexpr rb200_dtype value = RB200_U32; };
template<> struct multi_dtype<long> { static constexpr rb200_dtype value = RB200_I64; };
template<> struct multi_dtype<unsigned long> { static constexpr rb200_dtype value = RB200_U64; };
template<> struct multi_dtype<long long> { static constexpr rb200_dtype value = RB200_I64; };
template<> struct multi_dtype<unsigned long long> { static constexpr rb200_dtype value = RB200_U64; };
template<> struct multi_dtype<float> { static constexpr rb200_dtype value = RB200_F32; };
template<> struct multi_dtype<double> { static constexpr rb200_dtype value = RB200_F64; };
template<typename Fn, typename T> struct multi_op;
template<typename T> struct multi_op<RAJA::operators::plus<T, T, T>, T> { static constexpr rb200_op value = RB200_PLUS; };
template<typename T> struct multi_op<RAJA::operators::minimum<T, T, T>, T> { static constexpr rb200_op value = RB200_MINIMUM; };
template<typename T> struct multi_op<RAJA::operators::maximum<T, T, T>, T> { static constexpr rb200_op value = RB200_MAXIMUM; };

template<bool Inclusive, bool Async, typename InIter, typename OutIter, typename Function, typename T>
inline resources::EventProxy<resources::Cuda> run_multi(resources::Cuda res, InIter begin, InIter end, OutIter out, Function, T init)
{
  using Op = typename ::RAJA::b200_detail::kernel_op<Function, T>::type;
  static_assert(std::is_same<std::remove_cv_t<std::remove_reference_t<decltype(*out)>>, T>::value, "scan: input and output element types must match");
  const ::RAJA::b200_detail::MultiState& m = ::RAJA::b200_detail::multi_state_or_throw("scan<b200_multi_exec>");
  const long long n   = (long long)(end - begin);
  cudaStream_t stream = res.get_stream();
  const T* in_p       = n > 0 ? &*begin : nullptr;
  T* out_p            = n > 0 ? &*out : nullptr;
  const T* carry      = nullptr;
  ::RAJA::b200_detail::Scratch small(64, stream);  // [total][carry]
  if (m.world > 1)
  {
    constexpr rb200_dtype dt = multi_dtype<T>::value;
    constexpr rb200_op op    = multi_op<Function, T>::value;
    T* total                 = static_cast<T*>(small.p);
    T* pre                   = reinterpret_cast<T*>(static_cast<char*>(small.p) + 16);
    // an empty shard contributes the identity (rb200_reduce of zero elements)
    ::RAJA::b200_detail::check_rc(rb200_reduce((rb200_reduce_kind)op, dt, in_p, n, total, nullptr, 1, stream), "scan shard total");
    ::RAJA::b200_detail::check_rc(rb200_comm_allreduce_small(m.comm, op, dt, total, pre, 1, 1, stream), "scan carry exchange");
    if (m.rank > 0) carry = pre;
  }
  if (n > 0)
  {
    using Shape = typename ::b200::scan::Tuning<T>::shape;
    ::RAJA::b200_detail::Scratch ws(::b200::scan::workspace_bytes<T, Shape>(n), stream);
    ::RAJA::b200_detail::check(::b200::scan::launch<T, Op, Inclusive>(stream, in_p, out_p, n, init, ws.p, 0, 0, nullptr, carry), "scan launch");
  }
  if (!Async) res.wait();
  return resources::EventProxy<resources::Cuda>(res);
}
}  // namespace b200_detail

template<size_t BLOCK_SIZE, bool Async, typename InputIter, typename Function>
inline resources::EventProxy<resources::Cuda> inclusive_inplace(resources::Cuda res, ::RAJA::policy::b200::b200_multi_exec<BLOCK_SIZE, Async> const&, InputIter begin,
                                                                 InputIter end, Function binary_op)
{
  using T = std::remove_cv_t<std::remove_reference_t<decltype(*begin)>>;
  return b200_detail::run_multi<true, Async>(res, begin, end, begin, binary_op, T {});
}
template<size_t BLOCK_SIZE, bool Async, typename InputIter, typename Function, typename T>
inline resources::EventProxy<resources::Cuda> exclusive_inplace(resources::Cuda res, ::RAJA::policy::b200::b200_multi_exec<BLOCK_SIZE, Async> const&, InputIter begin,
                                                                 InputIter end, Function binary_op, T init)
{
  using V = std::remove_cv_t<std::remove_reference_t<decltype(*begin)>>;
  return b200_detail::run_multi<false, Async>(res, begin, end, begin, binary_op, (V)init);
}
template<size_t BLOCK_SIZE, bool Async, typename InputIter, typename OutputIter, typename Function>
inline resources::EventProxy<resources::Cuda> inclusive(resources::Cuda res, ::RAJA::policy::b200::b200_multi_exec<BLOCK_SIZE, Async> const&, InputIter begin,
                                                         InputIter end, OutputIter out, Function binary_op)
{
  using T = std::remove_cv_t<std::remove_reference_t<decltype(*begin)>>;
  return b200_detail::run_multi<true, Async>(res, begin, end, out, binary_op, T {});
}
template<size_t BLOCK_SIZE, bool Async, typename InputIter, typename OutputIter, typename Function, typename T>
inline resources::EventProxy<resources::Cuda> exclusive(resources::Cuda res, ::RAJA::policy::b200::b200_multi_exec<BLOCK_SIZE, Async> const&, InputIter begin,
                                                         InputIter end, OutputIter out, Function binary_op, T init)
{
  using V = std::remove_cv_t<std::remove_reference_t<decltype(*begin)>>;
  return b200_detail::run_multi<false, Async>(res, begin, end, out, binary_op, (V)init);
}

template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename InputIter, typename Function>
inline resources::EventProxy<resources::Cuda> inclusive_inplace(resources::Cuda res,
                                                                 ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&,
                                                                 InputIter begin,
                                                                 InputIter end,
                                                                 Function binary_op)
{
  using T = std::remove_cv_t<std::remove_reference_t<decltype(*begin)>>;
  return b200_detail::run<true, BLOCK_SIZE, Async>(res, begin, end, begin, binary_op, T {});
}

template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename InputIter, typename Function, typename T>
inline resources::EventProxy<resources::Cuda> exclusive_inplace(resources::Cuda res,
                                                                 ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&,
                                                                 InputIter begin,
                                                                 InputIter end,
                                                                 Function binary_op,
                                                                 T init)
{
  using V = std::remove_cv_t<std::remove_reference_t<decltype(*begin)>>;
  return b200_detail::run<false, BLOCK_SIZE, Async>(res, begin, end, begin, binary_op, (V)init);
}

template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename InputIter, typename OutputIter, typename Function>
inline resources::EventProxy<resources::Cuda> inclusive(resources::Cuda res,
                                                         ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&,
                                                         InputIter begin,
                                                         InputIter end,
                                                         OutputIter out,
                                                         Function binary_op)
{
  using T = std::remove_cv_t<std::remove_reference_t<decltype(*begin)>>;
  return b200_detail::run<true, BLOCK_SIZE, Async>(res, begin, end, out, binary_op, T {});
}

template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename InputIter, typename OutputIter, typename Function, typename T>
inline resources::EventProxy<resources::Cuda> exclusive(resources::Cuda res,
                                                         ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&,
                                                         InputIter begin,
                                                         InputIter end,
                                                         OutputIter out,
                                                         Function binary_op,
                                                         T init)
{
  using V = std::remove_cv_t<std::remove_reference_t<decltype(*begin)>>;
  return b200_detail::run<false, BLOCK_SIZE, Async>(res, begin, end, out, binary_op, (V)init);
}

}  // namespace scan
}  // namespace impl
}  // namespace RAJA

#endif
