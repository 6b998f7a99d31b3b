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
// ---- element types the TMA-staged kernel is not written for (1-, 2-, 16-byte ..., any trivially copyable T with any associative
//      functor; cub::DeviceScan in the reference takes them all, policy/cuda/scan.hpp:52-235): a plain three-kernel scan.  Every
//      block owns a contiguous chunk, every thread a contiguous piece of it, all folds run left to right (associativity is all the
//      functor needs), in place is fine (a thread reads x[i] before it writes out[i], pieces are disjoint, the partials are complete
//      before the second sweep starts).  Not a roofline path: pieces are strided across a warp.
constexpr int GENERIC_SCAN_BLOCK = 256;
template<typename T, typename Fn>
__global__ void __launch_bounds__(GENERIC_SCAN_BLOCK) generic_scan_partials(const T* in, long long n, long long chunk, Fn op, T* thread_partial, T* block_partial)
{
  constexpr int BLOCK = GENERIC_SCAN_BLOCK;
  const long long piece = chunk / BLOCK;
  const long long lo    = (long long)blockIdx.x * chunk + (long long)threadIdx.x * piece;
  long long hi          = lo + piece;
  if (hi > n) hi = n;
  if (lo < hi)
  {
    T acc = in[lo];
    for (long long i = lo + 1; i < hi; ++i) acc = op(acc, in[i]);
    thread_partial[(long long)blockIdx.x * BLOCK + threadIdx.x] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const long long base = (long long)blockIdx.x * chunk;
    T acc                = thread_partial[(long long)blockIdx.x * BLOCK];  // a launched block owns at least one element
    for (int t = 1; t < BLOCK && base + t * piece < n; ++t) acc = op(acc, thread_partial[(long long)blockIdx.x * BLOCK + t]);
    block_partial[blockIdx.x] = acc;
  }
}
template<bool Inclusive, typename T, typename Fn>
__global__ void generic_scan_carries(T* block_partial, int nblocks, Fn op, T init)
{
  // block_partial[b] becomes the fold of everything in front of block b (seeded with init for an exclusive scan; slot 0 of an
  // inclusive scan is never read)
  T run = Inclusive ? block_partial[0] : init;
  for (int b = Inclusive ? 1 : 0; b < nblocks; ++b)
  {
    const T mine     = block_partial[b];
    block_partial[b] = run;
    run              = op(run, mine);
  }
}
template<bool Inclusive, typename T, typename Fn>
__global__ void __launch_bounds__(GENERIC_SCAN_BLOCK) generic_scan_sweep(const T* in, T* out, long long n, long long chunk, Fn op, const T* thread_partial, const T* block_carry)
{
  constexpr int BLOCK = GENERIC_SCAN_BLOCK;
  __shared__ alignas(16) unsigned char raw[BLOCK * sizeof(T)];
  T* const pre          = reinterpret_cast<T*>(raw);  // fold of everything in front of thread t's piece (t > 0, or any t with a carry)
  const long long piece = chunk / BLOCK;
  const long long base  = (long long)blockIdx.x * chunk;
  const bool seeded     = !Inclusive || blockIdx.x > 0;  // is there anything in front of this block's first element?
  if (threadIdx.x == 0)
  {
    T run = seeded ? block_carry[blockIdx.x] : thread_partial[(long long)blockIdx.x * BLOCK];
    if (seeded) pre[0] = run;
    for (int t = seeded ? 0 : 1; t < BLOCK && base + (long long)t * piece < n; ++t)
    {
      pre[t] = run;
      run    = op(run, thread_partial[(long long)blockIdx.x * BLOCK + t]);
    }
  }
  __syncthreads();
  const long long lo = base + (long long)threadIdx.x * piece;
  long long hi       = lo + piece;
  if (hi > n) hi = n;
  if (lo >= hi) return;
  long long i = lo;
  T acc;
  if (seeded || threadIdx.x > 0) acc = pre[threadIdx.x];
  else
  {  // the very first element of an inclusive scan has nothing in front of it
    acc    = in[i];
    out[i] = acc;
    ++i;
  }
  for (; i < hi; ++i)
  {
    const T x = in[i];
    if (Inclusive)
    {
      acc    = op(acc, x);
      out[i] = acc;
    }
    else
    {
      out[i] = acc;
      acc    = op(acc, x);
    }
  }
}
template<bool Inclusive, typename T, typename Function>
inline void run_generic(cudaStream_t stream, const T* in, T* out, long long n, Function op, T init)
{
  static_assert(std::is_trivially_copyable<T>::value, "b200 scan: element type must be trivially copyable");
  constexpr int BLOCK = GENERIC_SCAN_BLOCK;
  long long blocks    = (n + (long long)BLOCK * 16 - 1) / ((long long)BLOCK * 16);
  if (blocks > 2048) blocks = 2048;
  long long chunk = (n + blocks - 1) / blocks;
  chunk           = (chunk + BLOCK - 1) / BLOCK * BLOCK;
  blocks          = (n + chunk - 1) / chunk;  // every launched block owns at least one element
  ::RAJA::b200_detail::Scratch tp(sizeof(T) * (size_t)blocks * BLOCK, stream);
  ::RAJA::b200_detail::Scratch bp(sizeof(T) * (size_t)blocks, stream);
  generic_scan_partials<T, Function><<<(unsigned)blocks, BLOCK, 0, stream>>>(in, n, chunk, op, static_cast<T*>(tp.p), static_cast<T*>(bp.p));
  generic_scan_carries<Inclusive, T, Function><<<1, 1, 0, stream>>>(static_cast<T*>(bp.p), (int)blocks, op, init);
  generic_scan_sweep<Inclusive, T, Function><<<(unsigned)blocks, BLOCK, 0, stream>>>(in, out, n, chunk, op, static_cast<const T*>(tp.p), static_cast<const T*>(bp.p));
  ::RAJA::b200_detail::check(cudaGetLastError(), "generic scan launch");
}

template<bool Inclusive, size_t BLOCK_SIZE, bool Async, typename InIter, typename OutIter, typename Function, typename T>
inline resources::EventProxy<resources::Cuda> run(resources::Cuda res, InIter begin, InIter end, OutIter out, Function binary_op, T init)
{
  static_assert(std::is_same<std::remove_cv_t<std::remove_reference_t<decltype(*out)>>, T>::value, "scan: input and output element types must match");
  const long long n = (long long)(end - begin);  // 64-bit (the reference truncates to int: policy/cuda/scan.hpp:65)
  if constexpr (!((sizeof(T) == 4 || sizeof(T) == 8) && std::is_arithmetic<T>::value))
  {
    if (n > 0)
    {
      run_generic<Inclusive>(res.get_stream(), (const T*)&*begin, (T*)&*out, n, binary_op, init);
      if (!Async) res.wait();
    }
    return resources::EventProxy<resources::Cuda>(res);
  }
  else if (n > 0)
  {
    using Op = typename ::RAJA::b200_detail::kernel_op<Function, T>::type;
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
