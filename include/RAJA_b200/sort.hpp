// RAJA::impl::sort::* for b200_exec: replaces policy/cuda/sort.hpp:101-653 (cub::DeviceRadixSort +
// copy-back) with the onesweep sort of b200/radix_sort.cuh.  Contract kept: arithmetic keys behind
// pointers, operators::less / operators::greater only, stable in every variant (the reference's
// "unstable" CUDA sorts are the stable radix sort too: sort.hpp:285-321, 612-653).
// Must be seen before RAJA/pattern/sort.hpp.
#ifndef RAJA_B200_SORT_HPP
#define RAJA_B200_SORT_HPP

namespace RAJA
{
namespace impl
{
namespace sort
{
namespace b200_detail
{
template<typename K> struct key_bits
{
  static_assert(std::is_arithmetic<K>::value && (sizeof(K) == 1 || sizeof(K) == 2 || sizeof(K) == 4 || sizeof(K) == 8), "b200 sort: keys must be arithmetic types of 1, 2, 4 or 8 bytes");
  static constexpr bool widened = sizeof(K) < 4;  // 1- and 2-byte keys are sorted as 32-bit words (sign- or zero-extended: the order is kept)
  using type = std::conditional_t<sizeof(K) <= 4, uint32_t, unsigned long long>;
  static constexpr int kind = std::is_floating_point<K>::value ? ::b200::rsort::KEY_FLOAT
                              : std::is_signed<K>::value       ? ::b200::rsort::KEY_SIGNED
                                                               : ::b200::rsort::KEY_UNSIGNED;
};
// Widths the onesweep kernels are not instantiated for go through small helper kernels around the 4- / 8-byte sorts (the reference
// hands any key / value type to cub::DeviceRadixSort, policy/cuda/sort.hpp:391-544):
//   * 1- and 2-byte keys: widened to 32-bit words (their constant upper bytes are degenerate digit passes, which the sweep copies
//     through), sorted, narrowed back;
//   * values of any other size (1, 2, 16, ... bytes, any trivially copyable type): the keys are sorted together with their
//     positions, the values are gathered through the sorted positions.
template<typename K, typename W> __global__ void widen_keys_kernel(const K* in, W* out, long long n)
{
  using S = std::conditional_t<std::is_signed<K>::value, int32_t, uint32_t>;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (W)(S)in[i];
}
template<typename K, typename W> __global__ void narrow_keys_kernel(const W* in, K* out, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (K)in[i];
}
template<typename I> __global__ void iota_kernel(I* out, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (I)i;
}
template<typename V, typename I> __global__ void gather_kernel(const V* in, const I* pos, V* out, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[pos[i]];
}
inline unsigned helper_grid(long long n)
{
  const long long blocks = (n + 255) / 256;
  return (unsigned)(blocks < 148 * 16 ? (blocks > 0 ? blocks : 1) : 148 * 16);
}
template<typename Compare, typename K> struct compare_order
{
  static constexpr bool is_less    = std::is_same<Compare, operators::less<K>>::value;
  static constexpr bool is_greater = std::is_same<Compare, operators::greater<K>>::value;
  static_assert(is_less || is_greater, "b200 sort: comparator must be RAJA::operators::less or RAJA::operators::greater (as for cuda_exec)");
  static constexpr int descending = is_greater ? 1 : 0;
};

template<bool Async, typename KeyIter, typename Compare>
inline resources::EventProxy<resources::Cuda> sort_keys(resources::Cuda res, KeyIter begin, KeyIter end, Compare)
{
  static_assert(std::is_pointer<KeyIter>::value, "b200 sort: keys must be passed as pointers");
  using K = std::remove_cv_t<std::remove_pointer_t<KeyIter>>;
  using U = typename key_bits<K>::type;
  const long long n = (long long)(end - begin);
  if (n > 1)
  {
    cudaStream_t stream = res.get_stream();
    const ::b200::rsort::KeyOrder order {key_bits<K>::kind, compare_order<Compare, K>::descending};
    ::RAJA::b200_detail::Scratch alt(sizeof(U) * (size_t)n, stream);
    ::RAJA::b200_detail::Scratch ws(::b200::rsort::workspace_bytes<U, ::b200::rsort::NoValue>(n), stream);
    if constexpr (key_bits<K>::widened)
    {
      ::RAJA::b200_detail::Scratch wide(sizeof(U) * (size_t)n, stream);
      widen_keys_kernel<K, U><<<helper_grid(n), 256, 0, stream>>>(begin, (U*)wide.p, n);
      ::RAJA::b200_detail::check(::b200::rsort::sort<U, ::b200::rsort::NoValue>(stream, (U*)wide.p, (U*)alt.p, nullptr, nullptr, n, order, ws.p), "sort launch");
      narrow_keys_kernel<K, U><<<helper_grid(n), 256, 0, stream>>>((const U*)wide.p, begin, n);
      ::RAJA::b200_detail::check(cudaGetLastError(), "sort helper launch");
    }
    else
    {
      ::RAJA::b200_detail::check(::b200::rsort::sort<U, ::b200::rsort::NoValue>(stream, (U*)begin, (U*)alt.p, nullptr, nullptr, n, order, ws.p), "sort launch");
    }
    if (!Async) res.wait();
  }
  return resources::EventProxy<resources::Cuda>(res);
}

template<typename I, typename U, typename V>
inline void sort_pairs_by_position(cudaStream_t stream, U* keys, U* keys_alt, V* vals, long long n, const ::b200::rsort::KeyOrder& order)
{
  ::RAJA::b200_detail::Scratch pos(sizeof(I) * (size_t)n, stream);
  ::RAJA::b200_detail::Scratch palt(sizeof(I) * (size_t)n, stream);
  ::RAJA::b200_detail::Scratch moved(sizeof(V) * (size_t)n, stream);
  ::RAJA::b200_detail::Scratch ws(::b200::rsort::workspace_bytes<U, I>(n), stream);
  iota_kernel<I><<<helper_grid(n), 256, 0, stream>>>((I*)pos.p, n);
  ::RAJA::b200_detail::check(::b200::rsort::sort<U, I>(stream, keys, keys_alt, (I*)pos.p, (I*)palt.p, n, order, ws.p), "sort_pairs launch");
  gather_kernel<V, I><<<helper_grid(n), 256, 0, stream>>>(vals, (const I*)pos.p, (V*)moved.p, n);
  ::RAJA::b200_detail::check(cudaMemcpyAsync(vals, moved.p, sizeof(V) * (size_t)n, cudaMemcpyDeviceToDevice, stream), "sort_pairs value copy");
}

template<bool Async, typename KeyIter, typename ValIter, typename Compare>
inline resources::EventProxy<resources::Cuda> sort_pairs(resources::Cuda res, KeyIter kb, KeyIter ke, ValIter vb, Compare)
{
  static_assert(std::is_pointer<KeyIter>::value && std::is_pointer<ValIter>::value, "b200 sort_pairs: keys and values must be passed as pointers");
  using K = std::remove_cv_t<std::remove_pointer_t<KeyIter>>;
  using V = std::remove_cv_t<std::remove_pointer_t<ValIter>>;
  using U = typename key_bits<K>::type;
  static_assert(std::is_trivially_copyable<V>::value, "b200 sort_pairs: values must be trivially copyable");
  constexpr bool direct_vals = (sizeof(V) == 4 || sizeof(V) == 8) && alignof(V) >= sizeof(V);  // values travel with their keys through the sweeps
  const long long n          = (long long)(ke - kb);
  if (n > 1)
  {
    cudaStream_t stream = res.get_stream();
    const ::b200::rsort::KeyOrder order {key_bits<K>::kind, compare_order<Compare, K>::descending};
    ::RAJA::b200_detail::Scratch kalt(sizeof(U) * (size_t)n, stream);
    // the keys the sweeps work on: the caller's, or their widened copy
    ::RAJA::b200_detail::Scratch wide(key_bits<K>::widened ? sizeof(U) * (size_t)n : 16, stream);
    U* keys = (U*)kb;
    if constexpr (key_bits<K>::widened)
    {
      keys = (U*)wide.p;
      widen_keys_kernel<K, U><<<helper_grid(n), 256, 0, stream>>>(kb, keys, n);
    }
    if constexpr (direct_vals)
    {
      using VT = std::conditional_t<sizeof(V) == 4, uint32_t, unsigned long long>;
      ::RAJA::b200_detail::Scratch valt(sizeof(VT) * (size_t)n, stream);
      ::RAJA::b200_detail::Scratch ws(::b200::rsort::workspace_bytes<U, VT>(n), stream);
      ::RAJA::b200_detail::check(::b200::rsort::sort<U, VT>(stream, keys, (U*)kalt.p, (VT*)vb, (VT*)valt.p, n, order, ws.p), "sort_pairs launch");
    }
    else
    {
      if (n <= 0xffffffffll) sort_pairs_by_position<uint32_t, U, V>(stream, keys, (U*)kalt.p, vb, n, order);
      else sort_pairs_by_position<unsigned long long, U, V>(stream, keys, (U*)kalt.p, vb, n, order);
    }
    if constexpr (key_bits<K>::widened) narrow_keys_kernel<K, U><<<helper_grid(n), 256, 0, stream>>>((const U*)keys, kb, n);
    ::RAJA::b200_detail::check(cudaGetLastError(), "sort_pairs helper launch");
    if (!Async) res.wait();
  }
  return resources::EventProxy<resources::Cuda>(res);
}
}  // namespace b200_detail

template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename Iter, typename Compare>
inline resources::EventProxy<resources::Cuda> unstable(resources::Cuda res, ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, Iter begin, Iter end, Compare comp)
{
  return b200_detail::sort_keys<Async>(res, begin, end, comp);
}
template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename Iter, typename Compare>
inline resources::EventProxy<resources::Cuda> stable(resources::Cuda res, ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, Iter begin, Iter end, Compare comp)
{
  return b200_detail::sort_keys<Async>(res, begin, end, comp);
}
template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename KeyIter, typename ValIter, typename Compare>
inline resources::EventProxy<resources::Cuda> unstable_pairs(resources::Cuda res, ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, KeyIter kb, KeyIter ke, ValIter vb, Compare comp)
{
  return b200_detail::sort_pairs<Async>(res, kb, ke, vb, comp);
}
template<size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename KeyIter, typename ValIter, typename Compare>
inline resources::EventProxy<resources::Cuda> stable_pairs(resources::Cuda res, ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async> const&, KeyIter kb, KeyIter ke, ValIter vb, Compare comp)
{
  return b200_detail::sort_pairs<Async>(res, kb, ke, vb, comp);
}


// b200_multi_exec: a sharded sort returns a piece of data-dependent size, which the in-place RAJA::sort signature cannot express
template<typename> struct multi_sort_not_in_place : std::false_type {};
#define RAJA_B200_NO_MULTI_SORT(NAME, ...)                                                                                          \
  template<size_t BLOCK_SIZE, bool Async, typename Iter, typename... Rest>                                                         \
  inline resources::EventProxy<resources::Cuda> NAME(resources::Cuda res, ::RAJA::policy::b200::b200_multi_exec<BLOCK_SIZE, Async> const&, Iter, Rest...) \
  {                                                                                                                                 \
    static_assert(multi_sort_not_in_place<Iter>::value, "RAJA::sort with b200_multi_exec: use RAJA::b200::multi_sort_pairs / multi_sort (RAJA_b200/multi.hpp)"); \
    return resources::EventProxy<resources::Cuda>(res);                                                                             \
  }
RAJA_B200_NO_MULTI_SORT(unstable)
RAJA_B200_NO_MULTI_SORT(stable)
RAJA_B200_NO_MULTI_SORT(unstable_pairs)
RAJA_B200_NO_MULTI_SORT(stable_pairs)
#undef RAJA_B200_NO_MULTI_SORT

}  // namespace sort
}  // namespace impl
}  // namespace RAJA

#endif
