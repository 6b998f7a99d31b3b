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
  static_assert(std::is_arithmetic<K>::value && (sizeof(K) == 4 || sizeof(K) == 8), "b200 sort: keys must be 4- or 8-byte arithmetic types");
  using type = std::conditional_t<sizeof(K) == 4, uint32_t, unsigned long long>;
  static constexpr int kind = std::is_floating_point<K>::value ? ::b200::rsort::KEY_FLOAT
                              : std::is_signed<K>::value       ? ::b200::rsort::KEY_SIGNED
                                                               : ::b200::rsort::KEY_UNSIGNED;
};
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
    ::RAJA::b200_detail::check(::b200::rsort::sort<U, ::b200::rsort::NoValue>(stream, (U*)begin, (U*)alt.p, nullptr, nullptr, n, order, ws.p), "sort launch");
    if (!Async) res.wait();
  }
  return resources::EventProxy<resources::Cuda>(res);
}

template<bool Async, typename KeyIter, typename ValIter, typename Compare>
inline resources::EventProxy<resources::Cuda> sort_pairs(resources::Cuda res, KeyIter kb, KeyIter ke, ValIter vb, Compare)
{
  static_assert(std::is_pointer<KeyIter>::value && std::is_pointer<ValIter>::value, "b200 sort_pairs: keys and values must be passed as pointers");
  using K = std::remove_cv_t<std::remove_pointer_t<KeyIter>>;
  using V = std::remove_cv_t<std::remove_pointer_t<ValIter>>;
  using U = typename key_bits<K>::type;
  static_assert(std::is_trivially_copyable<V>::value && (sizeof(V) == 4 || sizeof(V) == 8), "b200 sort_pairs: values must be trivially copyable, 4 or 8 bytes");
  using VT          = std::conditional_t<sizeof(V) == 4, uint32_t, unsigned long long>;
  const long long n = (long long)(ke - kb);
  if (n > 1)
  {
    cudaStream_t stream = res.get_stream();
    const ::b200::rsort::KeyOrder order {key_bits<K>::kind, compare_order<Compare, K>::descending};
    ::RAJA::b200_detail::Scratch kalt(sizeof(U) * (size_t)n, stream);
    ::RAJA::b200_detail::Scratch valt(sizeof(VT) * (size_t)n, stream);
    ::RAJA::b200_detail::Scratch ws(::b200::rsort::workspace_bytes<U, VT>(n), stream);
    ::RAJA::b200_detail::check(::b200::rsort::sort<U, VT>(stream, (U*)kb, (U*)kalt.p, (VT*)vb, (VT*)valt.p, n, order, ws.p), "sort_pairs launch");
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
