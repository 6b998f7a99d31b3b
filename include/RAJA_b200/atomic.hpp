// RAJA::atomicX(b200_atomic_explicit<host_policy>, ...): replaces policy/cuda/atomic.hpp:665-860.
// Device code issues the hardware operation wherever the ISA has one for the type (directly, or on the same-width
// integer word: add / sub / and / or / xor / exch / min / max / inc / dec / cas; load and store are relaxed
// ld / st at gpu scope, never a read-modify-write), and a CAS loop on the same-width unsigned word otherwise
// (floating-point min / max, wrap-around inc / dec of types other than unsigned int) -- the selection the reference makes in
// policy/cuda/atomic.hpp:185-220, 328-402, 597-609.  Host code falls through to host_policy like the reference (:690-700).
//   b200_atomic                       plain hardware atomics
//   b200_atomic_aggregated            atomicAdd: lanes of a warp that hit the same address are combined first
//                                     (match.any; one atomic per distinct address per warp) -- pays on contended bins
// Must be seen before RAJA/pattern/atomic.hpp (its calls are qualified).
#ifndef RAJA_B200_ATOMIC_HPP
#define RAJA_B200_ATOMIC_HPP

namespace RAJA
{
namespace policy
{
namespace b200
{
// warp-aggregated adds (b200/atomic.cuh); everything else as b200_atomic_explicit
template<typename host_policy>
struct b200_atomic_aggregated_explicit : b200_atomic_explicit<host_policy>
{};
}  // namespace b200
}  // namespace policy
using b200_atomic_aggregated = policy::b200::b200_atomic_aggregated_explicit<RAJA::seq_atomic>;

namespace b200_detail
{
template<size_t N> struct cas_word;
template<> struct cas_word<4> { using type = unsigned int; };
template<> struct cas_word<8> { using type = unsigned long long; };

template<typename T> using word_t = typename cas_word<sizeof(T)>::type;
template<typename T> __device__ __forceinline__ word_t<T> to_word(T v)
{
  word_t<T> w;
  memcpy(&w, &v, sizeof(T));
  return w;
}
template<typename T> __device__ __forceinline__ T from_word(word_t<T> w)
{
  T v;
  memcpy(&v, &w, sizeof(T));
  return v;
}

// relaxed load / store at gpu scope on a generic address (global or shared): one ld / st, no read-modify-write
template<typename T> __device__ __forceinline__ T dev_atomic_load(T* acc)
{
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "b200_atomic supports 4- and 8-byte types");
  if constexpr (sizeof(T) == 4)
  {
    unsigned w;
    asm volatile("ld.relaxed.gpu.b32 %0, [%1];" : "=r"(w) : "l"(acc) : "memory");
    return from_word<T>(w);
  }
  else
  {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.b64 %0, [%1];" : "=l"(w) : "l"(acc) : "memory");
    return from_word<T>(w);
  }
}
template<typename T> __device__ __forceinline__ void dev_atomic_store(T* acc, T v)
{
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "b200_atomic supports 4- and 8-byte types");
  if constexpr (sizeof(T) == 4) asm volatile("st.relaxed.gpu.b32 [%0], %1;" ::"l"(acc), "r"(to_word(v)) : "memory");
  else asm volatile("st.relaxed.gpu.b64 [%0], %1;" ::"l"(acc), "l"(to_word(v)) : "memory");
}

// old = *acc; *acc = f(old); return old -- atomically, for any 4- or 8-byte trivially copyable T
template<typename T, typename F>
__device__ __forceinline__ T atomic_rmw(T* acc, F f)
{
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "b200_atomic supports 4- and 8-byte types");
  using W = word_t<T>;
  W* addr = reinterpret_cast<W*>(acc);
  W seen  = to_word(dev_atomic_load(acc));
  while (true)
  {
    const T old_t = from_word<T>(seen);
    const W prev  = ::atomicCAS(addr, seen, to_word(f(old_t)));
    if (prev == seen) return old_t;  // compared as words: NaNs cannot loop forever
    seen = prev;
  }
}

template<typename T> struct is_hw_int : std::integral_constant<bool, std::is_integral<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)>
{};

template<typename T>
__device__ __forceinline__ T dev_atomic_add(T* acc, T v)
{
  if constexpr (is_hw_int<T>::value) return from_word<T>(::atomicAdd(reinterpret_cast<word_t<T>*>(acc), to_word(v)));  // two's complement
  else if constexpr (std::is_same<T, float>::value || std::is_same<T, double>::value) return ::atomicAdd(acc, v);
  else return atomic_rmw(acc, [v](T o) { return (T)(o + v); });
}
template<typename T>
__device__ __forceinline__ T dev_atomic_sub(T* acc, T v)
{
  if constexpr (is_hw_int<T>::value) return from_word<T>(::atomicAdd(reinterpret_cast<word_t<T>*>(acc), (word_t<T>)(0) - to_word(v)));
  else if constexpr (std::is_same<T, float>::value || std::is_same<T, double>::value) return ::atomicAdd(acc, -v);
  else return atomic_rmw(acc, [v](T o) { return (T)(o - v); });
}
template<typename T>
__device__ __forceinline__ T dev_atomic_min(T* acc, T v)
{
  if constexpr (is_hw_int<T>::value)
  {
    if constexpr (sizeof(T) == 4 && std::is_signed<T>::value) return (T)::atomicMin(reinterpret_cast<int*>(acc), (int)v);
    else if constexpr (sizeof(T) == 4) return (T)::atomicMin(reinterpret_cast<unsigned*>(acc), (unsigned)v);
    else if constexpr (std::is_signed<T>::value) return (T)::atomicMin(reinterpret_cast<long long*>(acc), (long long)v);
    else return (T)::atomicMin(reinterpret_cast<unsigned long long*>(acc), (unsigned long long)v);
  }
  else return atomic_rmw(acc, [v](T o) { return v < o ? v : o; });
}
template<typename T>
__device__ __forceinline__ T dev_atomic_max(T* acc, T v)
{
  if constexpr (is_hw_int<T>::value)
  {
    if constexpr (sizeof(T) == 4 && std::is_signed<T>::value) return (T)::atomicMax(reinterpret_cast<int*>(acc), (int)v);
    else if constexpr (sizeof(T) == 4) return (T)::atomicMax(reinterpret_cast<unsigned*>(acc), (unsigned)v);
    else if constexpr (std::is_signed<T>::value) return (T)::atomicMax(reinterpret_cast<long long*>(acc), (long long)v);
    else return (T)::atomicMax(reinterpret_cast<unsigned long long*>(acc), (unsigned long long)v);
  }
  else return atomic_rmw(acc, [v](T o) { return o < v ? v : o; });
}
template<typename T> __device__ __forceinline__ T dev_atomic_and(T* acc, T v)
{
  static_assert(is_hw_int<T>::value, "atomicAnd: 4- or 8-byte integral types");
  return from_word<T>(::atomicAnd(reinterpret_cast<word_t<T>*>(acc), to_word(v)));
}
template<typename T> __device__ __forceinline__ T dev_atomic_or(T* acc, T v)
{
  static_assert(is_hw_int<T>::value, "atomicOr: 4- or 8-byte integral types");
  return from_word<T>(::atomicOr(reinterpret_cast<word_t<T>*>(acc), to_word(v)));
}
template<typename T> __device__ __forceinline__ T dev_atomic_xor(T* acc, T v)
{
  static_assert(is_hw_int<T>::value, "atomicXor: 4- or 8-byte integral types");
  return from_word<T>(::atomicXor(reinterpret_cast<word_t<T>*>(acc), to_word(v)));
}
template<typename T> __device__ __forceinline__ T dev_atomic_exchange(T* acc, T v)
{
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "b200_atomic supports 4- and 8-byte types");
  return from_word<T>(::atomicExch(reinterpret_cast<word_t<T>*>(acc), to_word(v)));
}
// wrap-around forms (CUDA atomicInc / atomicDec semantics, policy/sequential/atomic.hpp:89-112)
template<typename T> __device__ __forceinline__ T dev_atomic_inc_wrap(T* acc, T value)
{
  if constexpr (std::is_same<T, unsigned int>::value) return ::atomicInc(acc, value);
  else return atomic_rmw(acc, [value](T o) { return value <= o ? (T)0 : (T)(o + (T)1); });
}
template<typename T> __device__ __forceinline__ T dev_atomic_dec_wrap(T* acc, T value)
{
  if constexpr (std::is_same<T, unsigned int>::value) return ::atomicDec(acc, value);
  else return atomic_rmw(acc, [value](T o) { return (o == (T)0 || value < o) ? value : (T)(o - (T)1); });
}
template<typename T> __device__ __forceinline__ T dev_atomic_cas(T* acc, T compare, T value)
{
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "b200_atomic supports 4- and 8-byte types");
  return from_word<T>(::atomicCAS(reinterpret_cast<word_t<T>*>(acc), to_word(compare), to_word(value)));
}
}  // namespace b200_detail

#if defined(__CUDA_ARCH__)
#define RAJA_B200_DEVICE_OR_HOST(DEV, HOST) DEV
#else
#define RAJA_B200_DEVICE_OR_HOST(DEV, HOST) HOST
#endif

#define RAJA_B200_ATOMIC_RMW(NAME, DEVICE_EXPR)                                                        \
  RAJA_SUPPRESS_HD_WARN                                                                              \
  template<typename T, typename host_policy>                                                         \
  RAJA_INLINE RAJA_HOST_DEVICE T NAME(policy::b200::b200_atomic_explicit<host_policy>, T* acc, T value) \
  {                                                                                                  \
    RAJA_B200_DEVICE_OR_HOST(return DEVICE_EXPR;, return RAJA::NAME(host_policy {}, acc, value);)      \
  }

RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE T atomicLoad(policy::b200::b200_atomic_explicit<host_policy>, T* acc)
{
  RAJA_B200_DEVICE_OR_HOST(return b200_detail::dev_atomic_load(acc);, return RAJA::atomicLoad(host_policy {}, acc);)
}

RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE void atomicStore(policy::b200::b200_atomic_explicit<host_policy>, T* acc, T value)
{
  RAJA_B200_DEVICE_OR_HOST(b200_detail::dev_atomic_store(acc, value);, RAJA::atomicStore(host_policy {}, acc, value);)
}

RAJA_B200_ATOMIC_RMW(atomicAdd, b200_detail::dev_atomic_add(acc, value))
RAJA_B200_ATOMIC_RMW(atomicSub, b200_detail::dev_atomic_sub(acc, value))
RAJA_B200_ATOMIC_RMW(atomicMin, b200_detail::dev_atomic_min(acc, value))
RAJA_B200_ATOMIC_RMW(atomicMax, b200_detail::dev_atomic_max(acc, value))
RAJA_B200_ATOMIC_RMW(atomicAnd, b200_detail::dev_atomic_and(acc, value))
RAJA_B200_ATOMIC_RMW(atomicOr, b200_detail::dev_atomic_or(acc, value))
RAJA_B200_ATOMIC_RMW(atomicXor, b200_detail::dev_atomic_xor(acc, value))
RAJA_B200_ATOMIC_RMW(atomicExchange, b200_detail::dev_atomic_exchange(acc, value))
RAJA_B200_ATOMIC_RMW(atomicInc, b200_detail::dev_atomic_inc_wrap(acc, value))
RAJA_B200_ATOMIC_RMW(atomicDec, b200_detail::dev_atomic_dec_wrap(acc, value))

RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE T atomicInc(policy::b200::b200_atomic_explicit<host_policy>, T* acc)
{
  RAJA_B200_DEVICE_OR_HOST(return b200_detail::dev_atomic_add(acc, (T)1);, return RAJA::atomicInc(host_policy {}, acc);)
}

RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE T atomicDec(policy::b200::b200_atomic_explicit<host_policy>, T* acc)
{
  RAJA_B200_DEVICE_OR_HOST(return b200_detail::dev_atomic_sub(acc, (T)1);, return RAJA::atomicDec(host_policy {}, acc);)
}

RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE T atomicCAS(policy::b200::b200_atomic_explicit<host_policy>, T* acc, T compare, T value)
{
  RAJA_B200_DEVICE_OR_HOST(return b200_detail::dev_atomic_cas(acc, compare, value);, return RAJA::atomicCAS(host_policy {}, acc, compare, value);)
}

// ---- b200_atomic_aggregated: atomicAdd / atomicSub / atomicInc / atomicDec combine the lanes of a warp that share an
//      address before touching memory; every other operation is the plain one (overload resolution picks the base) ----
RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE T atomicAdd(policy::b200::b200_atomic_aggregated_explicit<host_policy>, T* acc, T value)
{
#if defined(__CUDA_ARCH__)
  if constexpr (std::is_arithmetic<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)) return ::b200::atomic_add_aggregated(acc, value);
  else return b200_detail::dev_atomic_add(acc, value);
#else
  return RAJA::atomicAdd(host_policy {}, acc, value);
#endif
}
RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE T atomicSub(policy::b200::b200_atomic_aggregated_explicit<host_policy>, T* acc, T value)
{
#if defined(__CUDA_ARCH__)
  if constexpr (std::is_arithmetic<T>::value && std::is_signed<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)) return ::b200::atomic_add_aggregated(acc, (T)(-value));
  else return b200_detail::dev_atomic_sub(acc, value);
#else
  return RAJA::atomicSub(host_policy {}, acc, value);
#endif
}

#undef RAJA_B200_ATOMIC_RMW

}  // namespace RAJA

#endif
