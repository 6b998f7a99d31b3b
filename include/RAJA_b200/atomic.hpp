// RAJA::atomicX(b200_atomic_explicit<host_policy>, ...): replaces policy/cuda/atomic.hpp:665-860.
// Device code: hardware atomics where the ISA has them for the type, otherwise a CAS loop on the
// same-width unsigned word (what the reference does, policy/cuda/atomic.hpp:328-402).  Host code
// falls through to host_policy exactly like the reference (:690-700).
// Must be seen before RAJA/pattern/atomic.hpp (its calls are qualified).
#ifndef RAJA_B200_ATOMIC_HPP
#define RAJA_B200_ATOMIC_HPP

namespace RAJA
{
namespace b200_detail
{
template<size_t N> struct cas_word;
template<> struct cas_word<4> { using type = unsigned int; };
template<> struct cas_word<8> { using type = unsigned long long; };

// old = *acc; *acc = f(old); return old -- atomically, for any 4- or 8-byte trivially copyable T
template<typename T, typename F>
__device__ __forceinline__ T atomic_rmw(T* acc, F f)
{
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "b200_atomic supports 4- and 8-byte types");
  using W = typename cas_word<sizeof(T)>::type;
  W* addr = reinterpret_cast<W*>(acc);
  W seen  = *reinterpret_cast<volatile W*>(addr);
  while (true)
  {
    T old_t;
    memcpy(&old_t, &seen, sizeof(T));
    const T new_t = f(old_t);
    W desired;
    memcpy(&desired, &new_t, sizeof(T));
    const W prev = ::atomicCAS(addr, seen, desired);
    if (prev == seen) return old_t;
    seen = prev;
  }
}

// add: builtins for the types the ISA covers (same list as b200/atomic.cuh)
template<typename T> struct has_hw_add : std::false_type {};
template<> struct has_hw_add<int> : std::true_type {};
template<> struct has_hw_add<unsigned int> : std::true_type {};
template<> struct has_hw_add<long> : std::true_type {};
template<> struct has_hw_add<unsigned long> : std::true_type {};
template<> struct has_hw_add<long long> : std::true_type {};
template<> struct has_hw_add<unsigned long long> : std::true_type {};
template<> struct has_hw_add<float> : std::true_type {};
template<> struct has_hw_add<double> : std::true_type {};

template<typename T>
__device__ __forceinline__ T dev_atomic_add(T* acc, T v)
{
  if constexpr (has_hw_add<T>::value) return ::b200::atomic_add(acc, v);
  else return atomic_rmw(acc, [v](T o) { return (T)(o + v); });
}
template<typename T>
__device__ __forceinline__ T dev_atomic_min(T* acc, T v)
{
  if constexpr (std::is_same<T, int>::value || std::is_same<T, unsigned int>::value || std::is_same<T, unsigned long long>::value ||
                std::is_same<T, long long>::value)
    return ::atomicMin(acc, v);
  else return atomic_rmw(acc, [v](T o) { return v < o ? v : o; });
}
template<typename T>
__device__ __forceinline__ T dev_atomic_max(T* acc, T v)
{
  if constexpr (std::is_same<T, int>::value || std::is_same<T, unsigned int>::value || std::is_same<T, unsigned long long>::value ||
                std::is_same<T, long long>::value)
    return ::atomicMax(acc, v);
  else return atomic_rmw(acc, [v](T o) { return o < v ? v : o; });
}
}  // namespace b200_detail

#define RAJA_B200_ATOMIC_RMW(NAME, DEVICE_EXPR)                                                        \
  RAJA_SUPPRESS_HD_WARN                                                                              \
  template<typename T, typename host_policy>                                                         \
  RAJA_INLINE RAJA_HOST_DEVICE T NAME(policy::b200::b200_atomic_explicit<host_policy>, T* acc, T value) \
  {                                                                                                  \
    RAJA_B200_DEVICE_OR_HOST(return DEVICE_EXPR;, return RAJA::NAME(host_policy {}, acc, value);)      \
  }

#if defined(__CUDA_ARCH__)
#define RAJA_B200_DEVICE_OR_HOST(DEV, HOST) DEV
#else
#define RAJA_B200_DEVICE_OR_HOST(DEV, HOST) HOST
#endif

RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE T atomicLoad(policy::b200::b200_atomic_explicit<host_policy>, T* acc)
{
  RAJA_B200_DEVICE_OR_HOST(return b200_detail::atomic_rmw(acc, [](T o) { return o; });, return RAJA::atomicLoad(host_policy {}, acc);)
}

RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE void atomicStore(policy::b200::b200_atomic_explicit<host_policy>, T* acc, T value)
{
  RAJA_B200_DEVICE_OR_HOST((void)b200_detail::atomic_rmw(acc, [value](T) { return value; });, RAJA::atomicStore(host_policy {}, acc, value);)
}

RAJA_B200_ATOMIC_RMW(atomicAdd, b200_detail::dev_atomic_add(acc, value))
RAJA_B200_ATOMIC_RMW(atomicSub, b200_detail::atomic_rmw(acc, [value](T o) { return (T)(o - value); }))
RAJA_B200_ATOMIC_RMW(atomicMin, b200_detail::dev_atomic_min(acc, value))
RAJA_B200_ATOMIC_RMW(atomicMax, b200_detail::dev_atomic_max(acc, value))
RAJA_B200_ATOMIC_RMW(atomicAnd, b200_detail::atomic_rmw(acc, [value](T o) { return (T)(o & value); }))
RAJA_B200_ATOMIC_RMW(atomicOr, b200_detail::atomic_rmw(acc, [value](T o) { return (T)(o | value); }))
RAJA_B200_ATOMIC_RMW(atomicXor, b200_detail::atomic_rmw(acc, [value](T o) { return (T)(o ^ value); }))
RAJA_B200_ATOMIC_RMW(atomicExchange, b200_detail::atomic_rmw(acc, [value](T) { return value; }))
// wrap-around forms (CUDA atomicInc / atomicDec semantics, policy/sequential/atomic.hpp:89-112)
RAJA_B200_ATOMIC_RMW(atomicInc, b200_detail::atomic_rmw(acc, [value](T o) { return value <= o ? (T)0 : (T)(o + (T)1); }))
RAJA_B200_ATOMIC_RMW(atomicDec, b200_detail::atomic_rmw(acc, [value](T o) { return (o == (T)0 || value < o) ? value : (T)(o - (T)1); }))

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
  RAJA_B200_DEVICE_OR_HOST(return b200_detail::atomic_rmw(acc, [](T o) { return (T)(o - (T)1); });, return RAJA::atomicDec(host_policy {}, acc);)
}

RAJA_SUPPRESS_HD_WARN
template<typename T, typename host_policy>
RAJA_INLINE RAJA_HOST_DEVICE T atomicCAS(policy::b200::b200_atomic_explicit<host_policy>, T* acc, T compare, T value)
{
#if defined(__CUDA_ARCH__)
  using W = typename b200_detail::cas_word<sizeof(T)>::type;
  W c, v;
  memcpy(&c, &compare, sizeof(T));
  memcpy(&v, &value, sizeof(T));
  const W prev = ::atomicCAS(reinterpret_cast<W*>(acc), c, v);
  T r;
  memcpy(&r, &prev, sizeof(T));
  return r;
#else
  return RAJA::atomicCAS(host_policy {}, acc, compare, value);
#endif
}

#undef RAJA_B200_ATOMIC_RMW

}  // namespace RAJA

#endif
