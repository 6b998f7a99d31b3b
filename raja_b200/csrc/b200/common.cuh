// raja_b200 core: shared definitions for the sm_100a kernels.
//
// Everything in b200/*.cuh is header-only device code with no dependency on the
// reference (RAJA) headers.  Two consumers instantiate it:
//   * csrc/capi_*.cu          -> libraja_b200.so (the C ABI in include/raja_b200.h)
//   * include/RAJA_b200/*.hpp -> the RAJA::b200_exec policy adapter, instantiated
//                                in the user's TU with the user's lambda.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <type_traits>

namespace b200 {

constexpr int WARP = 32;
constexpr unsigned FULL_MASK = 0xffffffffu;

#define B200_HD __host__ __device__ __forceinline__
#define B200_D __device__ __forceinline__

// ---------------------------------------------------------------------------
// Device properties, cached per device (host side).
// ---------------------------------------------------------------------------
struct DeviceInfo {
  int sm_count;
  int max_smem_optin;
};

inline const DeviceInfo& device_info()
{
  // One slot per device ordinal; filled lazily.  Races are benign (same value).
  static DeviceInfo info[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (info[dev].sm_count == 0) {
    int sms = 0, smem = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    info[dev].max_smem_optin = smem;
    info[dev].sm_count = sms > 0 ? sms : 148;
  }
  return info[dev];
}

// A per-call-site cache with one slot per device ordinal: function attributes (the > 48 KB shared-memory
// opt-in) and occupancy are properties of (kernel, device), so a host thread that moves between devices
// with cudaSetDevice must not reuse what it learned on another one.
struct PerDevice {
  int v[64] = {};
  int& current()
  {
    int dev = 0;
    cudaGetDevice(&dev);
    return v[(dev >= 0 && dev < 64) ? dev : 0];
  }
};

// ---------------------------------------------------------------------------
// Launch hook: every kernel launch of the back-end is bracketed by these calls so
// a host (libraja_b200's profiler, rb200_profile_*) can count launches and time
// each kernel with CUDA events ON THE LAUNCHING STREAM.  Null by default (free).
// ---------------------------------------------------------------------------
struct LaunchHook {
  void (*begin)(const char* kernel, cudaStream_t stream) = nullptr;
  void (*end)(const char* kernel, cudaStream_t stream) = nullptr;
};
inline LaunchHook& launch_hook()
{
  static LaunchHook h;
  return h;
}
struct LaunchScope {
  const char* name;
  cudaStream_t stream;
  LaunchScope(const char* n, cudaStream_t s) : name(n), stream(s)
  {
    if (launch_hook().begin) launch_hook().begin(name, stream);
  }
  ~LaunchScope()
  {
    if (launch_hook().end) launch_hook().end(name, stream);
  }
};

// ---------------------------------------------------------------------------
// Operators.  Semantics follow the reference's RAJA::operators
// (/root/reference/include/RAJA/util/Operators.hpp:247-375 plus/minimum/maximum,
//  :555-594 tie rule "keep lhs", :644-656 limits): identity(plus)=0,
//  identity(minimum)=limits::max, identity(maximum)=limits::min (= -FLT_MAX for
//  floating point, NOT -inf).
// ---------------------------------------------------------------------------
template <class T> struct limits;
template <> struct limits<int32_t>  { B200_HD static int32_t  min() { return INT32_MIN; } B200_HD static int32_t  max() { return INT32_MAX; } };
template <> struct limits<uint32_t> { B200_HD static uint32_t min() { return 0u; }        B200_HD static uint32_t max() { return UINT32_MAX; } };
template <> struct limits<int64_t>  { B200_HD static int64_t  min() { return INT64_MIN; } B200_HD static int64_t  max() { return INT64_MAX; } };
template <> struct limits<uint64_t> { B200_HD static uint64_t min() { return 0ull; }      B200_HD static uint64_t max() { return UINT64_MAX; } };
template <> struct limits<float>    { B200_HD static float    min() { return -3.402823466e+38f; } B200_HD static float  max() { return 3.402823466e+38f; } };
template <> struct limits<double>   { B200_HD static double   min() { return -1.7976931348623157e+308; } B200_HD static double max() { return 1.7976931348623157e+308; } };
// long / unsigned long / long long aliases that are distinct types on LP64
template <> struct limits<long long>          : limits<int64_t>  { B200_HD static long long min() { return INT64_MIN; } B200_HD static long long max() { return INT64_MAX; } };
template <> struct limits<unsigned long long> : limits<uint64_t> { B200_HD static unsigned long long min() { return 0ull; } B200_HD static unsigned long long max() { return UINT64_MAX; } };

template <class T> struct Plus {
  using value_type = T;
  B200_HD static T identity() { return T(0); }
  B200_HD T operator()(T a, T b) const { return a + b; }
};
template <class T> struct Minimum {
  using value_type = T;
  B200_HD static T identity() { return limits<T>::max(); }
  // ties keep lhs (Operators.hpp:555-570)
  B200_HD T operator()(T a, T b) const { return (b < a) ? b : a; }
};
template <class T> struct Maximum {
  using value_type = T;
  B200_HD static T identity() { return limits<T>::min(); }
  B200_HD T operator()(T a, T b) const { return (a < b) ? b : a; }
};
template <class T> struct BitOr {
  using value_type = T;
  B200_HD static T identity() { return T(0); }
  B200_HD T operator()(T a, T b) const { return a | b; }
};
template <class T> struct BitAnd {
  using value_type = T;
  B200_HD static T identity() { return ~T(0); }
  B200_HD T operator()(T a, T b) const { return a & b; }
};

// Operators known to be commutative may be folded in any fixed order; anything else (user functors wrapped by the
// policy layer) is only assumed associative and is folded strictly left to right.
template <class Op> struct is_commutative : std::false_type {};
template <class T> struct is_commutative<Plus<T>> : std::true_type {};
template <class T> struct is_commutative<Minimum<T>> : std::true_type {};
template <class T> struct is_commutative<Maximum<T>> : std::true_type {};
template <class T> struct is_commutative<BitOr<T>> : std::true_type {};
template <class T> struct is_commutative<BitAnd<T>> : std::true_type {};

// (value, location) pair for Min/MaxLoc.  The reference's sequential fold keeps
// the FIRST (lowest-iteration) extremum (pattern/detail/reduce.hpp:127-179 and
// Operators.hpp tie rule); a parallel tree has no iteration order, so the
// combiners below break value ties toward the smaller loc, which reproduces the
// sequential answer whenever loc increases with the iteration number.
template <class T, class I> struct ValLoc {
  T val;
  I loc;
};
template <class T, class I> struct MinLoc {
  using value_type = ValLoc<T, I>;
  B200_HD static value_type identity() { return {limits<T>::max(), I(-1)}; }
  B200_HD value_type operator()(value_type a, value_type b) const
  {
    // An identity operand (loc == -1) never wins a tie against a real element.
    bool take_b = (b.val < a.val) || (!(a.val < b.val) && b.loc != I(-1) && (a.loc == I(-1) || b.loc < a.loc));
    return take_b ? b : a;
  }
};
template <class T, class I> struct MaxLoc {
  using value_type = ValLoc<T, I>;
  B200_HD static value_type identity() { return {limits<T>::min(), I(-1)}; }
  B200_HD value_type operator()(value_type a, value_type b) const
  {
    bool take_b = (a.val < b.val) || (!(b.val < a.val) && b.loc != I(-1) && (a.loc == I(-1) || b.loc < a.loc));
    return take_b ? b : a;
  }
};

// ---------------------------------------------------------------------------
// Cache-hinted global memory access, expressed with builtins so the compiler
// keeps its freedom to hoist read-only loads above stores when it unrolls:
//   ld_stream           ld.global.nc   (read-only for the kernel's lifetime)
//   ld_stream_coherent  ld.global.cs   (buffer may also be written by this kernel)
//   st_stream           st.global.cs   (written once, not re-read by this kernel)
// ---------------------------------------------------------------------------
template <class T, class R> B200_D T bit_cast_(const R& r)
{
  static_assert(sizeof(T) == sizeof(R), "");
  T t;
  memcpy(&t, &r, sizeof(T));
  return t;
}
template <class T> B200_D T ld_stream(const T* p)
{
  if constexpr (sizeof(T) == 16) {
    return bit_cast_<T>(__ldg(reinterpret_cast<const uint4*>(p)));
  } else if constexpr (sizeof(T) == 8) {
    return bit_cast_<T>(__ldg(reinterpret_cast<const uint2*>(p)));
  } else if constexpr (sizeof(T) == 4) {
    return bit_cast_<T>(__ldg(reinterpret_cast<const unsigned*>(p)));
  } else {
    return *p;
  }
}
template <class T> B200_D T ld_stream_coherent(const T* p)
{
  if constexpr (sizeof(T) == 16) {
    return bit_cast_<T>(__ldcs(reinterpret_cast<const uint4*>(p)));
  } else if constexpr (sizeof(T) == 8) {
    return bit_cast_<T>(__ldcs(reinterpret_cast<const uint2*>(p)));
  } else if constexpr (sizeof(T) == 4) {
    return bit_cast_<T>(__ldcs(reinterpret_cast<const unsigned*>(p)));
  } else {
    return *p;
  }
}
template <class T> B200_D void st_stream(T* p, const T& v)
{
  if constexpr (sizeof(T) == 16) {
    __stcs(reinterpret_cast<uint4*>(p), bit_cast_<uint4>(v));
  } else if constexpr (sizeof(T) == 8) {
    __stcs(reinterpret_cast<uint2*>(p), bit_cast_<uint2>(v));
  } else if constexpr (sizeof(T) == 4) {
    __stcs(reinterpret_cast<unsigned*>(p), bit_cast_<unsigned>(v));
  } else {
    *p = v;
  }
}

// Release / acquire at gpu scope for inter-CTA flags (decoupled look-back).
B200_D void st_release_u32(uint32_t* p, uint32_t v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
B200_D uint32_t ld_acquire_u32(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
B200_D void st_release_u64(uint64_t* p, uint64_t v)
{
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
B200_D uint64_t ld_acquire_u64(const uint64_t* p)
{
  uint64_t v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
B200_D void st_relaxed_u64(uint64_t* p, uint64_t v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
B200_D uint64_t ld_relaxed_u64(const uint64_t* p)
{
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
B200_D void st_relaxed_u32(uint32_t* p, uint32_t v)
{
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
B200_D uint32_t ld_relaxed_u32(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

B200_D int lane_id() { return threadIdx.x & 31; }
B200_D unsigned lanemask_lt()
{
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// V consecutive elements moved as one 4/8/16-byte access
template <class T, int V> struct alignas(sizeof(T) * V) Vec {
  T e[V];
};

template <class T> B200_HD constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

}  // namespace b200
