// raja_b200 core: warp / block / grid reduction primitives.
//
// Replaces the reference's
//   policy/cuda/intrinsics.hpp:196-533  (shfl_*, warp_reduce, block_reduce)
//   policy/cuda/reduce.hpp:139-214      (grid_reduce_last_block)
//   policy/cuda/reduce.hpp:219-366      (expt::block_reduce / expt::grid_reduce)
// with: redux.sync (__reduce_*_sync) where the ISA has it, a fixed-pattern xor
// butterfly otherwise, one shared-memory hop per block, and a last-block-done
// combine over a PERSISTENT grid (a few hundred partials instead of the
// reference's one partial + one atomicInc per 256 elements).  Slot order is fixed,
// so the result is run-to-run deterministic for a given (N, grid).
#pragma once
#include "common.cuh"

namespace b200 {

// ---- shuffles for any trivially copyable T (32-bit chunks) ------------------
template <class T> B200_D T shfl_xor(const T& v, int lane_mask)
{
  static_assert(sizeof(T) % 4 == 0, "reduction value types must be a multiple of 4 bytes");
  constexpr int W = sizeof(T) / 4;
  uint32_t w[W];
  memcpy(w, &v, sizeof(T));
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = __shfl_xor_sync(FULL_MASK, w[i], lane_mask);
  T r;
  memcpy(&r, w, sizeof(T));
  return r;
}
template <class T> B200_D T shfl_up(const T& v, int delta)
{
  static_assert(sizeof(T) % 4 == 0, "");
  constexpr int W = sizeof(T) / 4;
  uint32_t w[W];
  memcpy(w, &v, sizeof(T));
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = __shfl_up_sync(FULL_MASK, w[i], delta);
  T r;
  memcpy(&r, w, sizeof(T));
  return r;
}
template <class T> B200_D T shfl_idx(const T& v, int src)
{
  static_assert(sizeof(T) % 4 == 0, "");
  constexpr int W = sizeof(T) / 4;
  uint32_t w[W];
  memcpy(w, &v, sizeof(T));
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = __shfl_sync(FULL_MASK, w[i], src);
  T r;
  memcpy(&r, w, sizeof(T));
  return r;
}

template <class T> B200_D T shfl_idx_mask(unsigned mask, const T& v, int src)
{
  constexpr int W = (sizeof(T) + 3) / 4;
  uint32_t w[W] = {};
  memcpy(w, &v, sizeof(T));
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = __shfl_sync(mask, w[i], src);
  T r;
  memcpy(&r, w, sizeof(T));
  return r;
}

// ---- warp all-reduce --------------------------------------------------------
template <class Op> B200_D typename Op::value_type warp_allreduce(typename Op::value_type v, Op op)
{
  using T = typename Op::value_type;
  // redux.sync paths (sm_80+): one instruction instead of 5 shuffle+op pairs.
  if constexpr (std::is_same<Op, Plus<int32_t>>::value) {
    return __reduce_add_sync(FULL_MASK, v);
  } else if constexpr (std::is_same<Op, Plus<uint32_t>>::value) {
    return __reduce_add_sync(FULL_MASK, v);
  } else if constexpr (std::is_same<Op, Minimum<int32_t>>::value || std::is_same<Op, Minimum<uint32_t>>::value) {
    return __reduce_min_sync(FULL_MASK, v);
  } else if constexpr (std::is_same<Op, Maximum<int32_t>>::value || std::is_same<Op, Maximum<uint32_t>>::value) {
    return __reduce_max_sync(FULL_MASK, v);
  } else if constexpr (std::is_same<Op, BitOr<uint32_t>>::value || std::is_same<Op, BitOr<int32_t>>::value) {
    return (T)__reduce_or_sync(FULL_MASK, (unsigned)v);
  } else if constexpr (std::is_same<Op, BitAnd<uint32_t>>::value || std::is_same<Op, BitAnd<int32_t>>::value) {
    return (T)__reduce_and_sync(FULL_MASK, (unsigned)v);
  } else {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      T o = shfl_xor(v, m);
      // keep operand order by lane so non-commutative tie rules stay well defined
      v = (lane_id() & m) ? op(o, v) : op(v, o);
    }
    return v;
  }
}

// strictly left-to-right (lane 0 leftmost) for operators that are associative but not commutative:
// adjacent lanes first, so every partial is a fold of a contiguous run of lanes
template <class Op> B200_D typename Op::value_type warp_allreduce_ordered(typename Op::value_type v, Op op)
{
  using T = typename Op::value_type;
#pragma unroll
  for (int m = 1; m <= 16; m <<= 1) {
    T o = shfl_xor(v, m);
    v = (lane_id() & m) ? op(o, v) : op(v, o);
  }
  return v;
}

// ---- block reduce: result valid in every thread of warp 0 ---------------------
// smem must hold 32 values of T and may be reused after the call returns only
// after a __syncthreads().
template <int BLOCK, class Op>
B200_D typename Op::value_type block_reduce(typename Op::value_type v, Op op, typename Op::value_type* smem)
{
  using T = typename Op::value_type;
  constexpr int NW = BLOCK / WARP;
  static_assert(BLOCK % WARP == 0 && NW <= 32, "BLOCK must be a multiple of 32, at most 1024");
  v = warp_allreduce(v, op);
  if constexpr (NW == 1) return v;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (l < NW) ? smem[l] : Op::identity();
    v = warp_allreduce(v, op);
  }
  return v;
}

// ---- L2-coherent access to another CTA's partial ------------------------------
template <class T> B200_D T ld_cg(const T* p)
{
  static_assert(sizeof(T) % 4 == 0, "");
  constexpr int W = sizeof(T) / 4;
  uint32_t w[W];
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = __ldcg(q + i);
  T r;
  memcpy(&r, w, sizeof(T));
  return r;
}
template <class T> B200_D void st_cg(T* p, const T& v)
{
  static_assert(sizeof(T) % 4 == 0, "");
  constexpr int W = sizeof(T) / 4;
  uint32_t w[W];
  memcpy(w, &v, sizeof(T));
  uint32_t* q = reinterpret_cast<uint32_t*>(p);
#pragma unroll
  for (int i = 0; i < W; ++i) __stcg(q + i, w[i]);
}

// ---- grid combine, last block done ---------------------------------------------
// Every block calls this with its block value (valid in thread 0).  Returns true
// in ALL threads of the one block that arrived last; in that block `result` holds
// the fold of all gridDim.x partials in slot order (valid in warp 0).
// `counter` must be 0 on entry; atomicInc wraps it back to 0, so the workspace can
// be reused by the next launch on the same stream without a memset (the
// "zeroed pool" contract of MemUtils_CUDA.hpp:109-111).
template <int BLOCK, class Op>
B200_D bool grid_combine_last_block(typename Op::value_type block_val,
                                    Op op,
                                    typename Op::value_type* partials,
                                    unsigned* counter,
                                    typename Op::value_type* smem,
                                    typename Op::value_type& result)
{
  using T = typename Op::value_type;
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    st_cg(partials + blockIdx.x, block_val);
    __threadfence();
    unsigned ticket = atomicInc(counter, gridDim.x - 1);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  T acc = Op::identity();
  for (unsigned i = threadIdx.x; i < gridDim.x; i += BLOCK) acc = op(acc, ld_cg(partials + i));
  __syncthreads();  // smem reuse
  result = block_reduce<BLOCK>(acc, op, smem);
  return true;
}

// ---- self-resetting atomic slot (the unordered grid combine) --------------------
// One 8-byte device word that is ZERO when idle: blocks combine into it with one hardware atomic each, the block
// that arrives last takes the result out with an exchange that leaves zero behind, so the word goes back to a
// zeroed pool without a memset and needs no per-launch initialisation with the operator's identity.
//   plus        the value itself (the identity's bit pattern is zero)
//   max / min   atomicMax on an order-preserving unsigned code of v (min: of its complement); 0 = "nothing yet"
//   or / and    atomicOr on v (and: on ~v, De Morgan)
// available == false for operators / value types the ISA has no atomic for (callers fall back to the ordered combine,
// as the reference's atomic tunings do: policy/cuda/reduce.hpp:1001-1030).
template <class T> struct order_code {
  using U = typename std::conditional<sizeof(T) == 4, unsigned, unsigned long long>::type;
  static constexpr U SIGN = U(1) << (8 * sizeof(T) - 1);
  B200_D static U enc(T v)
  {
    U b = bit_cast_<U>(v);
    if constexpr (std::is_floating_point<T>::value) return (b & SIGN) ? ~b : (b | SIGN);
    else if constexpr (std::is_signed<T>::value) return b ^ SIGN;
    else return b;
  }
  B200_D static T dec(U c)
  {
    if constexpr (std::is_floating_point<T>::value) return bit_cast_<T>((c & SIGN) ? (c ^ SIGN) : ~c);
    else if constexpr (std::is_signed<T>::value) return bit_cast_<T>(U(c ^ SIGN));
    else return bit_cast_<T>(c);
  }
};
template <class Op, class = void> struct AtomicSlot {
  static constexpr bool available = false;
};
template <class T> struct AtomicSlot<Plus<T>, std::enable_if_t<std::is_arithmetic<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)>> {
  static constexpr bool available = true;
  using U = typename order_code<T>::U;
  B200_D static void combine(void* slot, T v)
  {
    if constexpr (std::is_floating_point<T>::value) atomicAdd(reinterpret_cast<T*>(slot), v);
    else atomicAdd(reinterpret_cast<U*>(slot), bit_cast_<U>(v));  // two's complement: one add for signed and unsigned
  }
  B200_D static T take(void* slot) { return bit_cast_<T>(atomicExch(reinterpret_cast<U*>(slot), U(0))); }
};
template <class T> struct AtomicSlot<Maximum<T>, std::enable_if_t<std::is_arithmetic<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)>> {
  static constexpr bool available = true;
  using U = typename order_code<T>::U;
  B200_D static void combine(void* slot, T v) { atomicMax(reinterpret_cast<U*>(slot), order_code<T>::enc(v)); }
  B200_D static T take(void* slot)
  {
    const U c = atomicExch(reinterpret_cast<U*>(slot), U(0));
    return c == 0 ? Maximum<T>::identity() : order_code<T>::dec(c);
  }
};
template <class T> struct AtomicSlot<Minimum<T>, std::enable_if_t<std::is_arithmetic<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)>> {
  static constexpr bool available = true;
  using U = typename order_code<T>::U;
  B200_D static void combine(void* slot, T v) { atomicMax(reinterpret_cast<U*>(slot), U(~order_code<T>::enc(v))); }
  B200_D static T take(void* slot)
  {
    const U c = atomicExch(reinterpret_cast<U*>(slot), U(0));
    return c == 0 ? Minimum<T>::identity() : order_code<T>::dec(U(~c));
  }
};
template <class T> struct AtomicSlot<BitOr<T>, std::enable_if_t<std::is_integral<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)>> {
  static constexpr bool available = true;
  using U = typename order_code<T>::U;
  B200_D static void combine(void* slot, T v) { atomicOr(reinterpret_cast<U*>(slot), bit_cast_<U>(v)); }
  B200_D static T take(void* slot) { return bit_cast_<T>(atomicExch(reinterpret_cast<U*>(slot), U(0))); }
};
template <class T> struct AtomicSlot<BitAnd<T>, std::enable_if_t<std::is_integral<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)>> {
  static constexpr bool available = true;
  using U = typename order_code<T>::U;
  B200_D static void combine(void* slot, T v) { atomicOr(reinterpret_cast<U*>(slot), U(~bit_cast_<U>(v))); }
  B200_D static T take(void* slot) { return bit_cast_<T>(U(~atomicExch(reinterpret_cast<U*>(slot), U(0)))); }
};

// Unordered variant ("b200_reduce_atomic"): one atomic per block on the target.
// Only defined for the types/ops the ISA has atomics for.
template <class T> B200_D void atomic_combine(Plus<T>, T* target, T v) { atomicAdd(target, v); }
B200_D void atomic_combine(Plus<int64_t>, int64_t* t, int64_t v) { atomicAdd((unsigned long long*)t, (unsigned long long)v); }
B200_D void atomic_combine(Plus<uint64_t>, uint64_t* t, uint64_t v) { atomicAdd((unsigned long long*)t, (unsigned long long)v); }
B200_D void atomic_combine(Minimum<int32_t>, int32_t* t, int32_t v) { atomicMin(t, v); }
B200_D void atomic_combine(Maximum<int32_t>, int32_t* t, int32_t v) { atomicMax(t, v); }
B200_D void atomic_combine(Minimum<uint32_t>, uint32_t* t, uint32_t v) { atomicMin(t, v); }
B200_D void atomic_combine(Maximum<uint32_t>, uint32_t* t, uint32_t v) { atomicMax(t, v); }
B200_D void atomic_combine(Minimum<int64_t>, int64_t* t, int64_t v) { atomicMin((long long*)t, (long long)v); }
B200_D void atomic_combine(Maximum<int64_t>, int64_t* t, int64_t v) { atomicMax((long long*)t, (long long)v); }
B200_D void atomic_combine(Minimum<uint64_t>, uint64_t* t, uint64_t v) { atomicMin((unsigned long long*)t, (unsigned long long)v); }
B200_D void atomic_combine(Maximum<uint64_t>, uint64_t* t, uint64_t v) { atomicMax((unsigned long long*)t, (unsigned long long)v); }

}  // namespace b200
