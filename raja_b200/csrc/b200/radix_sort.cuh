// raja_b200 core: stable LSD radix sort ("onesweep": one read + one write of the
// data per 8-bit digit, cross-tile offsets by decoupled look-back).
//
// Replaces the reference's policy/cuda/sort.hpp:101-227 (keys) and :391-544
// (pairs), which delegate to cub::DeviceRadixSort over all key bits.  Contract
// kept: ascending (operators::less) or descending (operators::greater), stable,
// arithmetic keys only, values permuted alongside.  Differences by design:
//   * 64-bit lengths (reference: int, sort.hpp:115,407);
//   * the pass count is even for every supported key width, and the ping-pong
//     starts in the caller's buffer, so the sorted data always ends where the
//     caller expects it -- no copy-back (reference: sort.hpp:143-151, 438-451).
//
// Order semantics match the sequential oracle's operator< on the key type
// (util/sort.hpp merge_sort via policy/sequential/sort.hpp:95-107,137-157):
// signed keys flip the sign bit, floating keys use the usual sign-magnitude ->
// biased mapping, and -0.0 is ranked as +0.0 (they compare equal under <, so a
// stable sort must keep their input order) while the stored key keeps its sign.
// NaN keys are outside the parity domain (operator< is not a strict weak order).
//
// Kernels:
//   histogram_kernel  : one read of the keys, all digit histograms at once
//                       (shared-memory privatised, then global atomics)
//   scan_bins_kernel  : exclusive scan of each 256-bin histogram
//   onesweep_kernel   : per pass, one CTA per tile (tiles claimed by ticket, so a
//                       tile's predecessors are always running or done):
//       A  keys -> registers (warp-striped, coalesced); the tile's VALUES are pulled
//          toward L2 by one bulk prefetch (cp.async.bulk.prefetch.L2) so the late
//          value loads of phase C do not pay DRAM latency;
//       B  early counts: per-warp digit histograms by shared-memory atomics ->
//          per-digit tile counts, PUBLISHED for the successors before any ranking
//          work is done -> exclusive bases per (warp, digit);
//       C  compute warps: stable ranks by an 8-ballot match (the match.any
//          instruction is issue-limited on this part: v1 spent 50 % of its warp
//          samples waiting on it, profiles/r01_onesweep_v1_stalls.txt), leader
//          fetch-and-add on the warp's digit counter, scatter into shared memory
//          in rank order;
//          TWO LOOK-BACK WARPS, concurrently: batched decoupled look-back, one
//          128-bit load = 4 digits of one predecessor tile, 8 predecessors in
//          flight per lane -- off the critical path unless it outlasts phase C;
//       D  run-coalesced write-out with 32-bit index arithmetic.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace b200 {
namespace rsort {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;

enum KeyKind : int { KEY_UNSIGNED = 0, KEY_SIGNED = 1, KEY_FLOAT = 2 };

struct KeyOrder {
  int kind;
  int descending;
};

// order-preserving map onto unsigned bits
template <class U> B200_D U to_ordered(U k, KeyOrder o)
{
  constexpr int BITS = 8 * sizeof(U);
  constexpr U SIGN = U(1) << (BITS - 1);
  if (o.kind == KEY_FLOAT) {
    if (k == SIGN) k = 0;  // -0.0 ranks as +0.0
    k ^= (k & SIGN) ? ~U(0) : SIGN;
  } else if (o.kind == KEY_SIGNED) {
    k ^= SIGN;
  }
  if (o.descending) k = ~k;
  return k;
}
template <class U> B200_D int digit_of(U k, KeyOrder o, int shift) { return (int)((to_ordered(k, o) >> shift) & U(RADIX - 1)); }

// Branch-free digit extraction for the sweep kernel.  For integer keys the order-preserving
// map is an XOR with a constant, which commutes with the shift and mask: digit =
// ((k >> shift) & 255) ^ dxor with dxor precomputed on the host.  Floating keys need the
// sign-dependent mask (and -0.0 ranked as +0.0) before the same step.
template <class U> inline unsigned digit_xor(KeyOrder o, int shift)
{
  constexpr int BITS = 8 * sizeof(U);
  U x = 0;
  if (o.kind == KEY_SIGNED || o.kind == KEY_FLOAT) x = U(1) << (BITS - 1);
  if (o.descending) x = ~x;
  return (unsigned)((x >> shift) & U(RADIX - 1));
}
template <bool FLOAT, class U> B200_D unsigned sweep_digit(U k, int shift, unsigned dxor)
{
  if constexpr (FLOAT) {
    constexpr int BITS = 8 * sizeof(U);
    constexpr U SIGN = U(1) << (BITS - 1);
    using SU = typename std::make_signed<U>::type;
    if (k == SIGN) k = 0;                                  // -0.0 ranks as +0.0
    k ^= (U)((SU)k >> (BITS - 1)) & ~SIGN;                  // negative: flip the magnitude bits too
  }
  return ((unsigned)(k >> shift) & (RADIX - 1)) ^ dxor;
}


// ---------------------------------------------------------------------------
// 1. histograms of every digit position in one read of the keys
// ---------------------------------------------------------------------------
template <class U, int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK)
histogram_kernel(const U* __restrict__ keys, const long long n, unsigned long long* __restrict__ ghist, const KeyOrder order)
{
  constexpr int NPASS = sizeof(U);
  __shared__ uint32_t sh[NPASS][RADIX];
  for (int i = threadIdx.x; i < NPASS * RADIX; i += BLOCK) (&sh[0][0])[i] = 0;
  __syncthreads();

  constexpr long long TILE = (long long)BLOCK * UNROLL;
  const long long full_tiles = n / TILE;
  // a block flushes before any 32-bit shared counter could overflow
  for (long long t = blockIdx.x; t < full_tiles; t += gridDim.x) {
    U k[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) k[u] = ld_stream(keys + t * TILE + (long long)u * BLOCK + threadIdx.x);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      U ok = to_ordered(k[u], order);
#pragma unroll
      for (int p = 0; p < NPASS; ++p) atomicAdd(&sh[p][(int)((ok >> (p * RADIX_BITS)) & U(RADIX - 1))], 1u);
    }
  }
  if (blockIdx.x == (unsigned)(full_tiles % gridDim.x)) {
    for (long long i = full_tiles * TILE + threadIdx.x; i < n; i += BLOCK) {
      U ok = to_ordered(keys[i], order);
#pragma unroll
      for (int p = 0; p < NPASS; ++p) atomicAdd(&sh[p][(int)((ok >> (p * RADIX_BITS)) & U(RADIX - 1))], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NPASS * RADIX; i += BLOCK) {
    uint32_t c = (&sh[0][0])[i];
    if (c) atomicAdd(ghist + i, (unsigned long long)c);
  }
}

// ---------------------------------------------------------------------------
// 2. exclusive scan of each pass's 256 bins (one block per pass, 256 threads)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(RADIX) scan_bins_kernel(const unsigned long long* __restrict__ ghist, unsigned long long* __restrict__ bin_base)
{
  __shared__ unsigned long long s_w[RADIX / WARP];
  const int d = threadIdx.x, p = blockIdx.x;
  const unsigned long long c = ghist[p * RADIX + d];
  unsigned long long s = c;
#pragma unroll
  for (int k = 1; k < WARP; k <<= 1) {
    unsigned long long o = __shfl_up_sync(FULL_MASK, s, k);
    if ((d & 31) >= k) s += o;
  }
  if ((d & 31) == 31) s_w[d >> 5] = s;
  __syncthreads();
  unsigned long long pre = 0;
#pragma unroll
  for (int i = 0; i < RADIX / WARP; ++i)
    if (i < (d >> 5)) pre += s_w[i];
  bin_base[p * RADIX + d] = pre + s - c;
}

// ---------------------------------------------------------------------------
// 3. one onesweep pass
// ---------------------------------------------------------------------------
struct NoValue {};

template <class C> struct LookbackFlags;
template <> struct LookbackFlags<uint32_t> {
  static constexpr uint32_t LOCAL = 1u << 30, GLOBAL = 2u << 30, MASK = (1u << 30) - 1;
};
template <> struct LookbackFlags<unsigned long long> {
  static constexpr unsigned long long LOCAL = 1ull << 62, GLOBAL = 2ull << 62, MASK = (1ull << 62) - 1;
};

B200_D void st_lookback(uint32_t* p, uint32_t v) { st_relaxed_u32(p, v); }
B200_D void st_lookback(unsigned long long* p, unsigned long long v) { st_relaxed_u64((uint64_t*)p, v); }

// four consecutive digits of one tile, one load / one store
template <class C> struct Quad {
  C v[4];
};
B200_D Quad<uint32_t> ld_quad(const uint32_t* p)
{
  Quad<uint32_t> q;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.v[0]), "=r"(q.v[1]), "=r"(q.v[2]), "=r"(q.v[3]) : "l"(p) : "memory");
  return q;
}
B200_D void st_quad(uint32_t* p, const Quad<uint32_t>& q)
{
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(q.v[0]), "r"(q.v[1]), "r"(q.v[2]), "r"(q.v[3]) : "memory");
}
B200_D Quad<unsigned long long> ld_quad(const unsigned long long* p)
{
  Quad<unsigned long long> q;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(q.v[0]), "=l"(q.v[1]) : "l"(p) : "memory");
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(q.v[2]), "=l"(q.v[3]) : "l"(p + 2) : "memory");
  return q;
}
B200_D void st_quad(unsigned long long* p, const Quad<unsigned long long>& q)
{
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(q.v[0]), "l"(q.v[1]) : "memory");
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p + 2), "l"(q.v[2]), "l"(q.v[3]) : "memory");
}

// mask of the lanes whose 8-bit digit equals this lane's (all 32 lanes participate)
template <int MODE> B200_D unsigned match_digit(unsigned d)
{
  if constexpr (MODE == 1) {
    return __match_any_sync(FULL_MASK, d);
  } else {
    unsigned peers = FULL_MASK;
#pragma unroll
    for (int b = 0; b < RADIX_BITS; ++b) {
      const bool bit = (d >> b) & 1u;
      const unsigned m = __ballot_sync(FULL_MASK, bit);
      peers &= bit ? m : ~m;
    }
    return peers;
  }
}

// Compile-time shape of the sweep kernel: BLOCK compute threads (+ 64 look-back threads),
// ITEMS keys per compute thread, MIN_CTAS resident CTAs per SM the register budget is cut for,
// MATCH 0 = ballot loop / 1 = match.any / 2 = shared-memory atomicOr masks.
template <int BLOCK_, int ITEMS_, int MIN_CTAS_, int MATCH_ = 0> struct SweepShape {
  static constexpr int BLOCK = BLOCK_, ITEMS = ITEMS_, MIN_CTAS = MIN_CTAS_, MATCH = MATCH_;
  static constexpr int LB_WARPS = 2;
  static constexpr int THREADS = BLOCK + LB_WARPS * WARP;
  static constexpr int TILE = BLOCK * ITEMS;
  static_assert(BLOCK % WARP == 0 && BLOCK >= RADIX, "one digit-owner thread per digit");
  static_assert(TILE < 65536, "tile-local counts must fit the look-back word next to its flags");
};

// per (warp, digit): running rank of the digit for this warp, and the peer mask of the row
// being ranked (MATCH == 2).  8 bytes, read with one LDS.64.
struct WarpDigit {
  uint32_t count;
  uint32_t mask;
};

template <class U, class VT, class C, class S> struct OnesweepSmem {
  static constexpr int NW = S::BLOCK / WARP;
  static constexpr bool HAS_VALUES = !std::is_same<VT, NoValue>::value;
  // global positions fit 32 bits whenever the look-back counters do (n < 2^30)
  using G = typename std::conditional<sizeof(C) == 4, uint32_t, unsigned long long>::type;
  WarpDigit whist[NW][RADIX];  // per-warp digit counts -> running rank of (warp, digit)
  uint32_t count[RADIX];       // digit counts of this tile
  uint32_t texcl[RADIX];       // first rank of each digit inside this tile
  G gbase[RADIX];              // global position of tile-rank r of digit d = gbase[d] + r (mod 2^bits)
  uint32_t wtot[RADIX / WARP];
  long long tile;
  U keys[S::TILE];
  typename std::conditional<HAS_VALUES, VT, char>::type vals[HAS_VALUES ? S::TILE : 1];
};

template <class U, class VT, bool FLOAT, class C, class S>
__global__ void __launch_bounds__(S::THREADS, S::MIN_CTAS)
onesweep_kernel(const U* __restrict__ keys_in, U* __restrict__ keys_out,
                const VT* __restrict__ vals_in, VT* __restrict__ vals_out,
                const long long n,
                const unsigned long long* __restrict__ bin_base,  // [256] of this pass
                C* lookback,                                      // [ntiles][256], zero on entry
                unsigned* ticket,                                 // zero on entry
                const int shift, const unsigned dxor)
{
  using SM = OnesweepSmem<U, VT, C, S>;
  using F = LookbackFlags<C>;
  using G = typename SM::G;
  constexpr int BLOCK = S::BLOCK, ITEMS = S::ITEMS, TILE = S::TILE, NW = SM::NW;
  constexpr bool HAS_VALUES = SM::HAS_VALUES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& s = *reinterpret_cast<SM*>(smem_raw);

  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const bool compute = tid < BLOCK;
  if (tid == 0) s.tile = (long long)atomicAdd(ticket, 1u);
  for (int i = tid; i < NW * RADIX * 2; i += S::THREADS) (&s.whist[0][0].count)[i] = 0;
  __syncthreads();
  const long long tile = s.tile;
  const long long tile_base = tile * TILE;
  const int tile_count = (int)((n - tile_base) < (long long)TILE ? (n - tile_base) : (long long)TILE);
  const bool full = (tile_count == TILE);
  const int wofs = w * (WARP * ITEMS);  // this warp's first item within the tile

  // ---- phase A: keys, warp-striped: item (j, l) is tile item wofs + j*32 + l --------------
  U key[ITEMS];
  if (compute) {
    if (full) {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) key[j] = ld_stream(keys_in + tile_base + wofs + j * WARP + l);
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const int it = wofs + j * WARP + l;
        key[j] = (it < tile_count) ? ld_stream(keys_in + tile_base + it) : U(0);
      }
    }
  }
  if constexpr (HAS_VALUES) {
    if (tid == BLOCK) {  // pull the tile's values toward L2 while the keys are being ranked
      const uintptr_t lo = ((uintptr_t)(vals_in + tile_base) + 15u) & ~(uintptr_t)15u;
      const uintptr_t hi = (uintptr_t)(vals_in + tile_base + tile_count) & ~(uintptr_t)15u;
      if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned)(hi - lo)) : "memory");
    }
  }

  // ---- phase B: early counts ------------------------------------------------------------------
  if (compute) {
    WarpDigit* const wh = s.whist[w];
    if (full) {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) atomicAdd(&wh[sweep_digit<FLOAT>(key[j], shift, dxor)].count, 1u);
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j)
        if (wofs + j * WARP + l < tile_count) atomicAdd(&wh[sweep_digit<FLOAT>(key[j], shift, dxor)].count, 1u);
    }
  }
  __syncthreads();
  if (tid < RADIX) {
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const uint32_t t = s.whist[i][tid].count;
      s.whist[i][tid].count = cnt;
      cnt += t;
    }
    // successors only need this tile's counts: publish them before any ranking work
    st_lookback(lookback + tile * RADIX + tid, (C)cnt | (tile == 0 ? F::GLOBAL : F::LOCAL));
    s.count[tid] = cnt;
    uint32_t incl = cnt;
#pragma unroll
    for (int k = 1; k < WARP; k <<= 1) {
      const uint32_t o = __shfl_up_sync(FULL_MASK, incl, k);
      if (l >= k) incl += o;
    }
    if (l == 31) s.wtot[w] = incl;
    asm volatile("bar.sync 1, %0;" ::"n"(RADIX) : "memory");
    uint32_t pre = 0;
#pragma unroll
    for (int i = 0; i < RADIX / WARP; ++i)
      if (i < w) pre += s.wtot[i];
    const uint32_t tile_excl = pre + incl - cnt;
    s.texcl[tid] = tile_excl;
#pragma unroll
    for (int i = 0; i < NW; ++i) s.whist[i][tid].count += tile_excl;
  }
  __syncthreads();

  // ---- phase C ----------------------------------------------------------------------------------
  if (compute) {
    // stable ranks + scatter into shared memory in rank order
    WarpDigit* const wh = s.whist[w];
    const unsigned lt = lanemask_lt();
    uint16_t rank[ITEMS];
    auto rank_and_scatter = [&](auto full_c) {
      constexpr bool FULL = decltype(full_c)::value;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const bool valid = FULL || (wofs + j * WARP + l < tile_count);
        const unsigned d = sweep_digit<FLOAT>(key[j], shift, dxor);
        uint32_t r;
        if constexpr (S::MATCH == 2) {
          // peers through shared memory: every lane ORs its bit into the digit's row mask, reads
          // (running count, mask) back with one 64-bit load, and the lowest peer advances the
          // count and clears the mask for the next row
          if (valid) atomicOr(&wh[d].mask, 1u << l);
          __syncwarp();
          const uint2 cm = *reinterpret_cast<const uint2*>(&wh[d]);
          __syncwarp();
          if (valid && (cm.y & lt) == 0) *reinterpret_cast<uint2*>(&wh[d]) = make_uint2(cm.x + __popc(cm.y), 0u);
          __syncwarp();
          r = cm.x + __popc(cm.y & lt);
        } else {
          unsigned peers = match_digit<S::MATCH>(d);
          if (!FULL) peers &= __ballot_sync(FULL_MASK, valid);
          const int leader = __ffs(peers) - 1;
          uint32_t before = 0;
          if (l == leader && valid) before = atomicAdd(&wh[d].count, (uint32_t)__popc(peers));
          before = __shfl_sync(FULL_MASK, before, leader);
          r = before + __popc(peers & lt);
        }
        rank[j] = (uint16_t)r;
        if (valid) s.keys[r] = key[j];
      }
    };
    if (full) rank_and_scatter(std::true_type{});
    else rank_and_scatter(std::false_type{});
    if constexpr (HAS_VALUES) {
      if (full) {
        VT val[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) val[j] = ld_stream(vals_in + tile_base + wofs + j * WARP + l);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) s.vals[rank[j]] = val[j];
      } else {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          const int it = wofs + j * WARP + l;
          if (it < tile_count) s.vals[rank[j]] = ld_stream(vals_in + tile_base + it);
        }
      }
    }
  } else {
    // look-back warps: lane owns digits q .. q+3
    const int q = (tid - BLOCK) * 4;
    C prev[4] = {0, 0, 0, 0};
    if (tile != 0) {
      constexpr int B = sizeof(C) == 4 ? 8 : 4;  // predecessor tiles in flight per lane
      bool done[4] = {false, false, false, false};
      long long p = tile - 1;
      while (true) {
        Quad<C> v[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
          if (p - b >= 0) v[b] = ld_quad(lookback + (p - b) * RADIX + q);
          else v[b] = Quad<C>{{F::GLOBAL, F::GLOBAL, F::GLOBAL, F::GLOBAL}};  // before the first tile: zero, final
        }
        int consumed = 0;
        bool all_done = false;
#pragma unroll
        for (int b = 0; b < B; ++b) {
          if (all_done) break;
          bool ok = true;
#pragma unroll
          for (int c = 0; c < 4; ++c) ok = ok && (done[c] || (v[b].v[c] & ~F::MASK) != 0);
          if (!__all_sync(FULL_MASK, ok)) break;  // a predecessor has not published yet: re-poll from here
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (!done[c]) {
              prev[c] += v[b].v[c] & F::MASK;
              if (v[b].v[c] & F::GLOBAL) done[c] = true;
            }
          }
          ++consumed;
          all_done = __all_sync(FULL_MASK, done[0] && done[1] && done[2] && done[3]);
        }
        if (all_done) break;
        p -= consumed;
        if (consumed < B) __nanosleep(consumed == 0 ? 400 : 100);  // do not steal issue slots from the ranking warps
      }
      Quad<C> g;
#pragma unroll
      for (int c = 0; c < 4; ++c) g.v[c] = (C)(prev[c] + s.count[q + c]) | F::GLOBAL;
      st_quad(lookback + tile * RADIX + q, g);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) s.gbase[q + c] = (G)(bin_base[q + c] + (unsigned long long)prev[c]) - (G)s.texcl[q + c];
  }
  __syncthreads();

  // ---- phase D: run-coalesced write-out ------------------------------------------------------------
  if (compute) {
    if (full) {
#pragma unroll
      for (int m = 0; m < ITEMS; ++m) {
        const int i = m * BLOCK + tid;
        const U k = s.keys[i];
        const G dst = s.gbase[sweep_digit<FLOAT>(k, shift, dxor)] + (G)i;
        keys_out[dst] = k;
        if constexpr (HAS_VALUES) vals_out[dst] = s.vals[i];
      }
    } else {
#pragma unroll
      for (int m = 0; m < ITEMS; ++m) {
        const int i = m * BLOCK + tid;
        if (i < tile_count) {
          const U k = s.keys[i];
          const G dst = s.gbase[sweep_digit<FLOAT>(k, shift, dxor)] + (G)i;
          keys_out[dst] = k;
          if constexpr (HAS_VALUES) vals_out[dst] = s.vals[i];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
template <class U, class VT> struct Tuning {
  static constexpr int ITEMS = sizeof(U) == 4 ? (sizeof(VT) <= 4 ? 24 : 16) : (std::is_same<VT, NoValue>::value ? 16 : 12);
  using shape = SweepShape<384, ITEMS, 2, 2>;
};

template <class U, class VT, class S = typename Tuning<U, VT>::shape> inline long long num_tiles(long long n)
{
  return (n + S::TILE - 1) / S::TILE;
}

// workspace layout: [ghist NPASS*256 u64][bin_base NPASS*256 u64][tickets NPASS u32, padded][lookback ntiles*256*C]
template <class U, class VT, class S = typename Tuning<U, VT>::shape> inline size_t workspace_bytes(long long n)
{
  constexpr int NPASS = sizeof(U);
  size_t header = (size_t)NPASS * RADIX * 8 * 2 + 256;
  size_t cbytes = (n < (1ll << 30)) ? 4 : 8;
  return header + (size_t)num_tiles<U, VT, S>(n) * RADIX * cbytes;
}

// opt-in to > 48 KB of dynamic shared memory, once per kernel (keyed on the kernel itself:
// kernels of one signature share a function-pointer TYPE) and host thread
template <auto Kern> inline cudaError_t configure_once(int smem_bytes)
{
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  return cudaSuccess;
}

template <class U, class VT, class C, class S>
inline cudaError_t launch_pass(cudaStream_t stream, const U* kin, U* kout, const VT* vin, VT* vout, long long n,
                               const unsigned long long* bin_base, void* lookback, unsigned* ticket, int shift, KeyOrder order)
{
  using SM = OnesweepSmem<U, VT, C, S>;
  const long long ntiles = num_tiles<U, VT, S>(n);
  const unsigned dxor = digit_xor<U>(order, shift);
  LaunchScope scope("onesweep_kernel", stream);
  if (order.kind == KEY_FLOAT) {
    cudaError_t e = configure_once<onesweep_kernel<U, VT, true, C, S>>((int)sizeof(SM));
    if (e != cudaSuccess) return e;
    onesweep_kernel<U, VT, true, C, S><<<(unsigned)ntiles, S::THREADS, sizeof(SM), stream>>>(kin, kout, vin, vout, n, bin_base, (C*)lookback, ticket, shift, dxor);
  } else {
    cudaError_t e = configure_once<onesweep_kernel<U, VT, false, C, S>>((int)sizeof(SM));
    if (e != cudaSuccess) return e;
    onesweep_kernel<U, VT, false, C, S><<<(unsigned)ntiles, S::THREADS, sizeof(SM), stream>>>(kin, kout, vin, vout, n, bin_base, (C*)lookback, ticket, shift, dxor);
  }
  return cudaGetLastError();
}

// Sorts n keys (and values) held in (keys, vals) using (keys_alt, vals_alt) as the
// ping-pong partner.  The pass count is even, so the result lands in (keys, vals).
template <class U, class VT, class S = typename Tuning<U, VT>::shape>
inline cudaError_t sort(cudaStream_t stream, U* keys, U* keys_alt, VT* vals, VT* vals_alt, long long n, KeyOrder order, void* workspace)
{
  if (n <= 1) return cudaSuccess;
  constexpr int NPASS = sizeof(U);
  static_assert(NPASS % 2 == 0, "even pass count keeps the result in the caller's buffer");
  const size_t header = (size_t)NPASS * RADIX * 8 * 2 + 256;
  unsigned long long* ghist = (unsigned long long*)workspace;
  unsigned long long* bin_base = ghist + NPASS * RADIX;
  unsigned* tickets = (unsigned*)(bin_base + NPASS * RADIX);
  void* lookback = (char*)workspace + header;
  const bool small = n < (1ll << 30);
  const size_t lb_bytes = (size_t)num_tiles<U, VT, S>(n) * RADIX * (small ? 4 : 8);

  cudaError_t e = cudaMemsetAsync(workspace, 0, header + lb_bytes, stream);
  if (e != cudaSuccess) return e;
  {
    constexpr int HB = 512, HU = 8;
    long long tiles = (n + HB * HU - 1) / (HB * HU);
    long long grid = (long long)device_info().sm_count * 4;
    if (grid > tiles) grid = tiles;
    {
      LaunchScope scope("histogram_kernel", stream);
      histogram_kernel<U, HB, HU><<<(unsigned)grid, HB, 0, stream>>>(keys, n, ghist, order);
    }
    {
      LaunchScope scope("scan_bins_kernel", stream);
      scan_bins_kernel<<<NPASS, RADIX, 0, stream>>>(ghist, bin_base);
    }
  }
  for (int p = 0; p < NPASS; ++p) {
    const U* kin = (p & 1) ? keys_alt : keys;
    U* kout = (p & 1) ? keys : keys_alt;
    const VT* vin = (p & 1) ? vals_alt : vals;
    VT* vout = (p & 1) ? vals : vals_alt;
    if (p > 0) {
      e = cudaMemsetAsync(lookback, 0, lb_bytes, stream);
      if (e != cudaSuccess) return e;
    }
    if (small)
      e = launch_pass<U, VT, uint32_t, S>(stream, kin, kout, vin, vout, n, bin_base + p * RADIX, lookback, tickets + p, p * RADIX_BITS, order);
    else
      e = launch_pass<U, VT, unsigned long long, S>(stream, kin, kout, vin, vout, n, bin_base + p * RADIX, lookback, tickets + p, p * RADIX_BITS, order);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace rsort
}  // namespace b200
