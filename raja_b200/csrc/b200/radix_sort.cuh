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
// Shared-memory layout: every digit counter is replicated COLS times and a lane only ever touches
// column (lane % COLS), so with COLS = 32 the NPASS atomics of one key row hit 32 distinct banks
// (the plain [pass][digit] layout costs ~3.5 conflict wavefronts per atomic on random keys, which
// made this kernel LSU-bound at 2.5 TB/s).  One persistent CTA per SM, 128 KB of counters.
template <class U> struct HistShape {
  static constexpr int NPASS = sizeof(U);
  // 4-byte keys: 32 columns, lane = column, every lane counts all four digit positions of its own keys.
  // 8-byte keys: eight digit positions x 256 bins x 32 columns do not fit; with 16 columns two lanes shared a column and half of
  // the atomics met a two-way bank conflict (0.55 ms for 2^28 keys, LSU-bound).  Now the two half-warps split the PASSES of the
  // same keys: lanes 0-15 count the even digit positions, lanes 16-31 the odd ones, and the parity of the pass selects the half
  // of the banks -- every atomic of a warp hits 32 distinct banks again.
  static constexpr bool SPLIT = sizeof(U) == 8;
  static constexpr int COLS = SPLIT ? 16 : 32;
  static constexpr int BLOCK = 1024;
  static constexpr int VEC = 16 / sizeof(U);  // keys per 128-bit load
  static constexpr int SMEM = NPASS * RADIX * COLS * 4;
  // counter of (pass p, digit d), column c
  B200_HD static int slot(int p, int d, int c) { return SPLIT ? ((((p >> 1) * RADIX + d) * 2 + (p & 1)) * COLS + c) : ((p * RADIX + d) * COLS + c); }
};

template <class U>
__global__ void __launch_bounds__(HistShape<U>::BLOCK, 1)
histogram_kernel(const U* __restrict__ keys, const long long n, unsigned long long* __restrict__ ghist, const KeyOrder order,
                 unsigned* __restrict__ negzero_flag = nullptr)
{
  using H = HistShape<U>;
  constexpr int NPASS = H::NPASS, COLS = H::COLS, BLOCK = H::BLOCK, VEC = H::VEC;
  constexpr bool SPLIT = H::SPLIT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* const sh = reinterpret_cast<uint32_t*>(smem_raw);  // H::slot(pass, digit, column)
  for (int i = threadIdx.x; i < NPASS * RADIX * COLS; i += BLOCK) sh[i] = 0;
  __syncthreads();

  const int hw = SPLIT ? ((threadIdx.x >> 4) & 1) : 0;  // SPLIT: which passes this lane counts (2 q + hw)
  uint32_t* const col = sh + (threadIdx.x & (COLS - 1)) + hw * COLS;
  // The order-preserving map is an XOR with one constant for integer keys (floating keys first flip the magnitude bits
  // of negatives and rank -0.0 as +0.0, as sweep_digit does), and a digit is ONE byte-permute: per key 1 + NPASS x
  // (PRMT, LEA, ATOMS) instructions instead of the ~20 a run-time KeyOrder switch and shift/mask pairs cost -- the kernel
  // was issue-bound (79 % issue slots, 63 % of the DRAM peak).
  constexpr int BITS = 8 * sizeof(U);
  constexpr U SIGN = U(1) << (BITS - 1);
  U xmask = (order.kind == KEY_SIGNED || order.kind == KEY_FLOAT) ? SIGN : U(0);
  if (order.descending) xmask = ~xmask;
  bool saw_negzero = false;
  auto run = [&](auto float_c) {
    constexpr bool FLOAT = decltype(float_c)::value;
    auto count = [&](U k) {
      if constexpr (FLOAT) {
        using SU = typename std::make_signed<U>::type;
        if (k == SIGN) {  // -0.0 ranks as +0.0; the sweeps only pay for this special case if it occurs at all
          k = 0;
          saw_negzero = true;
        }
        k ^= (U)((SU)k >> (BITS - 1)) & ~SIGN;
      }
      const U ok = k ^ xmask;
      if constexpr (SPLIT) {
        const uint32_t lo = (uint32_t)ok, hi = (uint32_t)(ok >> (BITS > 32 ? 32 : 0));
        const uint32_t sel0 = 0x4440u | (uint32_t)hw, sel1 = 0x4440u | (uint32_t)(2 + hw);
        // passes hw, 2 + hw (low word) and 4 + hw, 6 + hw (high word): q = 0 .. 3
        atomicAdd(col + (0 * RADIX + (int)__byte_perm(lo, 0u, sel0)) * (2 * COLS), 1u);
        atomicAdd(col + (1 * RADIX + (int)__byte_perm(lo, 0u, sel1)) * (2 * COLS), 1u);
        atomicAdd(col + (2 * RADIX + (int)__byte_perm(hi, 0u, sel0)) * (2 * COLS), 1u);
        atomicAdd(col + (3 * RADIX + (int)__byte_perm(hi, 0u, sel1)) * (2 * COLS), 1u);
      } else {
#pragma unroll
        for (int p = 0; p < NPASS; ++p) {
          const uint32_t half = (uint32_t)(p < 4 ? ok : (ok >> (BITS > 32 ? 32 : 0)));
          const uint32_t d = __byte_perm(half, 0u, 0x4440u | (uint32_t)(p & 3));
          atomicAdd(col + (p * RADIX + (int)d) * COLS, 1u);
        }
      }
    };
    // 128-bit loads over the aligned body; a CTA sees < 2^32 keys, so the 32-bit counters cannot overflow.
    // SPLIT: the two half-warps load the SAME vectors (lane l and l + 16 one address: one L1 request serves both), so a block
    // covers BLOCK / 2 vectors per load instruction.
    constexpr int LANES = SPLIT ? BLOCK / 2 : BLOCK;  // distinct vectors per block and load instruction
    const int vt = SPLIT ? (int)((threadIdx.x >> 5) * 16 + (threadIdx.x & 15)) : (int)threadIdx.x;
    const uintptr_t addr = (uintptr_t)keys;
    long long head = (long long)(((16 - (addr & 15)) & 15) / sizeof(U));
    if (head > n) head = n;
    const long long nvec = (n - head) / VEC;
    using V = Vec<U, VEC>;
    const V* const body = reinterpret_cast<const V*>(keys + head);
    constexpr int UNROLL = SPLIT ? 8 : 4;
    long long v = (long long)blockIdx.x * LANES * UNROLL + vt;
    const long long stride = (long long)gridDim.x * LANES * UNROLL;
    for (; v + (UNROLL - 1) * LANES < nvec; v += stride) {
      V x[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) x[u] = ld_stream(body + v + u * LANES);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int e = 0; e < VEC; ++e) count(x[u].e[e]);
    }
    for (int u = 0; u < UNROLL; ++u)
      if (v + u * LANES < nvec) {
        const V x = ld_stream(body + v + u * LANES);
#pragma unroll
        for (int e = 0; e < VEC; ++e) count(x.e[e]);
      }
    if (blockIdx.x == 0) {  // unaligned head and the tail shorter than one vector (SPLIT: both half-warps see every such key)
      for (long long i = vt; i < head; i += LANES) count(keys[i]);
      for (long long i = head + nvec * VEC + vt; i < n; i += LANES) count(keys[i]);
    }
  };
  if (order.kind == KEY_FLOAT) run(std::true_type{});
  else run(std::false_type{});
  if (saw_negzero && negzero_flag != nullptr) atomicOr(negzero_flag, 1u);
  __syncthreads();
  for (int b = threadIdx.x; b < NPASS * RADIX; b += BLOCK) {
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) c += sh[H::slot(b / RADIX, b % RADIX, (k + threadIdx.x) & (COLS - 1))];  // rotated: conflict-free
    if (c) atomicAdd(ghist + b, (unsigned long long)c);
  }
}

// ---------------------------------------------------------------------------
// 2. exclusive scan of each pass's 256 bins (one block per pass, 256 threads)
// ---------------------------------------------------------------------------
// also: degenerate[p] = 1 when ONE digit value holds all n keys of pass p -- the pass is then the identity permutation and
// the sweep copies its tiles straight through (low-entropy keys: constant high bytes, small ranges)
__global__ void __launch_bounds__(RADIX) scan_bins_kernel(const unsigned long long* __restrict__ ghist, unsigned long long* __restrict__ bin_base,
                                                          const unsigned long long n = 0, unsigned* __restrict__ degenerate = nullptr)
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
  if (degenerate != nullptr && n > 0 && c == n) degenerate[p] = 1u;
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

// The same match, written so that ptxas emits R2P (digit bits -> predicates) + per bit one VOTE,
// one SEL (bit ? 0 : ~0) and one LOP3 (peers & (ballot ^ sel)): 28 instructions per row instead of
// the ~48 the C++ form above compiles to.
template <int B> B200_D void match_bit(unsigned d, unsigned& peers)
{
  asm("{\n\t.reg .pred p;\n\t.reg .b32 m, t, e;\n\t"
      "and.b32 t, %1, %2;\n\t"
      "setp.ne.u32 p, t, 0;\n\t"
      "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
      "selp.b32 e, 0, 0xffffffff, p;\n\t"
      "lop3.b32 %0, %0, m, e, 0x60;\n\t"
      "}"
      : "+r"(peers)
      : "r"(d), "n"(1u << B));
}
B200_D unsigned match_digit_r2p(unsigned d)
{
  unsigned peers = FULL_MASK;
  match_bit<0>(d, peers);
  match_bit<1>(d, peers);
  match_bit<2>(d, peers);
  match_bit<3>(d, peers);
  match_bit<4>(d, peers);
  match_bit<5>(d, peers);
  match_bit<6>(d, peers);
  match_bit<7>(d, peers);
  return peers;
}

// whole 16-byte units inside [first, last) -> L2
B200_D void prefetch_l2(const void* first, const void* last)
{
  const uintptr_t lo = ((uintptr_t)first + 15u) & ~(uintptr_t)15u;
  const uintptr_t hi = (uintptr_t)last & ~(uintptr_t)15u;
  if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned)(hi - lo)) : "memory");
}

// Compile-time shape of the sweep kernel: BLOCK compute threads (+ 64 look-back threads),
// ITEMS keys per compute thread, MIN_CTAS resident CTAs per SM the register budget is cut for,
// MATCH 0 = ballot loop / 1 = match.any / 2 = shared-memory atomicOr masks.
template <int BLOCK_, int ITEMS_, int MIN_CTAS_, int MATCH_ = 0, bool PACK_ = false, int PREFETCH_ = 0, int LB_WARPS_ = 2, int LB_BATCH_ = 8, int LB_SLEEP_ = 400,
          int LB_GROUP_ = 8>
struct SweepShape {
  static constexpr int LB_GROUP = LB_GROUP_;  // late-count kernel: tiles per look-back group
  // late-count kernel: group words in flight per look-back lane and round trip, and (x4) the back-off
  // in ns when a round made no progress
  static constexpr int LB_BATCH = LB_BATCH_, LB_SLEEP = LB_SLEEP_;
  static constexpr int BLOCK = BLOCK_, ITEMS = ITEMS_, MIN_CTAS = MIN_CTAS_, MATCH = MATCH_;
  static constexpr bool PACK = PACK_;
  // late-count kernel: how many tiles ahead the keys are pulled toward L2 (0 = off); about the number
  // of CTAs resident on the device (148 SMs x MIN_CTAS), so the prefetch lands one tile-lifetime early
  static constexpr int PREFETCH_TILES = PREFETCH_;
  static constexpr int LB_WARPS = LB_WARPS_;  // look-back warps: 2 (4 digits per lane) or 1 (8 digits per lane)
  static constexpr int THREADS = BLOCK + LB_WARPS * WARP;
  static_assert(LB_WARPS == 1 || LB_WARPS == 2, "");
  static_assert(LB_WARPS == 2 || MATCH_ >= 3, "the early-count kernel is written for two look-back warps");
  static constexpr int TILE = BLOCK * ITEMS;
  static_assert(BLOCK % WARP == 0 && BLOCK_ + LB_WARPS_ * WARP >= RADIX, "one digit-owner thread per digit (look-back warps take part in the late-count kernel)");
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
// Digit functors of the late-count kernel (passed by value as a kernel argument).
//   RadixDigit : the 8-bit digit of one LSD pass (shift / xor precomputed on the host)
//   SplitDigit : multi-way split by up to 7 splitters (multi-GPU sample sort, raja_b200/multi.py):
//                digit = number of splitters that sort at or before the key, in the same total
//                order as the sort (KeyOrder); the splitters live in the constant bank, so the
//                compares cost no registers.
// ---------------------------------------------------------------------------
template <class U, bool FLOAT> struct RadixDigit {
  static constexpr bool PEER_OUT = false;
  static constexpr int BUCKETS = RADIX;
  static constexpr bool ENCODE_ON_LOAD = false, DECODE_ON_STORE = false;
  int shift;
  unsigned dxor;
  const unsigned* degenerate = nullptr;  // device flag of this pass: all keys share one digit value (scan_bins_kernel)
  B200_D bool identity_pass() const { return degenerate != nullptr && *degenerate != 0u; }
  B200_D unsigned operator()(U k) const { return sweep_digit<FLOAT>(k, shift, dxor); }
};

// Floating-point keys travel BETWEEN the passes of one sort in the order-preserving encoding (sign flipped for positives,
// all bits flipped for negatives): the first pass encodes on load, the last decodes on store, and in between a digit is a
// shift and a mask -- the sign-dependent transform used to be evaluated three times per key and pass, and the 64-bit sweep is
// ALU-bound (ALU pipe 75 %, profiles/r02_onesweep_f64_ncu.txt).  -0.0 must rank as +0.0 (they compare equal, a stable sort
// keeps their input order) while the stored key keeps its sign: its code (SIGN - 1) is read as +0.0's (SIGN) when a digit is
// taken -- but only if the histogram kernel, which sees every key first, found a -0.0 at all (one device flag; otherwise the
// compare is skipped).
enum FloatPass : int { FLOAT_FIRST = 1, FLOAT_MIDDLE = 2, FLOAT_LAST = 3 };
// NEGZERO = false is the functor the host passes; the kernel reads the flag once and runs the body instantiated for
// NEGZERO = true (its Twin) only when the keys do contain a -0.0 -- the compare + selects of that case cost 13 % of an
// ALU-bound 64-bit sweep when they were evaluated behind a run-time flag (3 digit evaluations per key and pass).
template <class U, int PASS, bool NEGZERO = false> struct FloatDigit {
  static constexpr bool PEER_OUT = false;
  static constexpr int BUCKETS = RADIX;
  static constexpr bool ENCODE_ON_LOAD = (PASS == FLOAT_FIRST), DECODE_ON_STORE = (PASS == FLOAT_LAST);
  static constexpr int BITS = 8 * sizeof(U);
  static constexpr U SIGN = U(1) << (BITS - 1);
  using SU = typename std::make_signed<U>::type;
  using Twin = FloatDigit<U, PASS, true>;
  int shift;
  unsigned dxor;                  // 0xff for descending order
  const unsigned* negzero_flag;   // device: non-zero if the keys contain a -0.0
  const unsigned* degenerate;     // device flag of this pass: all keys share one digit value
  B200_D bool use_twin() const { return !NEGZERO && *negzero_flag != 0u; }
  B200_D Twin twin() const { return Twin{shift, dxor, negzero_flag, degenerate}; }
  B200_D bool identity_pass() const { return degenerate != nullptr && *degenerate != 0u; }
  B200_D static U encode(U k) { return k ^ ((U)((SU)k >> (BITS - 1)) | SIGN); }
  B200_D static U decode(U o) { return o ^ ((U)(~(SU)o >> (BITS - 1)) | SIGN); }
  B200_D unsigned operator()(U o) const
  {
    if constexpr (NEGZERO) {
      if (o == U(SIGN - 1)) o = SIGN;
    }
    return ((unsigned)(o >> shift) & (RADIX - 1)) ^ dxor;
  }
};
template <class D, class = void> struct HasTwin : std::false_type {};
template <class D> struct HasTwin<D, std::void_t<typename D::Twin>> : std::true_type {};
constexpr int MAX_SPLITTERS = 7;
template <class U> struct SplitDigit {
  static constexpr bool PEER_OUT = false;
  static constexpr int BUCKETS = 8;  // MAX_SPLITTERS + 1: few, long runs per tile -> line-aligned write-out
  static constexpr bool ENCODE_ON_LOAD = false, DECODE_ON_STORE = false;
  U s[MAX_SPLITTERS];  // already mapped with to_ordered(), non-decreasing
  int nsplit;
  KeyOrder order;
  B200_D bool identity_pass() const { return false; }
  B200_D U splitter(int i) const { return s[i]; }
  B200_D unsigned operator()(U k) const
  {
    const U ok = to_ordered(k, order);
    unsigned d = 0;
#pragma unroll
    for (int i = 0; i < MAX_SPLITTERS; ++i)
      if (i < nsplit) d += (ok >= s[i]) ? 1u : 0u;
    return d;
  }
};

// SplitDigit whose buckets land in DIFFERENT buffers: bucket d is written to kdst[d] / vdst[d]
// (peer-mapped memory of GPU d over NVLink), at the positions the pass's bin bases give inside
// that buffer.  This is the fused "partition + all-to-all" of the multi-GPU sample sort: the
// write-out of the sweep IS the exchange.
template <class U, class VT, class Base = SplitDigit<U>> struct PeerSplitDigit : Base {
  static constexpr bool PEER_OUT = true;
  U* kdst[MAX_SPLITTERS + 1];
  VT* vdst[MAX_SPLITTERS + 1];
};

// ---------------------------------------------------------------------------
// 3b. one onesweep pass, "late counts" (S::MATCH == 3) -- the production kernel
//
// The early-counts kernel above touches shared memory ~9 times per key, 2 of them atomics, and ncu
// shows it LSU-bound (l1tex data-pipe wavefronts 86 % of peak, profiles/r01_onesweep_v3_*).  Here:
//   * every warp ranks its keys FIRST against warp-private running counters with no atomics at
//     all: peers by an 8-ballot match (R2P + VOTE/SEL/LOP3 per bit), one broadcast LDS of the
//     (warp, digit) counter, one STS by the lowest peer;
//   * the tile's digit counts are then read off the counters' final values by 256 digit-owner
//     threads (conflict-free sweep), published for the successors, and the counters are turned
//     into first tile ranks; 32-bit counter stride, so every bank is used;
//   * per key that leaves LDS + STS (rank), LDS (base) and the scatter: 22 shared-memory
//     wavefronts per 32 keys instead of 37;
//   * the keys of the tile PREFETCH_TILES ahead (about one tile-lifetime) and this tile's values
//     are pulled toward L2 by bulk prefetches issued from the otherwise idle look-back warps;
//   * the look-back warps run a TWO-LEVEL decoupled look-back next to the scatter: 16-bit tile counts
//     inside groups of 8 tiles, 32-bit group words (count, later inclusive prefix) across groups.
// PACK (u32 keys + 4-byte values): key and value cross shared memory as ONE 8-byte word.
// (A persistent, warp-specialised variant with the next tile's keys loaded during the write-out
//  was tried; ptxas 12.9 spills the key/rank registers in every loop form tried, which costs more
//  LSU wavefronts than the variant saves -- see DESIGN.md.)
// ---------------------------------------------------------------------------
// look-back words of one pass: the early-count kernel keeps one C word per (tile, digit); the late-count
// kernel one C word per (group of LB_GROUP tiles, digit) + one u16 per (tile, digit)
template <class S> B200_HD long long lookback_groups(long long ntiles)
{
  return S::MATCH >= 3 ? (ntiles + S::LB_GROUP - 1) / S::LB_GROUP : ntiles;
}
template <class S> inline size_t lookback_bytes(long long ntiles, size_t cbytes)
{
  return (size_t)lookback_groups<S>(ntiles) * RADIX * cbytes + (size_t)ntiles * RADIX * 2;
}

template <class U, class VT, class C, class S> struct LateSmem {
  static constexpr int NW = S::BLOCK / WARP;
  static constexpr bool HAS_VALUES = !std::is_same<VT, NoValue>::value;
  static constexpr bool PACK = S::PACK && HAS_VALUES && sizeof(U) == 4 && sizeof(VT) == 4;
  using G = typename std::conditional<sizeof(C) == 4, uint32_t, unsigned long long>::type;
  uint16_t cnt[NW][RADIX];  // running rank of (warp, digit) -> total -> first tile rank (< TILE < 65535; 16 bits: 4 KB more for keys)
  uint16_t count[RADIX];    // digit counts of this tile (TILE < 65536)
  uint32_t texcl[RADIX];    // first rank of each digit inside this tile
  G gbase[RADIX];           // global position of tile-rank r of digit d = gbase[d] + r (mod 2^bits)
  uint32_t wtot[RADIX / WARP];
  long long tile;
  int identity;  // this pass is an identity permutation (all keys share one digit value)
#ifdef RB200_LB_TRACE
  unsigned long long t_compute, t_lb;  // arrival times at the barrier before the write-out
#endif
  alignas(16) U keys[S::TILE];
  typename std::conditional<HAS_VALUES, VT, char>::type vals[HAS_VALUES ? S::TILE : 1];  // PACK: keys..vals is one uint2[TILE]
};

template <class U, class VT, class DigitFn, class C, class S>
B200_D void onesweep_late_body(const U* __restrict__ keys_in, U* __restrict__ keys_out,
                               const VT* __restrict__ vals_in, VT* __restrict__ vals_out,
                               const long long n,
                               const unsigned long long* __restrict__ bin_base,  // [256] of this pass
                               C* lookback,                                      // [ngroups][256] group words, then [ntiles][256] u16 tile counts; zero on entry
                               unsigned* ticket,                                 // zero on entry
                               const DigitFn& digit)
{
  using SM = LateSmem<U, VT, C, S>;
  using F = LookbackFlags<C>;
  using G = typename SM::G;
  constexpr int BLOCK = S::BLOCK, ITEMS = S::ITEMS, TILE = S::TILE, NW = SM::NW, LBT = S::LB_WARPS * WARP;
  constexpr bool HAS_VALUES = SM::HAS_VALUES, PACK = SM::PACK;
  static_assert(ITEMS % 2 == 0, "ranks are kept two per register");
  static_assert(S::MATCH == 3, "late-count kernel: ballot match");
  static_assert(TILE < 65535, "tile counts are published as 16-bit words (count + 1)");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& s = *reinterpret_cast<SM*>(smem_raw);

  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const bool compute = tid < BLOCK;
  if (tid == 0) {
    // the flag load travels together with the ticket atomic: no extra round trip in front of the key loads
    const bool ident = digit.identity_pass();
    s.tile = (long long)atomicAdd(ticket, 1u);
    s.identity = ident ? 1 : 0;
  }
#ifdef RB200_LB_TRACE
  if (tid == 0) {
    s.t_compute = 0;
    s.t_lb = 0;
  }
#endif
  for (int i = tid; i < NW * RADIX / 2; i += S::THREADS) reinterpret_cast<uint32_t*>(&s.cnt[0][0])[i] = 0;
  __syncthreads();
  const long long tile = s.tile;
  const long long tile_base = tile * TILE;
  const int tile_count = (int)((n - tile_base) < (long long)TILE ? (n - tile_base) : (long long)TILE);
  const bool full = (tile_count == TILE);
  if (s.identity) {
    // every key of this pass has the same digit: the stable permutation is the identity.  Copy the tile straight through
    // (no ranking, no look-back: every tile's destination is its own position), encoding / decoding floating-point keys
    // exactly as the general path would.
    for (int i = tid; i < tile_count; i += S::THREADS) {
      U k = ld_stream(keys_in + tile_base + i);
      if constexpr (DigitFn::ENCODE_ON_LOAD) k = DigitFn::encode(k);
      if constexpr (DigitFn::DECODE_ON_STORE) k = DigitFn::decode(k);
      keys_out[tile_base + i] = k;
      if constexpr (HAS_VALUES) vals_out[tile_base + i] = ld_stream(vals_in + tile_base + i);
    }
    return;
  }
  const int wofs = w * (WARP * ITEMS);  // this warp's first item within the tile
  // Look-back words.  A tile publishes its digit counts as 16-bit words (count + 1; 0 = not yet) in
  // local16 right after the ranking, and its inclusive prefixes (value | GLOBAL) in `lookback` once
  // its own look-back is done.  16-bit counts let a look-back lane keep 16 predecessors in flight in
  // the registers 8 took before: the rate at which finished prefixes propagate from tile to tile
  // (predecessors consumed per L2 round trip) is what bounds the pass, not HBM or the LSU.
  // Layout of one pass's look-back region: [groups][256] C group words, then [tiles][256] u16 counts.
  uint16_t* const local16 = reinterpret_cast<uint16_t*>(lookback + lookback_groups<S>((n + TILE - 1) / TILE) * RADIX);

  U key[ITEMS];            // warp-striped: item (j, l) is tile item wofs + j*32 + l
  uint32_t rp[ITEMS / 2];  // two 16-bit ranks per register: warp-local first, tile ranks later

  // ---- phase A + R: keys -> registers, warp-local stable ranks | look-back warps: prefetches -----
  if (compute) {
    if (full) {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) key[j] = ld_stream(keys_in + tile_base + wofs + j * WARP + l);
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const int i = wofs + j * WARP + l;
        key[j] = (i < tile_count) ? ld_stream(keys_in + tile_base + i) : U(0);
      }
    }
    if constexpr (DigitFn::ENCODE_ON_LOAD) {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) key[j] = DigitFn::encode(key[j]);
    }
  } else if (tid == BLOCK) {
    if constexpr (HAS_VALUES) prefetch_l2(vals_in + tile_base, vals_in + tile_base + tile_count);
  } else if (tid == BLOCK + 1) {
    if constexpr (S::PREFETCH_TILES > 0) {  // keys of the tile some CTA claims about one tile-lifetime from now
      const long long fb = tile_base + (long long)S::PREFETCH_TILES * TILE;
      if (fb < n) prefetch_l2(keys_in + fb, keys_in + ((n - fb) < (long long)TILE ? n : fb + TILE));
    }
  }
  if (compute) {
    uint16_t* const wc = s.cnt[w];
    const unsigned lt = lanemask_lt();
    auto rank_rows = [&](auto full_c) {
      constexpr bool FULL = decltype(full_c)::value;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const bool valid = FULL || (wofs + j * WARP + l < tile_count);
        const unsigned d = digit(key[j]);
        unsigned peers = match_digit_r2p(d);
        if (!FULL) peers &= __ballot_sync(FULL_MASK, valid);
        const uint32_t c = wc[d];  // all peers read the same word: broadcast
        const unsigned below = peers & lt;
        const uint32_t r = c + __popc(below);
        __syncwarp();
        if (valid && below == 0) wc[d] = (uint16_t)(c + __popc(peers));
        __syncwarp();
        if (j & 1) rp[j >> 1] |= r << 16;
        else rp[j >> 1] = r;
      }
    };
    if (full) rank_rows(std::true_type{});
    else rank_rows(std::false_type{});
  }
  __syncthreads();

  // ---- phase S: digit owners: counts of the tile, published; counters -> first tile ranks ---------
  if (tid < RADIX) {
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) cnt += s.cnt[i][tid];
    local16[tile * RADIX + tid] = (uint16_t)(cnt + 1);
    s.count[tid] = (uint16_t)cnt;
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
    uint32_t run = pre + incl - cnt;
    s.texcl[tid] = run;
#pragma unroll
    for (int i = 0; i < NW; ++i) {  // re-read rather than hold NW counts across the scan (register budget)
      const uint32_t c = s.cnt[i][tid];
      s.cnt[i][tid] = (uint16_t)run;
      run += c;
    }
  }
  __syncthreads();

  // ---- phase C: scatter into shared memory in tile-rank order | look-back -----------------------
  if (compute) {
    const uint16_t* const wc = s.cnt[w];
    uint2* const pairs = reinterpret_cast<uint2*>(s.keys);
    (void)pairs;
    auto scatter = [&](auto full_c) {
      constexpr bool FULL = decltype(full_c)::value;
      VT val[PACK ? ITEMS : 1];
      (void)val;
      if constexpr (PACK) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          const int i = wofs + j * WARP + l;
          val[j] = (FULL || i < tile_count) ? ld_stream(vals_in + tile_base + i) : VT(0);
        }
      }
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const bool valid = FULL || (wofs + j * WARP + l < tile_count);
        const unsigned d = digit(key[j]);
        const uint32_t r = wc[d] + ((j & 1) ? (rp[j >> 1] >> 16) : (rp[j >> 1] & 0xffffu));
        if constexpr (PACK) {
          if (valid) pairs[r] = make_uint2((uint32_t)key[j], bit_cast_<uint32_t>(val[j]));
        } else {
          if (valid) s.keys[r] = key[j];
          if (j & 1) rp[j >> 1] = (rp[j >> 1] & 0xffffu) | (r << 16);
          else rp[j >> 1] = (rp[j >> 1] & 0xffff0000u) | r;
        }
      }
    };
    if (full) scatter(std::true_type{});
    else scatter(std::false_type{});
    if constexpr (HAS_VALUES && !PACK) {
      if (full) {
        VT v[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) v[j] = ld_stream(vals_in + tile_base + wofs + j * WARP + l);
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) s.vals[(j & 1) ? (rp[j >> 1] >> 16) : (rp[j >> 1] & 0xffffu)] = v[j];
      } else {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          const int i = wofs + j * WARP + l;
          if (i < tile_count) s.vals[(j & 1) ? (rp[j >> 1] >> 16) : (rp[j >> 1] & 0xffffu)] = ld_stream(vals_in + tile_base + i);
        }
      }
    }
  } else {
    // look-back warps: lane owns D consecutive digits.  Two-level look-back: tiles form groups of 8;
    // the LAST tile of a group publishes, per digit, first the group's count (LOCAL flag: the sum of
    // its seven in-group predecessors' counts, which it reads anyway, plus its own) and later the
    // inclusive prefix through the group's end (GLOBAL flag) into lookback[group][digit].  A tile
    // then needs (a) the 16-bit counts of its <= 7 in-group predecessors and (b) group words walking
    // back until a GLOBAL one -- both fetched in the SAME round trip -- instead of walking ~40 tiles
    // back 8 at a time: the compute warps used to wait ~6.8 K cycles per tile for five to six look-back
    // round trips (measured with -DRB200_LB_TRACE: the look-back arrived last at the barrier below in
    // 100 % of the tiles).
    constexpr int NQ = RADIX / (LBT * 4);  // groups of 4 digits per lane: 1 (two look-back warps) or 2 (one)
    constexpr int D = 4 * NQ;
    constexpr int GT = S::LB_GROUP;        // tiles per group
    constexpr int NG = S::LB_BATCH;        // group words in flight per lane
    const int q = (tid - BLOCK) * D;
    const long long grp = tile / GT;
    const int k = (int)(tile - grp * GT);  // in-group predecessors: tile-1 .. tile-k
    C prev[D] = {};                        // counts of everything before this tile
    C ingroup[D] = {};                     // counts of the in-group predecessors
    {
      unsigned pending = (1u << k) - 1u;   // bit b: in-group predecessor tile-1-b not yet added
      bool gdone[D];
#pragma unroll
      for (int c = 0; c < D; ++c) gdone[c] = (grp == 0);
      bool groups_done = (grp == 0);
      bool owner_published = false;
      long long gh = grp - 1;              // next group word to consume
#ifdef RB200_LB_TRACE
      unsigned lb_rounds = 0;
#endif
      unsigned spins = 0;
      while (pending != 0u || !groups_done) {
#ifdef RB200_LB_TRACE
        ++lb_rounds;
#endif
        if (++spins > (1u << 24)) __trap();  // a predecessor never published: fail instead of hanging the device
        // ---- issue every load of this round
        uint2 v[GT - 1][NQ];
        Quad<C> gw[NG][NQ];
#pragma unroll
        for (int b = 0; b < GT - 1; ++b)
#pragma unroll
          for (int kq = 0; kq < NQ; ++kq) {
            v[b][kq] = make_uint2(0u, 0u);
            if (pending & (1u << b)) {
              const unsigned long long w64 = ld_relaxed_u64(reinterpret_cast<const uint64_t*>(local16 + (tile - 1 - b) * RADIX + q + 4 * kq));
              v[b][kq] = make_uint2((uint32_t)w64, (uint32_t)(w64 >> 32));
            }
          }
        if (!groups_done) {
#pragma unroll
          for (int b = 0; b < NG; ++b)
#pragma unroll
            for (int kq = 0; kq < NQ; ++kq) {
              gw[b][kq] = Quad<C>{{F::GLOBAL, F::GLOBAL, F::GLOBAL, F::GLOBAL}};  // before the first group: zero, final
              if (gh - b >= 0) gw[b][kq] = ld_quad(lookback + (gh - b) * RADIX + q + 4 * kq);
            }
        }
        // ---- in-group counts (any order: a sum)
        bool progress = false;
#pragma unroll
        for (int b = 0; b < GT - 1; ++b) {
          if (pending & (1u << b)) {
            uint32_t e[D];
#pragma unroll
            for (int kq = 0; kq < NQ; ++kq) {
              e[4 * kq + 0] = v[b][kq].x & 0xffffu;
              e[4 * kq + 1] = v[b][kq].x >> 16;
              e[4 * kq + 2] = v[b][kq].y & 0xffffu;
              e[4 * kq + 3] = v[b][kq].y >> 16;
            }
            bool ok = true;
#pragma unroll
            for (int c = 0; c < D; ++c) ok = ok && (e[c] != 0);
            if (__all_sync(FULL_MASK, ok)) {
#pragma unroll
              for (int c = 0; c < D; ++c) ingroup[c] += (C)(e[c] - 1);
              pending &= ~(1u << b);
              progress = true;
            }
          }
        }
        // the group's last tile: the group's count, as soon as its seven predecessors are in
        if (k == GT - 1 && pending == 0u && !owner_published) {
#pragma unroll
          for (int kq = 0; kq < NQ; ++kq) {
            Quad<C> g;
#pragma unroll
            for (int c = 0; c < 4; ++c) g.v[c] = (C)(ingroup[4 * kq + c] + s.count[q + 4 * kq + c]) | F::LOCAL;
            st_quad(lookback + grp * RADIX + q + 4 * kq, g);
          }
          owner_published = true;
        }
        // ---- group words, nearest first, until every digit has met a GLOBAL one
        if (!groups_done) {
          int consumed = 0;
#pragma unroll
          for (int b = 0; b < NG; ++b) {
            if (groups_done) break;
            bool ok = true;
#pragma unroll
            for (int c = 0; c < D; ++c) ok = ok && (gdone[c] || (gw[b][c >> 2].v[c & 3] & ~F::MASK) != 0);
            if (!__all_sync(FULL_MASK, ok)) break;  // this group has not published yet: re-poll from here
#pragma unroll
            for (int c = 0; c < D; ++c) {
              if (!gdone[c]) {
                prev[c] += gw[b][c >> 2].v[c & 3] & F::MASK;
                if (gw[b][c >> 2].v[c & 3] & F::GLOBAL) gdone[c] = true;
              }
            }
            ++consumed;
            bool mine = true;
#pragma unroll
            for (int c = 0; c < D; ++c) mine = mine && gdone[c];
            groups_done = __all_sync(FULL_MASK, mine);
          }
          gh -= consumed;
          progress = progress || consumed > 0;
        }
        if (!progress) __nanosleep(S::LB_SLEEP / 4);  // leave the issue slots to the scatter
      }
#pragma unroll
      for (int c = 0; c < D; ++c) prev[c] += ingroup[c];
#ifdef RB200_LB_TRACE
      if (tid == BLOCK) {  // this pass's ticket word + 8 / 16 / 24: look-back rounds, depth (tiles + groups read), tiles
        atomicAdd(ticket + 8, lb_rounds);
        atomicAdd(ticket + 16, (unsigned)(k + (grp - 1 - gh)));
        atomicAdd(ticket + 24, 1u);
      }
#endif
    }
    if (k == GT - 1) {  // the inclusive prefix through the group's end, for the groups behind
#pragma unroll
      for (int kq = 0; kq < NQ; ++kq) {
        Quad<C> g;
#pragma unroll
        for (int c = 0; c < 4; ++c) g.v[c] = (C)(prev[4 * kq + c] + s.count[q + 4 * kq + c]) | F::GLOBAL;
        st_quad(lookback + grp * RADIX + q + 4 * kq, g);
      }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) s.gbase[q + c] = (G)(bin_base[q + c] + (unsigned long long)prev[c]) - (G)s.texcl[q + c];
  }
#ifdef RB200_LB_TRACE
  // who arrives last at this barrier: the scatter or the look-back?  ticket + 32: tiles where the look-back was
  // last, + 40: by how much in total (units of 64 cycles), + 48: the reverse
  if (l == 0) {
    if (compute) atomicMax(&s.t_compute, (unsigned long long)clock64());
    else atomicMax(&s.t_lb, (unsigned long long)clock64());
  }
#endif
  __syncthreads();
#ifdef RB200_LB_TRACE
  if (tid == 0) {
    if (s.t_lb > s.t_compute) {
      atomicAdd(ticket + 32, 1u);
      atomicAdd(ticket + 40, (unsigned)((s.t_lb - s.t_compute) >> 6));
    } else {
      atomicAdd(ticket + 48, (unsigned)((s.t_compute - s.t_lb) >> 6));
    }
  }
#endif

  // ---- phase D: run-coalesced write-out ------------------------------------------------------------
  if constexpr (DigitFn::BUCKETS <= 16 && !PACK) {
    // Few buckets (splitter sweeps): a tile holds a handful of LONG runs.  Walk them one by one with the threads
    // mapped onto the DESTINATION's 128-byte lines (thread t writes line offset t mod 32), so every warp store is one
    // whole, aligned line except at the two ends of a run.  The rank-indexed loop below starts every run at an
    // arbitrary element, i.e. every warp store straddles two lines and ends in partial sectors -- which is what
    // costs over NVLink when the destination is a peer (the fused partition + all-to-all of the multi-GPU sort).
    if (compute) {
#pragma unroll 1
      for (int d = 0; d < DigitFn::BUCKETS; ++d) {
        const uint32_t cnt = s.count[d];
        if (cnt == 0) continue;
        const uint32_t r0 = s.texcl[d];
        const G g0 = s.gbase[d] + (G)r0;  // position of the run's first element inside bucket d's destination
        U* kd;
        VT* vd = nullptr;
        if constexpr (DigitFn::PEER_OUT) {
          kd = digit.kdst[d] + g0;
          if constexpr (HAS_VALUES) vd = digit.vdst[d] + g0;
        } else {
          kd = keys_out + g0;
          if constexpr (HAS_VALUES) vd = vals_out + g0;
        }
        const uint32_t lead = (uint32_t)(((uintptr_t)kd / sizeof(U)) & 31u);
        for (uint32_t i = (uint32_t)tid; i < lead + cnt; i += BLOCK) {
          if (i >= lead) {
            const uint32_t off = i - lead;
            kd[off] = s.keys[r0 + off];
            if constexpr (HAS_VALUES) vd[off] = s.vals[r0 + off];
          }
        }
      }
    }
    return;
  }
  if (compute) {
    const uint2* const pairs = reinterpret_cast<const uint2*>(s.keys);
    (void)pairs;
    auto write_out = [&](auto full_c) {
      constexpr bool FULL = decltype(full_c)::value;
#pragma unroll
      for (int m = 0; m < ITEMS; ++m) {
        const int i = m * BLOCK + tid;
        if (FULL || i < tile_count) {
          if constexpr (PACK) {
            const uint2 kv = pairs[i];
            const unsigned d = digit((U)kv.x);
            const G dst = s.gbase[d] + (G)i;
            U kx = (U)kv.x;
            if constexpr (DigitFn::DECODE_ON_STORE) kx = DigitFn::decode(kx);
            if constexpr (DigitFn::PEER_OUT) {
              digit.kdst[d][dst] = kx;
              digit.vdst[d][dst] = bit_cast_<VT>(kv.y);
            } else {
              keys_out[dst] = kx;
              vals_out[dst] = bit_cast_<VT>(kv.y);
            }
          } else {
            U k = s.keys[i];
            const unsigned d = digit(k);
            const G dst = s.gbase[d] + (G)i;
            if constexpr (DigitFn::DECODE_ON_STORE) k = DigitFn::decode(k);
            if constexpr (DigitFn::PEER_OUT) {
              digit.kdst[d][dst] = k;
              if constexpr (HAS_VALUES) digit.vdst[d][dst] = s.vals[i];
            } else {
              keys_out[dst] = k;
              if constexpr (HAS_VALUES) vals_out[dst] = s.vals[i];
            }
          }
        }
      }
    };
    if (full) write_out(std::true_type{});
    else write_out(std::false_type{});
  }
}

template <class U, class VT, class DigitFn, class C, class S>
__global__ void __launch_bounds__(S::THREADS, S::MIN_CTAS)
onesweep_late_kernel(const U* __restrict__ keys_in, U* __restrict__ keys_out, const VT* __restrict__ vals_in, VT* __restrict__ vals_out, const long long n,
                     const unsigned long long* __restrict__ bin_base, C* lookback, unsigned* ticket, const DigitFn digit)
{
  if constexpr (HasTwin<DigitFn>::value) {
    // the rare variant of the digit functor (keys contain a -0.0): same body, instantiated a second time; the branch is
    // uniform over the grid
    if (digit.use_twin()) {
      onesweep_late_body<U, VT, typename DigitFn::Twin, C, S>(keys_in, keys_out, vals_in, vals_out, n, bin_base, lookback, ticket, digit.twin());
      return;
    }
  }
  onesweep_late_body<U, VT, DigitFn, C, S>(keys_in, keys_out, vals_in, vals_out, n, bin_base, lookback, ticket, digit);
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
// Shapes picked on B200 with tools/tune/tune_sort.cu (profiles/r01_onesweep_tuning_v3.txt): 256 compute
// threads + 2 look-back warps, 3 CTAs per SM, late-count kernel, keys of the tile 300 ahead
// prefetched to L2; ITEMS is what the 64-register budget of that occupancy holds per key/value width.
template <class U, class VT> struct Tuning {
  static constexpr int VB = std::is_same<VT, NoValue>::value ? 0 : (int)sizeof(VT);
  static constexpr int ITEMS = sizeof(U) == 4 ? (VB <= 4 ? 32 : 20) : (VB == 0 ? 18 : (VB == 4 ? 16 : 14));
  using shape = SweepShape<256, ITEMS, 3, 3, false, 300, 2, 2>;
};
// 64-bit keys move twice the bytes per key: with the 3-CTA shapes a tile holds only 4-4.6 K keys and the per-tile costs
// (look-back round trips, barriers) are paid twice as often per byte as for the u32 pairs.  Larger tiles at 2 CTAs per SM
// (r02 sweep, profiles/r02_sort_tuning.txt): keys only 10.89 -> 9.85 ms at 2^28 (384 x 26 = 9984 keys per tile),
// u64 + 4-byte values 13.90 -> 12.99 ms (384 x 18); u64 + 8-byte values is flat and keeps the default.
template <> struct Tuning<unsigned long long, NoValue> {
  using shape = SweepShape<384, 26, 2, 3, false, 200, 2, 2>;
};
template <> struct Tuning<unsigned long long, uint32_t> {
  using shape = SweepShape<384, 18, 2, 3, false, 200, 2, 2>;
};

template <class U, class VT, class S = typename Tuning<U, VT>::shape> inline long long num_tiles(long long n)
{
  return (n + S::TILE - 1) / S::TILE;
}

// workspace layout: [ghist NPASS*256 u64][bin_base NPASS*256 u64][tickets NPASS u32, padded][NPASS look-back regions]
// (one region per pass, so that ONE memset up front clears the whole sort)
template <class U, class VT, class S = typename Tuning<U, VT>::shape> inline size_t workspace_bytes(long long n)
{
  constexpr int NPASS = sizeof(U);
  size_t header = (size_t)NPASS * RADIX * 8 * 2 + 512;
  size_t cbytes = (n < (1ll << 30)) ? 4 : 8;
  return header + NPASS * lookback_bytes<S>(num_tiles<U, VT, S>(n), cbytes);
}

// opt-in to > 48 KB of dynamic shared memory, once per kernel (keyed on the kernel itself:
// kernels of one signature share a function-pointer TYPE), host thread and device
template <auto Kern> inline cudaError_t configure_once(int smem_bytes)
{
  static thread_local PerDevice cache;
  int& configured = cache.current();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    configured = 1;
  }
  return cudaSuccess;
}

template <class U, class VT, bool FLOAT, class C, class S> struct SweepKernel {
  // MATCH < 3: early-counts kernel; MATCH >= 3: late-counts kernel
  using SM = typename std::conditional<(S::MATCH >= 3), LateSmem<U, VT, C, S>, OnesweepSmem<U, VT, C, S>>::type;
  static constexpr auto kernel()
  {
    if constexpr (S::MATCH >= 3) return &onesweep_late_kernel<U, VT, RadixDigit<U, FLOAT>, C, S>;
    else return &onesweep_kernel<U, VT, FLOAT, C, S>;
  }
};

template <class U, class VT, bool FLOAT, class C, class S>
inline cudaError_t launch_pass_k(cudaStream_t stream, const U* kin, U* kout, const VT* vin, VT* vout, long long n,
                                 const unsigned long long* bin_base, void* lookback, unsigned* ticket, int shift, unsigned dxor, long long ntiles,
                                 const unsigned* degenerate = nullptr)
{
  using K = SweepKernel<U, VT, FLOAT, C, S>;
  constexpr auto kern = K::kernel();
  cudaError_t e = configure_once<kern>((int)sizeof(typename K::SM));
  if (e != cudaSuccess) return e;
  if constexpr (S::MATCH >= 3)
    kern<<<(unsigned)ntiles, S::THREADS, sizeof(typename K::SM), stream>>>(kin, kout, vin, vout, n, bin_base, (C*)lookback, ticket, RadixDigit<U, FLOAT>{shift, dxor, degenerate});
  else
    kern<<<(unsigned)ntiles, S::THREADS, sizeof(typename K::SM), stream>>>(kin, kout, vin, vout, n, bin_base, (C*)lookback, ticket, shift, dxor);
  return cudaGetLastError();
}

// one pass of a floating-point sort with the keys in the ordered encoding between the passes (late-count kernel only)
template <class U, class VT, int PASS, class C, class S>
inline cudaError_t launch_pass_float(cudaStream_t stream, const U* kin, U* kout, const VT* vin, VT* vout, long long n, const unsigned long long* bin_base,
                                     void* lookback, unsigned* ticket, int shift, bool descending, const unsigned* negzero_flag, long long ntiles,
                                     const unsigned* degenerate = nullptr)
{
  using D = FloatDigit<U, PASS>;
  constexpr auto kern = &onesweep_late_kernel<U, VT, D, C, S>;
  cudaError_t e = configure_once<kern>((int)sizeof(LateSmem<U, VT, C, S>));
  if (e != cudaSuccess) return e;
  D digit;
  digit.shift = shift;
  digit.dxor = descending ? (unsigned)(RADIX - 1) : 0u;
  digit.negzero_flag = negzero_flag;
  digit.degenerate = degenerate;
  kern<<<(unsigned)ntiles, S::THREADS, sizeof(LateSmem<U, VT, C, S>), stream>>>(kin, kout, vin, vout, n, bin_base, (C*)lookback, ticket, digit);
  return cudaGetLastError();
}

// pass_pos: 0 = first, 1 = middle, 2 = last pass of the sort; negzero_flag: device word the histogram kernel sets when the
// keys contain a -0.0 (floating-point keys only)
template <class U, class VT, class C, class S>
inline cudaError_t launch_pass(cudaStream_t stream, const U* kin, U* kout, const VT* vin, VT* vout, long long n,
                               const unsigned long long* bin_base, void* lookback, unsigned* ticket, int shift, KeyOrder order, int pass_pos = 1,
                               const unsigned* negzero_flag = nullptr, const unsigned* degenerate = nullptr)
{
  const long long ntiles = num_tiles<U, VT, S>(n);
  const unsigned dxor = digit_xor<U>(order, shift);
  LaunchScope scope("onesweep_kernel", stream);
  if constexpr (S::MATCH >= 3) {
    if (order.kind == KEY_FLOAT && negzero_flag != nullptr) {
      const bool desc = order.descending != 0;
      if (pass_pos == 0) return launch_pass_float<U, VT, FLOAT_FIRST, C, S>(stream, kin, kout, vin, vout, n, bin_base, lookback, ticket, shift, desc, negzero_flag, ntiles, degenerate);
      if (pass_pos == 2) return launch_pass_float<U, VT, FLOAT_LAST, C, S>(stream, kin, kout, vin, vout, n, bin_base, lookback, ticket, shift, desc, negzero_flag, ntiles, degenerate);
      return launch_pass_float<U, VT, FLOAT_MIDDLE, C, S>(stream, kin, kout, vin, vout, n, bin_base, lookback, ticket, shift, desc, negzero_flag, ntiles, degenerate);
    }
  }
  if (order.kind == KEY_FLOAT) return launch_pass_k<U, VT, true, C, S>(stream, kin, kout, vin, vout, n, bin_base, lookback, ticket, shift, dxor, ntiles, degenerate);
  return launch_pass_k<U, VT, false, C, S>(stream, kin, kout, vin, vout, n, bin_base, lookback, ticket, shift, dxor, ntiles, degenerate);
}

// Sorts n keys (and values) held in (keys, vals) using (keys_alt, vals_alt) as the
// ping-pong partner.  The pass count is even, so the result lands in (keys, vals).
template <class U, class VT, class S = typename Tuning<U, VT>::shape>
inline cudaError_t sort(cudaStream_t stream, U* keys, U* keys_alt, VT* vals, VT* vals_alt, long long n, KeyOrder order, void* workspace)
{
  if (n <= 1) return cudaSuccess;
  constexpr int NPASS = sizeof(U);
  static_assert(NPASS % 2 == 0, "even pass count keeps the result in the caller's buffer");
  const size_t header = (size_t)NPASS * RADIX * 8 * 2 + 512;
  unsigned long long* ghist = (unsigned long long*)workspace;
  unsigned long long* bin_base = ghist + NPASS * RADIX;
  unsigned* tickets = (unsigned*)(bin_base + NPASS * RADIX);
  const bool small = n < (1ll << 30);
  const size_t lb_bytes = lookback_bytes<S>(num_tiles<U, VT, S>(n), small ? 4 : 8);
  unsigned* const degenerate_flags = tickets + 64;  // [NPASS]: the pass is an identity permutation (scan_bins_kernel); inside the 512-byte ticket area
  unsigned* const negzero_flag = tickets + 72;      // the keys contain a -0.0 (histogram_kernel)

  cudaError_t e = cudaMemsetAsync(workspace, 0, header + NPASS * lb_bytes, stream);
  if (e != cudaSuccess) return e;
  {
    using H = HistShape<U>;
    e = configure_once<histogram_kernel<U>>(H::SMEM);
    if (e != cudaSuccess) return e;
    long long grid = (n + H::BLOCK * H::VEC - 1) / (H::BLOCK * H::VEC);
    if (grid > device_info().sm_count) grid = device_info().sm_count;
    {
      LaunchScope scope("histogram_kernel", stream);
      histogram_kernel<U><<<(unsigned)grid, H::BLOCK, H::SMEM, stream>>>(keys, n, ghist, order, negzero_flag);
    }
    {
      LaunchScope scope("scan_bins_kernel", stream);
      scan_bins_kernel<<<NPASS, RADIX, 0, stream>>>(ghist, bin_base, (unsigned long long)n, degenerate_flags);
    }
  }
  for (int p = 0; p < NPASS; ++p) {
    const U* kin = (p & 1) ? keys_alt : keys;
    U* kout = (p & 1) ? keys : keys_alt;
    const VT* vin = (p & 1) ? vals_alt : vals;
    VT* vout = (p & 1) ? vals : vals_alt;
    void* lookback = (char*)workspace + header + p * lb_bytes;
    const int pos = p == 0 ? 0 : (p == NPASS - 1 ? 2 : 1);
    if (small)
      e = launch_pass<U, VT, uint32_t, S>(stream, kin, kout, vin, vout, n, bin_base + p * RADIX, lookback, tickets + p, p * RADIX_BITS, order, pos, negzero_flag, degenerate_flags + p);
    else
      e = launch_pass<U, VT, unsigned long long, S>(stream, kin, kout, vin, vout, n, bin_base + p * RADIX, lookback, tickets + p, p * RADIX_BITS, order, pos,
                                                    negzero_flag, degenerate_flags + p);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------------------
// Stable multi-way partition by splitters (one late-count sweep with SplitDigit): element i goes
// to bucket b = number of splitters that sort at or before keys[i]; inside a bucket the input
// order is kept.  counts[0 .. nsplit] (device, u64) receives the bucket sizes.  Used by the
// multi-GPU sample sort: partition -> all-to-all -> ONE local sort (raja_b200/multi.py).
// ---------------------------------------------------------------------------
// NS = number of splitters, a compile-time constant inside the loop (the kernel switches on it once)
template <class U, class DigitFn, int BLOCK, int NS>
B200_D void bucket_count_body(const U* __restrict__ keys, const long long n, unsigned long long* __restrict__ ghist, const DigitFn& digit)
{
  // Per thread: c[i] = number of keys at or after splitter i (one compare + one add per splitter and key; evaluating
  // the bucket first and then comparing it against all 8 bucket ids made this kernel instruction-bound at 2.1 TB/s).
  // Bucket sizes are differences of these cumulative counts.  Warp shuffles, then one global atomic per warp and bucket.
  // 128-bit loads, four per thread in flight.
  unsigned c[NS > 0 ? NS : 1] = {};
  unsigned seen = 0;
  auto count = [&](U k) {
    const U ok = to_ordered(k, digit.order);
    ++seen;
#pragma unroll
    for (int i = 0; i < NS; ++i) c[i] += (ok >= digit.splitter(i)) ? 1u : 0u;
  };
  constexpr int VEC = 16 / sizeof(U), UNROLL = 4;
  using V = Vec<U, VEC>;
  long long head = (long long)(((16 - ((uintptr_t)keys & 15)) & 15) / sizeof(U));
  if (head > n) head = n;
  const long long nvec = (n - head) / VEC;
  const V* const body = reinterpret_cast<const V*>(keys + head);
  long long v = (long long)blockIdx.x * BLOCK * UNROLL + threadIdx.x;
  const long long stride = (long long)gridDim.x * BLOCK * UNROLL;
  for (; v + (UNROLL - 1) * BLOCK < nvec; v += stride) {
    V x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) x[u] = ld_stream(body + v + u * BLOCK);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int e = 0; e < VEC; ++e) count(x[u].e[e]);
  }
  for (int u = 0; u < UNROLL; ++u)
    if (v + u * BLOCK < nvec) {
      const V x = ld_stream(body + v + u * BLOCK);
#pragma unroll
      for (int e = 0; e < VEC; ++e) count(x.e[e]);
    }
  if (blockIdx.x == 0) {  // unaligned head and the tail shorter than one vector
    for (long long i = threadIdx.x; i < head; i += BLOCK) count(keys[i]);
    for (long long i = head + nvec * VEC + threadIdx.x; i < n; i += BLOCK) count(keys[i]);
  }
  // bucket b = [splitter b-1, splitter b): seen - c[0], c[0] - c[1], ..., c[NS-1]
#pragma unroll
  for (int b = 0; b <= NS; ++b) {
    unsigned v2 = (b == 0 ? seen : c[b > 0 ? b - 1 : 0]) - (b < NS ? c[b < NS ? b : 0] : 0u);
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) v2 += __shfl_xor_sync(FULL_MASK, v2, k);
    if ((threadIdx.x & 31) == 0 && v2) atomicAdd(ghist + b, (unsigned long long)v2);
  }
}

template <class U, class DigitFn, int BLOCK>
__global__ void __launch_bounds__(BLOCK) bucket_count_kernel(const U* __restrict__ keys, const long long n, unsigned long long* __restrict__ ghist, const DigitFn digit)
{
  switch (digit.nsplit) {
    case 0: bucket_count_body<U, DigitFn, BLOCK, 0>(keys, n, ghist, digit); break;
    case 1: bucket_count_body<U, DigitFn, BLOCK, 1>(keys, n, ghist, digit); break;
    case 2: bucket_count_body<U, DigitFn, BLOCK, 2>(keys, n, ghist, digit); break;
    case 3: bucket_count_body<U, DigitFn, BLOCK, 3>(keys, n, ghist, digit); break;
    case 4: bucket_count_body<U, DigitFn, BLOCK, 4>(keys, n, ghist, digit); break;
    case 5: bucket_count_body<U, DigitFn, BLOCK, 5>(keys, n, ghist, digit); break;
    case 6: bucket_count_body<U, DigitFn, BLOCK, 6>(keys, n, ghist, digit); break;
    default: bucket_count_body<U, DigitFn, BLOCK, MAX_SPLITTERS>(keys, n, ghist, digit); break;
  }
}

template <class U, class VT> struct PartitionTuning {
  // the splitter digit costs ~16 instructions where the radix digit costs 2: a smaller ITEMS keeps
  // the three digit evaluations per key inside the register budget
  static constexpr int VB = std::is_same<VT, NoValue>::value ? 0 : (int)sizeof(VT);
  static constexpr int ITEMS = sizeof(U) == 4 ? (VB <= 4 ? 24 : 16) : (VB == 0 ? 14 : 12);
  using shape = SweepShape<256, ITEMS, 3, 3, false, 300, 2, 2>;
};

// workspace: same layout as sort's ([ghist 256 u64 x NPASS][bin_base ...][tickets][lookback]); only pass 0 is used
template <class U, class VT, class S = typename PartitionTuning<U, VT>::shape> inline size_t partition_workspace_bytes(long long n)
{
  return workspace_bytes<U, VT, S>(n);
}

// bucket sizes of a split: counts_dev[0 .. MAX_SPLITTERS] (u64, device); workspace as for partition()
template <class U, class VT, class S = typename PartitionTuning<U, VT>::shape, class DigitFn = SplitDigit<U>>
inline cudaError_t bucket_counts(cudaStream_t stream, const U* keys_in, long long n, const DigitFn& digit, unsigned long long* counts_dev, void* workspace)
{
  constexpr int NPASS = sizeof(U);
  const size_t header = (size_t)NPASS * RADIX * 8 * 2 + 512;
  unsigned long long* ghist = (unsigned long long*)workspace;
  const bool small = n < (1ll << 30);
  const long long ntiles = num_tiles<U, VT, S>(n > 0 ? n : 1);
  const size_t lb_bytes = lookback_bytes<S>(ntiles, small ? 4 : 8);
  cudaError_t e = cudaMemsetAsync(workspace, 0, header + lb_bytes, stream);
  if (e != cudaSuccess) return e;
  if (n > 0) {
    constexpr int CB = 512;
    long long grid = (n + CB * 16 - 1) / (CB * 16);
    const long long cap = (long long)device_info().sm_count * 4;
    if (grid > cap) grid = cap;
    LaunchScope scope("bucket_count_kernel", stream);
    bucket_count_kernel<U, DigitFn, CB><<<(unsigned)grid, CB, 0, stream>>>(keys_in, n, ghist, digit);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaMemcpyAsync(counts_dev, ghist, (MAX_SPLITTERS + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream);
}

// the sweep of a split; bucket_counts() must have run on the same workspace.  bases_host == nullptr:
// buckets back to back in (keys_out, vals_out) (exclusive scan of the counts on the device);
// otherwise bucket d starts at element bases_host[d] of ITS destination (PeerSplitDigit).
template <class U, class VT, class DigitFn, class S = typename PartitionTuning<U, VT>::shape>
inline cudaError_t split_sweep(cudaStream_t stream, const U* keys_in, U* keys_out, const VT* vals_in, VT* vals_out, long long n, const DigitFn& digit,
                               const unsigned long long* bases_host, void* workspace)
{
  if (n <= 0) return cudaSuccess;
  constexpr int NPASS = sizeof(U);
  const size_t header = (size_t)NPASS * RADIX * 8 * 2 + 512;
  unsigned long long* ghist = (unsigned long long*)workspace;
  unsigned long long* bin_base = ghist + NPASS * RADIX;
  unsigned* tickets = (unsigned*)(bin_base + NPASS * RADIX);
  void* lookback = (char*)workspace + header;
  const bool small = n < (1ll << 30);
  const long long ntiles = num_tiles<U, VT, S>(n);
  cudaError_t e;
  if (bases_host) {
    e = cudaMemcpyAsync(bin_base, bases_host, (MAX_SPLITTERS + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
  } else {
    LaunchScope scope("scan_bins_kernel", stream);
    scan_bins_kernel<<<1, RADIX, 0, stream>>>(ghist, bin_base);
  }
  LaunchScope scope("onesweep_kernel", stream);
  if (small) {
    constexpr auto kern = &onesweep_late_kernel<U, VT, DigitFn, uint32_t, S>;
    e = configure_once<kern>((int)sizeof(LateSmem<U, VT, uint32_t, S>));
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)ntiles, S::THREADS, sizeof(LateSmem<U, VT, uint32_t, S>), stream>>>(keys_in, keys_out, vals_in, vals_out, n, bin_base, (uint32_t*)lookback, tickets, digit);
  } else {
    constexpr auto kern = &onesweep_late_kernel<U, VT, DigitFn, unsigned long long, S>;
    e = configure_once<kern>((int)sizeof(LateSmem<U, VT, unsigned long long, S>));
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)ntiles, S::THREADS, sizeof(LateSmem<U, VT, unsigned long long, S>), stream>>>(keys_in, keys_out, vals_in, vals_out, n, bin_base, (unsigned long long*)lookback, tickets, digit);
  }
  return cudaGetLastError();
}

template <class U, class VT, class S = typename PartitionTuning<U, VT>::shape>
inline cudaError_t partition(cudaStream_t stream, const U* keys_in, U* keys_out, const VT* vals_in, VT* vals_out, long long n,
                             const SplitDigit<U>& digit, unsigned long long* counts_dev, void* workspace)
{
  cudaError_t e = bucket_counts<U, VT, S>(stream, keys_in, n, digit, counts_dev, workspace);
  if (e != cudaSuccess) return e;
  return split_sweep<U, VT, SplitDigit<U>, S>(stream, keys_in, keys_out, vals_in, vals_out, n, digit, nullptr, workspace);
}

}  // namespace rsort
}  // namespace b200
