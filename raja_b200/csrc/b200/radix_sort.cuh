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
//   onesweep_kernel   : per pass; tile claimed by ticket -> warp-striped key load
//                       -> match.any ranking -> per-digit look-back -> scatter to
//                       shared memory in rank order -> run-coalesced global write
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

B200_D uint32_t ld_lookback(const uint32_t* p) { return ld_relaxed_u32(p); }
B200_D unsigned long long ld_lookback(const unsigned long long* p) { return ld_relaxed_u64((const uint64_t*)p); }
B200_D void st_lookback(uint32_t* p, uint32_t v) { st_relaxed_u32(p, v); }
B200_D void st_lookback(unsigned long long* p, unsigned long long v) { st_relaxed_u64((uint64_t*)p, v); }

template <class U, class VT, int BLOCK, int ITEMS, class C> struct OnesweepSmem {
  static constexpr int NW = BLOCK / WARP;
  static constexpr int TILE = BLOCK * ITEMS;
  static constexpr bool HAS_VALUES = !std::is_same<VT, NoValue>::value;
  uint32_t whist[NW][RADIX];   // per-warp digit counts -> per-warp digit bases
  long long gofs[RADIX];       // global position of rank i of digit d = gofs[d] + i
  uint32_t wtot[RADIX / WARP];
  long long tile;
  U keys[TILE];
  typename std::conditional<HAS_VALUES, VT, char>::type vals[HAS_VALUES ? TILE : 1];
};

template <class U, class VT, int BLOCK, int ITEMS, class C, int MIN_CTAS>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS)
onesweep_kernel(const U* __restrict__ keys_in, U* __restrict__ keys_out,
                const VT* __restrict__ vals_in, VT* __restrict__ vals_out,
                const long long n,
                const unsigned long long* __restrict__ bin_base,  // [256] of this pass
                C* lookback,                                      // [ntiles][256], zero on entry
                unsigned* ticket,                                 // zero on entry
                const int shift, const KeyOrder order)
{
  using SM = OnesweepSmem<U, VT, BLOCK, ITEMS, C>;
  using F = LookbackFlags<C>;
  constexpr int NW = SM::NW;
  constexpr int TILE = SM::TILE;
  constexpr bool HAS_VALUES = SM::HAS_VALUES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& s = *reinterpret_cast<SM*>(smem_raw);

  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  if (tid == 0) s.tile = (long long)atomicAdd(ticket, 1u);
  for (int i = tid; i < NW * RADIX; i += BLOCK) (&s.whist[0][0])[i] = 0;
  __syncthreads();
  const long long tile = s.tile;
  const long long tile_base = tile * TILE;
  const int tile_count = (int)((n - tile_base) < (long long)TILE ? (n - tile_base) : (long long)TILE);
  const bool full = (tile_count == TILE);
  const int wofs = w * (WARP * ITEMS);  // this warp's first item within the tile

  // ---- load keys, warp-striped: item (j, l) is tile item wofs + j*32 + l ---------
  U key[ITEMS];
  if (full) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) key[j] = ld_stream(keys_in + tile_base + wofs + j * WARP + l);
  } else {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      int it = wofs + j * WARP + l;
      key[j] = (it < tile_count) ? ld_stream(keys_in + tile_base + it) : U(0);
    }
  }

  // ---- phase A: stable ranks within the warp + warp digit counts -------------------
  uint32_t* wh = s.whist[w];
  const unsigned lt = lanemask_lt();
  uint16_t rank[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const bool valid = full || (wofs + j * WARP + l < tile_count);
    const int d = valid ? digit_of(key[j], order, shift) : RADIX;  // invalid lanes group apart
    const unsigned peers = __match_any_sync(FULL_MASK, d);
    const int leader = __ffs(peers) - 1;
    uint32_t before = 0;
    if (l == leader && valid) {
      before = wh[d];
      wh[d] = before + __popc(peers);
    }
    before = __shfl_sync(FULL_MASK, before, leader);
    rank[j] = (uint16_t)(before + __popc(peers & lt));
    __syncwarp();
  }
  __syncthreads();

  // ---- phase B: per-digit bases across warps, tile scan over digits, look-back -------
  uint32_t count = 0, incl = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      uint32_t t = s.whist[i][tid];
      s.whist[i][tid] = count;
      count += t;
    }
    // publish the tile-local count as early as possible
    if (tile != 0) st_lookback(lookback + tile * RADIX + tid, (C)count | F::LOCAL);
    incl = count;
#pragma unroll
    for (int k = 1; k < WARP; k <<= 1) {
      uint32_t o = __shfl_up_sync(FULL_MASK, incl, k);
      if (l >= k) incl += o;
    }
    if (l == 31) s.wtot[w] = incl;
  }
  __syncthreads();
  if (tid < RADIX) {
    uint32_t pre = 0;
#pragma unroll
    for (int i = 0; i < RADIX / WARP; ++i)
      if (i < w) pre += s.wtot[i];
    const uint32_t tile_excl = pre + incl - count;  // first rank of digit tid inside this tile

    // decoupled look-back over predecessor tiles for this digit
    C prev = 0;
    if (tile != 0) {
      long long t = tile - 1;
      while (true) {
        C v = ld_lookback(lookback + t * RADIX + tid);
        while ((v & ~F::MASK) == 0) {
          __nanosleep(20);
          v = ld_lookback(lookback + t * RADIX + tid);
        }
        prev += v & F::MASK;
        if (v & F::GLOBAL) break;
        --t;
      }
    }
    st_lookback(lookback + tile * RADIX + tid, (C)(prev + count) | F::GLOBAL);
    s.gofs[tid] = (long long)(bin_base[tid] + (unsigned long long)prev) - (long long)tile_excl;
#pragma unroll
    for (int i = 0; i < NW; ++i) s.whist[i][tid] += tile_excl;
  }
  __syncthreads();

  // ---- phase C: scatter into shared memory in rank order ------------------------------
  if constexpr (HAS_VALUES) {
    VT val[ITEMS];
    if (full) {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) val[j] = ld_stream(vals_in + tile_base + wofs + j * WARP + l);
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        int it = wofs + j * WARP + l;
        if (it < tile_count) val[j] = ld_stream(vals_in + tile_base + it);
      }
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (full || (wofs + j * WARP + l < tile_count)) {
        const int pos = (int)wh[digit_of(key[j], order, shift)] + rank[j];
        s.keys[pos] = key[j];
        s.vals[pos] = val[j];
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (full || (wofs + j * WARP + l < tile_count)) {
        const int pos = (int)wh[digit_of(key[j], order, shift)] + rank[j];
        s.keys[pos] = key[j];
      }
    }
  }
  __syncthreads();

  // ---- phase D: run-coalesced write-out -------------------------------------------------
#pragma unroll
  for (int m = 0; m < ITEMS; ++m) {
    const int i = m * BLOCK + tid;
    if (full || i < tile_count) {
      const U k = s.keys[i];
      const long long dst = s.gofs[digit_of(k, order, shift)] + i;
      keys_out[dst] = k;
      if constexpr (HAS_VALUES) vals_out[dst] = s.vals[i];
    }
  }
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
template <class U, class VT> struct Tuning {
  static constexpr int BLOCK = 384;
  static constexpr int ITEMS = sizeof(U) == 4 ? (sizeof(VT) <= 4 ? 20 : 16) : (std::is_same<VT, NoValue>::value ? 16 : 12);
  static constexpr int MIN_CTAS = 2;
};

template <class U, class VT> inline long long num_tiles(long long n)
{
  using Tn = Tuning<U, VT>;
  return (n + (long long)Tn::BLOCK * Tn::ITEMS - 1) / ((long long)Tn::BLOCK * Tn::ITEMS);
}

// workspace layout: [ghist NPASS*256 u64][bin_base NPASS*256 u64][tickets NPASS u32, padded][lookback ntiles*256*C]
template <class U, class VT> inline size_t workspace_bytes(long long n)
{
  constexpr int NPASS = sizeof(U);
  size_t header = (size_t)NPASS * RADIX * 8 * 2 + 256;
  size_t cbytes = (n < (1ll << 30)) ? 4 : 8;
  return header + (size_t)num_tiles<U, VT>(n) * RADIX * cbytes;
}

template <class U, class VT, class C>
inline cudaError_t launch_pass(cudaStream_t stream, const U* kin, U* kout, const VT* vin, VT* vout, long long n,
                               const unsigned long long* bin_base, void* lookback, unsigned* ticket, int shift, KeyOrder order)
{
  using Tn = Tuning<U, VT>;
  using SM = OnesweepSmem<U, VT, Tn::BLOCK, Tn::ITEMS, C>;
  auto kern = onesweep_kernel<U, VT, Tn::BLOCK, Tn::ITEMS, C, Tn::MIN_CTAS>;
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long long ntiles = num_tiles<U, VT>(n);
  LaunchScope scope("onesweep_kernel", stream);
  kern<<<(unsigned)ntiles, Tn::BLOCK, sizeof(SM), stream>>>(kin, kout, vin, vout, n, bin_base, (C*)lookback, ticket, shift, order);
  return cudaGetLastError();
}

// Sorts n keys (and values) held in (keys, vals) using (keys_alt, vals_alt) as the
// ping-pong partner.  The pass count is even, so the result lands in (keys, vals).
template <class U, class VT>
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
  const size_t lb_bytes = (size_t)num_tiles<U, VT>(n) * RADIX * (small ? 4 : 8);

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
      e = launch_pass<U, VT, uint32_t>(stream, kin, kout, vin, vout, n, bin_base + p * RADIX, lookback, tickets + p, p * RADIX_BITS, order);
    else
      e = launch_pass<U, VT, unsigned long long>(stream, kin, kout, vin, vout, n, bin_base + p * RADIX, lookback, tickets + p, p * RADIX_BITS, order);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace rsort
}  // namespace b200
