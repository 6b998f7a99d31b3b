// raja_b200 core: single-pass prefix scan with decoupled look-back.
//
// Replaces the reference's policy/cuda/scan.hpp:52-235 (four entry points that
// each delegate to cub::DeviceScan).  One kernel reads N and writes N:
//   * tiles are claimed through an atomic ticket, so a tile's predecessors are
//     always resident or finished (forward progress without relying on blockIdx
//     scheduling order);
//   * data stays in registers in a warp-striped vector layout (lane l of warp w
//     owns V consecutive items at row k): loads and stores are 128-bit and fully
//     coalesced with no shared-memory transpose; the in-warp scan costs K
//     shuffle scans per thread, which is noise next to the HBM time;
//   * tile status + value travel in one 64-bit word for 4-byte T; 8-byte T uses
//     separate value arrays published with st.release / read with ld.acquire;
//   * 64-bit lengths (the reference truncates to int: scan.hpp:65).
//
// Semantics follow the reference's sequential definition
// (policy/sequential/scan.hpp:41-152): inclusive out[i] = x0 (+) ... (+) xi;
// exclusive out[i] = init (+) x0 (+) ... (+) x(i-1).  in == out is allowed
// (the *_inplace variants); partial overlap is not.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace b200 {
namespace scan {

enum : uint32_t { TS_INVALID = 0, TS_AGGREGATE = 1, TS_INCLUSIVE = 2 };

template <class T, bool Packed = (sizeof(T) == 4)> struct TileState;

template <class T> struct TileState<T, true> {
  uint64_t* words;
  static size_t bytes(long long ntiles) { return (size_t)ntiles * 8; }
  void bind(void* base, long long) { words = (uint64_t*)base; }
  B200_D void publish(long long tile, uint32_t status, T v) const
  {
    uint32_t bits;
    memcpy(&bits, &v, 4);
    st_relaxed_u64(words + tile, ((uint64_t)status << 32) | bits);
  }
  B200_D uint32_t peek(long long tile, T& v) const
  {
    uint64_t w = ld_relaxed_u64(words + tile);
    uint32_t bits = (uint32_t)w;
    memcpy(&v, &bits, 4);
    return (uint32_t)(w >> 32);
  }
};

template <class T> struct TileState<T, false> {
  static_assert(sizeof(T) == 8, "scan supports 4- and 8-byte types");
  uint32_t* status;
  uint64_t* agg;
  uint64_t* incl;
  static size_t bytes(long long ntiles)
  {
    size_t st = ((size_t)ntiles * 4 + 127) & ~size_t(127);
    return st + (size_t)ntiles * 16;
  }
  void bind(void* base, long long ntiles)
  {
    size_t st = ((size_t)ntiles * 4 + 127) & ~size_t(127);
    status = (uint32_t*)base;
    agg = (uint64_t*)((char*)base + st);
    incl = agg + ntiles;
  }
  B200_D void publish(long long tile, uint32_t s, T v) const
  {
    uint64_t bits;
    memcpy(&bits, &v, 8);
    st_relaxed_u64((s == TS_AGGREGATE ? agg : incl) + tile, bits);
    st_release_u32(status + tile, s);
  }
  B200_D uint32_t peek(long long tile, T& v) const
  {
    uint32_t s = ld_acquire_u32(status + tile);
    if (s != TS_INVALID) {
      uint64_t bits = ld_relaxed_u64((s == TS_AGGREGATE ? agg : incl) + tile);
      memcpy(&v, &bits, 8);
    }
    return s;
  }
};

// BLOCK threads, each owning K rows of V consecutive items.
template <class T, class Op, bool Inclusive, int BLOCK, int K, int V>
__global__ void __launch_bounds__(BLOCK)
scan_kernel(const T* in, T* out, const long long n, const T init, const TileState<T> ts, unsigned* ticket)
{
  constexpr int ITEMS = K * V;
  constexpr int TILE = BLOCK * ITEMS;
  constexpr int NW = BLOCK / WARP;
  constexpr int WARP_ITEMS = WARP * ITEMS;
  using VT = Vec<T, V>;

  __shared__ long long s_tile;
  __shared__ T s_warp[NW];
  __shared__ T s_prefix;

  const int tid = threadIdx.x;
  if (tid == 0) s_tile = (long long)atomicAdd(ticket, 1u);
  __syncthreads();
  const long long tile = s_tile;
  const long long tile_base = tile * TILE;
  const int w = tid >> 5, l = tid & 31;
  const long long wbase = tile_base + (long long)w * WARP_ITEMS;
  const bool full = (tile_base + TILE <= n);
  Op op;
  const T ident = Op::identity();

  // ---- load (in may alias out: coherent loads) --------------------------------
  T x[K][V];
  if (full) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      VT v = ld_stream_coherent(reinterpret_cast<const VT*>(in + wbase + (long long)(k * WARP + l) * V));
#pragma unroll
      for (int j = 0; j < V; ++j) x[k][j] = v.e[j];
    }
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        long long e = wbase + (long long)(k * WARP + l) * V + j;
        x[k][j] = (e < n) ? ld_stream_coherent(in + e) : ident;
      }
    }
  }

  // ---- thread-local inclusive scan of each row --------------------------------
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int j = 1; j < V; ++j) x[k][j] = op(x[k][j - 1], x[k][j]);
  }

  // ---- warp scan across lanes per row, rows chained in registers ---------------
  T lane_excl[K];  // (+) of everything in this warp before x[k][0]
  T run = ident;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    T s = x[k][V - 1];
#pragma unroll
    for (int d = 1; d < WARP; d <<= 1) {
      T o = shfl_up(s, d);
      if (l >= d) s = op(o, s);
    }
    T row_total = shfl_idx(s, WARP - 1);
    T e = shfl_up(s, 1);
    if (l == 0) e = ident;
    lane_excl[k] = op(run, e);
    run = op(run, row_total);
  }
  if (l == 0) s_warp[w] = run;
  __syncthreads();

  T wprefix = ident, block_agg = ident;
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    T t = s_warp[i];
    if (i < w) wprefix = op(wprefix, t);
    block_agg = op(block_agg, t);
  }

  // ---- decoupled look-back (warp 0) ---------------------------------------------
  if (w == 0) {
    T excl = ident;
    if (tile == 0) {
      excl = Inclusive ? ident : init;
      if (l == 0) ts.publish(0, TS_INCLUSIVE, op(excl, block_agg));
    } else {
      if (l == 0) ts.publish(tile, TS_AGGREGATE, block_agg);
      long long pred = tile - 1;
      while (true) {
        const long long idx = pred - l;
        uint32_t st;
        T val = ident;
        int spins = 0;
        do {
          st = (idx >= 0) ? ts.peek(idx, val) : (uint32_t)TS_INCLUSIVE;
          if (++spins > 4) __nanosleep(32);
        } while (__any_sync(FULL_MASK, st == TS_INVALID));
        if (idx < 0) val = ident;
        const unsigned incl_mask = __ballot_sync(FULL_MASK, st == TS_INCLUSIVE);
        const int first = __ffs(incl_mask) - 1;  // nearest predecessor holding an inclusive prefix
        const bool contrib = (first < 0) || (l <= first);
        T c = contrib ? val : ident;
        // fold so that farther tiles stay on the left: reverse lane order
        T rc = shfl_idx(c, WARP - 1 - l);
        T window = warp_allreduce(rc, op);
        excl = op(window, excl);
        if (first >= 0) break;
        pred -= WARP;
      }
      if (l == 0) ts.publish(tile, TS_INCLUSIVE, op(excl, block_agg));
    }
    if (l == 0) s_prefix = excl;
  }
  __syncthreads();
  const T tile_prefix = op(s_prefix, wprefix);

  // ---- apply prefix, store ----------------------------------------------------------
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const T base = op(tile_prefix, lane_excl[k]);
    VT v;
    if (Inclusive) {
#pragma unroll
      for (int j = 0; j < V; ++j) v.e[j] = op(base, x[k][j]);
    } else {
      v.e[0] = base;
#pragma unroll
      for (int j = 1; j < V; ++j) v.e[j] = op(base, x[k][j - 1]);
    }
    const long long e0 = wbase + (long long)(k * WARP + l) * V;
    if (full) {
      st_stream(reinterpret_cast<VT*>(out + e0), v);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (e0 + j < n) out[e0 + j] = v.e[j];
    }
  }
}

// Workspace: [ticket (128 B)] [tile state].  Must be zero before the launch.
template <class T> inline size_t workspace_bytes(long long ntiles) { return 128 + TileState<T>::bytes(ntiles); }

template <class T> struct Tuning {
  // 4-byte: 384 thr x 20 items = 7680 items (30 KB) per tile; 8-byte: 384 x 12 (36 KB)
  static constexpr int BLOCK = 384;
  static constexpr int V = 16 / sizeof(T);
  static constexpr int K = sizeof(T) == 4 ? 5 : 6;
  static constexpr int K_UNALIGNED = sizeof(T) == 4 ? 16 : 10;  // V = 1 fallback
};

template <class T> inline long long num_tiles(long long n, bool aligned)
{
  using Tn = Tuning<T>;
  long long tile = aligned ? (long long)Tn::BLOCK * Tn::K * Tn::V : (long long)Tn::BLOCK * Tn::K_UNALIGNED;
  return (n + tile - 1) / tile;
}

template <class T> inline bool vector_aligned(const void* in, const void* out)
{
  return (((uintptr_t)in | (uintptr_t)out) & 15u) == 0;
}

// Enqueue one scan on `stream`.  `workspace` must be workspace_bytes<T>(num_tiles)
// bytes of device memory; it is zeroed here (stream-ordered).
template <class T, class Op, bool Inclusive>
inline cudaError_t launch(cudaStream_t stream, const T* in, T* out, long long n, T init, void* workspace)
{
  if (n <= 0) return cudaSuccess;
  using Tn = Tuning<T>;
  const bool aligned = vector_aligned<T>(in, out);
  const long long ntiles = num_tiles<T>(n, aligned);
  cudaError_t e = cudaMemsetAsync(workspace, 0, workspace_bytes<T>(ntiles), stream);
  if (e != cudaSuccess) return e;
  TileState<T> ts;
  ts.bind((char*)workspace + 128, ntiles);
  unsigned* ticket = (unsigned*)workspace;
  LaunchScope scope("scan_kernel", stream);
  if (aligned)
    scan_kernel<T, Op, Inclusive, Tn::BLOCK, Tn::K, Tn::V><<<(unsigned)ntiles, Tn::BLOCK, 0, stream>>>(in, out, n, init, ts, ticket);
  else
    scan_kernel<T, Op, Inclusive, Tn::BLOCK, Tn::K_UNALIGNED, 1><<<(unsigned)ntiles, Tn::BLOCK, 0, stream>>>(in, out, n, init, ts, ticket);
  return cudaGetLastError();
}

}  // namespace scan
}  // namespace b200
