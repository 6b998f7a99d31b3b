// raja_b200 core: single-pass prefix scan -- persistent CTAs, TMA bulk staging,
// early aggregates, full-round decoupled look-back.
//
// Replaces the reference's policy/cuda/scan.hpp:52-235 (four entry points that
// each delegate to cub::DeviceScan).  One kernel reads N and writes N.
//
//   * persistent grid (SMs x resident CTAs); a CTA takes a virtual id from an
//     atomic ticket and then owns tiles vid, vid+G, vid+2G, ...  A CTA that holds
//     virtual id v was started after every id < v.  A tile of a CTA's SECOND round looks
//     back at tiles of higher virtual ids, so multi-round grids need every CTA resident:
//     they are launched cooperatively (gang-scheduled, fails instead of hanging when the
//     grid cannot fit) and the look-back spin is bounded (trap after ~10 s);
//   * each tile is brought into shared memory by ONE bulk async copy
//     (cp.async.bulk global->shared, completion on an mbarrier) issued three tiles
//     ahead, and leaves through ONE bulk async copy shared->global: without the
//     look-back this pipeline alone streams at ~95 % of a device memcpy
//     (profiles/r01_scan_tuning.txt);
//   * in shared memory a thread owns VPT consecutive 16-byte vectors (VPT odd: the
//     128-bit accesses are bank-conflict free);
//   * the look-back is what bounds a chained scan on this part: under a saturated
//     HBM pipe a publish -> observe hop through L2 costs microseconds, and the
//     classic formulation (publish aggregate, then walk back 32 tiles per hop,
//     everybody waiting at a barrier) exposes several hops per tile (v1: 68 % of
//     warp samples at that barrier, 45 % of HBM peak).  Here the chain is taken
//     off the critical path altogether.  A CTA is BLOCK compute threads plus one
//     LOOK-BACK WARP, and iteration i of a CTA runs
//       compute warps : reduce tile i+2 and publish its aggregate (two iterations
//                       early), then scan tile i with its already-known prefix;
//       look-back warp: prefix(i+1) = inclusive(own tile i-1)
//                                     (+) aggregates of the 2G-1 tiles in between,
//                       fetched in ONE batch of loads (G = grid size), and the
//                       bulk copies.
//     Every aggregate the look-back warp needs was published a full iteration
//     before it asks, and its answer is not needed until the next iteration, so
//     neither the store nor the load latency is ever waited for; there are no
//     "inclusive" states at all, the fold order is fixed, and floating-point
//     scans are therefore run-to-run deterministic (CUB's are not);
//   * a tile's "published" flag + aggregate travel in one 64-bit word for 4-byte T;
//     an 8-byte T is split into two such words (flag + low half, flag + high half)
//     and is valid when both flags are set, so no acquire/release chain is needed;
//   * 64-bit lengths (the reference truncates to int: scan.hpp:65).
//
// Semantics follow the reference's sequential definition
// (policy/sequential/scan.hpp:41-152): inclusive out[i] = x0 (+) ... (+) xi;
// exclusive out[i] = init (+) x0 (+) ... (+) x(i-1).  in == out is allowed
// (the *_inplace variants); partial overlap is not.
#pragma once
#include "common.cuh"
#include "reduce.cuh"
#include "tma.cuh"

namespace b200 {
namespace scan {

template <class T, bool Packed = (sizeof(T) == 4)> struct TileState;

// 4-byte T: one 64-bit word per tile, (1 << 32 | aggregate); zero = not published.
template <class T> struct TileState<T, true> {
  uint64_t* words;
  static size_t bytes(long long ntiles) { return (size_t)ntiles * 8; }
  void bind(void* base, long long) { words = (uint64_t*)base; }
  B200_D void publish(long long tile, T v) const
  {
    uint32_t bits;
    memcpy(&bits, &v, 4);
    st_relaxed_u64(words + tile, (1ull << 32) | bits);
  }
  B200_D bool peek(long long tile, T& v) const
  {
    const uint64_t w = ld_relaxed_u64(words + tile);
    const uint32_t bits = (uint32_t)w;
    memcpy(&v, &bits, 4);
    return (w >> 32) != 0;
  }
};

// 8-byte T: two 64-bit words per tile, each (1 << 32 | half of the aggregate).  Each word
// is written once, so the value is complete exactly when both flags are set -- a torn
// observation just reads as "not published yet".
template <class T> struct TileState<T, false> {
  static_assert(sizeof(T) == 8, "scan supports 4- and 8-byte types");
  uint64_t* words;  // [ntiles][2], 16-byte aligned
  static size_t bytes(long long ntiles) { return (size_t)ntiles * 16; }
  void bind(void* base, long long) { words = (uint64_t*)base; }
  B200_D void publish(long long tile, T v) const
  {
    uint64_t bits;
    memcpy(&bits, &v, 8);
    const uint64_t lo = (1ull << 32) | (bits & 0xffffffffull);
    const uint64_t hi = (1ull << 32) | (bits >> 32);
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(words + 2 * tile), "l"(lo), "l"(hi) : "memory");
  }
  B200_D bool peek(long long tile, T& v) const
  {
    uint64_t lo, hi;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(words + 2 * tile) : "memory");
    const uint64_t bits = (hi << 32) | (lo & 0xffffffffull);
    memcpy(&v, &bits, 8);
    return ((lo >> 32) & (hi >> 32)) != 0;
  }
};

// Compile-time shape of one scan kernel.
//   BLOCK  compute threads per CTA (the CTA has BLOCK + 32 threads)
//   VPT    16-byte vectors per thread (odd => conflict-free LDS.128/STS.128)
//   ROWS   look-back batch = ROWS*32 predecessor aggregates per lane-row; the launcher
//          keeps 2*grid - 1 <= ROWS*32
//   SLOTS  shared-memory tile slots; LAG = how many outgoing copies may still be reading
//          their slot when the next load is issued.  Tile i+2 is reduced while tile i is
//          scanned, so SLOTS >= 4 + LAG; loads run SLOTS-1-LAG tiles ahead of the scan.
//   SERIAL the reduce step leaves every thread's total in shared memory and the scan step is ONE
//          running chain per thread seeded with the thread's exclusive prefix: 2 combines per element
//          (reduce + scan) instead of 3.5 -- for operators whose combine is expensive (FP64 adds)
template <int BLOCK_, int VPT_, int ROWS_, int SLOTS_ = 4, int LAG_ = 0, bool SERIAL_ = false> struct Shape {
  static constexpr int BLOCK = BLOCK_, VPT = VPT_, ROWS = ROWS_, SLOTS = SLOTS_, LAG = LAG_;
  static constexpr bool SERIAL = SERIAL_;
  static constexpr int THREADS = BLOCK + WARP;
  static constexpr int AHEAD = SLOTS - 1 - LAG;
  static_assert(BLOCK % WARP == 0 && BLOCK + WARP <= 1024, "");
  static_assert(SLOTS >= 4 + LAG, "");
};

template <class T, class S> struct Layout {
  static constexpr int V = 16 / sizeof(T);
  static constexpr int ITEMS = S::VPT * V;
  static constexpr int TILE = S::BLOCK * ITEMS;
  static constexpr unsigned TILE_BYTES = TILE * sizeof(T);
  static constexpr int NW = S::BLOCK / WARP;
  // [SLOTS tiles][SLOTS mbarriers][3 x 32 warp totals][2 prefixes][vid]
  static constexpr size_t OFF_BAR = (size_t)S::SLOTS * TILE_BYTES;
  static constexpr size_t OFF_WARP = OFF_BAR + 8 * S::SLOTS;
  static constexpr size_t OFF_PREFIX = OFF_WARP + 3 * 8 * 32;
  static constexpr size_t OFF_VID = OFF_PREFIX + 16;
  static constexpr size_t OFF_THREAD = OFF_VID + 8;  // SERIAL: thread totals of tile i at [(i % 3) * BLOCK + tid]
  static constexpr size_t SMEM_BYTES = OFF_THREAD + (S::SERIAL ? 3 * S::BLOCK * sizeof(T) : 0);
};

enum : int { F_BULK_IN = 1, F_BULK_OUT = 2,  // pointer is 16-byte aligned: bulk copies allowed
              F_DBG_NO_LOOKBACK = 4 };        // tools/tune only: timing experiment (wrong results)

template <int COUNT> B200_D void named_barrier_1() { asm volatile("bar.sync 1, %0;" ::"n"(COUNT) : "memory"); }

template <class T, class Op, bool Inclusive, class S>
__global__ void __launch_bounds__(S::THREADS)
scan_kernel(const T* in, T* out, const long long n, const T init, const T* carry_dev, const TileState<T> ts, unsigned* vid_ticket,
            const int flags, unsigned* trace)
{
  using L = Layout<T, S>;
  constexpr int BLOCK = S::BLOCK, VPT = S::VPT, SLOTS = S::SLOTS, ROWS = S::ROWS, AHEAD = S::AHEAD;
  constexpr int V = L::V, ITEMS = L::ITEMS, TILE = L::TILE, NW = L::NW;
  using VT = Vec<T, V>;

  constexpr bool COMMUTATIVE = is_commutative<Op>::value;
  constexpr unsigned SPIN_LIMIT = 1u << 26;  // ~10 s of 64 ns naps plus the poll itself
  extern __shared__ __align__(128) unsigned char smem[];
  T* const tiles = reinterpret_cast<T*>(smem);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  T* const s_warp = reinterpret_cast<T*>(smem + L::OFF_WARP);      // warp totals of tile i at [(i % 3) * 32 + warp]
  T* const s_prefix = reinterpret_cast<T*>(smem + L::OFF_PREFIX);  // exclusive prefix of tile i at [i & 1]
  unsigned* const s_vid = reinterpret_cast<unsigned*>(smem + L::OFF_VID);
  T* const s_thread = reinterpret_cast<T*>(smem + L::OFF_THREAD);
  (void)s_thread;

  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const bool is_lb = (w == NW);           // the look-back warp; its lane 0 also owns every bulk copy
  const bool producer = (tid == BLOCK);   // (bulk groups are per thread)
  Op op;
  const T ident = Op::identity();
  // The seed is folded in front of everything: `init` of an exclusive scan, and/or a carry
  // that lives in device memory (multi-GPU: the fold of the lower ranks' totals).
  T seed = Inclusive ? ident : init;
  if (carry_dev) seed = op(*carry_dev, seed);

  if (tid == 0) {
    *s_vid = atomicAdd(vid_ticket, 1u);
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) mbar_init(bars + s, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const long long G = gridDim.x;
  const long long vid = *s_vid;
  const long long ntiles = (n + TILE - 1) / TILE;
  const long long my_tiles = vid < ntiles ? (ntiles - vid + G - 1) / G : 0;
  if (my_tiles == 0) return;

  // Producer: put tile #i of this CTA into its slot.  Full tiles of an aligned input
  // travel by bulk copy; a ragged or unaligned tile is fetched cooperatively at consume
  // time, and its mbarrier phase is completed by a plain arrive so every slot sees
  // exactly one phase per use.
  auto produce = [&](long long i) {
    const int s = (int)(i % SLOTS);
    const long long base = (vid + i * G) * TILE;
    if ((flags & F_BULK_IN) && base + TILE <= n) {
      fence_proxy_async();
      mbar_arrive_expect_tx(bars + s, L::TILE_BYTES);
      bulk_g2s(tiles + (size_t)s * TILE, in + base, L::TILE_BYTES, bars + s);
    } else {
      mbar_arrive(bars + s);
    }
  };

  // Compute warps: wait for tile #i, reduce it, leave the warp totals in s_warp, publish
  // the tile's aggregate.
  auto reduce_and_publish = [&](long long i) {
    const int s = (int)(i % SLOTS);
    const long long tile = vid + i * G;
    const long long base = tile * TILE;
    T* const buf = tiles + (size_t)s * TILE;
    mbar_wait(bars + s, (unsigned)((i / SLOTS) & 1));
    if (!((flags & F_BULK_IN) && base + TILE <= n)) {
      const long long rem = n - base;
      for (int e = tid; e < TILE; e += BLOCK) buf[e] = (e < rem) ? in[base + e] : ident;
      named_barrier_1<BLOCK>();
    }
    // per-vector sums first (independent chains), then an ordered pairwise tree over the VPT sums:
    // depth (V-1) + ceil(log2 VPT) instead of ITEMS-1 -- a serial fold of 14 doubles is
    // latency-bound on the FP64 pipe (fp64 scans ran at 66 % of the int64 rate)
    T pv[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      VT t = reinterpret_cast<const VT*>(buf)[tid * VPT + v];
      pv[v] = t.e[0];
#pragma unroll
      for (int j = 1; j < V; ++j) pv[v] = op(pv[v], t.e[j]);
    }
#pragma unroll
    for (int st = 1; st < VPT; st <<= 1)
#pragma unroll
      for (int v = 0; v + st < VPT; v += 2 * st) pv[v] = op(pv[v], pv[v + st]);
    T acc = pv[0];
    if constexpr (S::SERIAL) s_thread[(i % 3) * BLOCK + tid] = acc;
    // ordered fold across lanes (lower lane on the left)
#pragma unroll
    for (int d = 1; d < WARP; d <<= 1) {
      T o = shfl_up(acc, d);
      if (l >= d) acc = op(o, acc);
    }
    T* const wt = s_warp + (i % 3) * 32;
    if (l == WARP - 1) wt[w] = acc;
    named_barrier_1<BLOCK>();
    if (tid == 0) {
      T a = ident;
#pragma unroll
      for (int k = 0; k < NW; ++k) a = op(a, wt[k]);
      ts.publish(tile, a);
    }
  };

  // Look-back warp: exclusive prefix of tile #i = base (+) aggregates of the `dist` tiles
  // before it, where base is this CTA's own inclusive prefix two rounds back (or the seed).
  // poll() issues the batch of loads, finish() consumes it (re-polling if a predecessor has
  // not published yet -- rare, they publish two iterations early).
  T lb_val[ROWS];
  bool lb_ok = true;
  auto poll = [&](long long i) {
    const long long tile = vid + i * G;
    const long long span = 2 * G - 1;
    const long long dist = tile < span ? tile : span;
    lb_ok = true;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      // commutative operators: lane-interleaved distances (coalesced polls); otherwise every lane owns a CONTIGUOUS
      // run of predecessors so that the fold below is strictly left to right
      const long long d = COMMUTATIVE ? (long long)r * WARP + l + 1 : (long long)l * ROWS + r + 1;
      lb_val[r] = ident;
      if (d <= dist) lb_ok = ts.peek(tile - d, lb_val[r]) && lb_ok;
    }
  };
  auto finish = [&](long long i, T base_2back) -> T {
    const long long tile = vid + i * G;
    if (flags & F_DBG_NO_LOOKBACK) return seed;
    unsigned polls = 1;
    while (!__all_sync(FULL_MASK, lb_ok)) {
      __nanosleep(64);
      poll(i);
      // a predecessor that never publishes (its CTA never became resident) must end in an error, not a hang
      if (++polls > SPIN_LIMIT) __trap();
    }
    // fixed fold order: per lane farthest row first, then lanes farthest (highest) first
    T c = ident;
#pragma unroll
    for (int r = ROWS - 1; r >= 0; --r) c = op(c, lb_val[r]);
    T rc = shfl_idx(c, WARP - 1 - l);
    T window;
    if constexpr (COMMUTATIVE) window = warp_allreduce(rc, op);
    else window = warp_allreduce_ordered(rc, op);
    if (trace && l == 0) trace[tile] = polls;
    return op((tile >= 2 * G) ? base_2back : seed, window);
  };
  auto block_aggregate = [&](long long i) {
    T a = ident;
#pragma unroll
    for (int k = 0; k < NW; ++k) a = op(a, s_warp[(i % 3) * 32 + k]);
    return a;
  };

  // ---- prologue: AHEAD tiles in flight, aggregates of tiles 0 and 1, prefix of tile 0 ----
  T incl_m1 = ident, incl_0 = ident;  // look-back warp: inclusive prefixes of own tiles i-1 and i
  if (producer) {
    for (long long i = 0; i < AHEAD && i < my_tiles; ++i) produce(i);
  }
  if (!is_lb) {
    reduce_and_publish(0);
    if (my_tiles > 1) reduce_and_publish(1);
  }
  __syncthreads();
  if (is_lb) {
    poll(0);
    const T p0 = finish(0, ident);
    if (l == 0) s_prefix[0] = p0;
    incl_0 = op(p0, block_aggregate(0));
  }
  __syncthreads();

  for (long long i = 0; i < my_tiles; ++i) {
    const int s = (int)(i % SLOTS);
    const long long tile = vid + i * G;
    const long long base = tile * TILE;
    const bool full = (base + TILE <= n);
    const int count = full ? TILE : (int)(n - base);
    T* const buf = tiles + (size_t)s * TILE;
    const bool bulk_out = (flags & F_BULK_OUT) && full;

    if (is_lb) {
      if (i + 1 < my_tiles) poll(i + 1);
      // the slot of tile i-1-LAG is free once its outgoing copy has finished reading it:
      // refill it with the tile AHEAD of the scan
      if (producer) {
        bulk_wait_read<S::LAG>();
        if (i + AHEAD < my_tiles) produce(i + AHEAD);
      }
      if (i + 1 < my_tiles) {
        const T p = finish(i + 1, incl_m1);
        if (l == 0) s_prefix[(i + 1) & 1] = p;
        incl_m1 = incl_0;
        incl_0 = op(p, block_aggregate(i + 1));
      }
    } else {
      if (i + 2 < my_tiles) reduce_and_publish(i + 2);

      // ---- scan tile i with its prefix, write it back into the slot ---------------------------
      // inclusive scan inside each 16-byte vector (independent chains), then a serial scan over the
      // VPT vector totals only: dependent depth (V-1) + (VPT-1) instead of ITEMS-1
      if constexpr (S::SERIAL) {
        T sc = s_thread[(i % 3) * BLOCK + tid];  // own total, left there by reduce_and_publish(i) (same thread)
#pragma unroll
        for (int d = 1; d < WARP; d <<= 1) {
          T o = shfl_up(sc, d);
          if (l >= d) sc = op(o, sc);
        }
        T lane_excl = shfl_up(sc, 1);
        if (l == 0) lane_excl = ident;
        T wprefix = ident;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
          T t = s_warp[(i % 3) * 32 + k];
          if (k < w) wprefix = op(wprefix, t);
        }
        T run = op(op(s_prefix[i & 1], wprefix), lane_excl);
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          VT t = reinterpret_cast<const VT*>(buf)[tid * VPT + v];
#pragma unroll
          for (int j = 0; j < V; ++j) {
            if (Inclusive) {
              run = op(run, t.e[j]);
              t.e[j] = run;
            } else {
              const T e = t.e[j];
              t.e[j] = run;
              run = op(run, e);
            }
          }
          reinterpret_cast<VT*>(buf)[tid * VPT + v] = t;
        }
      } else {
        T x[ITEMS];   // x[v*V + j]: inclusive scan of vector v up to element j
        T vpre[VPT];  // vpre[v]: total of the vectors before v (v >= 1)
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          VT t = reinterpret_cast<const VT*>(buf)[tid * VPT + v];
          x[v * V] = t.e[0];
#pragma unroll
          for (int j = 1; j < V; ++j) x[v * V + j] = op(x[v * V + j - 1], t.e[j]);
        }
        vpre[0] = ident;
        if (VPT > 1) vpre[1] = x[V - 1];
#pragma unroll
        for (int v = 2; v < VPT; ++v) vpre[v] = op(vpre[v - 1], x[(v - 1) * V + V - 1]);
        T sc = (VPT > 1) ? op(vpre[VPT - 1], x[ITEMS - 1]) : x[ITEMS - 1];
#pragma unroll
        for (int d = 1; d < WARP; d <<= 1) {
          T o = shfl_up(sc, d);
          if (l >= d) sc = op(o, sc);
        }
        T lane_excl = shfl_up(sc, 1);
        if (l == 0) lane_excl = ident;
        T wprefix = ident;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
          T t = s_warp[(i % 3) * 32 + k];
          if (k < w) wprefix = op(wprefix, t);
        }
        const T pre = op(op(s_prefix[i & 1], wprefix), lane_excl);
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          const T pre_v = (v == 0) ? pre : op(pre, vpre[v]);  // everything before this vector
          VT t;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            if (Inclusive) t.e[j] = op(pre_v, x[v * V + j]);
            else t.e[j] = (j == 0) ? pre_v : op(pre_v, x[v * V + j - 1]);
          }
          reinterpret_cast<VT*>(buf)[tid * VPT + v] = t;
        }
      }
      if (bulk_out) {
        fence_proxy_async();  // generic-proxy writes above -> async-proxy read by the bulk store
      } else {
        named_barrier_1<BLOCK>();
        for (int e = tid; e < count; e += BLOCK) out[base + e] = buf[e];
      }
    }
    __syncthreads();  // tile i is complete in its slot; prefix(i+1) is in s_prefix
    if (producer && bulk_out) {
      bulk_s2g(out + base, buf, L::TILE_BYTES);
      bulk_commit();
    }
  }
  if (producer) bulk_wait_all();  // shared memory must outlive the last outgoing copy
}

// Workspace: [virtual-id ticket (128 B)] [tile state].  Zeroed here, stream-ordered.
template <class T, class S> inline long long num_tiles(long long n) { return (n + Layout<T, S>::TILE - 1) / Layout<T, S>::TILE; }
template <class T, class S> inline size_t workspace_bytes(long long n) { return 128 + TileState<T>::bytes(num_tiles<T, S>(n)); }

// Default shapes (tools/tune/tune_scan.cu sweeps them on the GPU box)
template <class T> struct Tuning {
  // 384 compute threads x 7 vectors = 42 KB tiles, 5 slots = 210 KB: one CTA per SM, G = 148,
  // one look-back batch = 2G-1 = 295 <= 320 aggregates.  Measured 6.17 TB/s (int32) and
  // 6.32 TB/s (int64) at N = 2^28, 96 % of a device memcpy (profiles/r01_scan_tuning.txt).
  using shape = Shape<384, 7, 10, 5, 0>;
};
// FP64 scans are bound by the NUMBER of DADDs issued, not by their latency or by bandwidth: the
// ILP-friendly scan step above costs 2.5 combines per element on top of the reduce step's 1, and
// the same kernel with an integer add on the bit patterns runs at the int64 rate.  SERIAL (thread
// totals kept in shared memory, one running chain per thread: 2 combines per element) with the
// largest tile that leaves four slots: 5.89 TB/s = 90 % of the measured peak, against 5.27 TB/s
// for the same tile without SERIAL and 4.49 TB/s for the integer shape (profiles/r01_scan_tuning_v4.txt).
template <> struct Tuning<double> {
  using shape = Shape<384, 9, 10, 4, 0, true>;
};

template <class T, class Op, bool Inclusive, class S>
inline cudaError_t configure(int* ctas_per_sm)
{
  auto kern = scan_kernel<T, Op, Inclusive, S>;
  static thread_local PerDevice cache;  // the opt-in and the occupancy are per (kernel, device)
  int& cached = cache.current();
  if (cached == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Layout<T, S>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, S::THREADS, Layout<T, S>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cached = occ > 0 ? occ : 1;
  }
  *ctas_per_sm = cached;
  return cudaSuccess;
}

// Enqueue one scan on `stream`.  `workspace` must hold workspace_bytes<T,S>(n) bytes of device memory.
template <class T, class Op, bool Inclusive, class S = typename Tuning<T>::shape>
inline cudaError_t launch(cudaStream_t stream, const T* in, T* out, long long n, T init, void* workspace,
                          int flags_clear = 0, int flags_set = 0, unsigned* trace = nullptr, const T* carry_dev = nullptr)
{
  if (n <= 0) return cudaSuccess;
  int occ = 1;
  cudaError_t e = configure<T, Op, Inclusive, S>(&occ);
  if (e != cudaSuccess) return e;
  const long long ntiles = num_tiles<T, S>(n);
  e = cudaMemsetAsync(workspace, 0, workspace_bytes<T, S>(n), stream);
  if (e != cudaSuccess) return e;
  TileState<T> ts;
  ts.bind((char*)workspace + 128, ntiles);
  long long grid = (long long)device_info().sm_count * occ;
  const long long cap = (S::ROWS * WARP + 1) / 2;  // one batch of look-back loads spans two rounds of the grid
  if (grid > cap) grid = cap;
  if (grid > ntiles) grid = ntiles;
  int flags = 0;
  if (((uintptr_t)in & 15u) == 0) flags |= F_BULK_IN;
  if (((uintptr_t)out & 15u) == 0) flags |= F_BULK_OUT;
  flags = (flags & ~flags_clear) | flags_set;
  LaunchScope scope("scan_kernel", stream);
  if (ntiles > grid) {
    // A CTA that owns several tiles (vid, vid + G, ...) looks back at tiles of CTAs with HIGHER ticket numbers, so
    // every CTA of the grid must be resident: a cooperative launch is gang-scheduled (it starts only when the whole
    // grid fits, e.g. after another scan on another stream has drained) and fails instead of hanging when it cannot.
    unsigned* ticket = (unsigned*)workspace;
    void* args[] = {(void*)&in, (void*)&out, (void*)&n, (void*)&init, (void*)&carry_dev, (void*)&ts, (void*)&ticket, (void*)&flags, (void*)&trace};
    const void* kern = (const void*)scan_kernel<T, Op, Inclusive, S>;
    e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(S::THREADS), args, Layout<T, S>::SMEM_BYTES, stream);
    if (e != cudaErrorCooperativeLaunchTooLarge) return e;
    // The driver finds fewer co-resident CTAs than the cached occupancy promised (another context or runtime instance holds
    // part of the shared-memory carve-out, a smaller MPS partition, ...): ask again, now, and shrink the grid -- any grid
    // works, a CTA simply owns more tiles.  The workspace is still zero: the refused launch never ran.
    cudaGetLastError();
    int occ_now = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_now, scan_kernel<T, Op, Inclusive, S>, S::THREADS, Layout<T, S>::SMEM_BYTES) == cudaSuccess) {
      const long long g = (long long)device_info().sm_count * occ_now;
      if (occ_now > 0 && g < grid) {
        e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)g), dim3(S::THREADS), args, Layout<T, S>::SMEM_BYTES, stream);
        if (e != cudaErrorCooperativeLaunchTooLarge) return e;
        cudaGetLastError();
      }
    }
    // no cooperative launch possible on this device / context although the occupancy query says the grid fits: a plain launch, which is safe whenever the grid does
    // become fully resident (always on an otherwise idle GPU); the bounded look-back spin turns the other case into an error
    scan_kernel<T, Op, Inclusive, S><<<(unsigned)grid, S::THREADS, Layout<T, S>::SMEM_BYTES, stream>>>(in, out, n, init, carry_dev, ts, (unsigned*)workspace, flags,
                                                                                                        trace);
    return cudaGetLastError();
  }
  // one tile per CTA: a tile only looks back at lower tickets, i.e. at CTAs that have already started
  scan_kernel<T, Op, Inclusive, S><<<(unsigned)grid, S::THREADS, Layout<T, S>::SMEM_BYTES, stream>>>(in, out, n, init, carry_dev, ts,
                                                                                                      (unsigned*)workspace, flags, trace);
  return cudaGetLastError();
}

}  // namespace scan
}  // namespace b200
