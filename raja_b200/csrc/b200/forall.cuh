// raja_b200 core: the forall engine.
//
// Replaces the reference's four forallp_cuda_kernel variants and their
// dimension calculators (policy/cuda/forall.hpp:73-353, 383-542;
// MemUtils_CUDA.hpp:527-613) with ONE mapping: a persistent grid
// (SMs x resident CTAs) that walks tiles of BLOCK*UNROLL indices round-robin.
// Inside a tile the UNROLL bodies of a thread are BLOCK apart, so every load a
// body issues is a fully coalesced warp access and UNROLL of them are in flight
// per thread (B200 needs ~30-40 KB outstanding per SM to saturate HBM3e).
//
// Body is copied into each thread (the reference's thread_privatize,
// pattern/detail/privatizer.hpp:45-99): it may carry per-thread state such as
// reduction accumulators.  After the loop `fin(body)` runs once per thread; that
// is where block/grid reductions happen.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace b200 {

struct NoFinalize {
  template <class Body> B200_D void operator()(Body&) const {}
};

// Optional body hooks, detected at compile time:
//   template <int U> void tile(Index base, Index stride)
//        -- process indices base + u*stride, u in [0,U), in one call, so the body
//           can issue all of its loads before its first store (an opaque
//           operator() cannot be reordered across stores that may alias);
//   Index tail_len() const / void tail(Index j)
//        -- j in [0, tail_len()): scalar leftovers of a vectorised index space,
//           run once by one block after the main loop.
namespace detail {
template <class Body, class Index, int U, class = void> struct has_tile : std::false_type {};
template <class Body, class Index, int U>
struct has_tile<Body, Index, U, std::void_t<decltype(std::declval<Body&>().template tile<U>(std::declval<Index>(), std::declval<Index>()))>>
    : std::true_type {};
template <class Body, class Index, class = void> struct has_tail : std::false_type {};
template <class Body, class Index>
struct has_tail<Body, Index, std::void_t<decltype(std::declval<Body&>().tail(std::declval<Index>()))>> : std::true_type {};
}  // namespace detail

template <int BLOCK, int UNROLL, class Index, class Body, class Finalize>
__global__ void __launch_bounds__(BLOCK)
forall_kernel(const __grid_constant__ Body body_in, const Index len, const __grid_constant__ Finalize fin, const Index full_tiles,
              const unsigned tail_owner)
{
  // full_tiles = len / TILE and tail_owner = full_tiles % gridDim.x come from the host: a 64-bit modulo by a run-time
  // value is a ~100-instruction subroutine, and a loop of a few hundred indices then spends its time in instruction
  // fetches (the kernel of a small forall took 4.1 us against 2.8 us for the reference's one-element-per-thread kernel)
  Body body = body_in;
  constexpr Index TILE = Index(BLOCK) * UNROLL;
  if constexpr (detail::has_tile<Body, Index, UNROLL>::value) {
    // hand-written tile bodies (the streaming kernels of the C ABI): the compiler may software-pipeline across tiles;
    // the ragged last tile stays on one block (spreading it, as below, makes ptxas 12.9 settle on 32 registers and
    // split the tile's eight loads into two groups: -2 % on triad / daxpy at 2^28)
    for (Index t = blockIdx.x; t < full_tiles; t += gridDim.x) body.template tile<UNROLL>(t * TILE + threadIdx.x, Index(BLOCK));
    if (blockIdx.x == tail_owner) {
      for (Index i = full_tiles * TILE + threadIdx.x; i < len; i += BLOCK) body(i);
    }
  } else {
    // opaque bodies (user lambdas): the tile is already UNROLL-wide; unrolling this loop again costs a trip-count
    // division and 4x the code, which a forall of a few thousand indices pays for in instruction fetches
#pragma unroll 1
    for (Index t = blockIdx.x; t < full_tiles; t += gridDim.x) {
      const Index base = t * TILE + threadIdx.x;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) body(base + Index(u) * BLOCK);
    }
    // ragged last tile (< TILE indices): its UNROLL chunks of BLOCK indices go to UNROLL different blocks, starting with
    // the block that would own the next tile, so that a short loop costs ONE memory round trip, not UNROLL in sequence
    const Index rem_base = full_tiles * TILE;
    const unsigned first = blockIdx.x >= tail_owner ? blockIdx.x - tail_owner : blockIdx.x + gridDim.x - tail_owner;
#pragma unroll 1
    for (unsigned c = first; c < (unsigned)UNROLL; c += gridDim.x) {  // chunk c belongs to block (tail_owner + c) mod grid
      const Index i = rem_base + Index(c) * BLOCK + threadIdx.x;
      if (i < len) body(i);
    }
  }
  if constexpr (detail::has_tail<Body, Index>::value) {
    if (blockIdx.x == gridDim.x - 1) {
      for (Index j = threadIdx.x; j < body.tail_len(); j += BLOCK) body.tail(j);
    }
  }
  fin(body);
}

// Grid size for the persistent mapping: every SM gets its resident CTAs, never
// more blocks than tiles.
// `waves` > 1 oversubscribes the device: the hardware block scheduler then hands a fresh CTA to whichever SM frees a slot,
// which balances loop bodies whose cost is uneven or contended (atomics: a histogram of 2^28 keys into 64 K bins runs at
// 1.93 ms with one wave at any UNROLL and any number of resident CTAs, 1.77 ms with four waves, 1.708 ms with 32 waves
// (256 CTAs per SM), 1.703 ms with the reference's one block per 256 indices) and costs streaming bodies nothing (triad
// 6.60-6.63 TB/s from 1 to 8 waves): profiles/r02_forall_grid_sweep.txt.
inline int persistent_grid(int occ_per_sm, long long tiles, int waves = 1)
{
  long long g = (long long)device_info().sm_count * (occ_per_sm > 0 ? occ_per_sm : 1) * (waves > 0 ? waves : 1);
  if (tiles < 1) tiles = 1;
  if (g > tiles) g = tiles;
  return (int)g;
}

template <int BLOCK, int UNROLL, class Index, class Body, class Finalize>
inline int forall_occupancy()
{
  // occupancy is a property of the instantiation and the device; query once per host thread and device
  static thread_local PerDevice cache;
  int& cached = cache.current();
  if (cached == 0) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, forall_kernel<BLOCK, UNROLL, Index, Body, Finalize>, BLOCK, 0);
    cached = occ > 0 ? occ : 1;
  }
  return cached;
}

template <int BLOCK, int UNROLL, class Index, class Body, class Finalize>
inline int forall_grid(Index len, long long fixed_grid = 0, int waves = 1)
{
  // full tiles + the BLOCK-sized chunks of the ragged last tile (each chunk may go to a block of its own)
  const long long tile = (long long)BLOCK * UNROLL, full = (long long)len / tile;
  const long long units = full + ((long long)len - full * tile + BLOCK - 1) / BLOCK;
  if (fixed_grid > 0) return (int)(fixed_grid < units ? fixed_grid : (units < 1 ? 1 : units));  // a grid the caller fixed
  return persistent_grid(forall_occupancy<BLOCK, UNROLL, Index, Body, Finalize>(), units, waves);
}

template <int BLOCK, int UNROLL, class Index, class Body, class Finalize>
inline cudaError_t launch_forall(cudaStream_t stream, const Body& body, Index len, const Finalize& fin, int grid,
                                 const char* name = "forall_kernel", size_t dyn_smem = 0)
{
  LaunchScope scope(name, stream);
  if (len < 0) return cudaErrorInvalidValue;  // len == 0 still launches: the body may own tail work
  // dyn_smem (<= 48 KB): shared memory objects captured in the body asked for (multi-reducer bins)
  if (grid < 1) grid = 1;
  const Index full_tiles = len / (Index(BLOCK) * UNROLL);
  const unsigned tail_owner = (unsigned)(full_tiles % Index(grid));
  forall_kernel<BLOCK, UNROLL, Index, Body, Finalize><<<grid, BLOCK, dyn_smem, stream>>>(body, len, fin, full_tiles, tail_owner);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Reduction finalizer: folds the per-thread accumulator `body.acc` across the
// block and the grid and lets the last block write the result.
//   ordered == true : last-block-done, fixed slot order (deterministic)
//   ordered == false: one atomic per block on *out (out must hold the identity
//                     or a running value beforehand)
// ---------------------------------------------------------------------------
template <int BLOCK, class Op, bool Ordered = true> struct ReduceFinalize {
  using T = typename Op::value_type;
  T* partials;        // [grid]
  unsigned* counter;  // zero on entry, zero on exit
  T* out;             // device-accessible result slot
  template <class Body> B200_D void operator()(Body& body) const
  {
    __shared__ T smem[32];
    Op op;
    T v = block_reduce<BLOCK>(body.acc, op, smem);
    if constexpr (Ordered) {
      T r;
      if (grid_combine_last_block<BLOCK>(v, op, partials, counter, smem, r)) {
        if (threadIdx.x == 0) *out = r;
      }
    } else {
      if (threadIdx.x == 0) atomic_combine(op, out, v);
    }
  }
};

}  // namespace b200
