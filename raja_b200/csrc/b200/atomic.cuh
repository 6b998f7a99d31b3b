// raja_b200 core: device atomics.
//
// Replaces the reference's policy/cuda/atomic.hpp:328-402 (built-in vs CAS-loop
// selection) for the add path and adds the one lever an opaque
// atomicAdd(policy, T*, T) call has on contended addresses: warp aggregation.
// Lanes that target the same address are found with match.any; one leader issues
// a single atomic with the group's sum and the old value is redistributed so the
// call still returns what a per-lane atomic would have (fetch-and-add semantics,
// pattern/atomic.hpp:90-118).
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace b200 {

// ---- plain atomics for every arithmetic type the reference supports ------------
B200_D int32_t atomic_add(int32_t* a, int32_t v) { return atomicAdd(a, v); }
B200_D uint32_t atomic_add(uint32_t* a, uint32_t v) { return atomicAdd(a, v); }
B200_D unsigned long long atomic_add(unsigned long long* a, unsigned long long v) { return atomicAdd(a, v); }
B200_D unsigned long atomic_add(unsigned long* a, unsigned long v)
{
  return (unsigned long)atomicAdd((unsigned long long*)a, (unsigned long long)v);
}
B200_D long long atomic_add(long long* a, long long v)
{
  return (long long)atomicAdd((unsigned long long*)a, (unsigned long long)v);
}
B200_D long atomic_add(long* a, long v) { return (long)atomicAdd((unsigned long long*)a, (unsigned long long)v); }
B200_D float atomic_add(float* a, float v) { return atomicAdd(a, v); }
B200_D double atomic_add(double* a, double v) { return atomicAdd(a, v); }

// ---- warp-aggregated fetch-and-add ---------------------------------------------
template <class T> B200_D T atomic_add_aggregated(T* addr, T v)
{
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, (unsigned long long)addr);
  const int lane = lane_id();
  const int npeers = __popc(peers);
  // fast exit: nobody shares an address
  if (__all_sync(active, npeers == 1)) return atomic_add(addr, v);

  const int leader = __ffs(peers) - 1;
  // every active lane runs the same trip count so the shuffles stay convergent
  const int iters = __reduce_max_sync(active, npeers);
  T total = T(0), before = T(0);
  unsigned rem = peers;
  for (int it = 0; it < iters; ++it) {
    const int src = rem ? (__ffs(rem) - 1) : lane;
    const T o = shfl_idx_mask(active, v, src);
    if (rem) {
      total += o;
      if (src < lane) before += o;
      rem &= rem - 1;
    }
  }
  T old = T(0);
  if (lane == leader) old = atomic_add(addr, total);
  old = shfl_idx_mask(active, old, leader);
  return old + before;
}

}  // namespace b200
