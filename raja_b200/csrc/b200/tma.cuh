// raja_b200 core: bulk asynchronous copies (the TMA engine's 1-D "bulk" mode) and the
// mbarrier objects that track them.  sm_90+/sm_100a PTX; SASS shows UBLKCP / SYNCS.
//
// Proxy rules the callers follow:
//   * a bulk copy global->shared writes shared memory through the ASYNC proxy; its
//     completion is observed by waiting on the mbarrier it was given -- after that the
//     data may be read with ordinary (generic-proxy) loads;
//   * before a bulk copy shared->global reads data that was written with ordinary
//     stores, every writing thread executes fence_proxy_async() and the CTA
//     synchronises;
//   * a slot may be overwritten once bulk_wait_read<N>() says the outgoing copies that
//     read it are done with shared memory (groups are per issuing thread).
#pragma once
#include "common.cuh"

namespace b200 {

B200_D uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

B200_D void mbar_init(uint64_t* bar, unsigned arrivals)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
B200_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
B200_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

B200_D void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
B200_D void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// Blocks until the phase with the given parity has completed (try_wait suspends the
// warp in hardware for a bounded time; the loop re-arms it).
B200_D void mbar_wait(uint64_t* bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
B200_D void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// shared -> global (joins the issuing thread's current bulk group)
B200_D void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes)
               : "memory");
}
B200_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING shared memory
template <int N> B200_D void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// every bulk group of this thread has fully completed
B200_D void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace b200
