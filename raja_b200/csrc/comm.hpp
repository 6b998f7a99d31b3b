// raja_b200 peer-memory communicator (internal): one symmetric workspace per rank, mapped into
// every process of the node (CUDA IPC / fabric handles; the caller does the exchange, e.g. with
// torch symmetric memory).  The first COMM_HEADER_BYTES of each workspace hold the slots of the
// small all-gather; the rest is the caller's (the receive buffers of the sorts live there).
#pragma once
#include "runtime.hpp"

struct rb200_comm {
  int rank = 0, world = 1;
  char* ws[8] = {};
  size_t ws_bytes = 0;
  unsigned long long epoch = 0;         // number of small all-gathers issued so far (identical on every rank)
  unsigned long long sample_epoch = 0;  // number of sample all-gathers issued so far
  unsigned split_calls = 0;             // rotates the __constant__ splitter slots of the sample sort
};

namespace rb200 {
constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_SLOT_BYTES = 512;     // 504 B of payload + the 8-byte epoch flag
constexpr int COMM_PAYLOAD_MAX = 504;
// header layout: [0, 8 KB) slots of the small all-gather (2 parities x 8 ranks x 512 B) | [32 KB, +128 B) barrier scratch |
// [64 KB, 64 KB + 2 x 8 x 40 KB) sample slots of the sort (parity x source rank: flag, count, <= 4096 samples of 8 bytes)
constexpr size_t COMM_HEADER_BYTES = 1 << 20;
constexpr size_t COMM_BARRIER_OFFSET = 32768;
constexpr size_t COMM_SAMPLE_OFFSET = 65536;
constexpr size_t COMM_SAMPLE_SLOT_BYTES = 40960;
constexpr int COMM_MAX_SAMPLES = 4096;       // per rank
// enqueue one small all-gather on `st`: dst_dev[r*bytes .. ) = rank r's src_dev[0 .. bytes)
int comm_allgather_small(rb200_comm* c, const void* src_dev, int bytes, void* dst_dev, cudaStream_t st);
// device view of a communicator for kernels in other translation units
struct CommDev {
  char* ws[COMM_MAX_WORLD];
  int rank, world;
};
inline CommDev comm_dev(const rb200_comm* c)
{
  CommDev d;
  for (int i = 0; i < COMM_MAX_WORLD; ++i) d.ws[i] = c->ws[i];
  d.rank = c->rank;
  d.world = c->world;
  return d;
}
}  // namespace rb200
