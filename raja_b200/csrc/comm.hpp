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
  unsigned long long epoch = 0;  // number of small all-gathers issued so far (identical on every rank)
};

namespace rb200 {
constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_SLOT_BYTES = 512;     // 504 B of payload + the 8-byte epoch flag
constexpr int COMM_PAYLOAD_MAX = 504;
constexpr size_t COMM_HEADER_BYTES = 65536;  // 2 parities x 8 ranks x 512 B = 8 KB used
// enqueue one small all-gather on `st`: dst_dev[r*bytes .. ) = rank r's src_dev[0 .. bytes)
int comm_allgather_small(rb200_comm* c, const void* src_dev, int bytes, void* dst_dev, cudaStream_t st);
}  // namespace rb200
