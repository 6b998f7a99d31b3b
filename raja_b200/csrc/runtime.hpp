// raja_b200 runtime internals shared by the capi_*.cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <string>
#include "../../include/raja_b200.h"

namespace rb200 {

void set_error(const std::string& msg);
int fail(int code, const char* what);
int cuda_fail(cudaError_t e, const char* what);

enum PoolKind { POOL_DEVICE = 0, POOL_DEVICE_ZEROED = 1, POOL_PINNED = 2 };
void* pool_malloc(PoolKind kind, size_t bytes, cudaStream_t stream);
int pool_free(PoolKind kind, void* p, cudaStream_t stream);

// RAII scratch from the device pool, returned (stream-ordered) on scope exit.
struct Scratch {
  void* p = nullptr;
  cudaStream_t stream;
  PoolKind kind;
  Scratch(PoolKind k, size_t bytes, cudaStream_t s) : stream(s), kind(k) { p = pool_malloc(k, bytes, s); }
  ~Scratch() { if (p) pool_free(kind, p, stream); }
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
};

inline size_t dtype_size(int t)
{
  switch (t) {
    case RB200_I32: case RB200_U32: case RB200_F32: return 4;
    case RB200_I64: case RB200_U64: case RB200_F64: return 8;
    default: return 0;
  }
}

#define RB200_CUDA(call)                                   \
  do {                                                     \
    cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) return ::rb200::cuda_fail(e_, #call); \
  } while (0)

}  // namespace rb200
