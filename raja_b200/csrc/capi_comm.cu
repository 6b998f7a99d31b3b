// raja_b200 C ABI: peer-memory communicator over NVLink / NVSwitch.
//
// The sharded-range paths exchange very little data except for the sorts: one scalar per rank
// (reductions), one shard total per rank (scan carries), an 8 x world count matrix (sorts).  A NCCL
// collective costs ~100-300 us end to end for such a message (launch + proxy + torch dispatch);
// here every rank stores its payload straight into its peers' memory and spins on the peers'
// flags in its own memory: one 1-CTA kernel, a few microseconds, stream ordered, and the result is
// gathered in RANK ORDER, so folds over it are deterministic.
#include "comm.hpp"
#include "b200/common.cuh"

using namespace rb200;

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// slot of `src_rank` in the workspace of its owner, for the epoch's parity
__device__ __forceinline__ unsigned long long* slot_of(char* ws, unsigned long long epoch, int src_rank)
{
  return reinterpret_cast<unsigned long long*>(ws + ((epoch & 1) * COMM_MAX_WORLD + src_rank) * COMM_SLOT_BYTES);
}

// One CTA.  Phase 1: payload -> every rank's slot[parity][my rank] (plain stores over NVLink), system
// fence, then the slot's flag = epoch with release semantics.  Phase 2: wait for the world flags in
// MY workspace, copy the payloads out in rank order.  Slots are double-buffered by epoch parity: a
// rank can only be two epochs ahead of another after that one has read the older slot.
// fold of element e over the ranks [0, hi) of the gathered payloads, in rank order (deterministic)
template <class T> __device__ __forceinline__ T fold_ranks(const CommDev& c, unsigned long long epoch, int e, int hi, int op)
{
  T acc = op == 0 ? T(0) : (op == 1 ? b200::limits<T>::max() : b200::limits<T>::min());
  for (int p = 0; p < hi; ++p) {
    const T v = reinterpret_cast<const T*>(slot_of(c.ws[c.rank], epoch, p))[e];
    acc = op == 0 ? (T)(acc + v) : (op == 1 ? (v < acc ? v : acc) : (acc < v ? v : acc));
  }
  return acc;
}

// fold == 0: plain all-gather into dst[world][words].  fold == 1: dst[count] = fold over all ranks (all-reduce);
// fold == 2: dst[count] = fold over the ranks below this one (exclusive prefix: the carry of a sharded scan).
__global__ void __launch_bounds__(256) peer_allgather_kernel(const CommDev c, const unsigned long long* __restrict__ src, const int words,
                                                             unsigned long long* __restrict__ dst, const unsigned long long epoch,
                                                             const int fold, const int op, const int dtype, const int count, const int half_tail)
{
  const int tid = threadIdx.x;
  for (int i = tid; i < c.world * words; i += blockDim.x) {
    const int p = i / words, k = i - p * words;
    // `words` 8-byte words; a payload that ends in half a word (an odd number of 4-byte elements) is read as 32 bits there
    const unsigned long long v = (half_tail && k == words - 1) ? (unsigned long long)reinterpret_cast<const unsigned*>(src)[2 * k] : src[k];
    slot_of(c.ws[p], epoch, c.rank)[k] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < c.world) st_release_sys(slot_of(c.ws[tid], epoch, c.rank) + COMM_PAYLOAD_MAX / 8, epoch);
  if (tid < c.world) {
    const unsigned long long* flag = slot_of(c.ws[c.rank], epoch, tid) + COMM_PAYLOAD_MAX / 8;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != epoch) {
      __nanosleep(200);
      if (clock64() - t0 > 20000000000ll) __trap();  // ~10 s: a peer never arrived; fail instead of hanging the box
    }
  }
  __syncthreads();
  if (fold == 0) {
    for (int i = tid; i < c.world * words; i += blockDim.x) {
      const int p = i / words, k = i - p * words;
      dst[i] = slot_of(c.ws[c.rank], epoch, p)[k];
    }
  } else {
    const int hi = fold == 1 ? c.world : c.rank;
    for (int e = tid; e < count; e += blockDim.x) {
      switch (dtype) {
        case RB200_I32: reinterpret_cast<int32_t*>(dst)[e] = fold_ranks<int32_t>(c, epoch, e, hi, op); break;
        case RB200_U32: reinterpret_cast<uint32_t*>(dst)[e] = fold_ranks<uint32_t>(c, epoch, e, hi, op); break;
        case RB200_I64: reinterpret_cast<int64_t*>(dst)[e] = fold_ranks<int64_t>(c, epoch, e, hi, op); break;
        case RB200_U64: reinterpret_cast<uint64_t*>(dst)[e] = fold_ranks<uint64_t>(c, epoch, e, hi, op); break;
        case RB200_F32: reinterpret_cast<float*>(dst)[e] = fold_ranks<float>(c, epoch, e, hi, op); break;
        default: reinterpret_cast<double*>(dst)[e] = fold_ranks<double>(c, epoch, e, hi, op); break;
      }
    }
  }
}

}  // namespace

namespace rb200 {

static int launch_allgather(rb200_comm* c, const void* src_dev, int bytes, void* dst_dev, int fold, int op, int dtype, int count, cudaStream_t st)
{
  const CommDev d = comm_dev(c);
  const unsigned long long epoch = ++c->epoch;
  {
    b200::LaunchScope scope("peer_allgather_kernel", st);
    peer_allgather_kernel<<<1, 256, 0, st>>>(d, (const unsigned long long*)src_dev, (bytes + 7) / 8, (unsigned long long*)dst_dev, epoch, fold, op, dtype, count,
                                             (bytes & 7) ? 1 : 0);
  }
  RB200_CUDA(cudaGetLastError());
  return 0;
}

int comm_allgather_small(rb200_comm* c, const void* src_dev, int bytes, void* dst_dev, cudaStream_t st)
{
  if (!c || !src_dev || !dst_dev || bytes <= 0 || bytes > COMM_PAYLOAD_MAX || (bytes & 7))
    return fail(RB200_EINVAL, "rb200_comm_allgather_small: bytes must be a multiple of 8 in (0, 504]");
  return launch_allgather(c, src_dev, bytes, dst_dev, 0, 0, 0, 0, st);
}

}  // namespace rb200

extern "C" {

int rb200_comm_create(int rank, int world, void* const* ws_ptrs, size_t ws_bytes, rb200_comm_t* out)
{
  if (!out || !ws_ptrs || world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || ws_bytes < COMM_HEADER_BYTES)
    return fail(RB200_EINVAL, "rb200_comm_create: bad arguments (1 <= world <= 8, workspace >= rb200_comm_header_bytes())");
  rb200_comm* c = new rb200_comm();
  c->rank = rank;
  c->world = world;
  c->ws_bytes = ws_bytes;
  for (int i = 0; i < world; ++i) {
    if (!ws_ptrs[i]) {
      delete c;
      return fail(RB200_EINVAL, "rb200_comm_create: null workspace pointer");
    }
    c->ws[i] = (char*)ws_ptrs[i];
  }
  // the slots' flags start at 0 and epochs at 1; the caller synchronises the ranks after this call
  cudaError_t e = cudaMemset(c->ws[rank], 0, COMM_HEADER_BYTES);
  if (e != cudaSuccess) {
    delete c;
    return cuda_fail(e, "cudaMemset(comm header)");
  }
  // Pinned staging blocks (count matrices, reducer results: all in the smallest size class) are taken from the pinned pool
  // between collectives; a pool miss there is a cudaHostAlloc, which synchronises with every device of the process -- and a
  // peer device may be spinning in an exchange kernel that waits for this rank (ranks as threads of one process).  Stock the
  // pool now, while the ranks are still being set up (the caller synchronises them after this call).
  {
    void* warm[8];
    for (void*& w : warm) w = rb200::pool_malloc(rb200::POOL_PINNED, 512, nullptr);
    for (void* w : warm)
      if (w) rb200::pool_free(rb200::POOL_PINNED, w, nullptr);
  }
  *out = c;
  return 0;
}

int rb200_comm_destroy(rb200_comm_t c)
{
  delete c;
  return 0;
}

size_t rb200_comm_header_bytes(void) { return COMM_HEADER_BYTES; }

int rb200_comm_allgather_small(rb200_comm_t c, const void* src_dev, int bytes, void* dst_dev, rb200_stream_t stream)
{
  return comm_allgather_small(c, src_dev, bytes, dst_dev, (cudaStream_t)stream);
}

int rb200_comm_barrier(rb200_comm_t c, rb200_stream_t stream)
{
  if (!c) return fail(RB200_EINVAL, "rb200_comm_barrier: null communicator");
  // an 8-byte all-gather into the tail of the own header (bytes 32 KiB .. of the workspace are scratch)
  char* scratch = c->ws[c->rank] + COMM_BARRIER_OFFSET;
  return comm_allgather_small(c, scratch, 8, scratch + 64, (cudaStream_t)stream);
}

int rb200_comm_allreduce_small(rb200_comm_t c, rb200_op op, rb200_dtype t, const void* src_dev, void* dst_dev, int count, int exclusive_prefix,
                               rb200_stream_t stream)
{
  const size_t es = dtype_size(t);
  if (!c || !src_dev || !dst_dev || es == 0 || count <= 0 || (size_t)count * es > COMM_PAYLOAD_MAX || (int)op < 0 || (int)op > 2)
    return fail(RB200_EINVAL, "rb200_comm_allreduce_small: bad arguments (count * sizeof(dtype) must be in (0, 504])");
  if ((uintptr_t)src_dev & 7) return fail(RB200_EINVAL, "rb200_comm_allreduce_small: src_dev must be 8-byte aligned");
  const int bytes = (int)((size_t)count * es);  // travels as 8-byte words; a trailing half word is read as 32 bits
  return launch_allgather(c, src_dev, bytes, dst_dev, exclusive_prefix ? 2 : 1, (int)op, (int)t, count, (cudaStream_t)stream);
}

}  // extern "C"
