// raja_b200 C ABI: sorts (type-erased front for b200/radix_sort.cuh).
#include "runtime.hpp"
#include "comm.hpp"
#include "b200/radix_sort.cuh"

namespace b200 {
namespace rsort {
// The same split with splitters that were COMPUTED ON THE DEVICE (the multi-GPU sample sort selects them from the gathered
// samples without a round trip to the host): they live in a __constant__ array that a stream-ordered device-to-device
// cudaMemcpyToSymbolAsync fills before the kernels that use it, so the compares still cost no registers and no
// shared-memory traffic.  CONST_SPLIT_SLOTS slots are used round-robin, so up to that many splits may be in flight on
// different streams of one device (one per call is the normal case).  Entries are stored as 64-bit words, already in the
// order-preserving encoding (to_ordered); unused entries are all-ones.
constexpr int CONST_SPLIT_SLOTS = 4;
__constant__ unsigned long long c_splitters[CONST_SPLIT_SLOTS][MAX_SPLITTERS + 1];
template <class U> struct ConstSplitDigit {
  static constexpr bool PEER_OUT = false;
  static constexpr int BUCKETS = 8;
  static constexpr bool ENCODE_ON_LOAD = false, DECODE_ON_STORE = false;
  B200_D bool identity_pass() const { return false; }
  int slot;
  int nsplit;
  KeyOrder order;
  B200_D U splitter(int i) const { return (U)c_splitters[slot][i]; }
  B200_D unsigned operator()(U k) const
  {
    const U ok = to_ordered(k, order);
    unsigned d = 0;
#pragma unroll
    for (int i = 0; i < MAX_SPLITTERS; ++i)
      if (i < nsplit) d += (ok >= (U)c_splitters[slot][i]) ? 1u : 0u;
    return d;
  }
};

// stream-ordered fill of one splitter slot from DEVICE memory (MAX_SPLITTERS + 1 u64 words)
inline cudaError_t set_const_splitters(cudaStream_t stream, int slot, const unsigned long long* splitters_dev)
{
  return cudaMemcpyToSymbolAsync(c_splitters, splitters_dev, (MAX_SPLITTERS + 1) * sizeof(unsigned long long),
                                 (size_t)slot * (MAX_SPLITTERS + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream);
}

}  // namespace rsort
}  // namespace b200

using namespace b200;
using namespace rb200;

namespace {

rsort::KeyOrder order_of(int key_t, int descending)
{
  int kind = rsort::KEY_UNSIGNED;
  if (key_t == RB200_I32 || key_t == RB200_I64) kind = rsort::KEY_SIGNED;
  if (key_t == RB200_F32 || key_t == RB200_F64) kind = rsort::KEY_FLOAT;
  return {kind, descending ? 1 : 0};
}

template <class U, class VT>
int sort_uv(U* keys, U* keys_alt, VT* vals, VT* vals_alt, long long n, rsort::KeyOrder order, cudaStream_t st)
{
  Scratch ws(POOL_DEVICE, rsort::workspace_bytes<U, VT>(n), st);
  if (!ws.p) return RB200_ENOMEM;
  RB200_CUDA((rsort::sort<U, VT>(st, keys, keys_alt, vals, vals_alt, n, order, ws.p)));
  return 0;
}

template <class U> int sort_u(U* keys, U* keys_alt, void* vals, void* vals_alt, int val_bytes, long long n, rsort::KeyOrder order, cudaStream_t st)
{
  switch (val_bytes) {
    case 0: return sort_uv<U, rsort::NoValue>(keys, keys_alt, nullptr, nullptr, n, order, st);
    case 4: return sort_uv<U, uint32_t>(keys, keys_alt, (uint32_t*)vals, (uint32_t*)vals_alt, n, order, st);
    case 8: return sort_uv<U, unsigned long long>(keys, keys_alt, (unsigned long long*)vals, (unsigned long long*)vals_alt, n, order, st);
    default: return fail(RB200_ENOTSUP, "rb200_sort_pairs: val_bytes must be 4 or 8");
  }
}

int sort_any(int key_t, void* keys, void* keys_alt, void* vals, void* vals_alt, int val_bytes, long long n, int descending, cudaStream_t st)
{
  const size_t ks = dtype_size(key_t);
  if (ks == 0) return fail(RB200_EINVAL, "rb200_sort: bad key dtype");
  if (n <= 1) return 0;  // pattern/sort.hpp:76,143,219,301: only sorts when N > 1
  rsort::KeyOrder order = order_of(key_t, descending);
  if (ks == 4) return sort_u<uint32_t>((uint32_t*)keys, (uint32_t*)keys_alt, vals, vals_alt, val_bytes, n, order, st);
  return sort_u<unsigned long long>((unsigned long long*)keys, (unsigned long long*)keys_alt, vals, vals_alt, val_bytes, n, order, st);
}

// host copy of rsort::to_ordered (the device one is __device__ only)
template <class U> U to_ordered_host(U k, rsort::KeyOrder o)
{
  constexpr int BITS = 8 * sizeof(U);
  constexpr U SIGN = U(1) << (BITS - 1);
  if (o.kind == rsort::KEY_FLOAT) {
    if (k == SIGN) k = 0;
    k ^= (k & SIGN) ? ~U(0) : SIGN;
  } else if (o.kind == rsort::KEY_SIGNED) {
    k ^= SIGN;
  }
  if (o.descending) k = ~k;
  return k;
}

template <class U, class VT>
int partition_uv(const U* kin, U* kout, const VT* vin, VT* vout, long long n, const U* split_host, int nsplit, rsort::KeyOrder order,
                 unsigned long long* counts_dev, cudaStream_t st)
{
  rsort::SplitDigit<U> digit;
  digit.nsplit = nsplit;
  digit.order = order;
  for (int i = 0; i < rsort::MAX_SPLITTERS; ++i) digit.s[i] = i < nsplit ? to_ordered_host<U>(split_host[i], order) : ~U(0);
  for (int i = 1; i < nsplit; ++i)
    if (digit.s[i] < digit.s[i - 1]) return fail(RB200_EINVAL, "rb200_partition_pairs: splitters are not in sort order");
  Scratch ws(POOL_DEVICE, rsort::partition_workspace_bytes<U, VT>(n), st);
  if (!ws.p) return RB200_ENOMEM;
  RB200_CUDA((rsort::partition<U, VT>(st, kin, kout, vin, vout, n, digit, counts_dev, ws.p)));
  return 0;
}

template <class U>
int partition_u(const U* kin, U* kout, const void* vin, void* vout, int val_bytes, long long n, const void* split_host, int nsplit,
                rsort::KeyOrder order, unsigned long long* counts_dev, cudaStream_t st)
{
  switch (val_bytes) {
    case 0: return partition_uv<U, rsort::NoValue>(kin, kout, nullptr, nullptr, n, (const U*)split_host, nsplit, order, counts_dev, st);
    case 4: return partition_uv<U, uint32_t>(kin, kout, (const uint32_t*)vin, (uint32_t*)vout, n, (const U*)split_host, nsplit, order, counts_dev, st);
    case 8: return partition_uv<U, unsigned long long>(kin, kout, (const unsigned long long*)vin, (unsigned long long*)vout, n, (const U*)split_host, nsplit, order, counts_dev, st);
    default: return fail(RB200_ENOTSUP, "rb200_partition_pairs: val_bytes must be 0, 4 or 8");
  }
}

// ---- the exchange proper, for any splitter functor: bucket sizes -> count matrix (all ranks) -> split sweep that writes
//      bucket d straight into rank d's receive buffer -> barrier.  ONE host synchronisation: the count matrix must reach the
//      host because the caller needs the number of elements it received (sizes its views and the local sort with it).
template <class U, class VT, class Base>
int exchange_with(rb200_comm* c, rsort::PeerSplitDigit<U, VT, Base>& digit, const U* kin, const VT* vin, long long n, size_t koff, size_t voff,
                  long long capacity, long long* n_received, cudaStream_t st)
{
  constexpr int NB = rsort::MAX_SPLITTERS + 1;
  using Digit = rsort::PeerSplitDigit<U, VT, Base>;
  for (int d = 0; d < NB; ++d) {
    digit.kdst[d] = d < c->world ? (U*)(c->ws[d] + koff) : nullptr;
    digit.vdst[d] = d < c->world ? (VT*)(c->ws[d] + voff) : nullptr;
  }
  Scratch ws(POOL_DEVICE, rsort::partition_workspace_bytes<U, VT>(n), st);
  Scratch small(POOL_DEVICE, (size_t)(1 + COMM_MAX_WORLD) * NB * 8, st);
  Scratch pinned(POOL_PINNED, (size_t)COMM_MAX_WORLD * NB * 8, st);
  if (!ws.p || !small.p || !pinned.p) return RB200_ENOMEM;
  unsigned long long* counts_dev = (unsigned long long*)small.p;
  unsigned long long* matrix_dev = counts_dev + NB;
  RB200_CUDA((rsort::bucket_counts<U, VT, typename rsort::PartitionTuning<U, VT>::shape, Digit>(st, kin, n, digit, counts_dev, ws.p)));
  // every rank learns every (source, destination) count: also the point where a rank that is still
  // reading its receive buffer from the previous exchange holds the others back (stream order)
  int rc = comm_allgather_small(c, counts_dev, NB * 8, matrix_dev, st);
  if (rc) return rc;
  unsigned long long(*m)[NB] = reinterpret_cast<unsigned long long(*)[NB]>(pinned.p);
  RB200_CUDA(cudaMemcpyAsync(m, matrix_dev, (size_t)c->world * NB * 8, cudaMemcpyDeviceToHost, st));
  RB200_CUDA(cudaStreamSynchronize(st));
  unsigned long long base[NB] = {}, total[NB] = {};
  for (int d = 0; d < c->world; ++d)
    for (int s = 0; s < c->world; ++s) {
      if (s < c->rank) base[d] += m[s][d];
      total[d] += m[s][d];
    }
  for (int d = 0; d < c->world; ++d)
    if ((long long)total[d] > capacity)  // every rank sees the same matrix, so every rank returns here
      return fail(RB200_ENOMEM, "rb200_comm_exchange_pairs: a rank would receive more than recv_capacity elements (skewed keys: raise the capacity)");
  RB200_CUDA((rsort::split_sweep<U, VT, Digit>(st, kin, (U*)nullptr, vin, (VT*)nullptr, n, digit, base, ws.p)));
  // all ranks' sweeps have finished (kernel completion makes the peer stores visible) before anyone reads
  rc = rb200_comm_barrier(c, st);
  if (rc) return rc;
  *n_received = (long long)total[c->rank];
  return 0;
}

// splitters given by the caller (host memory)
template <class U, class VT>
int exchange_uv(rb200_comm* c, const U* kin, const VT* vin, long long n, const U* split_host, int nsplit, rsort::KeyOrder order,
                size_t koff, size_t voff, long long capacity, long long* n_received, cudaStream_t st)
{
  rsort::PeerSplitDigit<U, VT> digit;
  digit.nsplit = nsplit;
  digit.order = order;
  for (int i = 0; i < rsort::MAX_SPLITTERS; ++i) digit.s[i] = i < nsplit ? to_ordered_host<U>(split_host[i], order) : ~U(0);
  for (int i = 1; i < nsplit; ++i)
    if (digit.s[i] < digit.s[i - 1]) return fail(RB200_EINVAL, "rb200_comm_exchange_pairs: splitters are not in sort order");
  return exchange_with<U, VT, rsort::SplitDigit<U>>(c, digit, kin, vin, n, koff, voff, capacity, n_received, st);
}

// ---- splitters selected ON THE DEVICE from a regular sample of every rank's unsorted shard --------------------------------
struct SampleSlot {  // one per (parity, source rank) in every rank's workspace header
  unsigned long long flag;   // epoch of the sample all-gather whose data the slot holds
  unsigned long long count;  // samples the source took (min(samples, shard length))
  unsigned long long data[1];
};
__device__ __forceinline__ SampleSlot* sample_slot(char* ws, unsigned long long epoch, int src_rank)
{
  return reinterpret_cast<SampleSlot*>(ws + COMM_SAMPLE_OFFSET + ((epoch & 1) * COMM_MAX_WORLD + src_rank) * COMM_SAMPLE_SLOT_BYTES);
}

// One CTA.  Takes `samples` regular samples of the unsorted shard, maps them with to_ordered (from here on they are plain
// unsigned words whose order is the sort's order) and stores them into every rank's slot over NVLink; then waits for the
// other ranks' slots in the own workspace and copies everything out: gathered[world][samples] (missing entries = all ones,
// they sort last) and counts[world].
template <class U>
__global__ void __launch_bounds__(1024) peer_sample_allgather_kernel(const CommDev c, const U* __restrict__ keys, const long long n, const int samples,
                                                                     const rsort::KeyOrder order, const unsigned long long epoch, U* __restrict__ gathered,
                                                                     unsigned long long* __restrict__ counts)
{
  const int tid = threadIdx.x;
  const long long taken = n < (long long)samples ? n : (long long)samples;
  for (int i = tid; i < (int)taken; i += blockDim.x) {
    const long long pos = (long long)(((2ull * (unsigned long long)i + 1ull) * (unsigned long long)n) / (2ull * (unsigned long long)taken));
    const U k = rsort::to_ordered(keys[pos], order);
    for (int p = 0; p < c.world; ++p) reinterpret_cast<U*>(sample_slot(c.ws[p], epoch, c.rank)->data)[i] = k;
  }
  if (tid < c.world) sample_slot(c.ws[tid], epoch, c.rank)->count = (unsigned long long)taken;
  __threadfence_system();
  __syncthreads();
  if (tid < c.world) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&sample_slot(c.ws[tid], epoch, c.rank)->flag), "l"(epoch) : "memory");
    const unsigned long long* flag = &sample_slot(c.ws[c.rank], epoch, tid)->flag;
    const long long t0 = clock64();
    unsigned long long seen;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
      if (seen == epoch) break;
      __nanosleep(200);
      if (clock64() - t0 > 20000000000ll) __trap();  // ~10 s: a peer never arrived; fail instead of hanging the box
    } while (true);
  }
  __syncthreads();
  for (int i = tid; i < c.world * samples; i += blockDim.x) {
    const int p = i / samples, k = i - p * samples;
    const SampleSlot* s = sample_slot(c.ws[c.rank], epoch, p);
    gathered[i] = (unsigned long long)k < s->count ? reinterpret_cast<const U*>(s->data)[k] : ~U(0);
  }
  if (tid < c.world) counts[tid] = sample_slot(c.ws[c.rank], epoch, tid)->count;
}

// world - 1 splitters at the ranks i * m / world of the m valid samples (sorted ascending in the ordered encoding);
// out[0 .. 7]: 64-bit words, unused entries all ones
template <class U>
__global__ void pick_splitters_kernel(const U* __restrict__ sorted, const unsigned long long* __restrict__ counts, const int world,
                                      unsigned long long* __restrict__ out)
{
  const int i = threadIdx.x;
  if (i > rsort::MAX_SPLITTERS) return;
  unsigned long long m = 0;
  for (int p = 0; p < world; ++p) m += counts[p];
  unsigned long long v = ~0ull;
  if (i < world - 1 && m > 0) v = (unsigned long long)sorted[((unsigned long long)(i + 1) * m) / (unsigned long long)world];
  out[i] = v;
}

template <class U, class VT>
int sample_exchange_uv(rb200_comm* c, const U* kin, const VT* vin, long long n, int samples, rsort::KeyOrder order, size_t koff, size_t voff,
                       long long capacity, long long* n_received, cudaStream_t st)
{
  const long long total = (long long)c->world * samples;
  Scratch gathered(POOL_DEVICE, (size_t)total * sizeof(U) * 2 + 64 * 8, st);  // [gathered][alternate][counts 8][splitters 8]
  Scratch sort_ws(POOL_DEVICE, rsort::workspace_bytes<U, rsort::NoValue>(total), st);
  if (!gathered.p || !sort_ws.p) return RB200_ENOMEM;
  U* g = (U*)gathered.p;
  U* alt = g + total;
  unsigned long long* counts = (unsigned long long*)(alt + total);
  unsigned long long* split = counts + 8;
  const unsigned long long epoch = ++c->sample_epoch;
  {
    b200::LaunchScope scope("peer_sample_allgather_kernel", st);
    peer_sample_allgather_kernel<U><<<1, 1024, 0, st>>>(comm_dev(c), kin, n, samples, order, epoch, g, counts);
  }
  RB200_CUDA(cudaGetLastError());
  // every rank sorts the same gathered samples with the same kernels: identical splitters everywhere, no broadcast
  RB200_CUDA((rsort::sort<U, rsort::NoValue>(st, g, alt, nullptr, nullptr, total, rsort::KeyOrder{rsort::KEY_UNSIGNED, 0}, sort_ws.p)));
  {
    b200::LaunchScope scope("pick_splitters_kernel", st);
    pick_splitters_kernel<U><<<1, 32, 0, st>>>(g, counts, c->world, split);
  }
  RB200_CUDA(cudaGetLastError());
  rsort::PeerSplitDigit<U, VT, rsort::ConstSplitDigit<U>> digit;
  digit.slot = (int)(c->split_calls++ % rsort::CONST_SPLIT_SLOTS);
  digit.nsplit = c->world - 1;
  digit.order = order;
  RB200_CUDA(rsort::set_const_splitters(st, digit.slot, split));
  return exchange_with<U, VT, rsort::ConstSplitDigit<U>>(c, digit, kin, vin, n, koff, voff, capacity, n_received, st);
}

template <class U>
int sample_exchange_u(rb200_comm* c, const U* kin, const void* vin, int val_bytes, long long n, int samples, rsort::KeyOrder order, size_t koff, size_t voff,
                      long long capacity, long long* n_received, cudaStream_t st)
{
  switch (val_bytes) {
    case 0: return sample_exchange_uv<U, rsort::NoValue>(c, kin, nullptr, n, samples, order, koff, voff, capacity, n_received, st);
    case 4: return sample_exchange_uv<U, uint32_t>(c, kin, (const uint32_t*)vin, n, samples, order, koff, voff, capacity, n_received, st);
    case 8: return sample_exchange_uv<U, unsigned long long>(c, kin, (const unsigned long long*)vin, n, samples, order, koff, voff, capacity, n_received, st);
    default: return fail(RB200_ENOTSUP, "rb200_comm_sample_exchange_pairs: val_bytes must be 0, 4 or 8");
  }
}

template <class U>
int exchange_u(rb200_comm* c, const U* kin, const void* vin, int val_bytes, long long n, const void* split_host, int nsplit, rsort::KeyOrder order,
               size_t koff, size_t voff, long long capacity, long long* n_received, cudaStream_t st)
{
  switch (val_bytes) {
    case 0: return exchange_uv<U, rsort::NoValue>(c, kin, nullptr, n, (const U*)split_host, nsplit, order, koff, voff, capacity, n_received, st);
    case 4: return exchange_uv<U, uint32_t>(c, kin, (const uint32_t*)vin, n, (const U*)split_host, nsplit, order, koff, voff, capacity, n_received, st);
    case 8: return exchange_uv<U, unsigned long long>(c, kin, (const unsigned long long*)vin, n, (const U*)split_host, nsplit, order, koff, voff, capacity, n_received, st);
    default: return fail(RB200_ENOTSUP, "rb200_comm_exchange_pairs: val_bytes must be 0, 4 or 8");
  }
}

}  // namespace

extern "C" {

int rb200_comm_exchange_pairs(rb200_comm_t c, rb200_dtype key_t, const void* keys_in, const void* vals_in, int val_bytes, int64_t n,
                              const void* splitters_host, int nsplit, int descending, size_t recv_keys_offset, size_t recv_vals_offset,
                              int64_t recv_capacity, int64_t* n_received, rb200_stream_t stream)
{
  const size_t ks = dtype_size(key_t);
  if (!c || !n_received || n < 0 || ks == 0 || nsplit != c->world - 1 || (nsplit > 0 && !splitters_host) || recv_capacity < 0 ||
      (n > 0 && (!keys_in || (val_bytes && !vals_in))))
    return fail(RB200_EINVAL, "rb200_comm_exchange_pairs: bad arguments (nsplit must be world - 1)");
  if ((recv_keys_offset & 15) || (recv_vals_offset & 15) || recv_keys_offset < COMM_HEADER_BYTES ||
      recv_keys_offset + ks * (size_t)recv_capacity > c->ws_bytes ||
      (val_bytes && (recv_vals_offset < COMM_HEADER_BYTES || recv_vals_offset + (size_t)val_bytes * (size_t)recv_capacity > c->ws_bytes)))
    return fail(RB200_EINVAL, "rb200_comm_exchange_pairs: receive buffers must lie inside the workspace, after its header, 16-byte aligned");
  const rsort::KeyOrder order = order_of(key_t, descending);
  cudaStream_t st = (cudaStream_t)stream;
  long long got = 0;
  int rc;
  if (ks == 4)
    rc = exchange_u<uint32_t>(c, (const uint32_t*)keys_in, vals_in, val_bytes, n, splitters_host, nsplit, order, recv_keys_offset, recv_vals_offset,
                              recv_capacity, &got, st);
  else
    rc = exchange_u<unsigned long long>(c, (const unsigned long long*)keys_in, vals_in, val_bytes, n, splitters_host, nsplit, order, recv_keys_offset,
                                        recv_vals_offset, recv_capacity, &got, st);
  *n_received = got;
  return rc;
}

int rb200_comm_sample_exchange_pairs(rb200_comm_t c, rb200_dtype key_t, const void* keys_in, const void* vals_in, int val_bytes, int64_t n,
                                     int samples_per_rank, int descending, size_t recv_keys_offset, size_t recv_vals_offset, int64_t recv_capacity,
                                     int64_t* n_received, rb200_stream_t stream)
{
  const size_t ks = dtype_size(key_t);
  if (!c || !n_received || n < 0 || ks == 0 || samples_per_rank < 1 || samples_per_rank > COMM_MAX_SAMPLES || recv_capacity < 0 ||
      (n > 0 && (!keys_in || (val_bytes && !vals_in))))
    return fail(RB200_EINVAL, "rb200_comm_sample_exchange_pairs: bad arguments (1 <= samples_per_rank <= 4096)");
  if ((recv_keys_offset & 15) || (recv_vals_offset & 15) || recv_keys_offset < COMM_HEADER_BYTES ||
      recv_keys_offset + ks * (size_t)recv_capacity > c->ws_bytes ||
      (val_bytes && (recv_vals_offset < COMM_HEADER_BYTES || recv_vals_offset + (size_t)val_bytes * (size_t)recv_capacity > c->ws_bytes)))
    return fail(RB200_EINVAL, "rb200_comm_sample_exchange_pairs: receive buffers must lie inside the workspace, after its header, 16-byte aligned");
  const rsort::KeyOrder order = order_of(key_t, descending);
  cudaStream_t st = (cudaStream_t)stream;
  long long got = 0;
  int rc;
  if (ks == 4)
    rc = sample_exchange_u<uint32_t>(c, (const uint32_t*)keys_in, vals_in, val_bytes, n, samples_per_rank, order, recv_keys_offset, recv_vals_offset, recv_capacity,
                                     &got, st);
  else
    rc = sample_exchange_u<unsigned long long>(c, (const unsigned long long*)keys_in, vals_in, val_bytes, n, samples_per_rank, order, recv_keys_offset,
                                               recv_vals_offset, recv_capacity, &got, st);
  *n_received = got;
  return rc;
}

int rb200_partition_pairs(rb200_dtype key_t, const void* keys_in, void* keys_out, const void* vals_in, void* vals_out, int val_bytes,
                          int64_t n, const void* splitters_host, int nsplit, int descending, uint64_t* counts_dev, rb200_stream_t stream)
{
  if (n < 0 || nsplit < 0 || nsplit > rsort::MAX_SPLITTERS || !counts_dev || (nsplit > 0 && !splitters_host) ||
      (n > 0 && (!keys_in || !keys_out || keys_in == keys_out || (val_bytes && (!vals_in || !vals_out || vals_in == vals_out)))))
    return fail(RB200_EINVAL, "rb200_partition_pairs: bad arguments");
  const size_t ks = dtype_size(key_t);
  if (ks == 0) return fail(RB200_EINVAL, "rb200_partition_pairs: bad key dtype");
  const rsort::KeyOrder order = order_of(key_t, descending);
  cudaStream_t st = (cudaStream_t)stream;
  if (ks == 4)
    return partition_u<uint32_t>((const uint32_t*)keys_in, (uint32_t*)keys_out, vals_in, vals_out, val_bytes, n, splitters_host, nsplit, order,
                                 (unsigned long long*)counts_dev, st);
  return partition_u<unsigned long long>((const unsigned long long*)keys_in, (unsigned long long*)keys_out, vals_in, vals_out, val_bytes, n,
                                         splitters_host, nsplit, order, (unsigned long long*)counts_dev, st);
}

int rb200_sort_pairs_buffered(rb200_dtype key_t, void* keys, void* keys_alt, void* vals, void* vals_alt, int val_bytes,
                              int64_t n, int descending, rb200_stream_t stream)
{
  if (n < 0 || (n > 1 && (!keys || !keys_alt || (val_bytes && (!vals || !vals_alt)))))
    return fail(RB200_EINVAL, "rb200_sort_pairs_buffered: bad arguments");
  return sort_any(key_t, keys, keys_alt, vals, vals_alt, val_bytes, n, descending, (cudaStream_t)stream);
}

int rb200_sort_pairs(rb200_dtype key_t, void* keys, void* vals, int val_bytes, int64_t n, int descending, rb200_stream_t stream)
{
  if (n < 0 || (n > 1 && (!keys || !vals))) return fail(RB200_EINVAL, "rb200_sort_pairs: bad arguments");
  if (val_bytes != 4 && val_bytes != 8) return fail(RB200_ENOTSUP, "rb200_sort_pairs: val_bytes must be 4 or 8");
  if (n <= 1) return 0;
  const size_t ks = dtype_size(key_t);
  if (ks == 0) return fail(RB200_EINVAL, "rb200_sort_pairs: bad key dtype");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch kalt(POOL_DEVICE, ks * (size_t)n, st);
  Scratch valt(POOL_DEVICE, (size_t)val_bytes * (size_t)n, st);
  if (!kalt.p || !valt.p) return RB200_ENOMEM;
  return sort_any(key_t, keys, kalt.p, vals, valt.p, val_bytes, n, descending, st);
}

int rb200_sort(rb200_dtype key_t, void* keys, int64_t n, int descending, rb200_stream_t stream)
{
  if (n < 0 || (n > 1 && !keys)) return fail(RB200_EINVAL, "rb200_sort: bad arguments");
  if (n <= 1) return 0;
  const size_t ks = dtype_size(key_t);
  if (ks == 0) return fail(RB200_EINVAL, "rb200_sort: bad key dtype");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch kalt(POOL_DEVICE, ks * (size_t)n, st);
  if (!kalt.p) return RB200_ENOMEM;
  return sort_any(key_t, keys, kalt.p, nullptr, nullptr, 0, n, descending, st);
}

}  // extern "C"
