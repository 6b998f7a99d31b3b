// raja_b200 C ABI: sorts (type-erased front for b200/radix_sort.cuh).
#include "runtime.hpp"
#include "comm.hpp"
#include "b200/radix_sort.cuh"

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

// fused partition + all-to-all: the split sweep writes bucket d straight into rank d's receive buffer
template <class U, class VT>
int exchange_uv(rb200_comm* c, const U* kin, const VT* vin, long long n, const U* split_host, int nsplit, rsort::KeyOrder order,
                size_t koff, size_t voff, long long capacity, long long* n_received, cudaStream_t st)
{
  constexpr int NB = rsort::MAX_SPLITTERS + 1;
  rsort::PeerSplitDigit<U, VT> digit;
  digit.nsplit = nsplit;
  digit.order = order;
  for (int i = 0; i < rsort::MAX_SPLITTERS; ++i) digit.s[i] = i < nsplit ? to_ordered_host<U>(split_host[i], order) : ~U(0);
  for (int i = 1; i < nsplit; ++i)
    if (digit.s[i] < digit.s[i - 1]) return fail(RB200_EINVAL, "rb200_comm_exchange_pairs: splitters are not in sort order");
  for (int d = 0; d < NB; ++d) {
    digit.kdst[d] = d < c->world ? (U*)(c->ws[d] + koff) : nullptr;
    digit.vdst[d] = d < c->world ? (VT*)(c->ws[d] + voff) : nullptr;
  }
  Scratch ws(POOL_DEVICE, rsort::partition_workspace_bytes<U, VT>(n), st);
  Scratch small(POOL_DEVICE, (size_t)(1 + COMM_MAX_WORLD) * NB * 8, st);
  if (!ws.p || !small.p) return RB200_ENOMEM;
  unsigned long long* counts_dev = (unsigned long long*)small.p;
  unsigned long long* matrix_dev = counts_dev + NB;
  RB200_CUDA((rsort::bucket_counts<U, VT>(st, kin, n, digit, counts_dev, ws.p)));
  // every rank learns every (source, destination) count: also the point where a rank that is still
  // reading its receive buffer from the previous exchange holds the others back (stream order)
  int rc = comm_allgather_small(c, counts_dev, NB * 8, matrix_dev, st);
  if (rc) return rc;
  unsigned long long m[COMM_MAX_WORLD][NB];
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
  RB200_CUDA((rsort::split_sweep<U, VT, rsort::PeerSplitDigit<U, VT>>(st, kin, (U*)nullptr, vin, (VT*)nullptr, n, digit, base, ws.p)));
  // all ranks' sweeps have finished (kernel completion makes the peer stores visible) before anyone reads
  rc = rb200_comm_barrier(c, st);
  if (rc) return rc;
  *n_received = (long long)total[c->rank];
  return 0;
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
    return fail(RB200_EINVAL, "rb200_comm_exchange_pairs: receive buffers must lie inside the workspace, after its 64 KiB header, 16-byte aligned");
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
