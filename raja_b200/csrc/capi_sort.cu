// raja_b200 C ABI: sorts (type-erased front for b200/radix_sort.cuh).
#include "runtime.hpp"
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

}  // namespace

extern "C" {

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
