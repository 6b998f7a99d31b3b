// raja_b200 C ABI: scans (type-erased front for b200/scan.cuh).
#include "runtime.hpp"
#include "b200/scan.cuh"

using namespace b200;
using namespace rb200;

namespace {

template <class T, class Op> int scan_t(int inclusive, const T* in, T* out, long long n, const void* init_host, const void* carry_dev,
                                        cudaStream_t st)
{
  using Shape = typename scan::Tuning<T>::shape;
  Scratch ws(POOL_DEVICE, scan::workspace_bytes<T, Shape>(n), st);
  if (!ws.p) return RB200_ENOMEM;
  T init = Op::identity();
  if (!inclusive && init_host) init = *(const T*)init_host;
  if (inclusive) RB200_CUDA((scan::launch<T, Op, true>(st, in, out, n, init, ws.p, 0, 0, nullptr, (const T*)carry_dev)));
  else RB200_CUDA((scan::launch<T, Op, false>(st, in, out, n, init, ws.p, 0, 0, nullptr, (const T*)carry_dev)));
  return 0;
}

template <class T> int scan_op(int inclusive, int op, const T* in, T* out, long long n, const void* init_host, const void* carry_dev,
                               cudaStream_t st)
{
  switch (op) {
    case RB200_PLUS: return scan_t<T, Plus<T>>(inclusive, in, out, n, init_host, carry_dev, st);
    case RB200_MINIMUM: return scan_t<T, Minimum<T>>(inclusive, in, out, n, init_host, carry_dev, st);
    case RB200_MAXIMUM: return scan_t<T, Maximum<T>>(inclusive, in, out, n, init_host, carry_dev, st);
    default: return fail(RB200_EINVAL, "rb200_scan: bad op");
  }
}

}  // namespace

namespace {
int scan_any(int inclusive, int op, int t, const void* in, void* out, long long n, const void* init_host, const void* carry_dev, cudaStream_t st)
{
  switch (t) {
    case RB200_I32: return scan_op<int32_t>(inclusive, op, (const int32_t*)in, (int32_t*)out, n, init_host, carry_dev, st);
    case RB200_U32: return scan_op<uint32_t>(inclusive, op, (const uint32_t*)in, (uint32_t*)out, n, init_host, carry_dev, st);
    case RB200_I64: return scan_op<int64_t>(inclusive, op, (const int64_t*)in, (int64_t*)out, n, init_host, carry_dev, st);
    case RB200_U64: return scan_op<uint64_t>(inclusive, op, (const uint64_t*)in, (uint64_t*)out, n, init_host, carry_dev, st);
    case RB200_F32: return scan_op<float>(inclusive, op, (const float*)in, (float*)out, n, init_host, carry_dev, st);
    case RB200_F64: return scan_op<double>(inclusive, op, (const double*)in, (double*)out, n, init_host, carry_dev, st);
    default: return fail(RB200_EINVAL, "rb200_scan: bad dtype");
  }
}
}  // namespace

extern "C" int rb200_scan(int inclusive, rb200_op op, rb200_dtype t, const void* in, void* out, int64_t n,
                          const void* init_host, rb200_stream_t stream)
{
  if (n < 0 || (n > 0 && (!in || !out))) return fail(RB200_EINVAL, "rb200_scan: bad arguments");
  if (n == 0) return 0;  // pattern/scan.hpp: empty range is a no-op
  return scan_any(inclusive, op, t, in, out, n, init_host, nullptr, (cudaStream_t)stream);
}

extern "C" int rb200_scan_carry(int inclusive, rb200_op op, rb200_dtype t, const void* in, void* out, int64_t n,
                                const void* init_host, const void* carry_dev, rb200_stream_t stream)
{
  if (n < 0 || (n > 0 && (!in || !out))) return fail(RB200_EINVAL, "rb200_scan_carry: bad arguments");
  if (n == 0) return 0;
  return scan_any(inclusive, op, t, in, out, n, init_host, carry_dev, (cudaStream_t)stream);
}
