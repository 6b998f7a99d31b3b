// Pools with the reference's surface (RAJA::cuda::*_mempool_type, policy/cuda/MemUtils_CUDA.hpp:162-167):
// getInstance().malloc<T>(n) / free(p), backed by the stream-ordered pools of libraja_b200.so.
#ifndef RAJA_B200_POOLS_HPP
#define RAJA_B200_POOLS_HPP

namespace RAJA
{
namespace b200
{
namespace detail
{
enum class PoolKind { device, device_zeroed, pinned };

template<PoolKind K>
class MemPool
{
public:
  static MemPool& getInstance()
  {
    static MemPool pool;
    return pool;
  }
  // stream the next malloc/free is ordered on (default: the legacy default stream)
  void set_stream(cudaStream_t s) { m_stream = s; }
  template<typename T>
  T* malloc(size_t nTs)
  {
    void* p = nullptr;
    if (K == PoolKind::device) p = rb200_device_malloc(nTs * sizeof(T), m_stream);
    else if (K == PoolKind::device_zeroed) p = rb200_device_zeroed_malloc(nTs * sizeof(T), m_stream);
    else p = rb200_pinned_malloc(nTs * sizeof(T));
    if (!p) throw std::runtime_error(std::string("raja_b200 pool: ") + rb200_last_error());
    return static_cast<T*>(p);
  }
  void free(const void* cptr)
  {
    void* p = const_cast<void*>(cptr);
    if (K == PoolKind::device) rb200_device_free(p, m_stream);
    else if (K == PoolKind::device_zeroed) rb200_device_zeroed_free(p, m_stream);
    else rb200_pinned_free(p);
  }
  void free_chunks() { rb200_pool_trim(); }
  size_t cached_bytes() const { return rb200_pool_bytes_cached(); }

private:
  cudaStream_t m_stream = 0;
};
}  // namespace detail
using device_mempool_type        = detail::MemPool<detail::PoolKind::device>;
using device_zeroed_mempool_type = detail::MemPool<detail::PoolKind::device_zeroed>;
using pinned_mempool_type        = detail::MemPool<detail::PoolKind::pinned>;
}  // namespace b200
}  // namespace RAJA

#endif
