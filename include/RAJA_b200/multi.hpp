// Multi-GPU execution for the b200 back-end: RAJA::b200_multi_exec<BLOCK, Async>.
//
// RAJA has no multi-GPU policy (the reference runs one device per process and leaves the rest to MPI); the surface here is
// the minimal one.  One RANK per GPU -- a process, or a host thread that owns a device.  Each rank owns a contiguous shard of
// every global index space and keeps its piece of every array in its own HBM; nothing is gathered onto one device.
//
//     RAJA::b200::MultiResource mres(rank, world, workspaces, ws_bytes, res);   // becomes current on this host thread
//     auto seg = mres.shard(0L, N);                                            // my piece of [0, N), GLOBAL index values
//     RAJA::b200::ShardedSpan<double> a(a_local, seg);                          // a[i] with a global i
//     RAJA::forall<RAJA::b200_multi_exec<256>>(mres.resource(), seg, RAJA::expt::Reduce<RAJA::operators::plus>(&dot),
//                    [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<double, RAJA::operators::plus>& d) { d += a[i] * b[i]; });
//     RAJA::inclusive_scan<RAJA::b200_multi_exec<256>>(mres.resource(), RAJA::make_span(in_local, n), RAJA::make_span(out_local, n), plus);
//     auto piece = RAJA::b200::multi_sort_pairs(mres, keys_local, vals_local, n);   // piece.keys / piece.vals / piece.n
//
//   forall              on the shard, no exchange
//   reducer parameters  grid reduce per GPU -> all-gather of the scalars through peer memory -> folded in rank order into the
//                       target on every rank (deterministic; MinLoc / MaxLoc keep the first extremum: ranks own ascending ranges)
//   scans               shard total -> exclusive prefix over the ranks (one peer-memory kernel) -> scan seeded with that carry
//                       (a rank whose shard may be EMPTY calls RAJA::b200::multi_inclusive_scan / multi_exclusive_scan below)
//   sort_pairs          sample sort: regular sample of the unsorted shard, splitters selected on the device, ONE sweep that writes
//                       every bucket straight into its destination GPU (fused partition + all-to-all), one local onesweep sort
// The exchanges use the peer-memory communicator of libraja_b200.so (include/raja_b200.h, rb200_comm_*): `workspaces` are the
// world ranks' symmetric buffers as THIS rank can address them -- CUDA IPC / fabric handles between processes (raja_b200/multi.py
// lets torch symmetric memory do that exchange), plain device pointers with peer access enabled between host threads of one
// process (LocalPeerGroup below).  Classic b200_reduce objects are per rank; combine them explicitly if needed.
#ifndef RAJA_B200_MULTI_HPP
#define RAJA_B200_MULTI_HPP

#include <condition_variable>
#include <mutex>
#include <vector>

namespace RAJA
{
namespace b200
{

class MultiResource
{
public:
  MultiResource(int rank, int world, void* const* workspaces, size_t ws_bytes, resources::Cuda res) : m_res(res)
  {
    ::RAJA::b200_detail::check_rc(rb200_comm_create(rank, world, workspaces, ws_bytes, &m_state.comm), "MultiResource");
    m_state.rank     = rank;
    m_state.world    = world;
    m_state.ws_bytes = ws_bytes;
    m_state.ws       = static_cast<char*>(workspaces[rank]);
    make_current();
  }
  ~MultiResource()
  {
    if (::RAJA::b200_detail::tl_multi().comm == m_state.comm) ::RAJA::b200_detail::tl_multi() = ::RAJA::b200_detail::MultiState {};
    rb200_comm_destroy(m_state.comm);
  }
  MultiResource(const MultiResource&)            = delete;
  MultiResource& operator=(const MultiResource&) = delete;

  //! the b200_multi_exec policies on this host thread act as this rank from now on
  void make_current() { ::RAJA::b200_detail::tl_multi() = m_state; }

  int rank() const { return m_state.rank; }
  int world() const { return m_state.world; }
  rb200_comm_t comm() const { return m_state.comm; }
  resources::Cuda& resource() { return m_res; }

  //! this rank's contiguous piece of [begin, end): near-equal shards, the first (end - begin) % world ranks get one more
  template<typename IdxT>
  TypedRangeSegment<IdxT> shard(IdxT begin, IdxT end) const
  {
    const IdxT n = end - begin, base = n / (IdxT)m_state.world, rem = n % (IdxT)m_state.world;
    const IdxT r  = (IdxT)m_state.rank;
    const IdxT lo = begin + r * base + (r < rem ? r : rem);
    return TypedRangeSegment<IdxT>(lo, lo + base + (r < rem ? 1 : 0));
  }

  //! bytes a workspace must have so that a sort of `elems` (key, value) pairs per rank fits (1.25 x a shard is received at most)
  static size_t workspace_bytes(size_t elems, size_t key_bytes, size_t val_bytes)
  {
    const size_t cap = elems + elems / 4 + 4096;
    return rb200_comm_header_bytes() + ((cap * key_bytes + 127) / 128) * 128 + cap * val_bytes + 256;
  }

  //! waits, on this rank's stream, until every rank has reached the same point (one peer-memory kernel; no host sync)
  void barrier() { ::RAJA::b200_detail::check_rc(rb200_comm_barrier(m_state.comm, m_res.get_stream()), "MultiResource::barrier"); }

private:
  ::RAJA::b200_detail::MultiState m_state;
  resources::Cuda m_res;
};

//! a rank's piece of a global array, indexed with GLOBAL indices (what the lambda of a sharded forall receives)
template<typename T, typename IdxT = Index_type>
struct ShardedSpan
{
  T* data;
  IdxT first;
  ShardedSpan(T* local, IdxT first_global) : data(local), first(first_global) {}
  template<typename Seg>
  ShardedSpan(T* local, const Seg& shard) : data(local), first(*shard.begin())
  {}
  RAJA_HOST_DEVICE RAJA_INLINE T& operator[](IdxT global_index) const { return data[global_index - first]; }
};

//! this rank's piece of a globally sorted sequence: views into the communicator's receive buffers, valid until the next sort
template<typename K, typename V>
struct SortedPiece
{
  K* keys;
  V* vals;
  long long n;
};

namespace detail
{
template<typename K> struct sort_dtype
{
  static_assert(std::is_arithmetic<K>::value && (sizeof(K) == 4 || sizeof(K) == 8), "multi-GPU sort: 4- or 8-byte arithmetic keys");
  static constexpr rb200_dtype value = std::is_floating_point<K>::value ? (sizeof(K) == 4 ? RB200_F32 : RB200_F64)
                                       : std::is_signed<K>::value       ? (sizeof(K) == 4 ? RB200_I32 : RB200_I64)
                                                                        : (sizeof(K) == 4 ? RB200_U32 : RB200_U64);
};
}  // namespace detail

//! Stable sample sort of the concatenation of all ranks' (keys, vals); rank r receives the keys that sort before rank r + 1's.
//! vals may be nullptr (keys only).  The inputs are left untouched.  Collective: every rank calls it.
template<typename K, typename V>
inline SortedPiece<K, V> multi_sort_pairs(MultiResource& m, const K* keys, const V* vals, long long n, bool descending = false, int samples_per_rank = 4096)
{
  static_assert(std::is_trivially_copyable<V>::value && (sizeof(V) == 4 || sizeof(V) == 8), "multi-GPU sort: values must be trivially copyable, 4 or 8 bytes");
  const ::RAJA::b200_detail::MultiState& st = ::RAJA::b200_detail::multi_state_or_throw("multi_sort_pairs");
  cudaStream_t stream  = m.resource().get_stream();
  const int vb         = vals ? (int)sizeof(V) : 0;
  const size_t header  = rb200_comm_header_bytes();
  const long long cap  = (long long)((st.ws_bytes - header - 256) / (sizeof(K) + (size_t)vb));
  const size_t koff    = header;
  const size_t voff    = koff + (((size_t)cap * sizeof(K) + 127) / 128) * 128;
  int64_t got          = 0;
  const rb200_dtype kt = detail::sort_dtype<K>::value;
  ::RAJA::b200_detail::check_rc(rb200_comm_sample_exchange_pairs(m.comm(), kt, keys, vals, vb, n, samples_per_rank, descending ? 1 : 0, koff, voff,
                                                                 (int64_t)cap, &got, stream),
                                "multi_sort_pairs exchange");
  SortedPiece<K, V> out {reinterpret_cast<K*>(st.ws + koff), vb ? reinterpret_cast<V*>(st.ws + voff) : nullptr, (long long)got};
  // the runs arrive concatenated in source-rank order, each in its source order: one stable local sort finishes the job
  if (vb) ::RAJA::b200_detail::check_rc(rb200_sort_pairs(kt, out.keys, out.vals, vb, got, descending ? 1 : 0, stream), "multi_sort_pairs local sort");
  else ::RAJA::b200_detail::check_rc(rb200_sort(kt, out.keys, got, descending ? 1 : 0, stream), "multi_sort_pairs local sort");
  return out;
}
template<typename K>
inline SortedPiece<K, unsigned> multi_sort(MultiResource& m, const K* keys, long long n, bool descending = false)
{
  return multi_sort_pairs<K, unsigned>(m, keys, nullptr, n, descending);
}

//! Collective scans that are safe for EMPTY shards.  RAJA::inclusive_scan<b200_multi_exec>(...) & co. work as long as every rank's
//! shard holds at least one element: RAJA's front-end returns before it reaches any back-end when the range is empty
//! (pattern/scan.hpp:72, 137, 212, 291), so a rank with an empty shard would skip the carry exchange its peers are waiting in.
//! These entry points go to the back-end directly; in place (out == in) is fine.  Collective: every rank calls them.
template<typename T, typename Function>
inline void multi_inclusive_scan(MultiResource& m, const T* in, T* out, long long n, Function op)
{
  ::RAJA::impl::scan::b200_detail::run_multi<true, false>(m.resource(), in, in + n, out, op, T {});
}
template<typename T, typename Function>
inline void multi_exclusive_scan(MultiResource& m, const T* in, T* out, long long n, Function op, T init)
{
  ::RAJA::impl::scan::b200_detail::run_multi<false, false>(m.resource(), in, in + n, out, op, init);
}

//! Ranks that are host threads of ONE process: every thread sets its device, then attach()es -- allocates its workspace, enables
//! peer access to the other devices and learns everyone's pointer.  (Between processes the pointers come from CUDA IPC.)
//! The exchange kernels wait for their peers ON THE DEVICE: in this one-process mode run with CUDA_MODULE_LOADING=EAGER, so
//! that no host thread has to synchronise a context (lazy module loading does, at a kernel's first launch) behind a kernel
//! that is waiting for that very thread's rank.
class LocalPeerGroup
{
public:
  LocalPeerGroup(int world, size_t ws_bytes) : m_world(world), m_bytes(ws_bytes), m_ptrs((size_t)world, nullptr), m_dev((size_t)world, -1) {}
  ~LocalPeerGroup()
  {
    for (int r = 0; r < m_world; ++r)
      if (m_ptrs[(size_t)r])
      {
        cudaSetDevice(m_dev[(size_t)r]);
        cudaFree(m_ptrs[(size_t)r]);
      }
  }
  size_t bytes() const { return m_bytes; }
  //! called by rank `rank`'s thread after cudaSetDevice; returns the world pointers once every rank has attached
  void* const* attach(int rank)
  {
    int dev = 0;
    ::RAJA::b200_detail::check(cudaGetDevice(&dev), "LocalPeerGroup");
    void* p = nullptr;
    ::RAJA::b200_detail::check(cudaMalloc(&p, m_bytes), "LocalPeerGroup workspace");
    ::RAJA::b200_detail::check(cudaMemset(p, 0, rb200_comm_header_bytes()), "LocalPeerGroup workspace");
    {
      std::unique_lock<std::mutex> lock(m_mu);
      m_ptrs[(size_t)rank] = p;
      m_dev[(size_t)rank]  = dev;
      ++m_arrived;
      m_cv.notify_all();
      m_cv.wait(lock, [&] { return m_arrived >= m_world; });
    }
    for (int r = 0; r < m_world; ++r)
      if (m_dev[(size_t)r] != dev)
      {
        cudaError_t e = cudaDeviceEnablePeerAccess(m_dev[(size_t)r], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else ::RAJA::b200_detail::check(e, "LocalPeerGroup peer access");
      }
    rendezvous();
    return m_ptrs.data();
  }
  //! host-side barrier of the group's threads
  void rendezvous()
  {
    std::unique_lock<std::mutex> lock(m_mu);
    const unsigned long gen = m_gen;
    if (++m_waiting == m_world)
    {
      m_waiting = 0;
      ++m_gen;
      m_cv.notify_all();
    }
    else m_cv.wait(lock, [&] { return m_gen != gen; });
  }

private:
  int m_world;
  size_t m_bytes;
  std::vector<void*> m_ptrs;
  std::vector<int> m_dev;
  std::mutex m_mu;
  std::condition_variable m_cv;
  int m_arrived = 0, m_waiting = 0;
  unsigned long m_gen = 0;
};

}  // namespace b200
}  // namespace RAJA

#endif
