// TypedIndexSet foralls for b200_exec: ExecPolicy<seq_segit, b200_exec<...>>.
//
// The generic dispatch (pattern/forall.hpp:279-299) runs every segment through the SEGMENT policy as given, so a
// synchronous policy waits for the stream after every segment.  Here the segments of one index set are enqueued
// back to back with the ASYNCHRONOUS form of the policy on the caller's resource and a synchronous policy waits
// ONCE, after the last segment -- what the reference's CUDA back-end states as its intent in
// policy/cuda/forall.hpp:682-706 (async launch per segment, one synchronize).  Segments run in index-set order on
// one stream, so the results are those of the sequential segment iteration.
//
// RAJA::wrap::forall is called QUALIFIED from pattern/forall.hpp, so this more specialised overload must be
// visible before that header is first included (same rule as scan / sort / atomic, see RAJA_b200/RAJA.hpp).
// ListSegments: the index array is read with fully coalesced, unrolled loads by the persistent engine
// (b200/forall.cuh: BLOCK x UNROLL tiles over the POSITION in the list); what the body then does with the index
// is a gather the back-end cannot see.
#ifndef RAJA_B200_INDEXSET_HPP
#define RAJA_B200_INDEXSET_HPP

#include "RAJA/index/IndexSet.hpp"

namespace RAJA
{
namespace b200_detail
{
// one segment of an index set -> forall_impl of the segment policy (found by ADL at instantiation)
struct CallSegment
{
  template<typename Segment, typename ExecPol, typename Body, typename Res, typename ForallParams>
  RAJA_INLINE void operator()(Segment const& seg, ExecPol pol, Body body, Res r, ForallParams f_params) const
  {
    forall_impl(r, pol, seg, body, f_params);
  }
};
}  // namespace b200_detail

namespace wrap
{
template<typename LoopBody, size_t BLOCK_SIZE, size_t GRID_SIZE, bool Async, typename... SegmentTypes, typename ForallParams>
RAJA_INLINE resources::EventProxy<resources::Cuda> forall(resources::Cuda r,
                                                          ExecPolicy<seq_segit, ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, Async>>,
                                                          const TypedIndexSet<SegmentTypes...>& iset,
                                                          LoopBody loop_body,
                                                          ForallParams f_params)
{
  using SegmentPolicy = ::RAJA::policy::b200::b200_exec_explicit<BLOCK_SIZE, GRID_SIZE, true>;
  const int num_seg   = (int)iset.getNumSegments();
  for (int isi = 0; isi < num_seg; ++isi) iset.segmentCall(isi, b200_detail::CallSegment {}, SegmentPolicy {}, loop_body, r, f_params);
  if (!Async && num_seg > 0)
  {
    r.wait();
    rb200_stream_mark_synced(r.get_stream());
  }
  return resources::EventProxy<resources::Cuda>(r);
}
}  // namespace wrap
}  // namespace RAJA

#endif
