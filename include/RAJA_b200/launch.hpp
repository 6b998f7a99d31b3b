// RAJA::launch for the b200 back-end: LaunchExecute<b200_launch_t<Async>> and the expt::Reduce /
// ReduceLoc / Name parameter hooks for it.
// Replaces policy/cuda/launch.hpp:31-215 (launch_global_fcn, launch_new_reduce_global_fcn and
// LaunchExecute<cuda_launch_explicit_t<async, unspecified, unspecified>>).  Teams and threads are the
// caller's (LaunchParams), as in the reference; what changes is the grid combine behind the
// reducer parameters: warp shuffles over the FLATTENED thread id (any 1-3 D block shape, partial
// last warp allowed) -> one shared-memory hop -> last-block-done over the per-team partials in slot
// order (deterministic), instead of the reference's cuda_flatten_global_xyz_direct grid_reduce.
// Loops inside the launch body use the reference's own loop policies (cuda_global_thread_x,
// cuda_block_x_direct, cuda_thread_x_loop, ...): they are plain device code and independent of the
// launch policy.
#ifndef RAJA_B200_LAUNCH_HPP
#define RAJA_B200_LAUNCH_HPP

namespace RAJA
{
namespace policy
{
namespace b200
{
template<bool Async = false>
struct b200_launch_t
    : public RAJA::make_policy_pattern_launch_platform_t<policy_id,
                                                         RAJA::Pattern::region,
                                                         (Async ? RAJA::Launch::async : RAJA::Launch::sync),
                                                         RAJA::Platform::cuda>
{
  static constexpr bool async = Async;
};
}  // namespace b200
}  // namespace policy
using policy::b200::b200_launch_t;

namespace resources
{
template<bool Async>
struct get_resource<::RAJA::policy::b200::b200_launch_t<Async>>
{
  using type = camp::resources::Cuda;
};
}  // namespace resources

namespace b200_detail
{
// Block + grid combine of one value per thread for a launch of ANY shape (shaped_grid_reduce, RAJA_b200/reduce.hpp).
// slot[0] receives the result, slot[1 ..] the per-team partials; `counter` is a zeroed, self-resetting ticket.
template<class Op>
__device__ __forceinline__ void launch_grid_reduce_to_slot(typename Op::value_type v, typename Op::value_type* slot, unsigned* counter)
{
  typename Op::value_type r;
  if (shaped_grid_reduce<Op>(v, slot + 1, counter, r)) ::b200::st_cg(slot, r);
}
}  // namespace b200_detail

namespace expt
{
namespace detail
{
template<bool Async, typename OP, typename T, typename VOp>
inline void param_init(::RAJA::policy::b200::b200_launch_t<Async> const&, Reducer<OP, T, VOp>& red, ::RAJA::policy::b200::LaunchInfo& li)
{
  red.devicetarget = static_cast<T*>(rb200_device_malloc(sizeof(T) * ((size_t)li.grid + 1), li.stream));
  red.device_count = static_cast<unsigned int*>(rb200_device_zeroed_malloc(sizeof(unsigned int), li.stream));
  if (!red.devicetarget || !red.device_count) throw std::runtime_error(std::string("raja_b200: reducer scratch: ") + rb200_last_error());
  li.expect(red.devicetarget, sizeof(T));
}

template<bool Async, typename OP, typename T, typename VOp>
RAJA_HOST_DEVICE inline void param_combine(::RAJA::policy::b200::b200_launch_t<Async> const&, Reducer<OP, T, VOp>& red)
{
#if defined(__CUDA_ARCH__)
  using Op = typename ::RAJA::b200_detail::kernel_op<OP, T>::type;
  ::RAJA::b200_detail::launch_grid_reduce_to_slot<Op>(red.getVal(), red.devicetarget, red.device_count);
#else
  RAJA_UNUSED_VAR(red);
#endif
}

template<bool Async, typename OP, typename T, typename VOp>
inline void param_resolve(::RAJA::policy::b200::b200_launch_t<Async> const&, Reducer<OP, T, VOp>& red, ::RAJA::policy::b200::LaunchInfo& li)
{
  b200_resolve::resolve(red, li);
}

template<bool Async>
inline void param_init(::RAJA::policy::b200::b200_launch_t<Async> const&, RAJA::detail::Name& kn, const ::RAJA::policy::b200::LaunchInfo&)
{
  if (kn.name != nullptr) rb200_range_push(kn.name);
}
template<bool Async>
RAJA_HOST_DEVICE inline void param_combine(::RAJA::policy::b200::b200_launch_t<Async> const&, RAJA::detail::Name&)
{}
template<bool Async>
inline void param_resolve(::RAJA::policy::b200::b200_launch_t<Async> const&, RAJA::detail::Name& kn, const ::RAJA::policy::b200::LaunchInfo&)
{
  if (kn.name != nullptr) rb200_range_pop();
}
}  // namespace detail
}  // namespace expt

namespace b200_detail
{
template<typename BODY>
__global__ void launch_kernel(const __grid_constant__ BODY body_in)
{
  LaunchContext ctx;
  using RAJA::internal::thread_privatize;
  auto privatizer = thread_privatize(body_in);
  auto& body      = privatizer.get_priv();
  extern __shared__ char raja_b200_shmem_ptr[];
  ctx.shared_mem_ptr = raja_b200_shmem_ptr;
  body(ctx);
}

template<typename POL, typename BODY, typename ReduceParams>
__global__ void launch_reduce_kernel(const __grid_constant__ BODY body_in, ReduceParams reduce_params)
{
  LaunchContext ctx;
  using RAJA::internal::thread_privatize;
  auto privatizer = thread_privatize(body_in);
  auto& body      = privatizer.get_priv();
  extern __shared__ char raja_b200_shmem_ptr[];
  ctx.shared_mem_ptr = raja_b200_shmem_ptr;
  RAJA::expt::invoke_body(reduce_params, body, ctx);
  RAJA::expt::ParamMultiplexer::parampack_combine(POL {}, reduce_params);
}

inline bool launch_dims(const LaunchParams& p, dim3& grid, dim3& block)
{
  grid  = dim3((unsigned)p.teams.value[0], (unsigned)p.teams.value[1], (unsigned)p.teams.value[2]);
  block = dim3((unsigned)p.threads.value[0], (unsigned)p.threads.value[1], (unsigned)p.threads.value[2]);
  // only launch if there is something to iterate over (policy/cuda/launch.hpp:104-107)
  return p.teams.value[0] > 0 && p.teams.value[1] > 0 && p.teams.value[2] > 0 && p.threads.value[0] > 0 && p.threads.value[1] > 0 && p.threads.value[2] > 0;
}
}  // namespace b200_detail

template<bool async>
struct LaunchExecute<::RAJA::policy::b200::b200_launch_t<async>>
{
  // no reducer parameters
  template<typename BODY_IN, typename ReduceParams>
  static concepts::enable_if_t<resources::EventProxy<resources::Resource>,
                               RAJA::expt::type_traits::is_ForallParamPack<ReduceParams>,
                               RAJA::expt::type_traits::is_ForallParamPack_empty<ReduceParams>>
  exec(RAJA::resources::Resource res, const LaunchParams& params, BODY_IN&& body_in, ReduceParams&)
  {
    using BODY = camp::decay<BODY_IN>;
    resources::Cuda cuda_res = res.get<RAJA::resources::Cuda>();
    dim3 grid, block;
    if (b200_detail::launch_dims(params, grid, block))
    {
      cudaStream_t stream = cuda_res.get_stream();
      auto kern = &b200_detail::launch_kernel<BODY>;
      if (params.shared_mem_size > 48 * 1024)
        b200_detail::check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)params.shared_mem_size), "launch shared memory");
      // classic reducer objects captured by the body set themselves up for this launch while the body is copied
      auto make_body = [&]() {
        b200_detail::LaunchScope scope((int)(grid.x * grid.y * grid.z), stream);
        return BODY(body_in);
      };
      BODY body = make_body();
      {
        ::b200::LaunchScope hook("launch_kernel", stream);
        kern<<<grid, block, params.shared_mem_size, stream>>>(body);
      }
      b200_detail::check(cudaGetLastError(), "launch");
      if (!async)
      {
        cuda_res.wait();
        rb200_stream_mark_synced(stream);
      }
    }
    return resources::EventProxy<resources::Resource>(res);
  }

  // with expt::Reduce / ReduceLoc / Name parameters
  template<typename BODY_IN, typename ReduceParams>
  static concepts::enable_if_t<resources::EventProxy<resources::Resource>,
                               RAJA::expt::type_traits::is_ForallParamPack<ReduceParams>,
                               concepts::negate<RAJA::expt::type_traits::is_ForallParamPack_empty<ReduceParams>>>
  exec(RAJA::resources::Resource res, const LaunchParams& params, BODY_IN&& body_in, ReduceParams& launch_reducers)
  {
    using BODY = camp::decay<BODY_IN>;
    using POL  = ::RAJA::policy::b200::b200_launch_t<async>;
    resources::Cuda cuda_res = res.get<RAJA::resources::Cuda>();
    dim3 grid, block;
    if (b200_detail::launch_dims(params, grid, block))
    {
      cudaStream_t stream = cuda_res.get_stream();
      auto kern = &b200_detail::launch_reduce_kernel<POL, BODY, camp::decay<ReduceParams>>;
      if (params.shared_mem_size > 48 * 1024)
        b200_detail::check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)params.shared_mem_size), "launch shared memory");
      ::RAJA::policy::b200::LaunchInfo info {(int)(grid.x * grid.y * grid.z), (int)(block.x * block.y * block.z), stream, cuda_res};
      RAJA::expt::ParamMultiplexer::parampack_init(POL {}, launch_reducers, info);
      auto make_body = [&]() {
        b200_detail::LaunchScope scope(info.grid, stream);
        return BODY(body_in);
      };
      BODY body = make_body();
      {
        ::b200::LaunchScope hook("launch_reduce_kernel", stream);
        kern<<<grid, block, params.shared_mem_size, stream>>>(body, launch_reducers);
      }
      b200_detail::check(cudaGetLastError(), "launch");
      RAJA::expt::ParamMultiplexer::parampack_resolve(POL {}, launch_reducers, info);
      if (!async)
      {
        cuda_res.wait();
        rb200_stream_mark_synced(stream);
      }
    }
    return resources::EventProxy<resources::Resource>(res);
  }
};

}  // namespace RAJA

#endif
