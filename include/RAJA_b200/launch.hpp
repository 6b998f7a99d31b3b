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
// Block + grid combine of one value per thread for a launch of ANY shape.  slot[0] receives the result,
// slot[1 ..] the per-team partials; `counter` is a zeroed, self-resetting ticket (atomicInc wraps).
// Fold order is fixed: lanes ascending inside a warp, warps ascending inside a team, teams ascending.
template<class Op>
__device__ __forceinline__ void launch_grid_reduce_to_slot(typename Op::value_type v, typename Op::value_type* slot, unsigned* counter)
{
  using T = typename Op::value_type;
  __shared__ alignas(16) unsigned char raw[32 * sizeof(T)];
  __shared__ bool s_last;
  T* smem = reinterpret_cast<T*>(raw);
  Op op;
  const unsigned nthreads = blockDim.x * blockDim.y * blockDim.z;
  const unsigned tid      = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const unsigned nblocks  = gridDim.x * gridDim.y * gridDim.z;
  const unsigned bid      = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const unsigned lane = tid & 31u, w = tid >> 5, nw = (nthreads + 31u) >> 5;

  // ordered tree over the lanes that exist in this warp (the last warp of a team may be partial)
  auto warp_fold = [&](T x, unsigned count) {
    const unsigned mask = count >= 32u ? 0xffffffffu : ((1u << count) - 1u);
#pragma unroll
    for (unsigned m = 1; m < 32u; m <<= 1)
    {
      T o = ::b200::shfl_idx_mask(mask, x, (int)((lane + m) < count ? lane + m : lane));
      if (lane + m < count && (lane & (2u * m - 1u)) == 0u) x = op(x, o);
    }
    return x;  // valid in lane 0
  };
  const unsigned in_warp = (nthreads - w * 32u) < 32u ? (nthreads - w * 32u) : 32u;
  v = warp_fold(v, in_warp);
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0)
  {
    if (nw > 1)  // then warp 0 is a full warp
    {
      v = (lane < nw) ? smem[lane] : Op::identity();
      v = warp_fold(v, 32u);
    }
    if (lane == 0)
    {
      ::b200::st_cg(slot + 1 + bid, v);
      __threadfence();
      const unsigned ticket = ::atomicInc(counter, nblocks - 1u);
      s_last = (ticket == nblocks - 1u);
    }
  }
  __syncthreads();
  if (s_last)
  {
    __threadfence();
    // every thread folds a CONTIGUOUS range of partials, so the final order is team order
    const unsigned per = (nblocks + nthreads - 1u) / nthreads;
    const unsigned lo = tid * per, hi = (lo + per < nblocks) ? lo + per : nblocks;
    T acc = Op::identity();
    for (unsigned i = lo; i < hi; ++i) acc = op(acc, ::b200::ld_cg(slot + 1 + i));
    acc = warp_fold(acc, in_warp);
    __syncthreads();  // smem reuse
    if (lane == 0) smem[w] = acc;
    __syncthreads();
    if (w == 0)
    {
      if (nw > 1)
      {
        acc = (lane < nw) ? smem[lane] : Op::identity();
        acc = warp_fold(acc, 32u);
      }
      if (lane == 0) ::b200::st_cg(slot, acc);
    }
  }
  __syncthreads();  // smem and s_last are reused by the next reducer of the pack
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
  // folded into the user's target on the host, as the reference does after res.wait() (policy/cuda/params/reduce.hpp:46-60)
  alignas(alignof(T)) unsigned char raw[sizeof(T)];
  ::RAJA::b200_detail::check(cudaMemcpyAsync(raw, red.devicetarget, sizeof(T), cudaMemcpyDeviceToHost, li.stream), "reducer readback");
  ::RAJA::b200_detail::check(cudaStreamSynchronize(li.stream), "reducer readback");
  T value = *reinterpret_cast<T*>(raw);
  red.combineTarget(value);
  rb200_device_zeroed_free(red.device_count, li.stream);
  red.device_count = nullptr;
  rb200_device_free(red.devicetarget, li.stream);
  red.devicetarget = nullptr;
}

template<bool Async>
inline void param_init(::RAJA::policy::b200::b200_launch_t<Async> const&, RAJA::detail::Name&, const ::RAJA::policy::b200::LaunchInfo&)
{}
template<bool Async>
RAJA_HOST_DEVICE inline void param_combine(::RAJA::policy::b200::b200_launch_t<Async> const&, RAJA::detail::Name&)
{}
template<bool Async>
inline void param_resolve(::RAJA::policy::b200::b200_launch_t<Async> const&, RAJA::detail::Name&, const ::RAJA::policy::b200::LaunchInfo&)
{}
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
      BODY body = body_in;
      kern<<<grid, block, params.shared_mem_size, stream>>>(body);
      b200_detail::check(cudaGetLastError(), "launch");
      if (!async) cuda_res.wait();
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
      BODY body = body_in;
      kern<<<grid, block, params.shared_mem_size, stream>>>(body, launch_reducers);
      b200_detail::check(cudaGetLastError(), "launch");
      RAJA::expt::ParamMultiplexer::parampack_resolve(POL {}, launch_reducers, info);
      if (!async) cuda_res.wait();
    }
    return resources::EventProxy<resources::Resource>(res);
  }
};

}  // namespace RAJA

#endif
