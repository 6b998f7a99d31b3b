// RAJA::kernel for the b200 back-end: statement::B200Kernel<EnclosedStmts...> launches its enclosed statements as ONE
// kernel and combines RAJA::expt::Reduce / ReduceLoc kernel parameters (RAJA::kernel_param) with the b200 grid combine
// (SURVEY.md section 8 f2).  Replaces the launch half of policy/cuda/kernel/CudaKernel.hpp:214-235 (CudaKernelLauncher[Fixed],
// whose epilogue is combine_params<cuda_flatten_global_xyz_direct>) and :535-726 (StatementExecutor<CudaKernelExt>).
//
// What is ours: the global function, the launch sizing, the reducer-parameter life cycle (param_init / param_combine /
// param_resolve below: ordered last-block-done combine over blocks of ANY shape, one wait for all results) and the set-up of
// classic b200_reduce objects captured by the lambdas.  What stays the reference's: the device-side loop mappings INSIDE the
// kernel (statement::For<I, cuda_thread_x_loop, ...>, Tile, Lambda, ... through cuda_statement_list_executor_t and its
// calculateDimensions): plain device code that does not depend on who launched the kernel -- the same split as RAJA::launch
// (RAJA_b200/launch.hpp).
//
//   using Pol = RAJA::KernelPolicy<RAJA::statement::B200Kernel<
//                 RAJA::statement::For<1, RAJA::cuda_block_x_loop, RAJA::statement::For<0, RAJA::cuda_thread_x_loop, RAJA::statement::Lambda<0>>>>>;
//   RAJA::kernel_param<Pol>(segments, RAJA::make_tuple(RAJA::expt::Reduce<RAJA::operators::plus>(&sum)), lambda);
#ifndef RAJA_B200_KERNEL_HPP
#define RAJA_B200_KERNEL_HPP

namespace RAJA
{
namespace policy
{
namespace b200
{
// the policy the reducer-parameter hooks of a B200Kernel are looked up with
struct b200_kernel_policy
    : public RAJA::make_policy_pattern_launch_platform_t<policy_id, RAJA::Pattern::forall, RAJA::Launch::async, RAJA::Platform::cuda>
{};
}  // namespace b200
}  // namespace policy

namespace statement
{
// num_threads: upper bound on the threads per block the enclosed statements may ask for (also the kernel's __launch_bounds__)
template<bool Async, int num_threads, typename... EnclosedStmts>
struct B200KernelExt : public internal::Statement<::RAJA::policy::b200::b200_exec_explicit<(num_threads > 0 ? (size_t)num_threads : 1024), 0, Async>, EnclosedStmts...>
{};
template<typename... EnclosedStmts>
using B200Kernel = B200KernelExt<false, 1024, EnclosedStmts...>;
template<typename... EnclosedStmts>
using B200KernelAsync = B200KernelExt<true, 1024, EnclosedStmts...>;
template<int num_threads, typename... EnclosedStmts>
using B200KernelFixed = B200KernelExt<false, num_threads, EnclosedStmts...>;
template<int num_threads, typename... EnclosedStmts>
using B200KernelFixedAsync = B200KernelExt<true, num_threads, EnclosedStmts...>;
}  // namespace statement

namespace expt
{
namespace detail
{
// reducer parameters of a B200Kernel: the shape-agnostic combine of RAJA_b200/reduce.hpp
template<typename OP, typename T, typename VOp>
inline void param_init(::RAJA::policy::b200::b200_kernel_policy const&, Reducer<OP, T, VOp>& red, ::RAJA::policy::b200::LaunchInfo& li)
{
  red.devicetarget = static_cast<T*>(rb200_device_malloc(sizeof(T) * ((size_t)li.grid + 1), li.stream));
  red.device_count = static_cast<unsigned int*>(rb200_device_zeroed_malloc(sizeof(unsigned int), li.stream));
  if (!red.devicetarget || !red.device_count) throw std::runtime_error(std::string("raja_b200: reducer scratch: ") + rb200_last_error());
  li.expect(red.devicetarget, sizeof(T));
}
template<typename OP, typename T, typename VOp>
RAJA_HOST_DEVICE inline void param_combine(::RAJA::policy::b200::b200_kernel_policy const&, Reducer<OP, T, VOp>& red)
{
#if defined(__CUDA_ARCH__)
  using Op = typename ::RAJA::b200_detail::kernel_op<OP, T>::type;
  ::RAJA::b200_detail::launch_grid_reduce_to_slot<Op>(red.getVal(), red.devicetarget, red.device_count);
#else
  RAJA_UNUSED_VAR(red);
#endif
}
template<typename OP, typename T, typename VOp>
inline void param_resolve(::RAJA::policy::b200::b200_kernel_policy const&, Reducer<OP, T, VOp>& red, ::RAJA::policy::b200::LaunchInfo& li)
{
  b200_resolve::resolve(red, li);
}
}  // namespace detail
}  // namespace expt

namespace internal
{
template<int BlockSize, typename Data, typename Exec>
__launch_bounds__(BlockSize) __global__ void b200_kernel_launcher(const __grid_constant__ Data data)
{
  Data private_data = data;  // every thread's own loop data: offsets, thread-private parameters, reducer accumulators
  Exec::exec(private_data, true);
  RAJA::expt::detail::combine_params<::RAJA::policy::b200::b200_kernel_policy>(private_data.param_tuple);
}

template<bool Async, int num_threads, typename... EnclosedStmts, typename Types>
struct StatementExecutor<statement::B200KernelExt<Async, num_threads, EnclosedStmts...>, Types>
{
  using stmt_list_t = StatementList<EnclosedStmts...>;

  template<typename Data>
  static inline void exec(Data&& data)
  {
    using data_t     = camp::decay<Data>;
    using executor_t = cuda_statement_list_executor_t<stmt_list_t, data_t, Types>;
    using POL        = ::RAJA::policy::b200::b200_kernel_policy;
    constexpr int max_threads = num_threads > 0 ? num_threads : 1024;

    RAJA::resources::Cuda res = data.get_resource();
    cudaStream_t stream       = res.get_stream();

    // what the enclosed statements ask for (direct mappings fix a dimension, loop mappings only suggest one)
    LaunchDims dims = executor_t::calculateDimensions(data);
    const bool active_threads = dims.threads_are_active(), active_blocks = dims.blocks_are_active();
    // only launch if there is something to iterate over (policy/cuda/kernel/CudaKernel.hpp:560-566)
    if (!(active_threads || active_blocks) || (active_blocks && dims.num_blocks() <= 0) || (active_threads && dims.num_threads() <= 0)) return;

    auto kern = &b200_kernel_launcher<max_threads, data_t, executor_t>;
    dims.dims.threads = fitCudaDims(max_threads, dims.dims.threads, dims.min_dims.threads);
    if (dims.num_threads() > max_threads) RAJA_ABORT_OR_THROW("RAJA::kernel exceeds max num threads");
    // loop-mapped block dimensions are cut down to one wave of resident blocks (grid-stride inside); direct ones are kept
    static thread_local ::b200::PerDevice occ_cache, occ_threads;  // per instantiation, host thread and device; one block size cached
    int& occ       = occ_cache.current();
    int& occ_block = occ_threads.current();
    if (occ == 0 || occ_block != dims.num_threads())
    {
      RAJA::b200_detail::check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, dims.num_threads(), 0), "kernel occupancy");
      if (occ < 1) occ = 1;
      occ_block = dims.num_threads();
    }
    const int max_blocks = occ * ::b200::device_info().sm_count;
    dims.dims.blocks     = fitCudaDims(max_blocks, dims.dims.blocks, dims.min_dims.blocks);

    ::RAJA::policy::b200::LaunchInfo info {dims.num_blocks(), dims.num_threads(), stream, res};
    RAJA::expt::detail::init_params<POL>(data.param_tuple, info);
    {
      // classic b200_reduce objects captured by the lambdas set themselves up while the loop data is copied
      auto make_data = [&]() {
        RAJA::b200_detail::LaunchScope scope(info.grid, stream);
        return data_t(data);
      };
      data_t launch_data = make_data();
      ::b200::LaunchScope hook("kernel_launcher", stream);
      kern<<<dims.dims.blocks, dims.dims.threads, 0, stream>>>(launch_data);
    }
    RAJA::b200_detail::check(cudaGetLastError(), "kernel launch");
    RAJA::expt::detail::resolve_params<POL>(data.param_tuple, info);
    if (!Async)
    {
      res.wait();
      rb200_stream_mark_synced(stream);
    }
  }
};

}  // namespace internal
}  // namespace RAJA

#endif
