// RAJA_b200/RAJA.hpp -- the b200 execution back-end for RAJA, reference-side boundary.
//
// Include this INSTEAD of "RAJA/RAJA.hpp" (it includes it).  It adds to an UNMODIFIED RAJA
// (the reference, built with CUDA enabled) a new policy family
//
//     RAJA::b200_exec<BLOCK, Async=false>, RAJA::b200_exec_async<BLOCK>, RAJA::b200_exec_grid<BLOCK, GRID>
//     RAJA::b200_reduce, b200_reduce_atomic (classic ReduceSum/Min/Max/MinLoc/MaxLoc/BitOr/BitAnd objects)
//     RAJA::b200_multi_reduce               (MultiReduceSum/Min/Max/BitOr/BitAnd: run-time indexed bins)
//     RAJA::b200_atomic, b200_atomic_aggregated (atomicAdd & friends)
//     RAJA::b200_synchronize
//     RAJA::b200_launch_t<Async>            (RAJA::launch, with or without expt::Reduce params)
//     RAJA::statement::B200Kernel<...>      (RAJA::kernel / kernel_param, with or without expt::Reduce params)
//     RAJA::b200_multi_exec<BLOCK, Async>   (sharded forall / reducer params / scans across the GPUs of one node; RAJA::b200::MultiResource)
//     RAJA::b200_work<BLOCK, Async>, RAJA::unordered_b200_fused   (WorkPool / WorkGroup / WorkSite)
//     RAJA::b200::{device,device_zeroed,pinned}_mempool_type
//
// usable with RAJA::forall (with or without RAJA::expt::Reduce / ReduceLoc / Name params),
// RAJA::inclusive_scan / exclusive_scan[_inplace], RAJA::sort / stable_sort[_pairs] and the
// RAJA::atomic* functions.  User code changes only the policy template argument; the resource
// type is camp::resources::Cuda with its stream semantics (work is enqueued on res.get_stream();
// synchronous policies wait, Async ones do not; an EventProxy is returned).
//
// RAJA selects a back-end by overload resolution on the policy type:
//   * forall_impl / param_init / param_combine / param_resolve / synchronize_impl are found by ADL
//     and may be declared after RAJA.hpp;
//   * RAJA::impl::scan::*, RAJA::impl::sort::* and RAJA::atomicX(policy, ...) are called QUALIFIED
//     from pattern/{scan,sort,atomic}.hpp, so our overloads must be visible before those headers
//     are first included -- hence the order below (prerequisites, our declarations, RAJA.hpp, rest).
//
// The kernels are the header-only templates in raja_b200/csrc/b200/*.cuh (add that directory's
// parent, raja_b200/csrc, to the include path), instantiated here with the user's lambda and
// operator types; pools and the type-erased runtime come from libraja_b200.so (include/raja_b200.h).
// Compile with nvcc -std=c++17 -arch=sm_100a --extended-lambda --expt-relaxed-constexpr.
#ifndef RAJA_B200_RAJA_HPP
#define RAJA_B200_RAJA_HPP

#ifdef RAJA_HPP
#error "include RAJA_b200/RAJA.hpp before (instead of) RAJA/RAJA.hpp: scan/sort/atomic overloads must precede the pattern headers"
#endif

// ---- 1. prerequisites from the reference ---------------------------------------------------
#include "RAJA/config.hpp"
#if !defined(RAJA_ENABLE_CUDA)
#error "RAJA_b200 needs a CUDA-enabled RAJA (camp::resources::Cuda)"
#endif
#include "RAJA/util/macros.hpp"
#include "RAJA/util/types.hpp"
#include "RAJA/util/Operators.hpp"
#include "RAJA/util/concepts.hpp"
#include "RAJA/policy/PolicyBase.hpp"
#include "RAJA/util/resource.hpp"
#include "RAJA/policy/sequential/policy.hpp"
#include "RAJA/policy/atomic_builtin.hpp"
#include "RAJA/policy/sequential/atomic.hpp"

#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "raja_b200.h"
#include "b200/common.cuh"
#include "b200/reduce.cuh"
#include "b200/forall.cuh"
#include "b200/atomic.cuh"
#include "b200/scan.cuh"
#include "b200/radix_sort.cuh"

// ---- 2. policies, resource mapping, and the overloads the pattern headers look up qualified ----
#include "RAJA_b200/policy.hpp"
#include "RAJA_b200/pools.hpp"
#include "RAJA_b200/atomic.hpp"
#include "RAJA_b200/scan.hpp"
#include "RAJA_b200/sort.hpp"
#include "RAJA_b200/indexset.hpp"

// ---- 3. the reference -------------------------------------------------------------------------
#include "RAJA/RAJA.hpp"

// ---- 4. everything found by ADL -----------------------------------------------------------------
#include "RAJA_b200/forall.hpp"
#include "RAJA_b200/reduce.hpp"
#include "RAJA_b200/multi_reduce.hpp"
#include "RAJA_b200/launch.hpp"
#include "RAJA_b200/kernel.hpp"
#include "RAJA_b200/workgroup.hpp"
#include "RAJA_b200/multi.hpp"

#endif  // RAJA_B200_RAJA_HPP
