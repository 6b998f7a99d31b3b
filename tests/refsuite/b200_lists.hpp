// Force-included (-include) in front of every translation unit of the reference's OWN typed gtest
// suites (tests/refsuite/suites.py): brings in the b200 back-end and the "B200" policy lists the
// reference's test templates ask for once their @BACKEND@ placeholder is set to B200
// (the reference defines the Sequential / OpenMP / Cuda / Hip lists in test/include/RAJA_test-*.hpp).
// TEST INFRASTRUCTURE; nothing here is product code.
#ifndef RAJA_B200_REFSUITE_LISTS_HPP
#define RAJA_B200_REFSUITE_LISTS_HPP

#include "RAJA_b200/RAJA.hpp"
#include "camp/list.hpp"

// test/include/RAJA_test-camp.hpp
using B200ResourceList = camp::list<camp::resources::Cuda>;

// test/include/RAJA_test-forall-execpol.hpp:108-117 (block sizes and variants of the Cuda list)
using B200ForallExecPols = camp::list<RAJA::b200_exec<128>,
                                      RAJA::b200_exec<256>,
                                      RAJA::b200_exec_grid<256, 64>,
                                      RAJA::b200_exec<1024>>;
using B200ForallReduceExecPols = B200ForallExecPols;
using B200ForallAtomicExecPols = B200ForallExecPols;

// test/include/RAJA_test-forall-indexset-execpol.hpp:50-54
using B200ForallIndexSetExecPols = camp::list<RAJA::ExecPolicy<RAJA::seq_segit, RAJA::b200_exec<128>>, RAJA::ExecPolicy<RAJA::seq_segit, RAJA::b200_exec<256>>,
                                              RAJA::ExecPolicy<RAJA::seq_segit, RAJA::b200_exec_async<256>>>;
using B200ForallIndexSetReduceExecPols = B200ForallIndexSetExecPols;

// test/include/RAJA_test-forall-async-execpol.hpp
using B200AsyncForallExecPols = camp::list<RAJA::b200_exec_async<128>, RAJA::b200_exec_async<256>>;

// test/include/RAJA_test-reducepol.hpp
using B200ReducePols = camp::list<RAJA::b200_reduce, RAJA::b200_reduce_atomic>;

// test/include/RAJA_test-atomicpol.hpp
using B200AtomicPols = camp::list<RAJA::b200_atomic, RAJA::b200_atomic_explicit<RAJA::builtin_atomic>>;

// test/include/RAJA_test-multi-reducepol.hpp
using B200MultiReducePols = camp::list<RAJA::b200_multi_reduce>;

#endif
