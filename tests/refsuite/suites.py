#!/usr/bin/env python
"""Instantiate the reference's OWN typed gtest suites for the b200 back-end (SURVEY.md section 4 / 7
step 8: "zero new test logic").

The reference generates one translation unit per (suite, back-end) from `*.cpp.in` templates with
cmake's configure_file().  This script does the same substitution with BACKEND = B200, writes the
generated sources under tests/refsuite/_build/gen/ (git-ignored: they are derived from the reference
and are NOT part of this repository's history), and writes _build/Makefile, which compiles each with
nvcc against the reference headers, force-including tests/refsuite/b200_lists.hpp (the B200 policy
lists), and links googletest built from the reference's vendored copy.  Test logic, data and
assertions are the reference's, untouched; only the policy lists are ours.

Runs in the build container only (needs /root/reference); the binaries travel to the GPU box and are
run by tests/test_refsuite.py.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("REF", "/root/reference")
T = os.path.join(REF, "test")
GTEST = os.path.join(REF, "blt", "thirdparty_builtin", "googletest-1.16.0", "googletest")
BUILD = os.path.join(HERE, "_build")

SUITES = []


def suite(name, template, subs=None, inject=None, rename=None):
    """name: binary name; template: path under test/ (a .cpp.in or a .cpp); subs: @VAR@ values;
    inject: {anchor substring: text inserted after the first line containing it};
    rename: [(old, new)] plain text replacements (for the non-templated *-cuda.cpp files)"""
    SUITES.append(dict(name=name, template=template, subs=subs or {}, inject=inject or {}, rename=rename or []))


# ---- forall / reduce-basic: classic reducer objects (5 core + 2 bitwise) and expt::Reduce params (7 + 2)
D = "functional/forall/reduce-basic/"
for r in ("ReduceSum", "ReduceMin", "ReduceMax", "ReduceMinLoc", "ReduceMaxLoc"):
    suite(f"forall-basic-{r}", D + "test-forall-basic-reduce.cpp.in", dict(BACKEND="B200", REDUCETYPE=r, DATATYPES="CoreReductionDataTypeList"))
for r in ("ReduceBitAnd", "ReduceBitOr"):
    suite(f"forall-basic-{r}", D + "test-forall-basic-reduce.cpp.in", dict(BACKEND="B200", REDUCETYPE=r, DATATYPES="BitwiseReductionDataTypeList"))
for r in ("ReduceSum", "ReduceMin", "ReduceMax", "ReduceMaxLoc", "ReduceMinLoc", "ReduceMaxLocAlt", "ReduceMinLocAlt"):
    suite(f"forall-basic-expt-{r}", D + "test-forall-basic-expt-reduce.cpp.in", dict(BACKEND="B200", REDUCETYPE=r, DATATYPES="CoreReductionDataTypeList"))
for r in ("ReduceBitAnd", "ReduceBitOr"):
    suite(f"forall-basic-expt-{r}", D + "test-forall-basic-expt-reduce.cpp.in", dict(BACKEND="B200", REDUCETYPE=r, DATATYPES="BitwiseReductionDataTypeList"))


# ---- forall / reduce-multiple-segment: several classic reducers in one loop, repeated launches
D = "functional/forall/reduce-multiple-segment/"
for r in ("ReduceSum", "ReduceMin", "ReduceMax", "ReduceMinLoc", "ReduceMaxLoc"):
    suite(f"forall-segment-multiple-{r}", D + "test-forall-segment-multiple-reduce.cpp.in", dict(BACKEND="B200", REDUCETYPE=r))

# ---- forall / resource-segment: explicit camp resource, the three segment types, strongly typed indices
D = "functional/forall/resource-segment/"
for seg in ("ListSegment", "RangeSegment", "RangeStrideSegment"):
    suite(f"forall-resource-{seg}", D + "test-forall-resource-segment.cpp.in", dict(BACKEND="B200", SEGTYPE=seg))

# ---- forall / atomic-basic and atomic-ref
suite("forall-atomic-basic", "functional/forall/atomic-basic/test-forall-atomic-basic.cpp.in", dict(ATOMIC_BACKEND="B200"))
for t in ("AtomicRefAdd", "AtomicRefSub", "AtomicRefLoadStore", "AtomicRefCAS", "AtomicRefMinMax", "AtomicRefLogical"):
    suite(f"forall-{t}", "functional/forall/atomic-ref/test-forall-atomicref.cpp.in", dict(ATOMIC_BACKEND="B200", TEST=t))

# ---- forall / multi-reduce-basic
for r in ("Sum", "Min", "Max", "BitAnd", "BitOr"):
    suite(f"forall-basic-MultiReduce{r}", "functional/forall/multi-reduce-basic/test-forall-basic-multi-reduce.cpp.in", dict(BACKEND="B200", REDUCETYPE=r))

# ---- scans
for t in ("Exclusive", "ExclusiveInplace", "Inclusive", "InclusiveInplace"):
    suite(f"scan-{t}", "functional/scan/test-scan.cpp.in", dict(SCAN_BACKEND="B200", SCAN_TYPE=t))

# ---- unit / algorithm: sort and stable sort (keys and pairs, every key type, ascending / descending, resource and no-resource interfaces)
SORTERS = """
using B200SortSorters = camp::list<PolicySort<RAJA::b200_exec<128>>, PolicySortPairs<RAJA::b200_exec<128>>, PolicySort<RAJA::b200_exec_grid<128, 64>>>;
"""
STABLE_SORTERS = """
using B200StableSortSorters = camp::list<PolicyStableSort<RAJA::b200_exec<128>>, PolicyStableSortPairs<RAJA::b200_exec<128>>, PolicyStableSort<RAJA::b200_exec_grid<128, 64>>>;
"""
suite("algorithm-sort", "unit/algorithm/test-algorithm-sort.cpp.in", dict(SORT_BACKEND="B200"), inject={'#include "test-algorithm-sort.hpp"': SORTERS})
suite("algorithm-stable-sort", "unit/algorithm/test-algorithm-stable-sort.cpp.in", dict(SORT_BACKEND="B200"),
      inject={'#include "test-algorithm-stable-sort.hpp"': STABLE_SORTERS})

# ---- unit / resource: stream semantics of asynchronous foralls
for t in ("Depends", "MultiStream", "AsyncTime", "BasicAsyncSemantics", "JoinAsyncSemantics"):
    suite(f"resource-{t}", "unit/resource/test-resource.cpp.in", dict(BACKEND="B200", TESTTYPE=t))

# ---- unit / reducer: constructors (incl. RAJA::tuple loc types) and reset; the reference keeps one hand-written
#      file per back-end, so the CUDA one is re-targeted by renaming its lists
UNIT_LISTS = """
using B200ReducerPolicyList = camp::list<RAJA::b200_reduce, RAJA::b200_reduce_atomic>;
using B200UnitTestPolicyList = CudaUnitTestPolicyList;  // how ONE device thread runs a test body (forone<test_cuda>), not a back-end
"""
REN = [("CudaReducerPolicyList", "B200ReducerPolicyList"), ("CudaResourceList", "B200ResourceList"), ("CudaUnitTestPolicyList", "B200UnitTestPolicyList"),
       ("CudaBasic", "B200Basic"), ("CudaInit", "B200Init"), ("CudaReset", "B200Reset"), ("CudaReducerResetTypes", "B200ReducerResetTypes")]
suite("reducer-constructors", "unit/reducer/test-reducer-constructors-cuda.cpp", rename=REN, inject={'#include "tests/test-reducer-constructors.hpp"': UNIT_LISTS})
suite("reducer-reset", "unit/reducer/test-reducer-reset-cuda.cpp", rename=REN, inject={'#include "tests/test-reducer-reset.hpp"': UNIT_LISTS})

# ---- launch / reduce-params
LAUNCH_LISTS = """
using b200_policies = camp::list<RAJA::LaunchPolicy<RAJA::b200_launch_t<true>>, RAJA::LoopPolicy<RAJA::cuda_global_thread_x>>;
using b200_sync_policies = camp::list<RAJA::LaunchPolicy<RAJA::b200_launch_t<false>>, RAJA::LoopPolicy<RAJA::cuda_global_thread_x>>;
using B200_launch_policies = camp::list<b200_policies, b200_sync_policies>;
"""
D = "functional/launch/reduce-params/"
for r, dt in (("ReduceSum", "CoreReductionDataTypeList"), ("ReduceMin", "CoreReductionDataTypeList"), ("ReduceBitAnd", "BitwiseReductionDataTypeList")):
    suite(f"launch-basic-param-expt-{r}", D + "test-launch-basic-param-expt-reduce.cpp.in", dict(BACKEND="B200", REDUCETYPE=r, DATATYPES=dt),
          inject={'#include "RAJA_test-launch-execpol.hpp"': LAUNCH_LISTS})

# ---- kernel / reduce-params: RAJA::kernel_param with expt::Reduce / ReduceLoc parameters (the four loop nests of the CUDA list)
KERNEL_LISTS = """
using B200KernelLocExecPols =
  camp::list<
    RAJA::KernelPolicy<RAJA::statement::B200Kernel<RAJA::statement::For<1, RAJA::cuda_thread_x_loop, RAJA::statement::For<0, RAJA::cuda_thread_y_loop, RAJA::statement::Lambda<0>>>>>,
    RAJA::KernelPolicy<RAJA::statement::For<1, RAJA::seq_exec, RAJA::statement::B200Kernel<RAJA::statement::For<0, RAJA::cuda_thread_x_loop, RAJA::statement::Lambda<0>>>>>,
    RAJA::KernelPolicy<RAJA::statement::B200Kernel<RAJA::statement::For<1, RAJA::seq_exec, RAJA::statement::For<0, RAJA::cuda_thread_x_loop, RAJA::statement::Lambda<0>>>>>,
    RAJA::KernelPolicy<RAJA::statement::B200Kernel<RAJA::statement::For<1, RAJA::cuda_thread_x_loop, RAJA::statement::For<0, RAJA::seq_exec, RAJA::statement::Lambda<0>>>>>,
    RAJA::KernelPolicy<RAJA::statement::B200KernelFixed<256, RAJA::statement::For<1, RAJA::cuda_block_x_loop, RAJA::statement::For<0, RAJA::cuda_thread_x_loop, RAJA::statement::Lambda<0>>>>>
  >;
using B200KernelLocReducePols = camp::list<RAJA::b200_reduce>;
using B200KernelLocForallPols = camp::list<RAJA::b200_exec<1024, true>>;
"""
suite("kernel-basic-param", "functional/kernel/reduce-params/test-kernel-basic-param.cpp.in", dict(LOC_BACKEND="B200"),
      inject={'#include "test-kernel-basic-param.hpp"': KERNEL_LISTS})

# ---- workgroup (functional): ordered / reverse_ordered / fused unordered runners, three storage policies, dispatchers
WG_LISTS = """
using B200ExecPolicyList = camp::list<RAJA::b200_work<1024>, RAJA::b200_work<256>>;
using B200OrderedPolicyList = SequentialOrderedPolicyList;
using B200OrderPolicyList = camp::list<RAJA::ordered, RAJA::reverse_ordered, RAJA::unordered_b200_fused>;
using B200StoragePolicyList = SequentialStoragePolicyList;
using B200AllocatorList = CudaAllocatorList;  // pinned allocations through the camp Cuda resource, as the reference's CUDA tests
"""
D = "functional/workgroup/"
for kind in ("Ordered", "Unordered"):
    for sub in ("Single", "MultipleReuse"):
        for disp in ("IndirectFunction", "Direct"):
            suite(f"workgroup-{kind}-{sub}-{disp}", D + f"test-workgroup-{kind}.cpp.in", dict(BACKEND="B200", SUBTESTNAME=sub, DISPATCHER=disp),
                  inject={f'#include "test-workgroup-{kind}-': WG_LISTS})


# ---- forall / segment, segment-view: the plain loop over the three segment types, every index type, with and without Views
for seg in ("ListSegment", "RangeSegment", "RangeStrideSegment"):
    suite(f"forall-{seg}", "functional/forall/segment/test-forall-segment.cpp.in", dict(BACKEND="B200", SEGTYPE=seg))
for sv in ("ListSegmentView", "RangeSegmentView", "RangeSegment2DView", "RangeStrideSegmentView"):
    suite(f"forall-{sv}", "functional/forall/segment-view/test-forall-segment-view.cpp.in", dict(BACKEND="B200", SEGVIEWTYPE=sv))

# ---- forall over TypedIndexSets (ExecPolicy<seq_segit, b200_exec>): plain, icount, with Views, with an explicit resource,
#      with several reducers (the one-wait dispatch of RAJA_b200/indexset.hpp)
for t in ("IndexSet", "IcountIndexSet"):
    suite(f"forall-{t}", "functional/forall/indexset/test-forall-indexset.cpp.in", dict(BACKEND="B200", INDEXSETTESTTYPE=t))
for t in ("IndexSetView", "IcountIndexSetView"):
    suite(f"forall-{t}", "functional/forall/indexset-view/test-forall-indexset-view.cpp.in", dict(BACKEND="B200", INDEXSETTESTTYPE=t))
for t in ("ResourceIndexSet", "ResourceIcountIndexSet"):
    suite(f"forall-{t}", "functional/forall/resource-indexset/test-forall-resource-indexset.cpp.in", dict(BACKEND="B200", INDEXSETTESTTYPE=t))
for r in ("ReduceSum", "ReduceMin", "ReduceMax", "ReduceMinLoc", "ReduceMaxLoc"):
    suite(f"forall-indexset-multiple-{r}", "functional/forall/reduce-multiple-indexset/test-forall-indexset-multiple-reduce.cpp.in",
          dict(BACKEND="B200", REDUCETYPE=r))

# ---- forall / atomic-view: atomics through RAJA::make_atomic_view / MultiView
for t in ("AtomicView", "AtomicMultiView"):
    suite(f"forall-{t}", "functional/forall/atomic-view/test-forall-atomic-view.cpp.in", dict(ATOMIC_BACKEND="B200", TEST=t))

# ---- unit / multi_reducer: constructors and reset
MR_LISTS = """
using B200MultiReducerPolicyList = camp::list<RAJA::b200_multi_reduce>;
using B200UnitTestPolicyList = CudaUnitTestPolicyList;  // how ONE device thread runs a test body (forone<test_cuda>), not a back-end
"""
suite("multi-reducer-constructors", "unit/multi_reducer/test-multi-reducer-constructors.cpp.in", dict(BACKEND="B200"),
      inject={'#include "test-multi-reducer-constructors.hpp"': MR_LISTS})
suite("multi-reducer-reset", "unit/multi_reducer/test-multi-reducer-reset.cpp.in", dict(BACKEND="B200"),
      inject={'#include "test-multi-reducer-reset.hpp"': MR_LISTS})

# ---- unit / workgroup: constructors, enqueue, dispatcher
for disp in ("Direct", "IndirectFunction"):
    suite(f"workgroup-unit-Constructor-{disp}", "unit/workgroup/test-workgroup-Constructor.cpp.in", dict(BACKEND="B200", SUBTESTNAME="Single", DISPATCHER=disp),
          inject={'#include "test-workgroup-Constructor.hpp"': WG_LISTS})
    for sub in ("Single", "Multiple"):
        suite(f"workgroup-unit-Enqueue-{sub}-{disp}", "unit/workgroup/test-workgroup-Enqueue.cpp.in", dict(BACKEND="B200", SUBTESTNAME=sub, DISPATCHER=disp),
              inject={'#include "test-workgroup-Enqueue-': WG_LISTS})
    # (unit/workgroup Dispatcher is not instantiated: it drives RAJA::detail::Dispatcher<platform, ...> with the reference's
    #  internal call signature; the b200 fused runner passes the tile index as a leading call argument, RAJA_b200/workgroup.hpp.
    #  Dispatch itself is covered by the functional workgroup suites above.)


# ---- launch / reduce-basic: CLASSIC reducer objects captured by a RAJA::launch body and re-captured by RAJA::loop lambdas
#      (first-level device copies run the grid combine, nested copies fold into their parent); launch / segment
D = "functional/launch/reduce-basic/"
for r, dt in (("ReduceSum", "CoreReductionDataTypeList"), ("ReduceMin", "CoreReductionDataTypeList"), ("ReduceBitAnd", "BitwiseReductionDataTypeList")):
    suite(f"launch-basic-{r}", D + "test-launch-basic-reduce.cpp.in", dict(BACKEND="B200", REDUCETYPE=r, DATATYPES=dt),
          inject={'#include "RAJA_test-launch-execpol.hpp"': LAUNCH_LISTS})
for seg in ("ListSegment", "RangeSegment", "RangeStrideSegment"):
    suite(f"launch-{seg}", "functional/launch/segment/test-launch-segment.cpp.in", dict(BACKEND="B200", SEGTYPES=seg),
          inject={'#include "RAJA_test-launch-execpol.hpp"': LAUNCH_LISTS})

# ---- kernel / reduce-loc: classic ReduceMinLoc / ReduceMaxLoc (2-D index structs, Views, RAJA::tuple locations) captured by
#      RAJA::kernel lambdas inside a B200Kernel
for t in ("Min2D", "Max2D", "Min2DView", "Max2DView", "Min2DViewTuple", "Max2DViewTuple"):
    suite(f"kernel-reduceloc-{t}", "functional/kernel/reduce-loc/test-kernel-reduceloc.cpp.in", dict(LOC_BACKEND="B200", LOC_TYPE=t),
          inject={f'#include "test-kernel-reduceloc-{t}.hpp"': KERNEL_LISTS})


def generate():
    gen = os.path.join(BUILD, "gen")
    os.makedirs(gen, exist_ok=True)
    rules = []
    for s in SUITES:
        src = os.path.join(T, s["template"])
        text = open(src).read()
        for k, v in s["subs"].items():
            text = text.replace(f"@{k}@", v)
        for old, new in s["rename"]:
            text = text.replace(old, new)
        left = re.findall(r"@[A-Z_]+@", text)
        if left:
            raise SystemExit(f"{s['name']}: unresolved placeholders {sorted(set(left))}")
        for anchor, extra in s["inject"].items():
            lines = text.split("\n")
            for i, ln in enumerate(lines):
                if anchor in ln:
                    lines.insert(i + 1, extra)
                    break
            else:
                raise SystemExit(f"{s['name']}: anchor {anchor!r} not found in {src}")
            text = "\n".join(lines)
        out = os.path.join(gen, s["name"] + ".cpp")
        if not os.path.exists(out) or open(out).read() != text:
            open(out, "w").write(text)
        incs = [os.path.join(os.path.dirname(src), "tests"), os.path.dirname(src)]
        rules.append((s["name"], out, incs))
    mk = []
    mk.append(f"""# generated by tests/refsuite/suites.py -- do not edit
REF   := {REF}
ROOT  := {ROOT}
NVCC  := /usr/local/cuda/bin/nvcc
GTEST := {GTEST}
ARCH  := -gencode arch=compute_100a,code=sm_100a
FLAGS := -std=c++17 -O1 $(ARCH) -ccbin /usr/bin/g++ --extended-lambda --expt-relaxed-constexpr \\
         -Xcompiler -fopenmp -diag-suppress 20012,20011,177,550,186,128,2810 -w \\
         -I$(ROOT)/include -I$(ROOT)/raja_b200/csrc -I$(ROOT)/oracle/_ref/include_cuda -I$(REF)/include -I$(REF)/tpl/camp/include \\
         -I$(REF)/test/include -I$(GTEST)/include -include $(ROOT)/tests/refsuite/b200_lists.hpp
LIBS  := -L$(ROOT)/raja_b200/lib -lraja_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../../../../raja_b200/lib' -lgomp -lpthread
DEPS  := $(wildcard $(ROOT)/include/RAJA_b200/*.hpp) $(wildcard $(ROOT)/raja_b200/csrc/b200/*.cuh) $(ROOT)/tests/refsuite/b200_lists.hpp $(ROOT)/include/raja_b200.h
REFLIB := $(ROOT)/tests/cpp/_build/libref.a
BINS  := {' '.join('bin/' + r[0] for r in rules)}
all: $(BINS)
libgtest.a:
\tg++ -std=c++17 -O1 -I$(GTEST)/include -I$(GTEST) -c $(GTEST)/src/gtest-all.cc -o gtest-all.o
\tg++ -std=c++17 -O1 -I$(GTEST)/include -I$(GTEST) -c $(GTEST)/src/gtest_main.cc -o gtest_main.o
\tar rcs $@ gtest-all.o gtest_main.o
""")
    for name, out, incs in rules:
        inc = " ".join("-I" + i for i in incs)
        # compile, link, strip and drop the object: the objects would triple the snapshot that travels to the GPU box
        mk.append(f"bin/{name}: {out} libgtest.a $(REFLIB) $(DEPS)\n\t@mkdir -p bin obj\n"
                  f"\t$(NVCC) $(FLAGS) {inc} -x cu -c {out} -o obj/{name}.o\n"
                  f"\t$(NVCC) $(ARCH) -ccbin /usr/bin/g++ obj/{name}.o libgtest.a $(REFLIB) $(LIBS) -o $@\n"
                  f"\t@strip $@; rm -f obj/{name}.o\n")
    mk.append(".PHONY: all\n")
    path = os.path.join(BUILD, "Makefile")
    text = "\n".join(mk)
    if not os.path.exists(path) or open(path).read() != text:
        open(path, "w").write(text)
    with open(os.path.join(BUILD, "suites.txt"), "w") as f:
        f.write("\n".join(r[0] for r in rules) + "\n")
    return [r[0] for r in rules]


if __name__ == "__main__":
    if not os.path.exists(os.path.join(T, "include", "RAJA_test-base.hpp")):
        print("reference tree not present: keeping prebuilt tests/refsuite/_build (if any)")
        sys.exit(0)
    names = generate()
    print(f"refsuite: {len(names)} suites generated under {BUILD}")
