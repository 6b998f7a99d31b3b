#!/bin/bash
# compute-sanitizer passes over the parity programs (memcheck everywhere, racecheck + synccheck on the shared-memory heavy kernels)
mkdir -p gpurun_out
out=gpurun_out/sanitizer_r01.txt
: > $out
run() { echo "== $*" >> $out; timeout 900 "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|OK|FAILED|Error|error|hazard" | head -12 >> $out; }
run compute-sanitizer --tool memcheck python __graft_entry__.py smoke
run compute-sanitizer --tool memcheck tests/cpp/_build/test_launch_reduce
run compute-sanitizer --tool memcheck tests/cpp/_build/test_workgroup
run compute-sanitizer --tool memcheck tests/cpp/_build/test_multi_reduce
run compute-sanitizer --tool racecheck python __graft_entry__.py smoke
run compute-sanitizer --tool synccheck python __graft_entry__.py smoke
run compute-sanitizer --tool racecheck tests/cpp/_build/test_launch_reduce
cat $out
