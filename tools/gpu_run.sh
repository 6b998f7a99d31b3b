#!/bin/bash
# tools/gpu_run.sh WHAT...  -- run on the GPU box (via gpurun); writes everything under gpurun_out/<tag>/
#   tests     pytest -m gpu (everything)          bench     python bench.py (short)
#   refsuite  only tests/test_refsuite.py          adapter   only tests/test_raja_adapter.py
#   ncu-list  launch list of one bench step        paths     tools/run_path.py table
TAG=${TAG:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for what in "$@"; do
  case $what in
    tests)    timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "tests rc=$?" ;;
    tests-all) timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.txt 2>&1; echo "tests rc=$?" ;;
    refsuite) timeout 2400 python -m pytest tests/test_refsuite.py -m gpu -q > $OUT/pytest_refsuite.txt 2>&1; echo "refsuite rc=$?" ;;
    adapter)  timeout 1200 python -m pytest tests/test_raja_adapter.py -m gpu -q > $OUT/pytest_adapter.txt 2>&1; echo "adapter rc=$?" ;;
    parity)   timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q > $OUT/pytest_parity.txt 2>&1; echo "parity rc=$?" ;;
    smoke)    timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?" ;;
    bench)    timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench.json ;;
    bench-ref) timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench-ref rc=$?" ;;
    ncu-list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify --no-reference-gpu > $OUT/ncu_bench.log 2>&1; echo "ncu-list rc=$?" ;;
    multi-cpp) timeout 600 tests/cpp/_build/test_multi > $OUT/test_multi_world${GPUS:-2}.txt 2>&1; echo "multi-cpp rc=$?"; tail -3 $OUT/test_multi_world${GPUS:-2}.txt ;;
    multi-py) timeout 900 python -m pytest tests/test_multi_gloo.py -m gpu -q > $OUT/pytest_multi_gpu.txt 2>&1; echo "multi-py rc=$?"; tail -3 $OUT/pytest_multi_gpu.txt ;;
    bench-n)  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${GPUS:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${GPUS:-2} --steps ${STEPS:-10} --warmup 3 > $OUT/bench_n${GPUS:-2}.json 2> $OUT/bench_n${GPUS:-2}.err; echo "bench-n rc=$?"; tail -c 1500 $OUT/bench_n${GPUS:-2}.json; tail -5 $OUT/bench_n${GPUS:-2}.err ;;
    bench-n-nccl) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${GPUS:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus ${GPUS:-2} --steps ${STEPS:-10} --warmup 3 --nccl-exchange --no-e2e > $OUT/bench_n${GPUS:-2}_nccl.json 2> $OUT/bench_n${GPUS:-2}_nccl.err; echo "bench-n-nccl rc=$?"; tail -c 600 $OUT/bench_n${GPUS:-2}_nccl.json ;;
    sanitize) for t in forall-basic-ReduceSum forall-basic-expt-ReduceMinLoc forall-segment-multiple-ReduceMaxLoc scan-Inclusive algorithm-stable-sort forall-atomic-basic launch-basic-param-expt-ReduceSum kernel-basic-param reducer-reset resource-MultiStream workgroup-Unordered-Single-Direct; do
                echo "== memcheck $t"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 tests/refsuite/_build/bin/$t --gtest_brief=1 2>&1 | tail -4; done > $OUT/sanitizer_memcheck.txt 2>&1
              for t in forall-basic-ReduceSum scan-Inclusive kernel-basic-param; do
                echo "== racecheck $t"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 tests/refsuite/_build/bin/$t --gtest_brief=1 2>&1 | tail -4; done > $OUT/sanitizer_racecheck.txt 2>&1
              echo "sanitize done"; grep -c "ERROR SUMMARY: 0 errors" $OUT/sanitizer_memcheck.txt $OUT/sanitizer_racecheck.txt ;;
    ncu-full) B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-verify --no-reference-gpu"
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep_late_kernel -s 9 -c 1 -f -o $OUT/ncu_onesweep_u32 $B > $OUT/ncu_onesweep.log 2>&1; echo "ncu onesweep rc=$?"
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 6 -c 1 -f -o $OUT/ncu_scan_i32 $B > $OUT/ncu_scan.log 2>&1; echo "ncu scan rc=$?"
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:forall_kernel -s 6 -c 1 -f -o $OUT/ncu_forall_dot $B > $OUT/ncu_forall.log 2>&1; echo "ncu forall rc=$?"
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep_late_kernel -s 2 -c 1 -f -o $OUT/ncu_onesweep_f64 python tools/run_path.py --path sortf64 --reps 1 > $OUT/ncu_onesweep_f64.log 2>&1; echo "ncu onesweep f64 rc=$?"
              # the reports themselves (tens of MB each with sources) do not fit the 64 MiB that travels back: keep the text pages
              for r in $OUT/*.ncu-rep; do b=${r%.ncu-rep}; ncu -i $r --page details > ${b}_details.txt 2>&1; ncu -i $r --page raw --csv > ${b}_raw.csv 2>&1; rm -f $r; done
              du -sh $OUT ;;
    paths)    timeout 900 python tools/run_path.py > $OUT/paths.txt 2>&1; echo "paths rc=$?" ;;
    *)        echo "running: $what"; timeout 1800 bash -c "$what" > $OUT/cmd_$(date +%s).txt 2>&1; echo "rc=$?" ;;
  esac
done
