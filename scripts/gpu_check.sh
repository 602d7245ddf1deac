#!/bin/bash
# One GPU-box pass: parity tests, role profilers, bench, ncu launch list (+ optional full capture).
# usage: gpurun -- 'bash scripts/gpu_check.sh [tests] [prof] [bench] [ncu] [full]'
set -u
mkdir -p gpurun_out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DRWKVTTS_PROFILE"
C=rwkvtts_b200/csrc
for what in "$@"; do
case $what in
tests) timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/tests.log; tail -5 gpurun_out/tests.log ;;
prof)
  $NV -o tests/csrc/prof_tc_fwd tests/csrc/prof_tc_fwd.cu $C/wkv7_tc_fwd.cu 2> gpurun_out/prof_build.log && timeout 120 tests/csrc/prof_tc_fwd > gpurun_out/prof_tc_fwd.txt 2>&1
  $NV -o tests/csrc/prof_tc_bwd tests/csrc/prof_tc_bwd.cu $C/wkv7_tc_fwd.cu $C/wkv7_tc_bwd.cu 2>> gpurun_out/prof_build.log && timeout 120 tests/csrc/prof_tc_bwd > gpurun_out/prof_tc_bwd.txt 2>&1
  cat gpurun_out/prof_tc_fwd.txt gpurun_out/prof_tc_bwd.txt ;;
bench) timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json ;;
ncu) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-decode --no-model-step --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300 ;;
full) timeout 900 ncu --set full --clock-control none --import-source on -k regex:wkv7_tc -s 3 -c 3 -o gpurun_out/tc_full -f python scripts/run_pair.py 2 > gpurun_out/full.log 2>&1; tail -3 gpurun_out/full.log ;;
esac
done
