#!/bin/bash
# round 2, call 34 (1 GPU): finer cycle counters inside the training forward's epilogue; fused leg through CUDA-graph replay
mkdir -p gpurun_out; P=gpurun_out/c34
timeout 120 tests/csrc/_bin/prof_tc_fwd > ${P}_roles_fwd_infer.txt 2>&1; echo "fwd rc=$?" >> ${P}_summary.txt
timeout 120 tests/csrc/_bin/prof_tc_bwd > ${P}_roles_pair.txt 2>&1; echo "pair rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg fused_tmix_kernels > ${P}_fused_leg.json 2> ${P}_fused_leg.err; echo "fused leg rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -5 ${P}_roles_fwd_infer.txt; sed -n 1,22p ${P}_roles_pair.txt; cat ${P}_fused_leg.json | cut -c1-1800; tail -3 ${P}_fused_leg.err
