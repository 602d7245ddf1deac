#!/bin/bash
# round 2, call 35 (1 GPU): training forward with the U^T hand-off ahead of the Y staging / U stores in the epilogue:
# parity (op, packed, stress), op times, role counters; fused leg through CUDA-graph replay
mkdir -p gpurun_out; P=gpurun_out/c35
timeout 900 python -m pytest tests/test_wkv7_gpu.py tests/test_varlen_gpu.py tests/test_stress_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg wkv_ops > ${P}_wkv_ops.json 2> ${P}_wkv_ops.err; echo "wkv_ops rc=$?" >> ${P}_summary.txt
timeout 120 tests/csrc/_bin/prof_tc_bwd > ${P}_roles_pair.txt 2>&1; echo "pair rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg fused_tmix_kernels > ${P}_fused_leg.json 2> ${P}_fused_leg.err; echo "fused leg rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; cut -c1-330 ${P}_wkv_ops.json; sed -n 1,22p ${P}_roles_pair.txt; cut -c1-420 ${P}_fused_leg.json
