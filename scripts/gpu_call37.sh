#!/bin/bash
# round 2, call 37 (1 GPU): forward issues S^ += V^T K~ behind phase 1's commit (it runs while the MMA warp waits for U^T):
# parity (op, packed, stress), op times
mkdir -p gpurun_out; P=gpurun_out/c37
timeout 900 python -m pytest tests/test_wkv7_gpu.py tests/test_varlen_gpu.py tests/test_stress_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg wkv_ops > ${P}_wkv_ops.json 2> ${P}_wkv_ops.err; echo "wkv_ops rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; cut -c1-330 ${P}_wkv_ops.json
