#!/bin/bash
# round 2, call 27 (1 GPU): tensor-map (TMA) staging of the raw input tiles in the chunked WKV kernels: op parity
# (dense, packed, stress) + op timings
mkdir -p gpurun_out; P=gpurun_out/c27
timeout 900 python -m pytest tests/test_wkv7_gpu.py tests/test_varlen_gpu.py tests/test_stress_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg wkv_ops > ${P}_wkv_ops.json 2> ${P}_wkv_ops.err; echo "wkv_ops rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -5 ${P}_pytest.log | cut -c1-300; cut -c1-700 ${P}_wkv_ops.json; tail -3 ${P}_wkv_ops.err
