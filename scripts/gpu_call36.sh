#!/bin/bash
# round 2, call 36 (1 GPU): backward MMA warp issues the products that need neither group C1's tiles nor S0^T in C1's shadow:
# parity (op, packed, stress, fused incl. the C = 2048 ring case), op times, train step
mkdir -p gpurun_out; P=gpurun_out/c36
timeout 900 python -m pytest tests/test_wkv7_gpu.py tests/test_varlen_gpu.py tests/test_stress_gpu.py tests/test_fused_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg wkv_ops > ${P}_wkv_ops.json 2> ${P}_wkv_ops.err; echo "wkv_ops rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; cut -c1-330 ${P}_wkv_ops.json
grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json; grep -o '"loss": [0-9.]*' ${P}_bench.json
