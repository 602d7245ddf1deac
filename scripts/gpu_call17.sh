#!/bin/bash
# round 2, call 17 (1 GPU): shift_mix backward through a cp.async ring: parity + timing against the register-pipeline form
mkdir -p gpurun_out; P=gpurun_out/c17
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_varlen_gpu.py tests/test_model_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg fused_tmix_kernels > ${P}_leg_ring.json 2>&1
RWKVTTS_MIX_BWD=4 timeout 300 python bench.py --leg fused_tmix_kernels > ${P}_leg_bwd4.json 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; tail -1 ${P}_leg_ring.json | cut -c1-400; tail -1 ${P}_leg_bwd4.json | cut -c1-400; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json
