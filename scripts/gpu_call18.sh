#!/bin/bash
# round 2, call 18 (1 GPU): ring forms of shift_mix / prep backward: parity + timing (env switches select the earlier forms)
mkdir -p gpurun_out; P=gpurun_out/c18
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_varlen_gpu.py tests/test_model_gpu.py tests/test_layouts_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg fused_tmix_kernels > ${P}_leg_ring.json 2>&1
RWKVTTS_MIX_BWD=4 RWKVTTS_PREP_BWD=1 timeout 300 python bench.py --leg fused_tmix_kernels > ${P}_leg_old.json 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; tail -1 ${P}_leg_ring.json | cut -c1-420; tail -1 ${P}_leg_old.json | cut -c1-420; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json
