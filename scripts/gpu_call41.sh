#!/bin/bash
# round 2, call 41 (1 GPU): the whole bench line on the final tree (fused leg through CUDA-graph replay) + the fused-kernel tests
mkdir -p gpurun_out; P=gpurun_out/c41
timeout 300 python -m pytest tests/test_fused_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' ${P}_bench.json; tail -4 ${P}_bench.err | cut -c1-200
