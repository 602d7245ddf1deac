#!/bin/bash
# round 2, call 19 (1 GPU): gradients into the flat buffer per bucket: engine tests + step time
mkdir -p gpurun_out; P=gpurun_out/c19
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_bench_dist.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench2.json 2> ${P}_bench2.err
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json ${P}_bench2.json; grep -o '"loss": [0-9.]*' ${P}_bench.json; grep -o '"gpu_launches": [0-9]*' ${P}_bench.json
