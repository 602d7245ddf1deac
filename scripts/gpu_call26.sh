#!/bin/bash
# round 2, call 26 (1 GPU): residual add deferred into the next block's norm: model / layout / varlen parity + step time
mkdir -p gpurun_out; P=gpurun_out/c26
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_layouts_gpu.py tests/test_varlen_gpu.py tests/test_head_gpu.py tests/test_decode_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json; grep -o '"loss": [0-9.]*' ${P}_bench.json
