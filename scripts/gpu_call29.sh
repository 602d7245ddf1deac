#!/bin/bash
# round 2, call 29 (1 GPU): add+LayerNorm ring adjoint at up to 7 CTAs per SM, parameter-gradient conversions batched,
# mask conversion cached: fused / model / layout / engine parity, fused call times, train step time
mkdir -p gpurun_out; P=gpurun_out/c29
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_model_gpu.py tests/test_layouts_gpu.py tests/test_varlen_gpu.py tests/test_engine_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 200 python scripts/run_fused.py 4 > ${P}_fused_ring.txt 2>&1; echo "fused ring rc=$?" >> ${P}_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"add_ln" -c 4 python scripts/run_fused.py 1 > ${P}_ln_ncu.txt 2>&1; echo "ncu ln rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; tail -2 ${P}_fused_ring.txt; grep -E "add_ln|duration" ${P}_ln_ncu.txt | head -12
grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json; grep -o '"loss": [0-9.]*' ${P}_bench.json
