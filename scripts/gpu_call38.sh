#!/bin/bash
# round 2, call 38 (1 GPU): WKV-7 + its output stage as one autograd node (rwkvtts_wkv7_backward_acc adds dq / dk / dv onto the
# output stage's gradients): parity (op, packed, stress, fused, model, layouts, engine), op times, train step with and without
mkdir -p gpurun_out; P=gpurun_out/c38
timeout 1200 python -m pytest tests/test_wkv7_gpu.py tests/test_varlen_gpu.py tests/test_stress_gpu.py tests/test_fused_gpu.py tests/test_model_gpu.py tests/test_layouts_gpu.py tests/test_engine_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg wkv_ops > ${P}_wkv_ops.json 2> ${P}_wkv_ops.err; echo "wkv_ops rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
RWKVTTS_FUSE_WKV_OUT=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench_two_nodes.json 2> ${P}_bench_two_nodes.err; echo "bench two nodes rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; cut -c1-330 ${P}_wkv_ops.json
for f in ${P}_bench.json ${P}_bench_two_nodes.json; do grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' $f; grep -o '"loss": [0-9.]*' $f; grep -o '"wkv_bwd_ms": [0-9.]*' $f; done
