#!/bin/bash
# round 2, call 28 (1 GPU): add+LayerNorm adjoint through a cp.async ring (parity, time against the warp-per-row form),
# fresh ncu --set full of the tcgen05 pair with the tensor-map staging in the backward, train step time
mkdir -p gpurun_out; P=gpurun_out/c28
timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_model_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 200 python scripts/run_fused.py 4 > ${P}_fused_ring.txt 2>&1; echo "fused ring rc=$?" >> ${P}_summary.txt
RWKVTTS_LN_BWD=1 timeout 200 python scripts/run_fused.py 4 > ${P}_fused_warp.txt 2>&1; echo "fused warp rc=$?" >> ${P}_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wkv7_tc -s 3 -c 3 -o ${P}_tc_full -f python scripts/run_pair.py 2 > ${P}_full.log 2>&1; echo "ncu wkv rc=$?" >> ${P}_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"add_ln_bwd" -c 2 -o ${P}_ln_full -f python scripts/run_fused.py 1 > ${P}_ln.log 2>&1; echo "ncu ln rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; tail -2 ${P}_fused_ring.txt; tail -2 ${P}_fused_warp.txt
grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json; grep -o '"loss": [0-9.]*' ${P}_bench.json
