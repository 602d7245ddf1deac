#!/bin/bash
# round 2, call 31 (1 GPU): stage A of both chunked kernels with the conflict-free thread mapping (4 tokens x 8 channel quads per warp,
# opposite channel halves on odd / even tokens) and the per-channel decay prefix: parity (op, packed, stress), op and step times
mkdir -p gpurun_out; P=gpurun_out/c31
timeout 1200 python -m pytest tests/test_wkv7_gpu.py tests/test_varlen_gpu.py tests/test_stress_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg wkv_ops > ${P}_wkv_ops.json 2> ${P}_wkv_ops.err; echo "wkv_ops rc=$?" >> ${P}_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:"wkv7_tc|add_ln" -c 8 python scripts/run_pair.py 1 > ${P}_pair_ncu.txt 2>&1; echo "ncu pair rc=$?" >> ${P}_summary.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"add_ln" -c 4 python scripts/run_fused.py 1 > ${P}_ln_ncu.txt 2>&1; echo "ncu ln rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; cut -c1-330 ${P}_wkv_ops.json; grep -E "wkv7_tc|add_ln|duration|wavefronts|conflicts" ${P}_pair_ncu.txt ${P}_ln_ncu.txt | grep -v PROF | cut -c1-150 | head -30
grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json; grep -o '"loss": [0-9.]*' ${P}_bench.json
