#!/bin/bash
# round 2, call 42 (1 GPU): Adam update with the hardware square root / reciprocal approximations: parity against
# torch.optim.AdamW (engine tests), kernel time inside the train step
mkdir -p gpurun_out; P=gpurun_out/c42
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"adam_multi" -c 6 python bench.py --steps 1 --warmup 3 --no-legs > ${P}_adam_ncu.txt 2>&1; echo "ncu rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json; grep -o '"loss": [0-9.]*' ${P}_bench.json; grep -E "duration" ${P}_adam_ncu.txt | head -6
