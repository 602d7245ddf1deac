#!/bin/bash
# round 2, call 25 (1 GPU): full GPU suite on the final tree (+ row-phase stamps of the decode kernel)
mkdir -p gpurun_out; P=gpurun_out/c25
timeout 1500 python -m pytest tests -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python scripts/time_decode.py 300 32 --no-graph > ${P}_time.txt 2>&1; echo "time rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; grep -v deprecated ${P}_time.txt | cut -c1-250; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench.json
