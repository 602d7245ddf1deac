#!/bin/bash
# round 2, call 33 (1 GPU): the whole GPU suite and the whole bench line on the current tree
mkdir -p gpurun_out; P=gpurun_out/c33
timeout 1500 python -m pytest tests -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 1500 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.txt 2>&1; echo "smoke rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; tail -2 ${P}_smoke.txt; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' ${P}_bench.json; tail -12 ${P}_bench.err | cut -c1-200
