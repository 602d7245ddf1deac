#!/bin/bash
# round 2, call 43 (1 GPU): the whole GPU suite + smoke() on the final tree
mkdir -p gpurun_out; P=gpurun_out/c43
timeout 170 python -m pytest tests -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.txt 2>&1; echo "smoke rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; tail -1 ${P}_smoke.txt
