#!/bin/bash
# round 2, call 22 (1 GPU): decode kernel with four tiles per GEMM pass; L2 prefetch on / off
mkdir -p gpurun_out; P=gpurun_out/c22
timeout 600 python -m pytest tests/test_decode_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python scripts/time_decode.py 300 32 > ${P}_time.txt 2>&1; echo "time rc=$?" >> ${P}_summary.txt
echo "== no L2 prefetch" >> ${P}_time.txt
RWKVTTS_DECODE_SKIP=4 timeout 300 python scripts/time_decode.py 300 32 --no-graph >> ${P}_time.txt 2>&1
cat ${P}_summary.txt; grep -v deprecated ${P}_time.txt | cut -c1-260; tail -3 ${P}_pytest.log | cut -c1-220
