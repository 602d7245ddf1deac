#!/bin/bash
# round 2, call 11 (1 GPU): one-kernel decode step with descriptors in shared memory + L2 prefetch: tests, phase time line
mkdir -p gpurun_out; P=gpurun_out/c11
timeout 600 python -m pytest tests/test_decode_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python scripts/time_decode.py 300 32 > ${P}_time.txt 2>&1; echo "time rc=$?" >> ${P}_summary.txt
timeout 300 python scripts/time_decode.py 300 1 --no-graph >> ${P}_time.txt 2>&1
cat ${P}_summary.txt; grep -v deprecated ${P}_time.txt; tail -5 ${P}_pytest.log | cut -c1-220
