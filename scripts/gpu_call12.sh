#!/bin/bash
# round 2, call 12 (1 GPU): barrier anatomy of the one-kernel decode step, three barrier forms
mkdir -p gpurun_out; P=gpurun_out/c12
for mode in 0 1 2; do
  echo "== barrier mode $mode" >> ${P}_time.txt
  RWKVTTS_DECODE_BARRIER=$mode timeout 300 python scripts/time_decode.py 300 32 --no-graph >> ${P}_time.txt 2>&1; echo "mode $mode rc=$?" >> ${P}_summary.txt
done
RWKVTTS_DECODE_BARRIER=1 timeout 600 python -m pytest tests/test_decode_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest(mode 1) rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; grep -v deprecated ${P}_time.txt; tail -3 ${P}_pytest.log | cut -c1-220
