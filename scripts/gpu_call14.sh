#!/bin/bash
# round 2, call 14 (1 GPU): instruction-cache experiment: time line with phases skipped (results meaningless, timing only)
mkdir -p gpurun_out; P=gpurun_out/c14
for skip in 3 2 1; do
  echo "== skip mask $skip" >> ${P}_time.txt
  RWKVTTS_DECODE_SKIP=$skip timeout 300 python scripts/time_decode.py 200 32 --no-graph >> ${P}_time.txt 2>&1; echo "skip $skip rc=$?" >> ${P}_summary.txt
done
cat ${P}_summary.txt; grep -v deprecated ${P}_time.txt
