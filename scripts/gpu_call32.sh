#!/bin/bash
# round 2, call 32 (1 GPU): per-role cycle counters of the current chunked kernels (-DRWKVTTS_PROFILE builds of the same
# sources, tests/csrc/prof_tc_*.cu): who waits for whom, per 16-token chunk
mkdir -p gpurun_out; P=gpurun_out/c32
timeout 120 tests/csrc/_bin/prof_tc_fwd > ${P}_roles_fwd_infer.txt 2>&1; echo "fwd rc=$?" >> ${P}_summary.txt
timeout 120 tests/csrc/_bin/prof_tc_bwd > ${P}_roles_pair.txt 2>&1; echo "pair rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt ${P}_roles_fwd_infer.txt ${P}_roles_pair.txt
