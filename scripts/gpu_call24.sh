#!/bin/bash
# round 2, call 24 (1 GPU): fresh ncu --set full of the three tcgen05 kernels and of the ring adjoints on the final tree
mkdir -p gpurun_out; P=gpurun_out/c24
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wkv7_tc -s 3 -c 3 -o ${P}_tc_full -f python scripts/run_pair.py 2 > ${P}_full.log 2>&1; echo "ncu wkv rc=$?" >> ${P}_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"shift_mix_bwd_ring|prep_bwd_ring|add_ln_bwd|out_bwd" -c 6 -o ${P}_fused_full -f python scripts/run_fused.py 1 > ${P}_fused.log 2>&1; echo "ncu fused rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; ls -la gpurun_out/c24_*
