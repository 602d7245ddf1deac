#!/bin/bash
# round 2, call 8 (1 GPU): ncu --set full of the fused elementwise adjoints; full GPU suite on the current tree
mkdir -p gpurun_out; P=gpurun_out/c8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"shift_mix_bwd|prep_bwd|add_ln_bwd|out_bwd" -c 6 -o ${P}_fused_full -f python scripts/run_fused.py 1 > ${P}_full.log 2>&1; echo "ncu rc=$?" >> ${P}_summary.txt
timeout 200 python scripts/run_fused.py 3 > ${P}_fused_times.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; cat ${P}_fused_times.txt; tail -5 ${P}_pytest.log | cut -c1-200
