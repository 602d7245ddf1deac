#!/bin/bash
# round 2, call 21 (1 GPU): ncu launch list of one train step and of the decode kernel (final tree), full bench line
mkdir -p gpurun_out; P=gpurun_out/c21
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file ${P}_launches.csv python bench.py --steps 1 --warmup 1 --no-legs > ${P}_ncu_bench.log 2>&1; echo "ncu train rc=$?" >> ${P}_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed -k regex:decode_step --launch-skip 20 -c 2 --clock-control none --csv --log-file ${P}_decode_ncu.csv python scripts/time_decode.py 30 32 --no-graph > ${P}_ncu_decode.log 2>&1; echo "ncu decode rc=$?" >> ${P}_summary.txt
timeout 1500 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 1, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' ${P}_bench.json
