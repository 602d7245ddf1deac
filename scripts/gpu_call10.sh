#!/bin/bash
# round 2, call 10 (1 GPU): one-kernel decode step: parity tests, timings, light ncu of one step
mkdir -p gpurun_out; P=gpurun_out/c10
timeout 900 python -m pytest tests/test_decode_gpu.py -x -q -m gpu -s > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python scripts/time_decode.py 300 32 > ${P}_time.txt 2>&1; echo "time rc=$?" >> ${P}_summary.txt
timeout 300 python scripts/time_decode.py 300 8 --no-graph >> ${P}_time.txt 2>&1
timeout 300 python scripts/time_decode.py 300 1 --no-graph >> ${P}_time.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed_pipe_tensor.sum -k regex:decode_step --launch-skip 20 -c 2 --clock-control none --csv --log-file ${P}_ncu.csv python scripts/time_decode.py 30 32 --no-graph > ${P}_ncu.log 2>&1; echo "ncu rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; cat ${P}_time.txt; tail -30 ${P}_pytest.log | cut -c1-220
