#!/bin/bash
# round 2, call 6 (1 GPU): full GPU suite after the watchdog / varlen / exact fixes, kernel timings, short bench
mkdir -p gpurun_out; P=gpurun_out/c6
timeout 1500 python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 300 python bench.py --leg wkv_ops > ${P}_wkv_ops.json 2> ${P}_wkv_ops.err; echo "wkv_ops rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --leg ref_gpu_op > ${P}_ref_gpu.json 2> ${P}_ref_gpu.err; echo "ref_gpu rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -8 ${P}_pytest.log | cut -c1-200; grep value ${P}_bench.err
