#!/bin/bash
# round 2, call 9 (1 GPU): new adjoint kernels: parity tests, timings, short bench
mkdir -p gpurun_out; P=gpurun_out/c9
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_model_gpu.py tests/test_varlen_gpu.py tests/test_layouts_gpu.py -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 200 python scripts/run_fused.py 3 > ${P}_fused_times.txt 2>&1
timeout 300 python bench.py --leg fused_tmix_kernels > ${P}_fused_leg.json 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt

cat ${P}_summary.txt; cat ${P}_fused_times.txt; tail -5 ${P}_pytest.log | cut -c1-200; grep value ${P}_bench.err
