#!/bin/bash
# round 2, call 5 (1 GPU): full GPU suite, full bench with legs, ncu launch list of one train step, ncu --set full of the WKV pair
mkdir -p gpurun_out; P=gpurun_out/c5
timeout 1500 python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
RWKVTTS_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${P}_launches.csv python bench.py --steps 1 --warmup 3 --no-legs --e2e-steps 2 > ${P}_under_ncu.log 2>&1; echo "ncu launches rc=$?" >> ${P}_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wkv7_tc -s 3 -c 3 -o ${P}_tc_full -f python scripts/run_pair.py 2 > ${P}_full.log 2>&1; echo "ncu full rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -12 ${P}_pytest.log | cut -c1-200; tail -3 ${P}_bench.err
