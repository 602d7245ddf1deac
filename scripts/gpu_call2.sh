#!/bin/bash
# round 2, call 2 (1 GPU): hazard variants, GPU suite, the new whole-model bench (plain and under torchrun), decode diagnosis
mkdir -p gpurun_out; P=gpurun_out/c2
for v in oldbar delay; do
  timeout 150 python scripts/stress_wkv7.py --variant $v --shape c2 --pairs 300 --check-every 100 > ${P}_stress_$v.out 2> ${P}_stress_$v.err
  echo "variant $v rc=$?" >> ${P}_summary.txt
done
timeout 300 python scripts/stress_wkv7.py --shape c2 --pairs 5000 > ${P}_stress_main.out 2> ${P}_stress_main.err; echo "main c2 rc=$?" >> ${P}_summary.txt
timeout 1200 python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?" >> ${P}_summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 20 --warmup 5 --no-legs > ${P}_bench_torchrun.json 2> ${P}_bench_torchrun.err; echo "torchrun bench rc=$?" >> ${P}_summary.txt
OMP_NUM_THREADS=1 timeout 400 python bench.py --leg decode > ${P}_decode_omp1.json 2> ${P}_decode_omp1.err; echo "decode omp1 rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -5 ${P}_pytest.log; cat ${P}_stress_oldbar.out; tail -3 ${P}_bench.err
