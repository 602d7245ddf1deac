#!/bin/bash
# first GPU contact of round 2: hazard reproduction, stress, the driver's exact torchrun command, GPU suite
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/c1_gpu.txt
for v in oldbar delay; do
  timeout 150 python scripts/stress_wkv7.py --variant $v --shape c2 --pairs 300 --check-every 100 > gpurun_out/c1_stress_$v.out 2> gpurun_out/c1_stress_$v.err
  echo "variant $v rc=$?" >> gpurun_out/c1_summary.txt
done
for sh in c2 c5 small; do
  timeout 300 python scripts/stress_wkv7.py --shape $sh --pairs 5000 > gpurun_out/c1_stress_main_$sh.out 2> gpurun_out/c1_stress_main_$sh.err
  echo "main $sh rc=$?" >> gpurun_out/c1_summary.txt
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c1_bench_torchrun.json 2> gpurun_out/c1_bench_torchrun.err
echo "torchrun bench rc=$?" >> gpurun_out/c1_summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_summary.txt
cat gpurun_out/c1_summary.txt; tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_stress_oldbar.out
