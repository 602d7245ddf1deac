#!/bin/bash
# Control-plane check of bench.py at world size 2 on ONE GPU: both ranks use cuda:0 (LOCAL_RANK=0), rendezvous over
# 127.0.0.1.  Exercises the mixed gloo/NCCL process-group registration and the gloo barriers / max-over-ranks on a GPU
# box without paying for a second GPU; the numbers it prints are not bench values (two processes share the device).
export WORLD_SIZE=2 MASTER_ADDR=127.0.0.1 MASTER_PORT=${MASTER_PORT:-29511} LOCAL_RANK=0
mkdir -p gpurun_out
T=${1:-14}
RANK=1 timeout $T python bench.py --gpus 2 --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/n2_r1.log 2>&1 &
RANK=0 timeout $T python bench.py --gpus 2 --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/n2_r0.log 2>&1
rc=$?
wait
echo "rank0 rc=$rc"; tail -c 600 gpurun_out/n2_r0.log; tail -c 300 gpurun_out/n2_r1.log
