#!/bin/bash
# round 2, call 40 (8 GPUs): N = 8 bench through torchrun on the current tree (NCCL exchange)
mkdir -p gpurun_out; P=gpurun_out/c40
nvidia-smi -L > ${P}_gpus.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > ${P}_bench_n8.json 2> ${P}_bench_n8.err; echo "bench n8 rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 8, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench_n8.json; grep -v "NCCL INFO" ${P}_bench_n8.err | tail -4 | cut -c1-200
