#!/bin/bash
# round 2, call 23 (8 GPUs): the scaling bench line at N = 8 on the final engine (gradients per bucket in one launch)
mkdir -p gpurun_out; P=gpurun_out/c23
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --no-legs > ${P}_bench_n8.json 2> ${P}_bench_n8.err; echo "bench n8 rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; grep -ho '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 8, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench_n8.json; tail -3 ${P}_bench_n8.err | cut -c1-300
