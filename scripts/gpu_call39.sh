#!/bin/bash
# round 2, call 39 (2 GPUs): N = 2 bench through torchrun on the current tree (NCCL exchange) + the 2-rank engine tests
mkdir -p gpurun_out; P=gpurun_out/c39
nvidia-smi -L > ${P}_gpus.txt
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > ${P}_bench_n2.json 2> ${P}_bench_n2.err; echo "bench n2 rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 2, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench_n2.json; grep -v "NCCL INFO" ${P}_bench_n2.err | tail -6 | cut -c1-200
