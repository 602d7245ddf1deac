#!/bin/bash
# round 2, call 20 (2 GPUs): two-rank engine tests (NCCL overlap, fused exchange) + N=2 bench lines on the final engine
mkdir -p gpurun_out; P=gpurun_out/c20
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-legs > ${P}_bench_n2.json 2> ${P}_bench_n2.err; echo "bench n2 rc=$?" >> ${P}_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-legs --zero-p2p > ${P}_bench_n2_p2p.json 2> ${P}_bench_n2_p2p.err; echo "bench n2 p2p rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -3 ${P}_pytest.log | cut -c1-200; grep -ho '"value": [0-9.]*, "unit": "tokens/s", "n_gpus": 2, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' ${P}_bench_n2.json ${P}_bench_n2_p2p.json
