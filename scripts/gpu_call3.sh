#!/bin/bash
# round 2, call 3 (2 GPUs): NCCL engine test, N=2 bench through torchrun, new 1-GPU tests
mkdir -p gpurun_out; P=gpurun_out/c3
nvidia-smi -L > ${P}_gpus.txt
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_head_gpu.py tests/test_stress_gpu.py -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > ${P}_bench_n2.json 2> ${P}_bench_n2.err; echo "bench n2 rc=$?" >> ${P}_summary.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-legs > ${P}_bench_n1.json 2> ${P}_bench_n1.err; echo "bench n1 rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -15 ${P}_pytest.log; grep -v "NCCL INFO" ${P}_bench_n2.err | tail -20
