#!/bin/bash
# round 2, call 4 (8 GPUs): the 1/2/4/8 curve of the whole-model train step through torchrun, plus the 2-rank NCCL engine test
mkdir -p gpurun_out; P=gpurun_out/c4
nvidia-smi -L > ${P}_gpus.txt
timeout 300 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "nccl or one_gpu" > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
for n in 8 4 2 1; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-legs > ${P}_bench_n$n.json 2> ${P}_bench_n$n.err
  echo "bench n$n rc=$?" >> ${P}_summary.txt
done
cat ${P}_summary.txt; tail -5 ${P}_pytest.log; for n in 1 2 4 8; do grep -h "value" ${P}_bench_n$n.err | tail -1; done
