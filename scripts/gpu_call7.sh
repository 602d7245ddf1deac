#!/bin/bash
# round 2, call 7 (2 GPUs): NCCL / symmetric-memory engine tests, sequence split, N=2 bench with the fused exchange
mkdir -p gpurun_out; P=gpurun_out/c7
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_stress_gpu.py -q -m gpu > ${P}_pytest.log 2>&1; echo "pytest rc=$?" >> ${P}_summary.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --zero-p2p > ${P}_bench_n2_p2p.json 2> ${P}_bench_n2_p2p.err; echo "bench n2 p2p rc=$?" >> ${P}_summary.txt
RWKVTTS_ZERO_MULTICAST=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --zero-p2p > ${P}_bench_n2_p2p_uc.json 2> ${P}_bench_n2_p2p_uc.err; echo "bench n2 p2p unicast rc=$?" >> ${P}_summary.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > ${P}_bench_n2.json 2> ${P}_bench_n2.err; echo "bench n2 nccl rc=$?" >> ${P}_summary.txt
cat ${P}_summary.txt; tail -12 ${P}_pytest.log | cut -c1-220; for f in ${P}_bench_n2_p2p ${P}_bench_n2_p2p_uc ${P}_bench_n2; do grep -h "value" $f.err | tail -1; tail -1 $f.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d.get('collective'))" 2>/dev/null; done; grep -c "NCCL INFO" ${P}_bench_n2.err
