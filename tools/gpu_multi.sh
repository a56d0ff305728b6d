#!/bin/bash
# Multi-GPU session: sharded parity + bench at N ranks.  Usage: tools/gpu_multi.sh <tag> <N>
TAG=${1:-r2m}; N=${2:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -s > $O/${TAG}_pytest_multi.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/sharded_check.py > $O/${TAG}_sharded_check.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 50 --warmup 3 > $O/${TAG}_bench_${N}gpu.log 2>&1
tail -3 $O/${TAG}_pytest_multi.log; grep -v "^W\|^\*" $O/${TAG}_sharded_check.log | tail -8; tail -2 $O/${TAG}_bench_${N}gpu.log | cut -c1-3000
