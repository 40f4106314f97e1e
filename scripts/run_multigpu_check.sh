#!/bin/bash
# N=2 sanity of the data-parallel path: torchrun bench with 2 ranks over NCCL
set -e
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1
