#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --also none --no-parity > gpurun_out/bench_n4b.json 2> gpurun_out/bench_n4b.err; echo "rc=$?"
