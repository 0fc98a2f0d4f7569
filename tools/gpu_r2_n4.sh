#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu" > gpurun_out/pytest_multi_4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi_4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --also none > gpurun_out/bench_n4_bal.json 2> gpurun_out/bench_n4_bal.err; echo "rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --also none --no-balance --no-parity > gpurun_out/bench_n4_nobal.json 2> gpurun_out/bench_n4_nobal.err; echo "rc=$?"
timeout 600 python bench.py --also none --no-cpu-baseline --no-parity --repeats 3 > gpurun_out/bench_n1_dropin4.json 2> gpurun_out/bench_n1_dropin4.err; echo "n1 rc=$?"
