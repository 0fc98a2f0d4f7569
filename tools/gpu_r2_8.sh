#!/bin/bash
# round 2, 8 GPUs of one box: multi-GPU tests (in-process groups, the drop-in with PFT_NGPUS), torchrun bench at 8 and 4
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu" > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi_8.log
for n in 8 4; do
  EXTRA=""; [ $n -ne 8 ] && EXTRA="--also none"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n $EXTRA > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"
done
