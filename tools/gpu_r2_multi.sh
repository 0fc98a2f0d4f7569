#!/bin/bash
# round 2: multi-GPU paths on N GPUs of one box: in-process groups (tests), the drop-in with PFT_NGPUS, torchrun bench
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q -k "multi_gpu" > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_multi_$N.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n ${BENCH_ARGS} > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"
    tail -c 400 gpurun_out/bench_n$n.err
  fi
done
