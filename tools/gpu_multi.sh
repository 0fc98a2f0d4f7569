# multi-GPU check + bench (development aid); usage: bash tools/gpu_multi.sh N
N=${1:-2}
QUICK=${2:-}
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/multigpu_check_$N.log
[ -z "$QUICK" ] && PFT_COMM=nccl timeout 300 $RUN --master-port 29513 tools/multigpu_check.py --workloads 100-GradP > gpurun_out/multigpu_check_nccl_$N.log 2>&1; echo "rc=$?" >> gpurun_out/multigpu_check_nccl_$N.log
timeout 300 $RUN --master-port 29512 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_$N.log 2> gpurun_out/bench_$N.err; echo "rc=$?" >> gpurun_out/bench_$N.err
[ -z "$QUICK" ] && PFT_COMM=nccl timeout 300 $RUN --master-port 29514 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_nccl_$N.log 2> gpurun_out/bench_nccl_$N.err; echo "rc=$?" >> gpurun_out/bench_nccl_$N.err
timeout 300 $RUN --master-port 29515 bench.py --gpus $N --steps 200 --warmup 20 --workload cube-400 --no-cpu-baseline > gpurun_out/bench_cube400_$N.log 2> gpurun_out/bench_cube400_$N.err
grep -h "PASS\|FAIL\|rc=" gpurun_out/multigpu_check_$N.log
