# multi-GPU iteration: bit-identity check, thin-slab bench, device timeline (development aid); usage: bash tools/gpu_mg_iter.sh N
N=${1:-2}
set -x
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NZ=$((50*N))
timeout 300 $RUN --master-port 29511 tools/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/multigpu_check_$N.log
timeout 300 $RUN --master-port 29521 bench.py --gpus $N --steps 200 --warmup 20 --workload slab-200x$NZ --no-cpu-baseline > gpurun_out/diag_slab_$N.log 2> gpurun_out/diag_slab_$N.err
timeout 300 $RUN --master-port 29531 tools/trace_timeline.py --workload slab-200x$NZ --attempts 8 > gpurun_out/trace_slab_$N.log 2>&1
timeout 300 $RUN --master-port 29512 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_$N.log 2> gpurun_out/bench_$N.err; echo "rc=$?" >> gpurun_out/bench_$N.err
grep -h "PASS\|FAIL\|rc=" gpurun_out/multigpu_check_$N.log
