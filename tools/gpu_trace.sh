# device timeline at thin slabs (development aid); usage: bash tools/gpu_trace.sh N
N=${1:-2}
set -x
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NZ=$((50*N))
timeout 300 $RUN --master-port 29531 tools/trace_timeline.py --workload slab-200x$NZ --attempts 4 > gpurun_out/trace_slab_$N.log 2>&1
