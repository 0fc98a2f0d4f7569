# synthetic sphere-pack weak scaling, fixed-step mode (development aid); usage: bash tools/gpu_weak.sh N [more workloads]
N=${1:-1}
set -x
mkdir -p gpurun_out
if [ "$N" = 1 ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"; fi
timeout 600 $RUN bench.py --gpus $N --steps 100 --warmup 20 --workload cube-800w --fixed-step --no-cpu-baseline --profile-steps 10 > gpurun_out/weak_$N.log 2> gpurun_out/weak_$N.err; echo "rc=$?" >> gpurun_out/weak_$N.err
for w in "$@"; do
  [ "$w" = "$N" ] && continue
  timeout 600 $RUN bench.py --gpus $N --steps 200 --warmup 20 --workload $w --no-cpu-baseline > gpurun_out/scale_${w}_$N.log 2> gpurun_out/scale_${w}_$N.err
done
