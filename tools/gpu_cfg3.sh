# BASELINE configs[2]: PhysRevE-2025 200 cases slab-sharded over N GPUs (development aid); usage: bash tools/gpu_cfg3.sh N
N=${1:-2}
set -x
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29561 tools/multigpu_check.py --workloads "2025-200-GradP,2025-200-SigmaP1-P,2025-200-Temp-Gamma=10" > gpurun_out/cfg3_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/cfg3_check_$N.log
for w in 2025-200-GradP 2025-200-Temp-Gamma=10; do
timeout 300 $RUN --master-port 29562 bench.py --gpus $N --steps 300 --warmup 20 --workload $w --no-cpu-baseline > gpurun_out/cfg3_${w}_$N.log 2> gpurun_out/cfg3_$N.err
done
timeout 300 python bench.py --steps 300 --warmup 20 --workload 2025-200-GradP --no-cpu-baseline > gpurun_out/cfg3_2025-200-GradP_1.log 2>> gpurun_out/cfg3_$N.err
grep -h "PASS\|FAIL\|rc=" gpurun_out/cfg3_check_$N.log
