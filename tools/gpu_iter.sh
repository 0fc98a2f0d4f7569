# one iteration on 2 GPUs: GPU tests, 1-GPU bench A/B (env switch), 2-GPU check/bench/trace (development aid)
set -x
mkdir -p gpurun_out
SW=${1:-PFT_NO_PDL}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_on$i.log 2> gpurun_out/bench.err
env $SW=1 timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_off$i.log 2>> gpurun_out/bench.err
done
timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --workload slab-200x50 > gpurun_out/bench_thin_on.log 2>> gpurun_out/bench.err
env $SW=1 timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --workload slab-200x50 > gpurun_out/bench_thin_off.log 2>> gpurun_out/bench.err
N=2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/multigpu_check_$N.log
for v in on off; do
  E=""; [ $v = off ] && E="$SW=1"
  env $E timeout 300 $RUN --master-port 29521 bench.py --gpus $N --steps 300 --warmup 20 --workload slab-200x100 --no-cpu-baseline > gpurun_out/mg_slab_$v.log 2>> gpurun_out/bench.err
  env $E timeout 300 $RUN --master-port 29522 bench.py --gpus $N --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/mg_hr_$v.log 2>> gpurun_out/bench.err
done
grep -h "PASS\|FAIL\|rc=" gpurun_out/multigpu_check_$N.log
tail -3 gpurun_out/pytest_gpu.log
