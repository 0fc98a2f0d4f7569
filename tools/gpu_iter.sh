# one iteration on 2 GPUs: GPU tests, 1-GPU bench, 2-GPU check/bench/trace (development aid)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python tools/data_dep.py > gpurun_out/data_dep.log 2>&1
bash tools/gpu_mg_iter.sh 2
tail -3 gpurun_out/pytest_gpu.log
