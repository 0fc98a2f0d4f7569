# quick GPU iteration: parity tests + bench + tile sweep (development aid)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python tools/tune.py --quick > gpurun_out/tune.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
