# last check of the round after the z-chunk heuristic change (development aid)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/last_400.log 2> gpurun_out/last.err
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --workload 2025-200-GradP > gpurun_out/last_200.log 2>> gpurun_out/last.err
tail -n 2 gpurun_out/pytest_gpu.log
