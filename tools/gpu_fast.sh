set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k fast_variant > gpurun_out/fast_test.log 2>&1
timeout 600 python bench.py --variant fast --no-cpu-baseline > gpurun_out/bench_fast.log 2> gpurun_out/bench_fast.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_strict.log 2> gpurun_out/bench_strict.err
tail -2 gpurun_out/fast_test.log
