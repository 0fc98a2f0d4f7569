set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 900 python tools/tune.py > gpurun_out/tune.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-steps 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-steps 3 > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-steps 3 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:merson_stage -s 25 -c 6 -o gpurun_out/prof python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-steps 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
