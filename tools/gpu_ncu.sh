# ncu capture of the stage kernels (development aid): launch list + full sections of one attempt
set -x
mkdir -p gpurun_out
ARGS="--steps 10 --warmup 3 --no-cpu-baseline --profile-steps 3"
timeout 300 python bench.py $ARGS > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench.py $ARGS > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:merson_stage -s 25 -c 6 -o gpurun_out/prof python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
