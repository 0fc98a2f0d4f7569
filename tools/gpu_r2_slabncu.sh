#!/bin/bash
set -x
mkdir -p gpurun_out
ARGS="--workload slab-200x50 --steps 6 --warmup 3 --repeats 1 --also none --no-cpu-baseline --no-parity --no-dropin --profile-steps 2"
timeout 300 python bench.py $ARGS > gpurun_out/plain_slab.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,sm__warps_active.avg.pct_of_peak_sustained_active,launch__waves_per_multiprocessor,smsp__inst_executed.sum --clock-control none -s 20 -c 15 --csv --log-file gpurun_out/r02_ncu_slab_launches.csv python bench.py $ARGS > gpurun_out/ncu_slab.log 2>&1
