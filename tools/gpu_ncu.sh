#!/bin/bash
# ncu evidence for the stage kernels (one GPU): the launch list of a short bench run and one --set full capture of
# one attempt's five stage kernels.  Numbers printed by the runs under ncu are never bench values.
set -x
mkdir -p gpurun_out
ARGS="--steps 6 --warmup 3 --repeats 1 --also none --no-cpu-baseline --no-parity --no-dropin --profile-steps 2"
timeout 300 python bench.py $ARGS > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench.py $ARGS > gpurun_out/plain2.log 2>&1 && \
PFT_RUNAHEAD=1 ncu --set full --clock-control none --import-source on -k regex:merson_stage -s 25 -c 5 -o gpurun_out/r02_prof python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
