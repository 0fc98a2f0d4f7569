#!/bin/bash
# round 2, one GPU: default bench line, fast-variant line, thin-slab sweep, ncu evidence
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n1.err
timeout 300 python bench.py --variant fast --also none --no-cpu-baseline --no-parity --no-dropin --repeats 3 > gpurun_out/bench_n1_fast.json 2> gpurun_out/bench_n1_fast.err; echo "fast rc=$?"
for zc in 0 3 4 5 6 8; do
  PFT_ZCHUNKS=$zc TUNE_ONLY="v2 R1 nobar,v2 R1 bar" timeout 300 python tools/tune2.py slab-200x50 2>&1 | sed "s/^/zchunks=$zc /" >> gpurun_out/tune2_slab.txt
done
cat gpurun_out/tune2_slab.txt
bash tools/gpu_ncu.sh > gpurun_out/ncu_script.log 2>&1
