#!/bin/bash
# last check of the round on one GPU: smoke, all GPU tests, small-grid timing, the default bench line
set -x
mkdir -p gpurun_out; rm -f gpurun_out/tune2_last.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in 100-GradP slab-200x50 100-GradP; do
TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py $w 2>&1 | grep -v generic | grep -v runahead | sed "s/^/$w desc-prefetch /" | tee -a gpurun_out/tune2_last.txt
done
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
