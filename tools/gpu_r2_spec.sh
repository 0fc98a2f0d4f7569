#!/bin/bash
# speculative stage 1 (PFT_SPEC) on one GPU: parity tests, then interleaved A/B on four grids
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/tune2_spec.txt
for w in 400-SigmaP1-P-smallsigma slab-200x50 2025-200-GradP 100-GradP; do
reps="0 1 0 1"; case $w in 2025*|100*) reps="0 1";; esac
for f in $reps; do
PFT_SPEC=$f TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py $w 2>&1 | grep -v generic | grep -v runahead | sed "s/^/$w spec=$f /" | tee -a gpurun_out/tune2_spec.txt
done
done
for f in 0 1; do
PFT_SPEC=$f timeout 600 python bench.py --fixed-step --workload cube-800w --steps 20 --warmup 5 --repeats 3 --also none --no-cpu-baseline --no-parity --no-dropin 2>gpurun_out/fixed_$f.err | tee gpurun_out/fixed_$f.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fixed-step 800x800x100 spec=$f', d['value'], d['ms_per_step'], d['gpu_launches'])"
done
