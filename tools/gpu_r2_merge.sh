#!/bin/bash
set -x
mkdir -p gpurun_out; rm -f gpurun_out/tune2_merge.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "data_flow or adaptive or golden or interval" > gpurun_out/pytest_merge.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_merge.log
for w in 100-GradP 2025-200-GradP slab-200x50 400-SigmaP1-P-smallsigma 100-GradP 2025-200-GradP; do
TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py $w 2>&1 | grep -v generic | grep -v runahead | sed "s/^/$w merged-wait /" | tee -a gpurun_out/tune2_merge.txt
done
