#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "data_flow or adaptive_sequence or fused" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
rm -f gpurun_out/tune2_flow.txt
for w in 400-SigmaP1-P-smallsigma slab-200x50 2025-200-GradP 100-GradP; do
for f in 0 1 0 1; do
PFT_DATAFLOW=$f TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py $w 2>&1 | grep -v generic | grep -v runahead | sed "s/^/$w dataflow=$f /" | tee -a gpurun_out/tune2_flow.txt
done
done
