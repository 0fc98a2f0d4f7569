#!/bin/bash
# split (arrive / wait) barrier of stages 2, 4, 5 against the per-plane CTA barrier, 400 case, one box
mkdir -p gpurun_out; rm -f gpurun_out/tune2_splitbar.txt
for f in 0 1; do
PFT_SPLITBAR=$f TUNE_ONLY="v2 R1 nobar" TUNE_WARMUP=5 timeout 40 python tools/tune2.py 400-SigmaP1-P-smallsigma 2>&1 | grep -v generic | grep -v runahead | sed "s/^/400 splitbar=$f /" | tee -a gpurun_out/tune2_splitbar.txt
done
PFT_SPLITBAR=1 timeout 40 python -m pytest tests -m gpu -x -q -k "golden or adaptive_sequence or fused" 2>&1 | tail -2 | tee -a gpurun_out/tune2_splitbar.txt
