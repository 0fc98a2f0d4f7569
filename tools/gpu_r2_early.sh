#!/bin/bash
# S1' CTAs returning at once when the decision has already rejected the candidate (default) against always evaluating
set -x
mkdir -p gpurun_out; rm -f gpurun_out/tune2_early.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "data_flow or adaptive or golden or interval or nan" > gpurun_out/pytest_early.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_early.log
for w in 400-SigmaP1-P-smallsigma 2025-200-GradP; do
for f in 0 1 0 1; do
PFT_SPEC_EARLYOUT=$f TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py $w 2>&1 | grep -v generic | grep -v runahead | sed "s/^/$w earlyout=$f /" | tee -a gpurun_out/tune2_early.txt
done
done
