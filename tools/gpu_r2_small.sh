#!/bin/bash
set -x
mkdir -p gpurun_out; rm -f gpurun_out/tune2_small.txt
for zc in 0 20 25 33 50; do
PFT_ZCHUNKS=$zc TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py 100-GradP 2>&1 | grep -v generic | grep -v runahead | sed "s/^/100-GradP zchunks=$zc /" | tee -a gpurun_out/tune2_small.txt
done
for zc in 0 33 40 50 66; do
PFT_ZCHUNKS=$zc TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py 2025-200-GradP 2>&1 | grep -v generic | grep -v runahead | sed "s/^/2025-200 zchunks=$zc /" | tee -a gpurun_out/tune2_small.txt
done
for tx in 26 50; do for ty in 5 3; do
PFT_TX=$tx PFT_TY=$ty PFT_ZCHUNKS=25 TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py 100-GradP 2>&1 | grep -v generic | grep -v runahead | sed "s/^/100-GradP tile=${tx}x${ty} zchunks=25 /" | tee -a gpurun_out/tune2_small.txt
done; done
