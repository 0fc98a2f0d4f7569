#!/bin/bash
# mbarrier wait with a suspend-time hint (NANOSLEEP.SYNCS) against the plain try_wait spin loop (ab_build/libpft_b200_nohint.so)
set -x
mkdir -p gpurun_out; rm -f gpurun_out/tune2_hint.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or adaptive_sequence or data_flow or golden" > gpurun_out/pytest_hint.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_hint.log
for w in 400-SigmaP1-P-smallsigma slab-200x50; do
for f in nohint hint nohint hint; do
lib=""; [ $f = nohint ] && lib=$PWD/ab_build/libpft_b200_nohint.so
PFT_BENCH_LIB=$lib TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py $w 2>&1 | grep -v generic | grep -v runahead | sed "s/^/$w $f /" | tee -a gpurun_out/tune2_hint.txt
done
done
