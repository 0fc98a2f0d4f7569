#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
TUNE_ONLY="v2 R1 nobar" timeout 300 python tools/tune2.py slab-200x50 2>&1 | tee gpurun_out/tune2_slab_b.txt
TUNE_ONLY="v1,v2 R1 nobar" timeout 300 python tools/tune2.py 2>&1 | tee gpurun_out/tune2_400_b.txt
