#!/bin/bash
set -x
mkdir -p gpurun_out
PFT_ROWS=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rhs or fused or adaptive_sequence or fixed_step or fullsize or ragged" > gpurun_out/pytest_rows2.log 2>&1; echo "rows2 rc=$?"; tail -3 gpurun_out/pytest_rows2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rhs or fused or adaptive_sequence or fixed_step or fullsize or ragged" > gpurun_out/pytest_v2.log 2>&1; echo "v2 rc=$?"; tail -3 gpurun_out/pytest_v2.log
timeout 1500 python tools/tune2.py $1 > gpurun_out/tune2_${1:-400}.txt 2>&1; cat gpurun_out/tune2_${1:-400}.txt
