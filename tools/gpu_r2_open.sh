#!/bin/bash
# run-ahead window opened at once against opened from one (PFT_SLOW_OPEN=1), short timed regions as the driver uses
set -x
mkdir -p gpurun_out; rm -f gpurun_out/open_ab.txt
for w in 400-SigmaP1-P-smallsigma slab-200x50; do
for f in 1 0 1 0; do
PFT_SLOW_OPEN=$f timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --also none --no-cpu-baseline --no-parity --no-dropin 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w steps=20 slow_open=$f', round(d['value']), d['ms_per_step'], [round(x,3) for x in d['device_ms_regions']])" | tee -a gpurun_out/open_ab.txt
done
done
