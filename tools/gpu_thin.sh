# thin-slab tuning on one GPU (what each rank of an 8-GPU run owns); development aid
set -x
mkdir -p gpurun_out
for w in slab-200x50 slab-200x100 slab-400x50; do
  timeout 500 python tools/tune.py --workload $w --zsweep > gpurun_out/thin_$w.log 2>&1
done
