#!/bin/bash
# speculative stage 1 on N GPUs of one box: multi-GPU tests (N=2 only), then interleaved A/B of the torchrun bench
N=${1:-2}
set -x
mkdir -p gpurun_out
if [ $N -le 2 ]; then
timeout 1500 python -m pytest tests -m gpu -q -k "multi_gpu or ngpus or group or dropin" > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_multi.log
fi
for f in 0 1; do
PFT_SPEC=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$f bench.py --gpus $N --also none --no-cpu-baseline > gpurun_out/bench_n${N}_spec$f.json 2> gpurun_out/bench_n${N}_spec$f.err; echo "bench n=$N spec=$f rc=$?"
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_n${N}_spec$f.json") if l.startswith("{")][0]
    print("N=$N spec=$f", d["value"], d["ms_per_step"], "parity", d.get("parity",{}).get("ok"), "dropin", [(x["gpus"], round(x["value"])) for x in (d.get("e2e_dropin") or [])])
except Exception as e: print("no line", e)
PY
done
