python tools/tune.py --quick 2>&1 | head -3 > gpurun_out/share_a.log
PFT_SHARE=1 python - <<'PY' > gpurun_out/share_b.log 2>&1
import os, sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import tune
from bench import make_spec
from porousfreezethaw_b200.solver import SlabSolver
spec,_,_ = make_spec("400-SigmaP1-P-smallsigma", 1)
s0 = SlabSolver(spec); u,p,gl = spec.initial_fields(); s0.upload_interior(u,p,gl); s0.rasterise(spec.spheres); s0.solve(spec.final_time, max_attempts=40)
fields = s0.download_interior(); th=(s0.sys.t, s0.sys.h); s0.close()
for env, fl in (({},0), ({},16), ({"PFT_PROFILE_STAGES":1},16), ({"PFT_PROFILE_STAGES":1},0)):
    print(json.dumps(tune.run(spec, fields, env, fl, 60, th)), flush=True)
PY
cat gpurun_out/share_b.log
