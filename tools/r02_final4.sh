#!/bin/bash
# Verification of the lowered streamed-mode threshold: full suite, smoke, bench, default-mode step times at small batches
O=gpurun_out
TAG=r02f4
T="timeout -k 5"
$T 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_$TAG.log 2>&1; echo "full suite rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
$T 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err || tail -5 $O/bench_$TAG.err
$T 300 python tools/bench_rollout_cfg5.py > $O/${TAG}_cfg5_rollout.jsonl 2> $O/${TAG}_cfg5.err; cat $O/${TAG}_cfg5_rollout.jsonl
$T 200 python - <<'PY' 2>&1 | tee $O/${TAG}_default_mode_steps.jsonl
import json, sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from gym_lorenz_b200.vec_env import BatchedChaosVecEnv
for kind, N in (("lorenz_rk4", 4096), ("lorenz_rk4", 16384), ("hr_sync", 4096), ("hr_sync", 16384), ("pmsm_sync", 16384), ("lorenz_rk4", 65536)):
    env = BatchedChaosVecEnv(kind, N); env.reset()
    rng = np.random.default_rng(0)
    acts = [rng.uniform(-0.3, 0.3, (N, env.batch.act_dim)).astype(np.float32) for _ in range(8)]
    for k in range(60): env.step(acts[k % 8])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(400): env.step(acts[k % 8])
    torch.cuda.synchronize()
    print(json.dumps({"kind": kind, "envs": N, "mode": "default", "us_per_step": round((time.perf_counter() - t0) / 400 * 1e6, 2),
                      "fallbacks": int(env.batch.streamed_fallbacks)}), flush=True)
    env.close()
PY
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02f4.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["us_per_control_interval"], d["cpu_baseline"]["value"])
PY
