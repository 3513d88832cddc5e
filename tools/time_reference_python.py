#!/usr/bin/env python
"""Times the UNMODIFIED reference env classes (oracle/ref_loader.py) -- the reference's real CPU
gym path, single env, one core -- where /root/reference exists (the build container; it cannot
travel to the GPU box).  20,000 steps, best of 3.  Test/measurement infrastructure only."""
import contextlib, io, json, os, platform, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from oracle import ref_loader as R

CASES = [("dynamic.py::lorenzEnv_transient", R.lorenz3, 3, 0.05, False), ("dynamic.py nested", R.lorenz3_pair, 3, 0.05, False),
         ("lorenz_env_transient.py (4-D pair)", R.lorenz4_pair, 3, 1.0, False), ("lorenz_env_try.py::HRSyncEnv", R.hr_sync, 2, 1.0, True),
         ("lorenz_env_try_pmsm.py::PMSM_Sync_Env", R.pmsm_sync, 2, 1.0, True), ("lorenz_env_transient_pmsm.py", R.pmsm_classic, 2, 1.0, False),
         ("lorenz_env_transient2.py", R.memristive4_pair, 3, 0.01, False)]
out = {"host": platform.processor() or platform.machine(), "python": platform.python_version(), "numpy": np.__version__,
       "note": "build container, 1 core, NOT the GPU box", "steps_per_s": {}}
for name, ctor, na, amp, gymn in CASES:
    best = 0.0
    for rep in range(3):
        env = ctor(); np.random.seed(0)
        env.reset(seed=0) if gymn else env.reset()
        acts = np.random.default_rng(1).uniform(-amp, amp, (20000, na)).astype(np.float32)
        with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            for a in acts:
                env.step(a)
            el = time.perf_counter() - t0
        best = max(best, 20000 / el)
    out["steps_per_s"][name] = round(best)
print(json.dumps(out))
