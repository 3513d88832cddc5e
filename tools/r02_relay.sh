#!/bin/bash
# Relay as block 0 of the step kernel vs as its own kernel: parity, A/B, and the device-path step kernels unchanged
O=gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_gpu_vecenv.py tests/test_gpu_parity.py -q -m gpu -x > $O/pytest_relay.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_relay.log
$T 200 python tools/e2e_threads_ab.py lorenz_rk4 -3:32,3:32,-2:32,2:32 > $O/r02o_e2e_relay_lorenz.jsonl 2> $O/r02o.err; cat $O/r02o_e2e_relay_lorenz.jsonl
$T 200 python tools/e2e_threads_ab.py hr_sync -3:32,3:32 > $O/r02o_e2e_relay_hr.jsonl 2>> $O/r02o.err; cat $O/r02o_e2e_relay_hr.jsonl
$T 200 python tools/e2e_threads_ab.py pmsm_sync -3:32,3:32 > $O/r02o_e2e_relay_pmsm.jsonl 2>> $O/r02o.err; cat $O/r02o_e2e_relay_pmsm.jsonl
$T 300 python tools/sweep.py --sizes 1048576 --kinds lorenz3,hr_sync,pmsm_sync,lorenz4_pair 2>> $O/r02o.err | grep '"step"' | python -c "
import sys, json
for ln in sys.stdin:
    x = json.loads(ln); print(x['kind'], round(x['ms_per_launch']*1e3, 2), 'us', round(x['hbm_frac'], 3))"
tail -3 $O/r02o.err
