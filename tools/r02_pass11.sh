#!/bin/bash
O=gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_r02k.log 2>&1; echo "full suite rc=$?"; tail -3 $O/pytest_gpu_r02k.log
KINDS=lorenz3,lorenz3_pair,lorenz4_pair,hr_sync,pmsm_sync,pmsm_classic,pmsm_single,memristive4_pair,pmsm_free
$T 400 python tools/sweep.py --sizes 1048576 --kinds $KINDS > $O/sweep_r02k_auto.jsonl 2> $O/sweep_r02k.err
CHAOS_B200_BLOCK=64 $T 400 python tools/sweep.py --sizes 1048576 --kinds $KINDS > $O/sweep_r02k_b64.jsonl 2>> $O/sweep_r02k.err
CHAOS_B200_BLOCK=128 $T 400 python tools/sweep.py --sizes 1048576 --kinds $KINDS > $O/sweep_r02k_b128.jsonl 2>> $O/sweep_r02k.err
python - <<'PY'
import json
for tag in ("auto", "b64", "b128"):
    print(tag)
    for ln in open(f"gpurun_out/sweep_r02k_{tag}.jsonl"):
        d = json.loads(ln)
        if "kind" in d and d["mode"] == "step":
            print(f"  {d['kind']:18s} {d['ms_per_launch']*1e3:8.2f} us hbm {d['hbm_frac']:.3f} block {d['block']}")
PY
$T 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['us_per_control_interval'])"
