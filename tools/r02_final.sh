#!/bin/bash
# Last measurement pass of round 2 on the final build: parity suite, smoke, bench (both arms), single-step sweep of
# the parity kinds, rl_ops kernel times, cfg 5.
O=gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_r02_final.log 2>&1; echo "full suite rc=$?"; tail -3 $O/pytest_gpu_r02_final.log
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02_final.log 2>&1; echo "smoke rc=$?"
$T 600 python bench.py > $O/bench_r02_final.json 2> $O/bench_r02_final.err || tail -5 $O/bench_r02_final.err
$T 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_r02_final.json 2>> $O/bench_r02_final.err
KINDS=lorenz3,lorenz3_pair,lorenz4_pair,hr_sync,pmsm_sync,pmsm_classic,pmsm_single,memristive4_pair,pmsm_free
$T 400 python tools/sweep.py --sizes 1048576 --kinds $KINDS > $O/sweep_r02_final.jsonl 2> $O/sweep_r02_final.err
$T 600 ncu --set full --clock-control none -k regex:'k_gae|k_moments|k_normalize|k_frame_stack|k_eval|k_rms' -f -o $O/prof_rlops_final python tools/profile_hbm_kernels.py rl_ops > $O/ncu_rlops_final.log 2>&1
python tools/ncu_kernels_summary.py $O/prof_rlops_final.ncu-rep > $O/r02_rl_ops_ncu_metrics_final.txt 2>&1; rm -f $O/prof_rlops_final.ncu-rep
$T 300 python tools/bench_rollout_cfg5.py > $O/r02_cfg5_rollout.jsonl 2> $O/r02_cfg5.err; cat $O/r02_cfg5_rollout.jsonl
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_final.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["us_per_control_interval"], d["cpu_baseline"]["value"])
r = json.load(open("gpurun_out/bench_reference_r02_final.json")); print("reference", r["value"], r["config"] == d["config"])
for ln in open("gpurun_out/sweep_r02_final.jsonl"):
    x = json.loads(ln)
    if "kind" in x and x["mode"] == "step":
        print(f"  {x['kind']:18s} {x['ms_per_launch']*1e3:8.2f} us hbm {x['hbm_frac']:.3f} block {x['block']}")
PY
grep -E "^==|duration" $O/r02_rl_ops_ncu_metrics_final.txt
