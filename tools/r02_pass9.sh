#!/bin/bash
O=gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests -q -m gpu --durations=6 > $O/pytest_gpu_r02i.log 2>&1; echo "full suite rc=$?"; tail -12 $O/pytest_gpu_r02i.log
for k in lorenz_rk4 hr_sync pmsm_sync; do $T 200 python tools/e2e_modes.py $k 4096,16384,65536 dma:1,zerocopy:1,streamed:8,streamed:32 >> $O/r02i_e2e_host_modes.jsonl 2>> $O/r02i_e2e.err; done
cat $O/r02i_e2e_host_modes.jsonl
$T 600 python bench.py --steps 200 --warmup 5 > $O/bench_r02i.json 2> $O/bench_r02i.err || tail -5 $O/bench_r02i.err
python -c "
import json
d=json.load(open('$O/bench_r02i.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['e2e']['us_per_control_interval'], d['e2e']['us_per_control_interval_pinned_inputs_rank0'], d['cpu_baseline']['value'])
"
# smoke under ncu (what the driver does at round end): the streamed host mode must fall back, not fail
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/smoke_launches_r02i.csv python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_ncu_r02i.log 2>&1; echo "smoke under ncu rc=$?"; tail -2 $O/smoke_ncu_r02i.log
grep -c k_rollout_sm $O/smoke_launches_r02i.csv
$T 600 ncu --set full --clock-control none -k regex:'k_gae|k_moments|k_normalize|k_frame_stack|k_eval|k_rms' -f -o $O/prof_rlops_r02i python tools/profile_hbm_kernels.py rl_ops > $O/ncu_rlops_r02i.log 2>&1
python tools/ncu_kernels_summary.py $O/prof_rlops_r02i.ncu-rep > $O/r02i_rl_ops_ncu_metrics.txt 2>&1; rm -f $O/prof_rlops_r02i.ncu-rep
grep -E "^==|duration" $O/r02i_rl_ops_ncu_metrics.txt
