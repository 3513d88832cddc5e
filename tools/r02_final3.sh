#!/bin/bash
# Verification + evidence pass of the final build of round 2 (relay inside the step kernel): full parity suite, smoke,
# bench (both arms), ncu launch list of the bench command, smoke under ncu (kernel list), host-mode table.
O=gpurun_out
TAG=r02f3
T="timeout -k 5"
$T 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_$TAG.log 2>&1; echo "full suite rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
$T 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_$TAG.json 2> $O/bench_$TAG.err
$T 600 python bench.py > $O/bench_$TAG.json 2>> $O/bench_$TAG.err || tail -5 $O/bench_$TAG.err
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 20 --warmup 3 --e2e-chunks 1 --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1; echo "ncu launch list rc=$?"
$T 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/smoke_launches_$TAG.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_ncu_$TAG.log 2>&1; echo "smoke under ncu rc=$?"
$T 200 python tools/e2e_modes.py lorenz_rk4 65536 dma:1,zerocopy:1,streamed:32 > $O/${TAG}_e2e_host_modes.jsonl 2> $O/e2e_$TAG.err
$T 200 python tools/e2e_modes.py hr_sync 65536 dma:1,zerocopy:1,streamed:32 >> $O/${TAG}_e2e_host_modes.jsonl 2>> $O/e2e_$TAG.err
$T 200 python tools/e2e_modes.py pmsm_sync 65536 dma:1,zerocopy:1,streamed:32 >> $O/${TAG}_e2e_host_modes.jsonl 2>> $O/e2e_$TAG.err
$T 200 python tools/e2e_modes.py lorenz_rk4 4096,16384,262144 zerocopy:1,streamed:8,streamed:32 >> $O/${TAG}_e2e_host_modes.jsonl 2>> $O/e2e_$TAG.err
cat $O/${TAG}_e2e_host_modes.jsonl
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02f3.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["us_per_control_interval"], d["e2e"]["us_per_control_interval_pinned_inputs_rank0"], d["cpu_baseline"]["value"])
r = json.load(open("gpurun_out/bench_reference_r02f3.json")); print("reference", r["value"], r["config"] == d["config"])
PY
