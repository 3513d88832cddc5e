#!/bin/bash
# Verification pass of the last session of round 2: full parity suite, smoke, bench (both arms), key-stream rate.
O=gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_r02_final2.log 2>&1; echo "full suite rc=$?"; tail -3 $O/pytest_gpu_r02_final2.log
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02_final2.log 2>&1; echo "smoke rc=$?"
$T 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_r02_final2.json 2> $O/bench_r02_final2.err
$T 600 python bench.py > $O/bench_r02_final2.json 2>> $O/bench_r02_final2.err || tail -5 $O/bench_r02_final2.err
$T 300 python tools/keystream_rate.py > $O/r02_keystream_rate.jsonl 2> $O/r02_keystream_rate.err; cat $O/r02_keystream_rate.jsonl; tail -2 $O/r02_keystream_rate.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_final2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["us_per_control_interval"], d["cpu_baseline"]["value"])
r = json.load(open("gpurun_out/bench_reference_r02_final2.json")); print("reference", r["value"], r["config"] == d["config"])
PY
