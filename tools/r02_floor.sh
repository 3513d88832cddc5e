#!/bin/bash
# Key-stream GPU tests, breakdown of a host-buffer step at 65,536 / 256 envs, bench line (profiles/r02m_*)
O=gpurun_out
T="timeout -k 5"
$T 300 python -m pytest tests/test_keystream.py -q -m gpu -x > $O/pytest_keystream.log 2>&1; echo "keystream rc=$?"; tail -3 $O/pytest_keystream.log
$T 120 python tools/e2e_breakdown.py 65536 lorenz_rk4 2>&1 | tee $O/r02m_breakdown_65536.txt
$T 120 python tools/e2e_breakdown.py 256 lorenz_rk4 2>&1 | tee $O/r02m_breakdown_256.txt
$T 400 python bench.py > $O/bench_r02m.json 2> $O/bench_r02m.err; tail -2 $O/bench_r02m.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02m.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["us_per_control_interval"], d["e2e"]["us_per_control_interval_pinned_inputs_rank0"], d["cpu_baseline"]["value"])
PY
