#!/bin/bash
O=gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-2} --steps 100 --warmup 3 > $O/bench_r02m_${NG:-2}gpu.json 2> $O/bench_r02m_${NG:-2}gpu.err; echo "rc=$?"; tail -3 $O/bench_r02m_${NG:-2}gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02m_%sgpu.json" % __import__("os").environ.get("NG","2") + "").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["us_per_control_interval"], d["e2e"].get("us_per_control_interval_pinned_inputs_rank0"), d.get("cfg4"))
PY
nproc
