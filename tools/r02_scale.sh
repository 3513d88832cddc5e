#!/bin/bash
# bench.py at N GPUs the way the driver launches it; prints the headline numbers
N=$1
O=gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 5 > $O/bench_r02_${N}gpu.json 2> $O/bench_r02_${N}gpu.err
tail -2 $O/bench_r02_${N}gpu.err
python - <<PY
import json
d=json.loads(open("$O/bench_r02_${N}gpu.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["us_per_control_interval"], "cores", d["e2e"].get("host_cores_of_rank0"))
print(json.dumps(d.get("cfg4")))
PY
