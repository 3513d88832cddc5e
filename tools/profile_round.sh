#!/bin/bash
# One pass of the measurement protocol on the GPU box (run through gpurun): tests, bench (both arms),
# ncu launch list of the same command, one `ncu --set full` capture of the dominant kernel.
# Everything lands in gpurun_out/; tools/ncu_summarize.py turns it into profiles/ files afterwards.
TAG=${1:-r01b}
O=gpurun_out
python -m pytest tests -q -m gpu > $O/pytest_gpu_$TAG.log 2>&1; tail -2 $O/pytest_gpu_$TAG.log
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err || tail -5 $O/bench_$TAG.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_$TAG.json 2>> $O/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 20 --warmup 3 --e2e-chunks 1 --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_dyn -s 4 -c 2 -f -o $O/prof_dyn_$TAG \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1
ls -la $O/prof_dyn_$TAG.ncu-rep
python -c "
import json
d=json.load(open('$O/bench_$TAG.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])
"
