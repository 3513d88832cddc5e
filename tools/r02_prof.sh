#!/bin/bash
# Round-2 profiling pass (run through gpurun): ncu launch list + full capture of the benchmarked kernel,
# full captures of the HBM-bound single-step kernels and the rl_ops kernels, the all-kind sweep, the
# host-mode table and the tensor-path trace.  Everything lands in gpurun_out/.
TAG=${1:-r02}
O=gpurun_out
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err || tail -5 $O/bench_$TAG.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_$TAG.json 2>> $O/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 20 --warmup 3 --e2e-chunks 1 --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_sm -s 4 -c 2 -f -o $O/prof_dyn_$TAG \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_sm -s 4 -c 1 -f -o $O/prof_dyn_f32_$TAG \
    python bench.py --kind lorenz_rk4_f32 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_f32_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout_sm -s 4 -c 1 -f -o $O/prof_dyn_pmsm_$TAG \
    python bench.py --kind pmsm_rk4 --substeps 4 --param-jitter 0.1 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_pmsm_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_step|k_gae|k_moments|k_normalize|k_frame_stack|k_eval|k_rms' -f -o $O/prof_hbm_$TAG \
    python tools/profile_hbm_kernels.py > $O/ncu_hbm_$TAG.log 2>&1
ls -la $O/*.ncu-rep
python tools/sweep.py --sizes 1048576 > $O/sweep_$TAG.jsonl 2> $O/sweep_$TAG.err
python tools/trace_tensor_path.py > $O/trace_tensor_path_$TAG.json 2> $O/trace_$TAG.err
python -c "
import json
d=json.load(open('$O/bench_$TAG.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'], d['e2e'], d['cpu_baseline']['value'])
"
