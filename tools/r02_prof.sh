#!/bin/bash
# Round-2 measurement pass (run through gpurun): full parity suite, bench (both arms), ncu launch list of the
# same command, `ncu --set full` captures of the benchmarked kernel (f64 / f32 / PMSM), of the HBM-bound
# single-step kernels and of the rl_ops kernels, the all-kind sweep, the host-mode table and the tensor-path
# trace.  Everything lands in gpurun_out/; tools/ncu_summarize.py / ncu_kernels_summary.py turn the reports
# into the tracked files under profiles/.
TAG=${1:-r02}
O=gpurun_out
T="timeout -k 5"
if [ "$SKIP_TESTS" != 1 ]; then $T 900 python -m pytest tests -q -m gpu --durations=8 > $O/pytest_gpu_$TAG.log 2>&1; echo "full suite rc=$?"; tail -4 $O/pytest_gpu_$TAG.log; fi
$T 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err || tail -5 $O/bench_$TAG.err
$T 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_$TAG.json 2>> $O/bench_$TAG.err
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 20 --warmup 3 --e2e-chunks 1 --no-cpu-baseline > $O/ncu_launch_$TAG.log 2>&1
$T 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_sm -s 4 -c 1 -f -o $O/prof_dyn_$TAG \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1
$T 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_sm -s 4 -c 1 -f -o $O/prof_dyn_f32_$TAG \
    python bench.py --kind lorenz_rk4_f32 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_f32_$TAG.log 2>&1
$T 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_sm -s 4 -c 1 -f -o $O/prof_dyn_pmsm_$TAG \
    python bench.py --kind pmsm_rk4 --substeps 4 --param-jitter 0.1 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_pmsm_$TAG.log 2>&1
# the gpurun_out merge is capped at 64 MiB per call: no source import for the many small kernels
$T 900 ncu --set full --clock-control none -k regex:'k_step|k_gae|k_moments|k_normalize|k_frame_stack|k_eval|k_rms' -f -o $O/prof_hbm_$TAG \
    python tools/profile_hbm_kernels.py > $O/ncu_hbm_$TAG.log 2>&1
# the gpurun_out merge is capped at 64 MiB per call: summarise the reports HERE (ncu -i needs no GPU) and
# keep only the report of the benchmarked kernel
export NCU_SUMMARY_OUT=$O/profiles_$TAG
mkdir -p $NCU_SUMMARY_OUT
python tools/ncu_summarize.py $TAG > $O/ncu_summarize_$TAG.log 2>&1
python tools/ncu_kernels_summary.py $O/prof_dyn_f32_$TAG.ncu-rep > $NCU_SUMMARY_OUT/${TAG}_dyn_f32_ncu_metrics.txt 2>&1
python tools/ncu_kernels_summary.py $O/prof_dyn_pmsm_$TAG.ncu-rep > $NCU_SUMMARY_OUT/${TAG}_dyn_pmsm_jitter_ncu_metrics.txt 2>&1
python tools/ncu_kernels_summary.py $O/prof_hbm_$TAG.ncu-rep > $NCU_SUMMARY_OUT/${TAG}_hbm_kernels_ncu_metrics.txt 2>&1
ls -la $O/*.ncu-rep
rm -f $O/prof_dyn_f32_$TAG.ncu-rep $O/prof_dyn_pmsm_$TAG.ncu-rep $O/prof_hbm_$TAG.ncu-rep
$T 600 python tools/sweep.py --sizes 1048576 > $O/sweep_$TAG.jsonl 2> $O/sweep_$TAG.err
$T 120 python tools/trace_tensor_path.py > $O/trace_tensor_path_$TAG.json 2> $O/trace_$TAG.err
$T 200 python tools/e2e_modes.py lorenz_rk4 65536 zerocopy:1,streamed:16,streamed:32 > $O/e2e_slices_$TAG.jsonl 2>> $O/trace_$TAG.err
CHAOS_B200_COPY_THREADS=1 $T 200 python tools/e2e_modes.py lorenz_rk4 65536 zerocopy:1,streamed:16,streamed:32 > $O/e2e_slices_1thread_$TAG.jsonl 2>> $O/trace_$TAG.err
cat $O/e2e_slices_1thread_$TAG.jsonl
for cfg in "--kind pmsm_rk4 --substeps 4 --param-jitter 0.1" "--param-jitter 0.1" "--kind lorenz_rk4_f32" "--envs-per-gpu 1048576 --chunk 16" "--envs-per-gpu 1048576 --chunk 16 --kind lorenz_rk4_f32" "--envs-per-gpu 1048576 --chunk 16 --kind pmsm_rk4 --substeps 4 --param-jitter 0.1"; do
  $T 300 python bench.py --steps 300 --warmup 5 --no-e2e --no-cpu-baseline $cfg 2>/dev/null | tail -1 >> $O/cfg_1gpu_$TAG.jsonl
done
cat $O/e2e_slices_$TAG.jsonl
python -c "
import json
d=json.load(open('$O/bench_$TAG.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['roofline']['frac'], d['e2e'], d['cpu_baseline']['value'])
for ln in open('$O/cfg_1gpu_$TAG.jsonl'):
    l=json.loads(ln); print(l['config']['kind'], l['config']['envs_per_gpu'], l['config']['param_jitter'], l['value'], round(l['roofline']['frac'],4))
"
