#!/bin/bash
O=gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_northstar.py tests/test_gpu_vecenv.py -q -x -k "sm_local or bench_configuration or plain_rollout or host_step_modes or dynamic_rollout" > $O/pytest_gpu_r02d_sm.log 2>&1; echo "sm tests rc=$?"; tail -3 $O/pytest_gpu_r02d_sm.log
export BENCH_ARGS=""
tools/ab_variants.sh $O/r02d_ab_lorenz_f64.jsonl main main:CHAOS_B200_SM_CHUNK=4 main:CHAOS_B200_SM_CHUNK=16 main:CHAOS_B200_SM_WORKERS=16 2>&1 | tail -8
BENCH_ARGS="--kind lorenz_rk4_f32" tools/ab_variants.sh $O/r02d_ab_lorenz_f32.jsonl main main:CHAOS_B200_SM_CHUNK=16 main:CHAOS_B200_SM=0 2>&1 | tail -6
BENCH_ARGS="--kind pmsm_rk4 --substeps 4 --param-jitter 0.1" tools/ab_variants.sh $O/r02d_ab_pmsm_jit.jsonl main main:CHAOS_B200_SM_CHUNK=16 2>&1 | tail -4
for pm in 0 1; do CHAOS_B200_POLL=$pm timeout 200 python tools/e2e_modes.py lorenz_rk4 4096,65536 zerocopy:1,streamed:1,streamed:4,streamed:8,streamed:16 >> $O/r02d_e2e_poll$pm.jsonl 2>> $O/r02d_e2e.err; done
tail -30 $O/r02d_e2e_poll0.jsonl $O/r02d_e2e_poll1.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_sm -s 4 -c 1 -f -o $O/prof_sm_r02d \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_r02d.log 2>&1
ls -la $O/prof_sm_r02d.ncu-rep
