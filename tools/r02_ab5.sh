#!/bin/bash
O=gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_northstar.py tests/test_gpu_vecenv.py -q -x -k "sm_local or bench_configuration or plain_rollout or host_step_modes or dynamic_rollout" > $O/pytest_gpu_r02e_sm.log 2>&1; echo "sm tests rc=$?"; tail -3 $O/pytest_gpu_r02e_sm.log
export BENCH_ARGS=""
tools/ab_variants.sh $O/r02e_ab_lorenz_f64.jsonl main main:CHAOS_B200_SM_CHUNK=4 main:CHAOS_B200_SM_CHUNK=16 main:CHAOS_B200_SM_TMAP=0 main:CHAOS_B200_SM_WORKERS=8 2>&1 | tail -10
BENCH_ARGS="--kind lorenz_rk4_f32" tools/ab_variants.sh $O/r02e_ab_lorenz_f32.jsonl main main:CHAOS_B200_SM_CHUNK=16 main:CHAOS_B200_SM=0 2>&1 | tail -6
BENCH_ARGS="--kind pmsm_rk4 --substeps 4 --param-jitter 0.1" tools/ab_variants.sh $O/r02e_ab_pmsm_jit.jsonl main main:CHAOS_B200_SM_CHUNK=16 2>&1 | tail -4
BENCH_ARGS="--param-jitter 0.1" tools/ab_variants.sh $O/r02e_ab_lorenz_jit.jsonl main 2>&1 | tail -2
for k in lorenz_rk4 hr_sync pmsm_sync; do timeout 200 python tools/e2e_modes.py $k 4096,65536 dma:1,zerocopy:1,streamed:1,streamed:4,streamed:8,streamed:16 >> $O/r02e_e2e_host_modes.jsonl 2>> $O/r02e_e2e.err; done
cat $O/r02e_e2e_host_modes.jsonl
