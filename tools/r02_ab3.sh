#!/bin/bash
# Round-2 third GPU pass: the new-kernel parity tests first (bounded), then the full suite, then A/B.
O=gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_northstar.py -q -x -k "sm_local or bench_configuration or plain_rollout" > $O/pytest_gpu_r02c_sm.log 2>&1; echo "sm tests rc=$?"; tail -3 $O/pytest_gpu_r02c_sm.log
timeout -k 5 1500 python -m pytest tests -q -m gpu --durations=12 > $O/pytest_gpu_r02c.log 2>&1; echo "full suite rc=$?"; tail -25 $O/pytest_gpu_r02c.log
export BENCH_ARGS=""
tools/ab_variants.sh $O/r02c_ab_lorenz_f64.jsonl \
   main main:CHAOS_B200_SM_CHUNK=2 main:CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM_CHUNK=16 \
   main:CHAOS_B200_SM_WORKERS=16 main:CHAOS_B200_SM_WORKERS=16,CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM=0 2>&1 | tail -14
BENCH_ARGS="--kind lorenz_rk4_f32" tools/ab_variants.sh $O/r02c_ab_lorenz_f32.jsonl main main:CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM_CHUNK=16 main:CHAOS_B200_SM_WORKERS=16,CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM=0 2>&1 | tail -10
BENCH_ARGS="--kind pmsm_rk4 --substeps 4 --param-jitter 0.1" tools/ab_variants.sh $O/r02c_ab_pmsm_jit.jsonl main main:CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM_CHUNK=16 main:CHAOS_B200_SM_WORKERS=16,CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM=0 2>&1 | tail -10
BENCH_ARGS="--param-jitter 0.1" tools/ab_variants.sh $O/r02c_ab_lorenz_jit.jsonl main main:CHAOS_B200_SM_CHUNK=8 2>&1 | tail -4
BENCH_ARGS="--envs-per-gpu 1048576 --chunk 16" tools/ab_variants.sh $O/r02c_ab_1Mi.jsonl main 2>&1 | tail -2
BENCH_ARGS="--envs-per-gpu 1048576 --chunk 16 --kind lorenz_rk4_f32" tools/ab_variants.sh $O/r02c_ab_1Mi_f32.jsonl main 2>&1 | tail -2
BENCH_ARGS="--envs-per-gpu 1048576 --chunk 16 --kind pmsm_rk4 --substeps 4 --param-jitter 0.1" tools/ab_variants.sh $O/r02c_ab_1Mi_pmsm.jsonl main 2>&1 | tail -2
for k in lorenz_rk4 hr_sync pmsm_sync; do timeout 300 python tools/e2e_modes.py $k 4096,65536 >> $O/r02c_e2e_host_modes.jsonl 2>> $O/r02c_e2e.err; done
tail -40 $O/r02c_e2e_host_modes.jsonl
