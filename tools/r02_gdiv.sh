#!/bin/bash
O=gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_northstar.py -q -x -k "sm_local or bench_configuration" > $O/pytest_gpu_r02_gdiv.log 2>&1; echo "sm tests rc=$?"; tail -2 $O/pytest_gpu_r02_gdiv.log
export BENCH_ARGS=""
tools/ab_variants.sh $O/r02f_ab_gdiv_f64.jsonl main main:CHAOS_B200_SM_GDIV=2 main:CHAOS_B200_SM_GDIV=4 main:CHAOS_B200_SM_GDIV=8 2>&1 | tail -8
BENCH_ARGS="--kind lorenz_rk4_f32" tools/ab_variants.sh $O/r02f_ab_gdiv_f32.jsonl main main:CHAOS_B200_SM_GDIV=2 main:CHAOS_B200_SM_GDIV=4 2>&1 | tail -6
BENCH_ARGS="--kind pmsm_rk4 --substeps 4 --param-jitter 0.1" tools/ab_variants.sh $O/r02f_ab_gdiv_pmsm.jsonl main main:CHAOS_B200_SM_GDIV=2 main:CHAOS_B200_SM_GDIV=4 2>&1 | tail -6
