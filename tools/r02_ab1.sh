#!/bin/bash
# Round-2 first GPU pass: parity tests, then A/B of the rollout-kernel variants on the bench workload.
O=gpurun_out
python -m pytest tests -q -m gpu -x > $O/pytest_gpu_r02a.log 2>&1; tail -3 $O/pytest_gpu_r02a.log
export BENCH_ARGS=""
tools/ab_variants.sh $O/r02a_ab_lorenz_f64.jsonl \
   main main:CHAOS_B200_SM=0 main:CHAOS_B200_DYN=0 \
   main:CHAOS_B200_SM_WORKERS=16 main:CHAOS_B200_SM_WORKERS=16,CHAOS_B200_SM_CHUNK=1 main:CHAOS_B200_SM_WORKERS=16,CHAOS_B200_SM_CHUNK=2 \
   main:CHAOS_B200_SM_CHUNK=2 main:CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM_CHUNK=1 main:CHAOS_B200_SM_WORKERS=8 \
   noreuse 2>&1 | tail -30
BENCH_ARGS="--kind lorenz_rk4_f32" tools/ab_variants.sh $O/r02a_ab_lorenz_f32.jsonl main main:CHAOS_B200_SM=0 main:CHAOS_B200_SM_WORKERS=16 main:CHAOS_B200_SM_CHUNK=2 2>&1 | tail -10
BENCH_ARGS="--kind pmsm_rk4 --substeps 4 --param-jitter 0.1" tools/ab_variants.sh $O/r02a_ab_pmsm_jit.jsonl main main:CHAOS_B200_SM=0 main:CHAOS_B200_SM_WORKERS=16 2>&1 | tail -8
BENCH_ARGS="--param-jitter 0.1" tools/ab_variants.sh $O/r02a_ab_lorenz_jit.jsonl main main:CHAOS_B200_SM=0 2>&1 | tail -6
BENCH_ARGS="--envs-per-gpu 1048576 --chunk 16" tools/ab_variants.sh $O/r02a_ab_1Mi.jsonl main 2>&1 | tail -4
