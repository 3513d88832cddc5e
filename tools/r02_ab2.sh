#!/bin/bash
# Round-2 second GPU pass: full parity suite, A/B of task-chunk sizes after the fence removal.
O=gpurun_out
python -m pytest tests -q -m gpu > $O/pytest_gpu_r02b.log 2>&1; tail -5 $O/pytest_gpu_r02b.log
export BENCH_ARGS=""
tools/ab_variants.sh $O/r02b_ab_lorenz_f64.jsonl \
   main main:CHAOS_B200_SM_CHUNK=2 main:CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM_CHUNK=16 \
   main:CHAOS_B200_SM_WORKERS=16 main:CHAOS_B200_SM_WORKERS=16,CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM=0 2>&1 | tail -14
BENCH_ARGS="--kind lorenz_rk4_f32" tools/ab_variants.sh $O/r02b_ab_lorenz_f32.jsonl main main:CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM_CHUNK=16 main:CHAOS_B200_SM_WORKERS=16,CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM=0 2>&1 | tail -10
BENCH_ARGS="--kind pmsm_rk4 --substeps 4 --param-jitter 0.1" tools/ab_variants.sh $O/r02b_ab_pmsm_jit.jsonl main main:CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM_CHUNK=16 main:CHAOS_B200_SM_WORKERS=16,CHAOS_B200_SM_CHUNK=8 main:CHAOS_B200_SM=0 2>&1 | tail -10
BENCH_ARGS="--param-jitter 0.1" tools/ab_variants.sh $O/r02b_ab_lorenz_jit.jsonl main main:CHAOS_B200_SM_CHUNK=8 2>&1 | tail -4
BENCH_ARGS="--envs-per-gpu 1048576 --chunk 16" tools/ab_variants.sh $O/r02b_ab_1Mi.jsonl main 2>&1 | tail -2
BENCH_ARGS="--envs-per-gpu 1048576 --chunk 16 --kind lorenz_rk4_f32" tools/ab_variants.sh $O/r02b_ab_1Mi_f32.jsonl main 2>&1 | tail -2
BENCH_ARGS="--envs-per-gpu 1048576 --chunk 16 --kind pmsm_rk4 --substeps 4 --param-jitter 0.1" tools/ab_variants.sh $O/r02b_ab_1Mi_pmsm.jsonl main 2>&1 | tail -2
