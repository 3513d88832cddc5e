#!/bin/bash
export CUDA_LAUNCH_BLOCKING=1
for lib in gym_lorenz_b200/libchaos_b200.so build/variants/lib_nogc.so; do
 for tm in 1 0; do for w in "" 4 13; do
  echo "lib=$lib tmap=$tm workers=$w"
  CHAOS_B200_LIB=$lib CHAOS_B200_SM_TMAP=$tm CHAOS_B200_SM_WORKERS=$w timeout 120 python tools/debug_sm.py 65536 16 2>&1 | tail -1
 done; done
 CHAOS_B200_LIB=$lib timeout 120 python tools/debug_sm.py 65536 256 2>&1 | tail -1
 CHAOS_B200_LIB=$lib CHAOS_B200_DYN=1 timeout 120 python tools/debug_sm.py 4096 32 2>&1 | tail -1
done
