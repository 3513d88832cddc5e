#!/bin/bash
O=gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_gpu_rl_ops.py tests/test_gpu_vecenv.py tests/test_eval_and_derivative_goldens.py -q -m gpu > $O/pytest_gpu_r02j.log 2>&1; echo "rl_ops + vecenv rc=$?"; tail -3 $O/pytest_gpu_r02j.log
for arr in 4 16; do for th in 1 2; do
  E2E_ARRAYS=$arr CHAOS_B200_COPY_THREADS=$th $T 200 python tools/e2e_modes.py lorenz_rk4 65536 zerocopy:1,streamed:32 >> $O/r02j_e2e_threads.jsonl 2>> $O/r02j_e2e.err
  E2E_ARRAYS=$arr CHAOS_B200_COPY_THREADS=$th $T 200 python tools/e2e_modes.py hr_sync 65536 streamed:32 >> $O/r02j_e2e_threads.jsonl 2>> $O/r02j_e2e.err
done; done
cat $O/r02j_e2e_threads.jsonl
$T 600 ncu --set full --clock-control none -k regex:'k_gae|k_moments|k_normalize|k_frame_stack|k_eval|k_rms' -f -o $O/prof_rlops_r02j python tools/profile_hbm_kernels.py rl_ops > $O/ncu_rlops_r02j.log 2>&1
python tools/ncu_kernels_summary.py $O/prof_rlops_r02j.ncu-rep > $O/r02j_rl_ops_ncu_metrics.txt 2>&1; rm -f $O/prof_rlops_r02j.ncu-rep
grep -E "^==|duration" $O/r02j_rl_ops_ncu_metrics.txt
