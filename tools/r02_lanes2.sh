#!/bin/bash
# Helpers-first staging (stage_begin before the launches): parity, sync-launch behaviour, in-process A/B, stream-query sync A/B.
# Both variants measured worse and were reverted (dead ends 4 and 5; results: profiles/r02n_*); the script is kept
# as the record of what was run -- on the current tree it measures the shipped scheme.
O=gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_gpu_vecenv.py -q -m gpu -x > $O/pytest_lanes2.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_lanes2.log
$T 200 python tools/e2e_threads_ab.py lorenz_rk4 1:32,2:32,3:33,4:32,3:64,3:16 > $O/r02n_e2e_helpers_first_lorenz.jsonl 2> $O/r02n.err; cat $O/r02n_e2e_helpers_first_lorenz.jsonl
$T 200 python tools/e2e_threads_ab.py hr_sync 2:32,3:32 > $O/r02n_e2e_helpers_first_hr.jsonl 2>> $O/r02n.err; cat $O/r02n_e2e_helpers_first_hr.jsonl
CHAOS_B200_SYNC=query $T 200 python tools/e2e_threads_ab.py lorenz_rk4 3:32,3:16 > $O/r02n_e2e_sync_query_lorenz.jsonl 2>> $O/r02n.err; cat $O/r02n_e2e_sync_query_lorenz.jsonl
tail -3 $O/r02n.err
