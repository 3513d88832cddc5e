#!/bin/bash
# Staging lanes (round 2, session 3): parity of the host modes with 1..4 copy threads, the SB3-artefact GPU tests,
# and the in-process A/B of thread / slice counts.
O=gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_gpu_vecenv.py tests/test_sb3_artefacts.py -q -m gpu -x > $O/pytest_lanes.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_lanes.log
nproc; lscpu | grep -E "Model name|Thread|Core|Socket|NUMA" 
$T 200 python tools/e2e_threads_ab.py lorenz_rk4 > $O/r02l_e2e_lanes_lorenz.jsonl 2> $O/r02l_e2e_lanes.err; cat $O/r02l_e2e_lanes_lorenz.jsonl
$T 200 python tools/e2e_threads_ab.py hr_sync 2:32,4:32,4:64 > $O/r02l_e2e_lanes_hr.jsonl 2>> $O/r02l_e2e_lanes.err; cat $O/r02l_e2e_lanes_hr.jsonl
$T 200 python tools/e2e_threads_ab.py pmsm_sync 2:32,4:32 > $O/r02l_e2e_lanes_pmsm.jsonl 2>> $O/r02l_e2e_lanes.err; cat $O/r02l_e2e_lanes_pmsm.jsonl
tail -5 $O/r02l_e2e_lanes.err
