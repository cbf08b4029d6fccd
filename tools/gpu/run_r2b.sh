#!/bin/bash
# round 2, call B: fused Lanczos+QL+Ritz kernel: parity, then timings
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_lanczos or lanczos_matches or tridiag_ritz" > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2b_rc.txt
timeout 600 python tools/time_lanczos.py > gpurun_out/r2b_times.log 2>&1; echo "times rc=$?" >> gpurun_out/r2b_rc.txt
tail -30 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_rc.txt; cat gpurun_out/r2b_times.log
