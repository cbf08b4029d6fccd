#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_lanczos" > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2d_rc.txt
timeout 600 python tools/time_lanczos.py > gpurun_out/r2d_times.log 2>&1; echo "times rc=$?" >> gpurun_out/r2d_rc.txt
tail -12 gpurun_out/r2d_tests.log; cat gpurun_out/r2d_rc.txt; grep -v "want_T" gpurun_out/r2d_times.log | cut -c1-200
