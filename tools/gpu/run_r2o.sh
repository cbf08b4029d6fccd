#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_lanczos" > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2o_rc.txt
timeout 600 python tools/time_lanczos.py > gpurun_out/r2o_times.log 2>&1; echo "times rc=$?" >> gpurun_out/r2o_rc.txt
LNB_LANCZOS_TPG=1024 timeout 600 python tools/time_lanczos.py 1024 > gpurun_out/r2o_times1024.log 2>&1; echo "times1024 rc=$?" >> gpurun_out/r2o_rc.txt
tail -4 gpurun_out/r2o_tests.log; cat gpurun_out/r2o_rc.txt; grep -v "want_T\|want_ritz" gpurun_out/r2o_times.log | cut -c1-180; echo ---- TPG1024; grep -v "want_T\|want_ritz" gpurun_out/r2o_times1024.log | cut -c1-180
