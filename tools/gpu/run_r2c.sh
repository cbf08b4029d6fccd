#!/bin/bash
# round 2, call C: ncu of the fused Lanczos+QL+Ritz kernel (source-level stall sampling)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_lanczos_ritz_edges" > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c_rc.txt
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lanczos_ritz' -c 4 -o gpurun_out/r2c_prof_fused -f python tools/prof_lanczos.py > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2c_rc.txt
tail -5 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_rc.txt; tail -3 gpurun_out/r2c_ncu.log
