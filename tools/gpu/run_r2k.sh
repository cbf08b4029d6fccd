#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "ada" > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2k_rc.txt
timeout 600 python tools/bench_configs.py ada > gpurun_out/r2k_ada.log 2>&1; echo "ada rc=$?" >> gpurun_out/r2k_rc.txt
tail -8 gpurun_out/r2k_tests.log; cat gpurun_out/r2k_rc.txt; tail -4 gpurun_out/r2k_ada.log
