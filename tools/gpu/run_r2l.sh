#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q -k "graph_messages or ada or general or sibling or dcnn or cheby or power_filter" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2l_rc.txt
timeout 600 python tools/bench_configs.py ada general > gpurun_out/r2l_ada.log 2>&1; echo "ada rc=$?" >> gpurun_out/r2l_rc.txt
tail -12 gpurun_out/r2l_tests.log; cat gpurun_out/r2l_rc.txt; tail -4 gpurun_out/r2l_ada.log
