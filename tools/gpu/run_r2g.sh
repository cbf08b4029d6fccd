#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "sparse" > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2g_rc.txt
tail -25 gpurun_out/r2g_tests.log; cat gpurun_out/r2g_rc.txt
