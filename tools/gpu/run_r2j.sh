#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_models.py -m gpu -x -q -k "train or gradient or dropin_surface" > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2j_rc.txt
tail -30 gpurun_out/r2j_tests.log; cat gpurun_out/r2j_rc.txt
