#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fused_lanczos or online_ritz or ada" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2f_rc.txt
timeout 600 python tools/study_online_eigs.py > gpurun_out/r2f_study.md 2> gpurun_out/r2f_study.err; echo "study rc=$?" >> gpurun_out/r2f_rc.txt
tail -15 gpurun_out/r2f_tests.log; cat gpurun_out/r2f_rc.txt; cat gpurun_out/r2f_study.md; tail -3 gpurun_out/r2f_study.err
