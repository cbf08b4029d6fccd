#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2m_rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2m_rc.txt
timeout 600 python tools/exp_gcn.py > gpurun_out/r2m_siblings.log 2>&1; echo "siblings rc=$?" >> gpurun_out/r2m_rc.txt
tail -5 gpurun_out/r2m_tests.log; cat gpurun_out/r2m_rc.txt; tail -2 gpurun_out/r2m_smoke.log; cat gpurun_out/r2m_siblings.log | tail -5
