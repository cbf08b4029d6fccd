#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (memcheck: whole kernel test file; racecheck + synccheck:
# the shared-memory-heavy kernels at small sizes).  Logs -> gpurun_out/r2u_*.log
cd /root/repo
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r2u_memcheck.log 2>&1
echo "memcheck rc=$?" | tee gpurun_out/r2u_rc.txt
tail -4 gpurun_out/r2u_memcheck.log
timeout 900 $CS --tool memcheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "sparse or online or golden" > gpurun_out/r2u_memcheck_models.log 2>&1
echo "memcheck models rc=$?" | tee -a gpurun_out/r2u_rc.txt
tail -4 gpurun_out/r2u_memcheck_models.log
timeout 1200 $CS --tool racecheck --racecheck-report all --print-limit 200000 --error-exitcode 9 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_lanczos_ritz_matches or fused_lanczos_ritz_edges or graph_messages or operator_chain" > gpurun_out/r2u_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2u_rc.txt
grep -c "Race reported" gpurun_out/r2u_racecheck.log; tail -4 gpurun_out/r2u_racecheck.log
timeout 600 $CS --tool synccheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_lanczos or graph_messages or operator_chain or segment" > gpurun_out/r2u_synccheck.log 2>&1
echo "synccheck rc=$?" | tee -a gpurun_out/r2u_rc.txt
tail -4 gpurun_out/r2u_synccheck.log
