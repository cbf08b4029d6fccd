#!/bin/bash
cd /root/repo
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool racecheck --racecheck-report analysis --print-limit 2000 --error-exitcode 9 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_lanczos or lanczos_tridiag or tridiag or lanczos_matches" > gpurun_out/r2w_racecheck.log 2>&1
echo "racecheck rc=$?" | tee gpurun_out/r2w_rc.txt
grep -E "^========= (Error|Warning)" gpurun_out/r2w_racecheck.log | sed -E 's/0x[0-9a-f]+/X/g; s/\+X//g' | sort | uniq -c | sort -rn | head -20
tail -3 gpurun_out/r2w_racecheck.log
timeout 600 $CS --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_lanczos or lanczos_tridiag" 2>&1 | tail -3
gzip -f gpurun_out/r2w_racecheck.log
