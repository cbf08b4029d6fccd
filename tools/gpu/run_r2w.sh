#!/bin/bash
cd /root/repo
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "lanczos or tridiag or ritz" 2>&1 | tail -3
timeout 1500 $CS --tool racecheck --racecheck-report analysis --print-limit 2000 --error-exitcode 9 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_lanczos or graph_messages or operator_chain or segment or lanczos_tridiag or tridiag or gaussian or embedding" > gpurun_out/r2w_racecheck.log 2>&1
echo "racecheck rc=$?" | tee gpurun_out/r2w_rc.txt
grep -E "^========= (Error|Warning)" gpurun_out/r2w_racecheck.log | sed -E 's/0x[0-9a-f]+/X/g; s/\+X//g' | sort | uniq -c | sort -rn | head -40 > gpurun_out/r2w_race_lines.txt
tail -3 gpurun_out/r2w_racecheck.log; cat gpurun_out/r2w_race_lines.txt
true
gzip -f gpurun_out/r2w_racecheck.log
