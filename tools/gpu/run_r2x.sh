#!/bin/bash
# racecheck over everything (tcgen05 kernels included, to see what the tool says about them)
cd /root/repo
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool racecheck --racecheck-report analysis --print-limit 2000 --error-exitcode 9 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -q > gpurun_out/r2x_racecheck_kernels.log 2>&1
echo "racecheck kernels rc=$?" | tee gpurun_out/r2x_rc.txt
grep -E "^========= (Error|Warning)" gpurun_out/r2x_racecheck_kernels.log | sed -E 's/0x[0-9a-f]+/X/g; s/\+X//g' | sort | uniq -c | sort -rn | head -40
tail -3 gpurun_out/r2x_racecheck_kernels.log
timeout 1200 $CS --tool racecheck --racecheck-report analysis --print-limit 2000 --error-exitcode 9 \
  python -m pytest tests/test_gpu_models.py -m gpu -q -k "sparse or golden or online" > gpurun_out/r2x_racecheck_models.log 2>&1
echo "racecheck models rc=$?" | tee -a gpurun_out/r2x_rc.txt
grep -E "^========= (Error|Warning)" gpurun_out/r2x_racecheck_models.log | sed -E 's/0x[0-9a-f]+/X/g; s/\+X//g' | sort | uniq -c | sort -rn | head -40
tail -3 gpurun_out/r2x_racecheck_models.log
gzip -f gpurun_out/r2x_racecheck_kernels.log gpurun_out/r2x_racecheck_models.log
