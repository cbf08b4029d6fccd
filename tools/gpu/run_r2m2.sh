#!/bin/bash
# the whole GPU suite under memcheck on the final tree
cd /root/repo
timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q > gpurun_out/r2m2_memcheck_all.log 2>&1
echo "memcheck all rc=$?" | tee gpurun_out/r2m2_rc.txt
tail -6 gpurun_out/r2m2_memcheck_all.log
