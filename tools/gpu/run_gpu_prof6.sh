#!/bin/bash
# full ncu capture of one kernel (regex $1) of the bench forward -> gpurun_out/prof_$2.ncu-rep
mkdir -p gpurun_out
export LNB_NO_GRAPH=1
timeout 600 /usr/local/cuda/bin/ncu --set full --warp-sampling-interval 1 --clock-control none --import-source on --kernel-name-base demangled \
   -k regex:$1 -s 4 -c 1 -o gpurun_out/prof_$2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$2.log 2>&1
echo "$2 rc=$?"
