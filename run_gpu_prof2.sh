#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc4.txt
NCU=/usr/local/cuda/bin/ncu
timeout 900 $NCU --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:SpectralPolicy -s 50 -c 2 \
   -o gpurun_out/prof_fused_r1c python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
echo "fused rc=$?" >> gpurun_out/rc4.txt
cat gpurun_out/rc4.txt; tail -3 gpurun_out/ncu_fused.log
