#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc5.txt
NCU=/usr/local/cuda/bin/ncu
export LNB_NO_GRAPH=1
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -s 96 -c 34 --csv \
   --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch rc=$?" >> gpurun_out/rc5.txt
timeout 900 $NCU --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:SpectralPolicy -s 50 -c 2 \
   -o gpurun_out/prof_fused_r1e python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
echo "fused rc=$?" >> gpurun_out/rc5.txt
cat gpurun_out/rc5.txt
