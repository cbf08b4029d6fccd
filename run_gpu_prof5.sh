#!/bin/bash
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
export LNB_NO_GRAPH=1
timeout 600 $NCU --set full --warp-sampling-interval 0 --clock-control none --import-source on --kernel-name-base demangled -k regex:graph_prepare -s 4 -c 1 \
   -o gpurun_out/prof_prepare python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_prep.log 2>&1
echo "prep rc=$?"
timeout 600 $NCU --set full --warp-sampling-interval 0 --clock-control none --import-source on --kernel-name-base demangled -k regex:tile_assign -s 4 -c 1 \
   -o gpurun_out/prof_assign python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_assign.log 2>&1
echo "assign rc=$?"
