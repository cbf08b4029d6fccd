#!/bin/bash
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
export LNB_NO_GRAPH=1
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -s 350 -c 120 --csv \
   --log-file gpurun_out/launches_ada.csv python tools/bench_configs.py ada > gpurun_out/ncu_ada.log 2>&1
echo "ada rc=$?"
