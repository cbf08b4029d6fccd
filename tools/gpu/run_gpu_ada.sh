#!/bin/bash
# launch list of the AdaLanczosNet forward (eager launches)
mkdir -p gpurun_out
export LNB_NO_GRAPH=1
timeout 900 /usr/local/cuda/bin/ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_ada.csv python tools/bench_configs.py ada > gpurun_out/ncu_ada.log 2>&1
echo "ada rc=$?"
tail -5 gpurun_out/ncu_ada.log
