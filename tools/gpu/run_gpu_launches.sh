#!/bin/bash
# per-kernel launch list of one bench run (eager launches, ncu serialises: cold-cache times)
mkdir -p gpurun_out
TAG=${1:-tmp}
export LNB_NO_GRAPH=1
timeout 900 /usr/local/cuda/bin/ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch rc=$?"
