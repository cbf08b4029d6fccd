#!/bin/bash
# launch list of one bench run (eager launches) + full ncu captures of the two tcgen05 kernels
mkdir -p gpurun_out; rm -f gpurun_out/rc5.txt
NCU=/usr/local/cuda/bin/ncu
TAG=${1:-r1f}
export LNB_NO_GRAPH=1
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch rc=$?" >> gpurun_out/rc5.txt
timeout 900 $NCU --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:SpectralPolicy -s 4 -c 1 \
   -o gpurun_out/prof_stack_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stack.log 2>&1
echo "stack rc=$?" >> gpurun_out/rc5.txt
timeout 900 $NCU --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:mlp_chain -s 4 -c 1 \
   -o gpurun_out/prof_mlp_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mlp.log 2>&1
echo "mlp rc=$?" >> gpurun_out/rc5.txt
cat gpurun_out/rc5.txt
