#!/bin/bash
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
# 1) launch list of one warm forward (skip the warm-up launches: 3 warmup x2 x59 + check ~ 420)
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -s 400 -c 140 --csv \
   --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch rc=$?" >> gpurun_out/rc2.txt
# 2) full capture of the dominant kernel (2 launches) and of the FFMA batched GEMM (3 launches)
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:linear_tf32x3 -s 30 -c 2 \
   -o gpurun_out/prof_linear_r1a python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_linear.log 2>&1
echo "linear rc=$?" >> gpurun_out/rc2.txt
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:batched_gemm -s 200 -c 4 \
   -o gpurun_out/prof_bgemm_r1a python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bgemm.log 2>&1
echo "bgemm rc=$?" >> gpurun_out/rc2.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "segment" > gpurun_out/t_seg.log 2>&1; echo "seg rc=$?" >> gpurun_out/rc2.txt
cat gpurun_out/rc2.txt; ls -la gpurun_out
