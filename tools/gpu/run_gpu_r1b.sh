#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rc3.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "linear" -s > gpurun_out/t_linear.log 2>&1; echo "linear rc=$?" >> gpurun_out/rc3.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "fused" > gpurun_out/t_fused.log 2>&1; echo "fused rc=$?" >> gpurun_out/rc3.txt
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --tb=short > gpurun_out/t_models.log 2>&1; echo "models rc=$?" >> gpurun_out/rc3.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc3.txt
cat gpurun_out/rc3.txt
