#!/bin/bash
# first GPU pass: parity tests in isolated processes (a hung kernel only loses its own group)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "not linear" > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "linear" > gpurun_out/t_linear.log 2>&1; echo "linear rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --tb=short > gpurun_out/t_models.log 2>&1; echo "models rc=$?" >> gpurun_out/rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
tail -5 gpurun_out/t_kernels.log gpurun_out/t_linear.log gpurun_out/t_models.log gpurun_out/smoke.log gpurun_out/bench.log
