#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "conv or stack or linear or mlp or readout" > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2i_rc.txt
python tools/exp_stack_phases.py > gpurun_out/r2i_phases.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-workloads --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?" >> gpurun_out/r2i_rc.txt
tail -4 gpurun_out/r2i_tests.log; cat gpurun_out/r2i_rc.txt; cat gpurun_out/r2i_phases.log | head -22; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel ms',d['roofline']['avg_ms_per_launch'])"
