#!/bin/bash
# round 2, call A: tests + bench + launch list + ncu of the round-1 Lanczos kernels (baseline)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2a_rc.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lanczos|ritz' -c 8 -o gpurun_out/r2a_prof_lanczos -f python tools/prof_lanczos.py > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2a_rc.txt
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_rc.txt; cat gpurun_out/r2a_bench.json | head -c 6000
