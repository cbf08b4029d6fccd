#!/bin/bash
# launch list of the bench command (shares of the step) + a fresh full capture of the dominant kernel
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 4 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/r2s_b.log 2>&1; echo "launches rc=$?" > gpurun_out/r2s_rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 6 -c 1 -o gpurun_out/r2s_prof_stack -f python bench.py --steps 2 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/r2s_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2s_rc.txt
cat gpurun_out/r2s_rc.txt; tail -2 gpurun_out/r2s_ncu.log
