#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "graphed" 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2v_launches_train.csv python tools/prof_train.py 3 > gpurun_out/r2v_train.log 2>&1
tail -2 gpurun_out/r2v_train.log
