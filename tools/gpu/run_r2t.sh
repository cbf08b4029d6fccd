#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python - <<'PY' 2>&1 | tail -20
import sys, json, torch
sys.path.insert(0, 'tests')
import bench
print(json.dumps(bench.train_workload(torch.device('cuda:0')), indent=1))
PY
