#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "lanczos or tridiag or ritz" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "online or ada or pipeline" 2>&1 | tail -3
timeout 300 python tools/time_lanczos.py 2>&1 | grep -v want_
