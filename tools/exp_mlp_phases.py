"""MMA-warp wait breakdown of the filter-MLP chain kernel on the bench workload (profiling aid)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from lanczosnetwork_b200 import _lib, ops  # noqa: E402

dev = torch.device('cuda:0')
mod, params = bench.build_model()
mod = mod.to(dev).eval()
mod.use_cuda_graph = False
bt = bench.make_batches(1, bench.BATCH, 1000)[0]
t = {k: torch.from_numpy(bt[k]).to(dev) for k in ('node_feat', 'L', 'D', 'V', 'node_mask')}
prof = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
lib = _lib.load()
orig = ops.ritz_filter_mlp


def probed(*a, **kw):
  _lib.check(lib.lnb_debug_set_prof(ctypes.c_void_p(prof.data_ptr())), 'set_prof')
  r = orig(*a, **kw)
  torch.cuda.synchronize()
  lib.lnb_debug_set_prof(None)
  return r


with torch.no_grad():
  for _ in range(3):
    mod(t['node_feat'], t['L'], t['D'], t['V'], mask=t['node_mask'])
  ops.ritz_filter_mlp = probed
  mod(t['node_feat'], t['L'], t['D'], t['V'], mask=t['node_mask'])
  ops.ritz_filter_mlp = orig
p = prof.cpu().reshape(148, 32).double()
for i, nm in enumerate(['wait acc drained', 'wait W tile', 'wait A k-block', 'issue', 'total', 'A wait kb0', 'A wait kb1', 'A wait kb2', 'A wait kb3', 'W wait kb0', 'W wait kb1', 'W wait kb2', 'W wait kb3', '-', '-', '-', 'g0 wait acc_full', 'g0 drain', 'g0 parked flush', 'g0 first emit', 'g0 second chunk', 'g0 first stage', 'g0 output stage']):
  print('  %-18s cta0 %8d cta1 %8d cta100 %8d mean %8d max %8d' % (nm, p[0, i], p[1, i], p[100, i], p[:, i].mean(), p[:, i].max()))
