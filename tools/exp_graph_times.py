"""Steady-state time of each launch group of the LanczosNet forward, each captured in its own CUDA
graph and replayed back to back (warm caches, no host launch latency) on the bench workload."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from lanczosnetwork_b200 import ops  # noqa: E402
from lanczosnetwork_b200.spectral_conv import ritz_filter_coefficients  # noqa: E402

dev = torch.device('cuda:0')
mod, params = bench.build_model()
mod = mod.to(dev).eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.BATCH
bt = bench.make_batches(1, B, 1000)[0]
t = {k: torch.from_numpy(bt[k]).to(dev) for k in ('node_feat', 'L', 'D', 'V', 'node_mask')}
L, V, D = t['L'].float().contiguous(), t['V'].float().contiguous(), t['D'].float().contiguous()


def graph_time(fn, reps=50):
  s = torch.cuda.Stream()
  s.wait_stream(torch.cuda.current_stream())
  with torch.cuda.stream(s):
    for _ in range(3):
      fn()
  torch.cuda.current_stream().wait_stream(s)
  torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    fn()
  for _ in range(3):
    g.replay()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    g.replay()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps * 1e3


with torch.no_grad():
  prep = ops.graph_prepare(L, V)
  table = ops.ritz_power_table(D, mod.long_diffusion_dist)
  mlp = mod._filter_mlp_params()
  print('graph_prepare (2 kernels)   %.1f us' % graph_time(lambda: ops.graph_prepare(L, V)))
  print('ritz_power_table            %.1f us' % graph_time(lambda: ops.ritz_power_table(D, mod.long_diffusion_dist)))
  print('filter MLP chain            %.1f us' % graph_time(
      lambda: ritz_filter_coefficients(D, mod.long_diffusion_dist, mlp, mod._wcache, prep, table=table)))
  mod.use_cuda_graph = False
  print('whole forward (one graph)   %.1f us' % graph_time(
      lambda: mod(t['node_feat'], t['L'], t['D'], t['V'], mask=t['node_mask'])))
