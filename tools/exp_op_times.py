"""CUDA-event time of every kernel launch group of one LanczosNet forward on the bench workload
(eager launches behind a GPU spin so host launch latency is excluded; warm L2)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from lanczosnetwork_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
mod, params = bench.build_model()
mod = mod.to(dev).eval()
mod.use_cuda_graph = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.BATCH
bt = bench.make_batches(1, B, 1000)[0]
t = {k: torch.from_numpy(bt[k]).to(dev) for k in ('node_feat', 'L', 'D', 'V', 'node_mask')}
names = ['graph_prepare', 'ritz_power_table', 'ritz_filter_mlp', 'spectral_stack_forward']
orig = {n: getattr(ops, n) for n in names}
times = {n: [] for n in names}


def wrap(n):
  def f(*a, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(3000000)
    e0.record()
    r = orig[n](*a, **kw)
    e1.record()
    times[n].append((e0, e1))
    return r
  return f


with torch.no_grad():
  for _ in range(3):
    mod(t['node_feat'], t['L'], t['D'], t['V'], mask=t['node_mask'])
  for n in names:
    setattr(ops, n, wrap(n))
  for _ in range(10):
    mod(t['node_feat'], t['L'], t['D'], t['V'], mask=t['node_mask'])
  torch.cuda.synchronize()
tot = 0.0
for n in names:
  us = sorted(a.elapsed_time(b) * 1e3 for a, b in times[n])
  if us:
    print('%-24s median %.1f us  min %.1f  (%d calls)' % (n, us[len(us) // 2], us[0], len(us)))
    tot += us[len(us) // 2]
print('sum of medians %.1f us' % tot)
