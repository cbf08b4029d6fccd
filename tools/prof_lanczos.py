"""One launch of each Lanczos / Ritz kernel at the BASELINE.json sizes, for ncu captures
(profiling aid).  python tools/prof_lanczos.py [qm8] [64] [256] [1024]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from lanczosnetwork_b200 import data, ops  # noqa: E402

dev = torch.device('cuda:0')
which = sys.argv[1:] or ['qm8', '64', '256', '1024']
reps = int(os.environ.get('REPS', '2'))
if 'qm8' in which:
  b = data.synthetic_qm8_batch(1024, seed=1)
  A = torch.from_numpy(b['L'][..., 0].copy()).to(dev)
  mask = torch.from_numpy(b['node_mask']).to(dev)
  q1 = torch.randn(1024, 26, generator=torch.Generator().manual_seed(1)).to(dev)
  for _ in range(reps):
    ops.lanczos_ritz(A, mask, q1, 20)
for N, G in ((64, 10000), (256, 10000), (1024, 2000)):
  if str(N) not in which:
    continue
  rng = np.random.RandomState(1234 + N)
  base = np.stack([bench.gnp_operator(rng, N, min(0.5, 8.0 / N)) for _ in range(8)])
  Ad = torch.from_numpy(base).to(dev).repeat((G + 7) // 8, 1, 1)[:G].contiguous()
  q1 = torch.randn(G, N, generator=torch.Generator().manual_seed(1234)).to(dev)
  for _ in range(reps):
    ops.lanczos_ritz(Ad, None, q1, 40)
  del Ad
torch.cuda.synchronize()
print('done')
