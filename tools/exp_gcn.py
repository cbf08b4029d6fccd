"""Throughput of the sibling-model drop-ins (GCN, GCNFP, DCNN, ChebyNet) on the bench's QM8-shaped batches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from helpers import deterministic_state_dict  # noqa: E402
from lanczosnetwork_b200 import configs  # noqa: E402
from lanczosnetwork_b200.model import ChebyNet, DCNN, GCN, GCNFP  # noqa: E402

dev = torch.device('cuda:0')
B = bench.BATCH
batches = bench.make_batches(4, B, 1000)
res = [{k: torch.from_numpy(b[k]).to(dev) for k in ('node_feat', 'L', 'node_mask')} for b in batches]
for cls, cfg in ((GCN, configs.qm8_gcn()), (GCNFP, configs.qm8_gcn(name='GCNFP')), (DCNN, configs.qm8_dcnn()),
                 (ChebyNet, configs.qm8_cheby_net())):
  mod = cls(cfg)
  mod.load_state_dict(deterministic_state_dict(mod, 7))
  mod = mod.to(dev).eval()
  with torch.no_grad():
    for i in range(12):
      b = res[i % 4]
      mod(b['node_feat'], b['L'], mask=b['node_mask'])
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20):
      b = res[i % 4]
      mod(b['node_feat'], b['L'], mask=b['node_mask'])
    e.record()
    torch.cuda.synchronize()
  ms = a.elapsed_time(e) / 20
  print('%-6s QM8 forward B=%d: %.3f ms/step, %.2f M molecules/s' % (cls.__name__, B, ms, B / ms / 1e3))
