"""Eager LanczosNet training steps at B=64 for an ncu launch list (profiling aid)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import deterministic_state_dict
from lanczosnetwork_b200 import configs, data
from lanczosnetwork_b200.model import LanczosNet
dev = torch.device('cuda:0')
mod = LanczosNet(configs.qm8_lanczos_net())
mod.load_state_dict(deterministic_state_dict(mod, 7))
mod = mod.to(dev).train()
opt = torch.optim.Adam(mod.parameters(), lr=1e-4)
b = data.collate(data.synthetic_qm8_samples(64, seed=900), 20, num_nodes=27)
b['label'] = np.random.RandomState(0).randn(64, 16).astype(np.float32)
t = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
  if i == 2:
    torch.cuda.synchronize(); print('MARK')
  opt.zero_grad()
  _, loss = mod(t['node_feat'], t['L'], t['D'], t['V'], label=t['label'], mask=t['node_mask'])
  loss.backward()
  opt.step()
torch.cuda.synchronize()
print('done', float(loss.detach()))
