import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
from helpers import deterministic_state_dict, oracle_spec
from lanczosnetwork_b200 import configs, data
from lanczosnetwork_b200.model import AdaLanczosNet
from oracle import lanczos_oracle as orc
dev = torch.device('cuda:0')
cfg = configs.qm8_ada_lanczos_net()
mod = AdaLanczosNet(cfg); params = deterministic_state_dict(mod, 2024); mod.load_state_dict(params); mod = mod.to(dev).eval()
for seed, B in ((17, 6), (18, 12)):
  batch = data.synthetic_qm8_batch(B, seed=seed)
  N = batch['node_feat'].shape[1]
  torch.manual_seed(5); q1 = torch.randn(B, N, 1)
  spec = oracle_spec(mod, 'AdaLanczosNet')
  ref32 = orc.ada_lanczos_net_forward(params, spec, batch['node_feat'], batch['L'], batch['node_mask'], q1[:, :, 0]).numpy()
  ref64 = orc.ada_lanczos_net_forward(params, spec, batch['node_feat'], batch['L'], batch['node_mask'], q1[:, :, 0].double(), dtype=torch.float64).numpy()
  torch.manual_seed(5)
  with torch.no_grad():
    out = mod(torch.from_numpy(batch['node_feat']).to(dev), torch.from_numpy(batch['L']).to(dev), mask=torch.from_numpy(batch['node_mask']).to(dev)).cpu().numpy()
  print('B', B, 'scale', np.abs(ref64).max(), 'ours-ref64', np.abs(out - ref64).max(), 'ref32-ref64', np.abs(ref32 - ref64).max(), 'ours-ref32', np.abs(out - ref32).max())
