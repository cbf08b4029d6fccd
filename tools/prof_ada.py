"""One eager AdaLanczosNet forward at B=256 for an ncu launch list (profiling aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import deterministic_state_dict
from lanczosnetwork_b200 import configs, data
from lanczosnetwork_b200.model import AdaLanczosNet
dev = torch.device('cuda:0')
mod = AdaLanczosNet(configs.qm8_ada_lanczos_net())
mod.load_state_dict(deterministic_state_dict(mod, 2024))
mod = mod.to(dev).eval()
mod.use_cuda_graph = False
b = data.synthetic_qm8_batch(256, seed=3)
args = [torch.from_numpy(b[k]).to(dev) for k in ('node_feat', 'L')]
mask = torch.from_numpy(b['node_mask']).to(dev)
with torch.no_grad():
  for _ in range(3):
    mod(*args, mask=mask)
torch.cuda.synchronize()
print('done')
