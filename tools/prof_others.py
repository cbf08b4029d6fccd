"""One eager pass over the round-2 kernels that have no capture yet (profiling aid for ncu): the GPU-side
batch construction of a packed QM8 batch (B=1024), one AdaLanczosNet forward (B=256: Gaussian Laplacian,
fused Lanczos without QL, tridiagonal powers, the 4096-wide MLP GEMMs, symmetrise, graph_messages,
operator chain), one DCNN forward (B=256)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import deterministic_state_dict
from lanczosnetwork_b200 import configs, data
from lanczosnetwork_b200.model import AdaLanczosNet, DCNN, LanczosNet
dev = torch.device('cuda:0')

def build(cls, cfg, seed):
  m = cls(cfg)
  m.load_state_dict(deterministic_state_dict(m, seed))
  m = m.to(dev).eval()
  m.use_cuda_graph = False
  return m

with torch.no_grad():
  ln = build(LanczosNet, configs.qm8_lanczos_net(), 1)
  samples = data.synthetic_qm8_samples(1024, seed=5)
  pk = data.pack_sparse(data.sparse_collate(samples, 20))
  batch = {'blob': torch.from_numpy(pk['blob']).to(dev), 'B': pk['B'], 'N': pk['N'], 'K': pk['K']}
  for _ in range(2):
    ln.forward_sparse(batch)
  ada = build(AdaLanczosNet, configs.qm8_ada_lanczos_net(), 2024)
  b = data.synthetic_qm8_batch(256, seed=3)
  t = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
  for _ in range(2):
    ada(t['node_feat'], t['L'], mask=t['node_mask'])
  dc = build(DCNN, configs.qm8_dcnn(), 3)
  for _ in range(2):
    dc(t['node_feat'], t['L'], mask=t['node_mask'])
torch.cuda.synchronize()
print('done')
