"""Shared test utilities: deterministic (numpy-seeded, platform independent) parameters so
golden vectors need not store weights, spec extraction for the oracle, fixture loading."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def deterministic_state_dict(module, seed):
  """Same names/shapes as module.state_dict(); values from numpy RandomState(seed + i):
  2-D Linear weights Xavier-uniform, biases U(-0.1, 0.1), embeddings N(0,1)."""
  out = {}
  for i, (name, t) in enumerate(module.state_dict().items()):
    rng = np.random.RandomState(seed + i)
    shape = tuple(t.shape)
    if name.startswith('embedding'):
      v = rng.randn(*shape)
    elif len(shape) == 2:
      bound = np.sqrt(6.0 / (shape[0] + shape[1]))
      v = rng.uniform(-bound, bound, size=shape)
    else:
      v = rng.uniform(-0.1, 0.1, size=shape)
    out[name] = torch.from_numpy(v.astype(np.float32))
  return out


def oracle_spec(module, kind):
  from oracle.lanczos_oracle import make_spec
  return make_spec(module.short_diffusion_dist, module.long_diffusion_dist, module.num_edgetype,
                   module.num_layer, module.num_eig_vec, module.spectral_filter_kind, kind)


def load_golden(name):
  return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))
