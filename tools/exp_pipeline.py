"""Pipeline experiments for the tcgen05 skeleton (profiling aid, not part of the product path).
Times lnb_linear_tf32x3 and lnb_spectral_conv_fused with LNB_DBG debug bits set."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lanczosnetwork_b200 import data, ops  # noqa: E402

dev = torch.device('cuda:0')


def timeit(fn, iters=20):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / iters * 1e3   # us


M, N, K = 26624, 128, 1920
x = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev) / 40
bias = torch.randn(N, device=dev)
w_hi, w_lo = ops.split_tf32(w)

batch = data.synthetic_qm8_batch(1024, seed=1)
L = torch.from_numpy(batch['L']).to(dev)
V = torch.from_numpy(batch['V']).to(dev)
X = torch.randn(1024, 26, 128, device=dev)
coeff = torch.randn(1024, 20, 8, device=dev)
prep = ops.graph_prepare(L, V)

for flags in [0, 8, 1, 2, 4, 15]:
  os.environ['LNB_DBG'] = str(flags)
  t_lin = timeit(lambda: ops.linear_tf32x3(x, w_hi, w_lo, bias, True))
  t_fus = timeit(lambda: ops.spectral_conv_fused(X, V, coeff, prep, w_hi, w_lo, bias, True))
  print('LNB_DBG=%2d  linear %8.1f us   fused %8.1f us' % (flags, t_lin, t_fus), flush=True)
os.environ['LNB_DBG'] = '0'
