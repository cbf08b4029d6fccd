"""Per-phase clock64 timers of the fused convolution kernel (profiling aid)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lanczosnetwork_b200 import _lib, data, ops  # noqa: E402

dev = torch.device('cuda:0')
batch = data.synthetic_qm8_batch(1024, seed=1)
L = torch.from_numpy(batch['L']).to(dev)
V = torch.from_numpy(batch['V']).to(dev)
X = torch.randn(1024, 26, 128, device=dev)
coeff = torch.randn(1024, 20, 8, device=dev)
w = torch.randn(128, 1920, device=dev) / 40
bias = torch.randn(128, device=dev)
w_hi, w_lo = ops.split_tf32(w)
prep = ops.graph_prepare(L, V)
for _ in range(3):
  ops.spectral_conv_fused(X, V, coeff, prep, w_hi, w_lo, bias, True)
torch.cuda.synchronize()
def run(flag):
  os.environ['LNB_DBG'] = str(flag)
  for _ in range(2):
    ops.spectral_conv_fused(X, V, coeff, prep, w_hi, w_lo, bias, True)
  torch.cuda.synchronize()
  prof = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
  _lib.check(_lib.load().lnb_debug_set_prof(ctypes.c_void_p(prof.data_ptr())), 'set_prof')
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  ops.spectral_conv_fused(X, V, coeff, prep, w_hi, w_lo, bias, True)
  b.record()
  torch.cuda.synchronize()
  _lib.load().lnb_debug_set_prof(None)
  p = prof.cpu().reshape(148, 32).double()
  names = ['stage issue', 'stage wait', 'U', 'k-loop s0', 'pre_epi', 'acc wait s0', 'tmem ld', 'store', 'k-loop s1', 'acc wait s1']
  print('LNB_DBG=%d kernel %.1f us; per-CTA clock64 totals (cycles), CTAs 0..3 and mean over CTAs with 2 tiles:' % (flag, a.elapsed_time(b) * 1e3))
  for i in range(10):
    print('  %-10s cta0 %8d cta1 %8d cta120 %8d  mean(first 100) %8d' % (names[i], p[0, i], p[1, i], p[120, i], p[:100, i].mean()))
  print('  sum cta0 %d' % p[0, :10].sum())


for flag in (0, 1, 8, 9, 2, 4):
  run(flag)
os.environ['LNB_DBG'] = '0'
