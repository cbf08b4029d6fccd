"""Accuracy study of the online (D, V) provider: Ritz pairs from the fused
Lanczos+QL kernel against the reference's offline fp64 eigh top-K, on QM8-shaped molecules and on
larger G(n,p) graphs, and the effect on LanczosNet's scores.  Writes a markdown table to stdout
(copied into profiles/r2_online_eigs_study.md)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from lanczosnetwork_b200 import data, provider  # noqa: E402

dev = torch.device('cuda:0')
K = 20
mod, params = bench.build_model()
mod = mod.to(dev).eval()
print('| set | graphs | Ritz pairs returned / eigh pairs (non-zero) | max abs (theta - lambda) over matched values | '
      'median / max abs (V theta V^T - V_e lambda V_e^T) | LanczosNet score: median / max abs change |')
print('|---|---|---|---|---|---|')


def study(tag, batch):
  t = {k: torch.from_numpy(batch[k]).to(dev) for k in ('node_feat', 'L', 'D', 'V', 'node_mask')}
  th, V, info = provider.online_ritz_pairs(t['L'], t['node_mask'], K, generator=torch.Generator(device=dev).manual_seed(7))
  assert int((info['status'] & 1).sum()) == 0
  D_e, V_e = batch['D'].astype(np.float64), batch['V'].astype(np.float64)
  th_n, V_n = th.cpu().numpy().astype(np.float64), V.cpu().numpy().astype(np.float64)
  got = int((np.abs(V_n).sum(axis=1) > 0).sum())
  want = int((np.abs(V_e).sum(axis=1) > 0).sum())
  # every Ritz value against the nearest exact eigenvalue of the operator
  A = batch['L'][..., 0].astype(np.float64)
  errs = []
  for b in range(A.shape[0]):
    n = int(batch['node_mask'][b].sum())
    lam = np.linalg.eigvalsh(A[b, :n, :n])
    k = int(info['idx'][b].item())
    for v in th_n[b, :k]:
      errs.append(np.abs(lam - v).min())
  rec_r = np.einsum('bnk,bk,bmk->bnm', V_n, th_n, V_n)
  rec_e = np.einsum('bnk,bk,bmk->bnm', V_e, D_e, V_e)
  dr = np.abs(rec_r - rec_e).reshape(A.shape[0], -1).max(axis=1)
  with torch.no_grad():
    s_e = mod(t['node_feat'], t['L'], t['D'], t['V'], mask=t['node_mask']).cpu().numpy()
    s_r = mod(t['node_feat'], t['L'], th, V, mask=t['node_mask']).cpu().numpy()
  ds = np.abs(s_e - s_r).max(axis=1)
  print('| %s | %d | %d / %d | %.2e | %.2e / %.2e | %.2e / %.2e |' % (
      tag, A.shape[0], got, want, max(errs), np.median(dr), dr.max(), np.median(ds), ds.max()))


study('QM8-shaped, n_b in [3,26], K=20', data.synthetic_qm8_batch(1024, seed=11))
small = data.synthetic_qm8_batch(1024, seed=12, max_nodes=18)
study('QM8-shaped, n_b <= 18 < K', small)
