"""Online (D, V) provider: Ritz pairs of the simple-graph operator computed on the GPU by the fused
Lanczos -> QL -> Ritz-vector kernel, in place of the offline fp64 ``eigh`` of the reference's
preprocessing (utils/data_helper.py:169-226 called from dataset/get_qm8_data.py:63-83, truncated /
padded to K at collate, dataset/qm8.py:265-291).

This is SURVEY 8(f4) and the paper's actual algorithm; it is a MODEL-INPUT CHANGE, never "reference
MAE": K Lanczos steps from one start vector span a Krylov space, so

  * a graph with n_b <= K real nodes and simple eigenvalues gets all its eigenpairs (to fp32
    rounding) -- but ordered / signed by QL, and the reference's (D, V) are only defined up to sign
    and to rotations inside degenerate eigenspaces anyway;
  * a repeated eigenvalue (symmetric molecules) contributes ONE Ritz vector (the projection of the
    start vector onto its eigenspace), so fewer than min(n_b, K) non-zero pairs come back where
    ``eigh`` returns an arbitrary basis of the eigenspace;
  * a graph with n_b > K gets K Ritz pairs approximating the extremal part of the spectrum, not
    the exact top-K by |lambda|.

``tools/study_online_eigs.py`` measures all three effects and what they do to LanczosNet's scores
(profiles/r2_online_eigs_study.md).
"""
import torch

from . import ops

__all__ = ['online_ritz_pairs']


def online_ritz_pairs(L, mask, num_eigs, q1=None, generator=None):
  """(D [B,K], V [B,N,K], info) from channel 0 of the padded operator tensor L [B,N,N,E+1]
  (or a [B,N,N] operator), ready for ``LanczosNet.forward(node_feat, L, D, V, mask=mask)``.

  q1: start vectors [B,N] (device); default: standard normal draws from ``generator`` on the
  operator's device (masked and normalised by the kernel like model/ada_lanczos_net.py:159-167).
  info: dict(idx [B] retained Krylov directions, status [B] kernel status bits)."""
  A = L[..., 0] if L.dim() == 4 else L
  A = A.float().contiguous()
  B, N = A.shape[0], A.shape[1]
  if q1 is None:
    q1 = torch.randn(B, N, device=A.device, generator=generator)
  out = ops.lanczos_ritz(A, mask, q1, int(num_eigs), want_T=False, want_Q=False, proper=True)
  return out['theta'], out['V'], {'idx': out['idx'], 'status': out['status']}
