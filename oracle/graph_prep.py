"""Oracle for the graph preparation that feeds the hot path (numpy, fp64 like the
reference's offline preprocessors).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows:
  * utils/data_helper.py:92-116   normalize_adj
  * utils/data_helper.py:119-166  get_laplacian ('L4' branch :155-156)
  * utils/data_helper.py:169-226  get_graph_laplacian_eigs (dense eigh branch)
  * dataset/qm8.py:57-90,220-291  collate_fn (default branch)
"""
import numpy as np


def sym_normalize(mat, exponent=0.5):
  """D^-e M D^-e with D = rowsum(M); inf -> 0 (utils/data_helper.py:92-116)."""
  mat = np.asarray(mat, dtype=np.float64)
  deg = mat.sum(axis=1)
  with np.errstate(divide='ignore'):
    scale = np.power(deg, -exponent)
  scale[np.isinf(scale)] = 0.0
  # reference: r_mat_inv.dot(A).dot(r_mat_inv) with diagonal r_mat_inv
  return np.diag(scale).dot(mat).dot(np.diag(scale))


def laplacian_L4(adj):
  """GCN renormalisation D~^-1/2 (I + A) D~^-1/2 (utils/data_helper.py:155-156)."""
  adj = np.asarray(adj, dtype=np.float64)
  return sym_normalize(np.eye(adj.shape[0]) + adj)


def eig_topk_by_magnitude(lap, k=100):
  """Dense eigh, sort by -|lambda| (stable), keep k (utils/data_helper.py:199-223)."""
  vals, vecs = np.linalg.eigh(lap)
  order = np.argsort(-np.abs(vals), kind='mergesort')[:k]
  return vals[order], vecs[:, order]


def prepare_molecule(adjs):
  """What dataset/get_qm8_data.py:60-90 stores per molecule (keys the collate uses).

  adjs: n x n x E binary adjacency per bond type.
  """
  adjs = np.asarray(adjs, dtype=np.float64)
  simple = adjs.sum(axis=2)
  L_multi = np.stack([laplacian_L4(adjs[:, :, e]) for e in range(adjs.shape[2])],
                     axis=2)
  L_simple = laplacian_L4(simple)
  D, V = eig_topk_by_magnitude(L_simple)
  return {'L_multi': L_multi, 'L_simple_4': L_simple, 'D_simple': D, 'V_simple': V}


def collate(samples, num_eigs):
  """Pad to the batch-max node count (dataset/qm8.py:57-90,220-291).

  samples: list of dicts with node_feat (n,) or (n,d), L_multi (n,n,E),
           L_simple_4 (n,n), D_simple (<=100,), V_simple (n,<=100).
  returns float32/int64/uint8 numpy arrays keyed like the reference batch dict.
  """
  sizes = [s['L_simple_4'].shape[0] for s in samples]
  N = max(sizes)
  B = len(samples)
  E = samples[0]['L_multi'].shape[2]
  nf0 = np.asarray(samples[0]['node_feat'])
  if nf0.ndim == 1:
    node_feat = np.zeros((B, N), dtype=np.int64)
  else:
    node_feat = np.zeros((B, N, nf0.shape[1]), dtype=np.float32)
  mask = np.zeros((B, N), dtype=np.uint8)
  L = np.zeros((B, N, N, E + 1), dtype=np.float32)
  D = np.zeros((B, num_eigs), dtype=np.float32)
  V = np.zeros((B, N, num_eigs), dtype=np.float32)
  for b, s in enumerate(samples):
    n = sizes[b]
    node_feat[b, :n] = s['node_feat']
    mask[b, :n] = 1
    L[b, :n, :n, 0] = s['L_simple_4']       # dataset/qm8.py:262 cat(simple, multi)
    L[b, :n, :n, 1:] = s['L_multi']
    kk = min(num_eigs, len(s['D_simple']))  # dataset/qm8.py:268-287 truncate / zero-pad
    D[b, :kk] = s['D_simple'][:kk]
    V[b, :n, :kk] = s['V_simple'][:, :kk]
  return {'node_feat': node_feat, 'node_mask': mask, 'L': L, 'D': D, 'V': V}
