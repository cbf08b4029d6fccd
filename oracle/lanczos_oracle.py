"""Functional torch-CPU oracle of the LanczosNet / AdaLanczosNet / LanczosNetGeneral
forward path.  TEST INFRASTRUCTURE -- see oracle/__init__.py.

All functions take a flat ``params`` dict keyed exactly like the reference
``state_dict`` (embedding.weight, filter.{i}.weight, spectral_filter.{l}.{0,2,4,6}.weight,
att_func.0.weight, ...) and a small ``spec`` dict (see ``make_spec``).  ``dtype``
selects fp32 (parity with the reference) or fp64 (rounding budget).

Reference lines followed (relative to /root/reference):
  model/lanczos_net.py:95-123     spectral filters from Ritz pairs
  model/lanczos_net.py:125-199    LanczosNet.forward
  model/lanczos_net_general.py:127-201  LanczosNetGeneral.forward (float node features)
  model/ada_lanczos_net.py:101-137  Gaussian-kernel Laplacian
  model/ada_lanczos_net.py:139-247  batched Lanczos with double Gram-Schmidt + masking
  model/ada_lanczos_net.py:250-286  spectral filters from (T, Q)
  model/ada_lanczos_net.py:289-368  AdaLanczosNet.forward
"""
import numpy as np
import torch

EPS = float(np.finfo(np.float32).eps)   # model/lanczos_net.py:9
BETA_LOWER_BOUND = 1.0e-4               # model/ada_lanczos_net.py:169


def make_spec(short_dist, long_dist, num_edgetype, num_layer, num_eig_vec,
              spectral_filter_kind='MLP', kind='LanczosNet'):
  return {
      'short': list(short_dist), 'long': list(long_dist),
      'num_edgetype': int(num_edgetype), 'num_layer': int(num_layer),
      'K': int(num_eig_vec), 'filter_kind': spectral_filter_kind, 'kind': kind,
  }


def _cast(params, dtype):
  return {k: (v.detach().to('cpu').to(dtype) if v.is_floating_point() else v.detach().cpu())
          for k, v in params.items()}


def _linear(params, prefix, x):
  return x @ params[prefix + '.weight'].t() + params[prefix + '.bias']


def _filter_mlp(params, layer, x):
  """spectral_filter[layer]: Linear-ReLU x3 then Linear (model/lanczos_net.py:47-58)."""
  p = 'spectral_filter.%d.' % layer
  h = torch.relu(_linear(params, p + '0', x))
  h = torch.relu(_linear(params, p + '2', h))
  h = torch.relu(_linear(params, p + '4', h))
  return _linear(params, p + '6', h)


# ----------------------------------------------------------------------------
# LanczosNet (Ritz pairs are inputs)
# ----------------------------------------------------------------------------
def ritz_power_table(D, long_dist):
  """[D**p for p in long_dist] (model/lanczos_net.py:146-149) -> B x K x S."""
  return torch.stack([torch.pow(D, p) for p in long_dist], dim=2)


def ritz_filter_coefficients(params, spec, D, layer):
  """Per-Ritz-value multi-scale coefficients f (B x K x S), model/lanczos_net.py:109-113
  (MLP) or :118-121 (plain powers)."""
  table = ritz_power_table(D, spec['long'])
  if spec['filter_kind'] == 'MLP':
    B, K, S = table.shape
    return _filter_mlp(params, layer, table.reshape(B * K, S)).reshape(B, K, S)
  return table


def spectral_filters_from_ritz(params, spec, D, V, layer):
  """Lf[..., s] = (V * f_s) V^T  (model/lanczos_net.py:114-123) -> B x N x N x S."""
  f = ritz_filter_coefficients(params, spec, D, layer)
  out = []
  for s in range(len(spec['long'])):
    out.append(torch.bmm(V * f[:, :, s].unsqueeze(1), V.transpose(1, 2)))
  return torch.stack(out, dim=3)


def conv_layer(params, spec, layer, state, L, Lf):
  """One graph-convolution layer body shared by the three models
  (model/lanczos_net.py:157-182, ada_lanczos_net.py:321-347).
  Message order is short scales, long scales, edge types (weight-column order)."""
  msgs = []
  if spec['short']:
    walk = state
    for step in range(1, max(spec['short']) + 1):
      walk = torch.bmm(L[:, :, :, 0], walk)
      if step in spec['short']:
        msgs.append(walk)
  if spec['long']:
    for s in range(len(spec['long'])):
      msgs.append(torch.bmm(Lf[:, :, :, s], state))
  for e in range(spec['num_edgetype'] + 1):
    msgs.append(torch.bmm(L[:, :, :, e], state))
  B, N = state.shape[0], state.shape[1]
  cat = torch.cat(msgs, dim=2).reshape(B * N, -1)
  return torch.relu(_linear(params, 'filter.%d' % layer, cat)).reshape(B, N, -1)


def readout(params, spec, state, mask):
  """Gated per-node output, masked mean over nodes (model/lanczos_net.py:185-194;
  mask=None branch exists only in ada_lanczos_net.py:356-361)."""
  B, N, H = state.shape
  flat = state.reshape(B * N, H)
  y = _linear(params, 'filter.%d' % spec['num_layer'], flat)
  gate = torch.sigmoid(_linear(params, 'att_func.0', flat))
  y = (gate * y).reshape(B, N, -1)
  if mask is None:
    return y.mean(dim=1)
  m = mask.to(torch.bool)
  rows = []
  for b in range(B):
    rows.append(y[b, m[b], :].mean(dim=0))
  return torch.stack(rows)


def lanczos_net_forward(params, spec, node_feat, L, D, V, mask, dtype=torch.float32,
                        return_states=False):
  """LanczosNet.forward / LanczosNetGeneral.forward without the loss."""
  params = _cast(params, dtype)
  L = torch.as_tensor(L).to(dtype)
  D = torch.as_tensor(D).to(dtype)
  V = torch.as_tensor(V).to(dtype)
  node_feat = torch.as_tensor(node_feat)
  if spec['kind'] == 'LanczosNetGeneral':
    state = node_feat.to(dtype)                       # lanczos_net_general.py:156
  else:
    state = params['embedding.weight'][node_feat.long()]   # lanczos_net.py:154
  states = [state]
  for layer in range(spec['num_layer']):
    Lf = spectral_filters_from_ritz(params, spec, D, V, layer) if spec['long'] else None
    state = conv_layer(params, spec, layer, state, L, Lf)
    states.append(state)
  score = readout(params, spec, state, None if mask is None else torch.as_tensor(mask))
  return (score, states) if return_states else score


def gcn_forward(params, spec, node_feat, L, mask, dtype=torch.float32, binarize=False):
  """GCN.forward without the loss (model/gcn.py:64-119): the LanczosNet layer with no
  diffusion scales -- msg = [L_e X]_e (:84-88), Linear + ReLU (:90-91), gated readout (:95-110).
  binarize=True is GCNFP.forward (model/gcnfp.py:68-125): the same on L[L != 0] = 1.0 (:83)."""
  assert not spec['short'] and not spec['long']
  params = _cast(params, dtype)
  L = torch.as_tensor(L).to(dtype)
  if binarize:
    L = (L != 0).to(dtype)
  state = params['embedding.weight'][torch.as_tensor(node_feat).long()]   # gcn.py:81
  for layer in range(spec['num_layer']):
    state = conv_layer(params, spec, layer, state, L, None)
  return readout(params, spec, state, None if mask is None else torch.as_tensor(mask))


def dcnn_forward(params, diffusion_dist, num_edgetype, num_layer, node_feat, L, mask,
                 dtype=torch.float32):
  """DCNN.forward without the loss (model/dcnn.py:64-124): per layer the walk L_0^k X for k in
  diffusion_dist (:88-92), the edge-type products (:94-96), concatenated EDGES FIRST (:98),
  Linear + ReLU (:99); the gated readout shared with the other models (:103-118)."""
  params = _cast(params, dtype)
  L = torch.as_tensor(L).to(dtype)
  state = params['embedding.weight'][torch.as_tensor(node_feat).long()]   # dcnn.py:83
  B, N = state.shape[0], state.shape[1]
  for layer in range(num_layer):
    scales, walk = [], state
    for step in range(1, max(diffusion_dist) + 1):
      walk = torch.bmm(L[:, :, :, 0], walk)
      if step in diffusion_dist:
        scales.append(walk)
    msgs = [torch.bmm(L[:, :, :, e], state) for e in range(num_edgetype + 1)]
    cat = torch.cat(msgs + scales, dim=2).reshape(B * N, -1)
    state = torch.relu(_linear(params, 'filter.%d' % layer, cat)).reshape(B, N, -1)
  spec = {'num_layer': num_layer}
  return readout(params, spec, state, None if mask is None else torch.as_tensor(mask))


def cheby_net_forward(params, polynomial_order, num_edgetype, num_layer, node_feat, L, mask,
                      dtype=torch.float32):
  """ChebyNet.forward without the loss (model/cheby_net.py:64-124): per layer the Chebyshev chain
  on channel 0 -- state_scale[-1] = X, [0] = L_0 X, [k] = 2 L_0 [k-1] - [k-2] (:88-93; index -1 is
  the LAST slot, which is how k = 1 reaches X) --, the bond-type products for e >= 1 (:95-97),
  cat(edges + state_scale) (:99), Linear + ReLU (:100); the shared gated readout (:104-119)."""
  params = _cast(params, dtype)
  L = torch.as_tensor(L).to(dtype)
  state = params['embedding.weight'][torch.as_tensor(node_feat).long()]   # cheby_net.py:83
  B, N = state.shape[0], state.shape[1]
  for layer in range(num_layer):
    scale = [None] * (polynomial_order + 1)
    scale[-1] = state
    scale[0] = torch.bmm(L[:, :, :, 0], state)
    for kk in range(1, polynomial_order):
      scale[kk] = 2.0 * torch.bmm(L[:, :, :, 0], scale[kk - 1]) - scale[kk - 2]
    msgs = [torch.bmm(L[:, :, :, e], state) for e in range(1, num_edgetype + 1)]
    cat = torch.cat(msgs + scale, dim=2).reshape(B * N, -1)
    state = torch.relu(_linear(params, 'filter.%d' % layer, cat)).reshape(B, N, -1)
  return readout(params, {'num_layer': num_layer}, state,
                 None if mask is None else torch.as_tensor(mask))


# ----------------------------------------------------------------------------
# AdaLanczosNet pieces
# ----------------------------------------------------------------------------
def adjacency_from_laplacian(L0):
  """Binary mask of non-zeros of the simple-graph operator (ada_lanczos_net.py:310-311)."""
  return (L0 != 0).to(L0.dtype)


def gaussian_kernel_laplacian(x, adj):
  """model/ada_lanczos_net.py:101-137.  x: B x N x D, adj: B x N x N (binary, self loops).

  NB the reference's flat pair index p = r*N + c uses node c for the first gather and
  node r for the second (meshgrid default 'xy'); dist2 is symmetric so only the
  reshape order matters: entry (r, c) of the reshaped matrix is ||x_c - x_r||^2.
  sigma2 is the mean over ALL N^2 pairs including padded nodes (:126)."""
  diff = x.unsqueeze(1) - x.unsqueeze(2)            # [b, r, c, :] = x_c - x_r
  dist2 = (diff * diff).sum(dim=3)
  sigma2 = dist2.reshape(dist2.shape[0], -1).mean(dim=1).reshape(-1, 1, 1)
  A = torch.exp(-dist2 / sigma2) * adj
  row_sum = A.sum(dim=2, keepdim=True)
  pad = (row_sum == 0).to(A.dtype)
  d = 1.0 / (row_sum + pad).pow(0.5)
  return d * A * d.transpose(1, 2)


def lanczos_tridiagonalise(A, mask, q1, K, reorth=True):
  """model/ada_lanczos_net.py:139-247 with the start vector q1 (B x N, *before* masking
  and normalisation, i.e. the raw randn draw of :161) passed in.

  Returns dict with T (B x K x K), Q (B x N x K), alpha (B x K), beta (B x K-1... padded
  to K-1), idx (B,) int64 -- including every masking quirk:
    * valid_i = prod_{j<=i} [beta_j >= 1e-4]                          (:193-199)
    * idx = min(sum valid, sum mask); valid[idx:] = 0                  (:207-215)
    * alpha *= valid ; beta *= valid[:-1]                              (:218-219)
    * Q columns *= valid and Q rows >= idx zeroed                      (:229-237)
    * zero-pad to K when N < K                                         (:240-245)
  """
  B, N = A.shape[0], A.shape[1]
  iters = min(N, K)
  dtype = A.dtype
  eps = torch.tensor(EPS, dtype=dtype)
  q = q1.reshape(B, N, 1).to(dtype)
  if mask is not None:
    fmask = mask.reshape(B, N, 1).to(dtype)
    q = q * fmask
  q = q / torch.norm(q, 2, dim=1, keepdim=True)
  basis = [q]
  prev = torch.zeros_like(q)
  beta_prev = torch.zeros(B, 1, 1, dtype=dtype)
  alphas, betas, valids = [], [], []
  for i in range(iters):
    cur = basis[i]
    z = torch.bmm(A, cur)
    a = (cur * z).sum(dim=1, keepdim=True)
    z = z - a * cur - beta_prev * prev
    if reorth and i > 0:
      for _ in range(2):
        for j in range(i):
          qj = basis[j]
          z = z - (z * qj).sum(dim=1, keepdim=True) / ((qj * qj).sum(dim=1, keepdim=True) + eps) * qj
    b = torch.norm(z, p=2, dim=1, keepdim=True)
    ok = (b >= BETA_LOWER_BOUND).to(dtype)
    if valids:
      ok = valids[-1] * ok
    valids.append(ok)
    nxt = (z * ok) / (b + eps)
    alphas.append(a)
    betas.append(b)
    basis.append(nxt)
    prev, beta_prev = cur, b
  alpha = torch.cat(alphas, dim=1).squeeze(2)           # B x iters
  beta = torch.cat(betas[:-1], dim=1).squeeze(2) if iters > 1 else torch.zeros(B, 0, dtype=dtype)
  valid = torch.cat(valids, dim=1).squeeze(2)           # B x iters
  idx = valid.sum(dim=1).long()
  if mask is not None:
    idx = torch.minimum(idx, fmask.sum(dim=1).reshape(B).long())
  col = torch.arange(iters).unsqueeze(0)
  valid = valid * (col < idx.unsqueeze(1)).to(dtype)
  alpha = alpha * valid
  beta = beta * valid[:, :-1]
  T = torch.diag_embed(alpha) + torch.diag_embed(beta, offset=1) + torch.diag_embed(beta, offset=-1)
  Q = torch.cat(basis[:iters], dim=2)                   # B x N x iters
  row = torch.arange(N).reshape(1, N, 1)
  keep = valid.unsqueeze(1) * (row < idx.reshape(B, 1, 1)).to(dtype)
  Q = Q * keep
  if iters < K:
    T = torch.nn.functional.pad(T, (0, K - iters, 0, K - iters))
    Q = torch.nn.functional.pad(Q, (0, K - iters))
    alpha = torch.nn.functional.pad(alpha, (0, K - iters))
    beta = torch.nn.functional.pad(beta, (0, K - iters))
  return {'T': T, 'Q': Q, 'alpha': alpha, 'beta': beta, 'idx': idx}


def tridiag_ritz(alpha, beta, Q=None):
  """Ritz pairs of the Lanczos tridiagonal.  The reference has no executable counterpart
  (SURVEY fact 2); the oracle is LAPACK (numpy eigh, fp64) on the reference's T, ordered
  by the reference's own Ritz ordering rule -|lambda| stable (utils/data_helper.py:217-223).

  alpha: B x K, beta: B x (K-1).  Returns (theta B x K, S B x K x K[, V = Q S])."""
  a = np.asarray(alpha, dtype=np.float64)
  b = np.asarray(beta, dtype=np.float64)
  Bn, K = a.shape
  theta = np.zeros((Bn, K))
  S = np.zeros((Bn, K, K))
  for g in range(Bn):
    T = np.diag(a[g]) + np.diag(b[g, :K - 1], 1) + np.diag(b[g, :K - 1], -1)
    w, v = np.linalg.eigh(T)
    order = np.argsort(-np.abs(w), kind='mergesort')
    theta[g], S[g] = w[order], v[:, order]
  if Q is None:
    return theta, S
  return theta, S, np.einsum('bnk,bkj->bnj', np.asarray(Q, dtype=np.float64), S)


def tridiag_power_stack(T, long_dist):
  """T^p for p in long_dist by repeated bmm (ada_lanczos_net.py:262-270) -> list of BxKxK."""
  out = []
  cur = T
  for p in range(1, max(long_dist) + 1):
    if p in long_dist:
      out.append(cur)
    cur = torch.bmm(cur, T)
  return out


def spectral_filters_from_tridiag(params, spec, T, Q, layer):
  """model/ada_lanczos_net.py:250-286 -> B x N x N x S."""
  powers = tridiag_power_stack(T, spec['long'])
  B, K = T.shape[0], T.shape[1]
  S = len(spec['long'])
  out = []
  if spec['filter_kind'] == 'MLP':
    flat = torch.cat(powers, dim=2).reshape(B, -1)       # index r*S*K + s*K + c   (:274)
    G = _filter_mlp(params, layer, flat).reshape(B, K, K, S)   # index r*K*S + c*S + s (:275)
    G = (G + G.transpose(1, 2)) * 0.5
    for s in range(S):
      out.append(Q.bmm(G[:, :, :, s]).bmm(Q.transpose(1, 2)))
  else:
    for s in range(S):
      out.append(Q.bmm(powers[s]).bmm(Q.transpose(1, 2)))
  return torch.stack(out, dim=3)


def ada_lanczos_net_forward(params, spec, node_feat, L, mask, q1, dtype=torch.float32,
                            return_aux=False):
  """AdaLanczosNet.forward without the loss.  q1 is the raw B x N (x1) randn draw the
  reference makes on the CPU generator (ada_lanczos_net.py:161)."""
  params = _cast(params, dtype)
  L = torch.as_tensor(L).to(dtype)
  node_feat = torch.as_tensor(node_feat)
  state = params['embedding.weight'][node_feat.long()]
  aux = {}
  T = Q = None
  if spec['long']:
    adj = adjacency_from_laplacian(L[:, :, :, 0])
    Le = gaussian_kernel_laplacian(state, adj)
    lz = lanczos_tridiagonalise(Le, None if mask is None else torch.as_tensor(mask),
                                torch.as_tensor(q1), spec['K'])
    T, Q = lz['T'], lz['Q']
    aux.update(Le=Le, **lz)
  for layer in range(spec['num_layer']):
    Lf = spectral_filters_from_tridiag(params, spec, T, Q, layer) if spec['long'] else None
    state = conv_layer(params, spec, layer, state, L, Lf)
  score = readout(params, spec, state, None if mask is None else torch.as_tensor(mask))
  return (score, aux) if return_aux else score
