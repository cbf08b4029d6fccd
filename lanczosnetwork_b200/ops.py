"""torch-tensor front end of the C ABI (include/lanczosnet_b200.h).

PyTorch is plumbing here: device memory, streams, dtype/contiguity checks.  All arithmetic
runs in the hand-written CUDA kernels of liblanczosnet_b200.so.  Every function requires
CUDA tensors and raises otherwise -- there is no CPU path.
"""
import ctypes

import torch

from . import _lib
from ._lib import GemmDesc, SpectralStack

__all__ = [
    'bgemm', 'split_tf32', 'linear_tf32x3', 'linear_tf32x3_grouped', 'graph_prepare', 'spectral_conv_fused',
    'graph_prepare_sparse', 'graph_prepare_sparse_packed', 'fused_conv_supported', 'spectral_stack_forward', 'ritz_rowmap', 'ritz_filter_mlp', 'embedding_rows', 'ritz_power_table', 'readout',
    'operator_chain', 'operator_chain_supported', 'graph_messages', 'graph_messages_supported', 'gaussian_laplacian', 'lanczos_tridiag', 'lanczos_ritz', 'tridiag_ritz', 'tridiag_powers',
    'symmetrize_filters', 'segment_sum_forward', 'segment_sum_backward', 'launch_count',
]


def _need_cuda(*tensors):
  for t in tensors:
    if t is None:
      continue
    if not t.is_cuda:
      raise RuntimeError('lanczosnetwork_b200 ops run on CUDA (sm_100a) only; got a %s tensor. '
                         'There is no CPU fallback.' % t.device)


def _f32c(t):
  if t.dtype != torch.float32:
    t = t.float()
  return t.contiguous()


def _ptr(t):
  return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(t):
  return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ints(vals):
  arr = (ctypes.c_int * len(vals))(*[int(v) for v in vals])
  return arr


def launch_count():
  return _lib.launch_count()


# --------------------------------------------------------------------------------------------
def bgemm(A, a_str, Bm, b_str, C, c_str, batch, nz, M, N, K, kscale=None, s_str=(0, 0, 0),
          bias=None, bias_sz=0, relu=False, a_off=0, b_off=0, c_off=0, alpha=1.0, addend=None,
          add_str=(0, 0, 0, 0), add_off=0, beta=0.0):
  """C[b,z] = act(alpha * (A[b,z] * kscale[b,z]) @ B[b,z] + beta * addend[b,z] + bias); strides in
  elements, *_off element offsets into the (fp32, CUDA) storage of A / B / C / addend."""
  _need_cuda(A, Bm, C, kscale, bias)
  lib = _lib.load()
  d = GemmDesc()
  d.A = A.data_ptr() + 4 * a_off
  d.a_sb, d.a_sz, d.a_sm, d.a_sk = [int(v) for v in a_str]
  d.B = Bm.data_ptr() + 4 * b_off
  d.b_sb, d.b_sz, d.b_sk, d.b_sn = [int(v) for v in b_str]
  d.C = C.data_ptr() + 4 * c_off
  d.c_sb, d.c_sz, d.c_sm, d.c_sn = [int(v) for v in c_str]
  d.kscale = kscale.data_ptr() if kscale is not None else None
  d.s_sb, d.s_sz, d.s_sk = [int(v) for v in s_str]
  d.bias = bias.data_ptr() if bias is not None else None
  d.bias_sz = int(bias_sz)
  d.batch, d.nz, d.M, d.N, d.K, d.relu = int(batch), int(nz), int(M), int(N), int(K), int(bool(relu))
  d.alpha, d.beta = float(alpha), float(beta)
  d.addend = addend.data_ptr() + 4 * add_off if addend is not None else None
  d.d_sb, d.d_sz, d.d_sm, d.d_sn = [int(v) for v in add_str]
  with torch.cuda.device(C.device):
    _lib.check(lib.lnb_batched_gemm(_stream(C), ctypes.byref(d)), 'lnb_batched_gemm')
  return C


def split_tf32(x):
  """(hi, lo) tf32 split of an fp32 tensor: hi = rna_tf32(x), lo = rna_tf32(x - hi)."""
  _need_cuda(x)
  x = _f32c(x)
  hi = torch.empty_like(x)
  lo = torch.empty_like(x)
  with torch.cuda.device(x.device):
    _lib.check(_lib.load().lnb_split_tf32(_stream(x), _ptr(x), x.numel(), _ptr(hi), _ptr(lo)),
               'lnb_split_tf32')
  return hi, lo


def linear_tf32x3(x, w_hi, w_lo, bias=None, relu=False, out=None):
  """act(x @ W^T + bias) on tcgen05 tensor cores (3xTF32).  x [M,K], w_hi/w_lo [N,K]."""
  _need_cuda(x, w_hi, w_lo, bias)
  x = _f32c(x)
  M, K = x.shape
  N = w_hi.shape[0]
  if out is None:
    out = torch.empty((M, N), device=x.device, dtype=torch.float32)
  tiles = ((M + 127) // 128) * ((N + 127) // 128)
  nkb = (K + 31) // 32
  splits = 1
  if tiles * 2 <= _sm_count(x.device) and nkb >= 32:
    # few output tiles, deep K (the Ada filter MLP): split K so every SM streams weights
    splits = min(_sm_count(x.device) // tiles, 8, nkb // 16)
    while splits > 1 and ((nkb + splits - 1) // splits) * (splits - 1) >= nkb:
      splits -= 1
  with torch.cuda.device(x.device):
    if splits > 1:
      ws, counters = _splitk_workspace(x.device, tiles * splits * 128 * 128, tiles)
      _lib.check(_lib.load().lnb_linear_tf32x3_splitk(
          _stream(x), _ptr(x), _ptr(w_hi), _ptr(w_lo), _ptr(bias), M, N, K, int(bool(relu)),
          _ptr(out), splits, _ptr(ws), _ptr(counters)), 'lnb_linear_tf32x3_splitk')
    else:
      _lib.check(_lib.load().lnb_linear_tf32x3(_stream(x), _ptr(x), _ptr(w_hi), _ptr(w_lo),
                                               _ptr(bias), M, N, K, int(bool(relu)), _ptr(out)),
                 'lnb_linear_tf32x3')
  return out


_SPLITK_WS = {}
_SM_COUNT = {}


def _sm_count(device):
  idx = device.index if device.index is not None else torch.cuda.current_device()
  if idx not in _SM_COUNT:
    _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
  return _SM_COUNT[idx]


def _splitk_workspace(device, nfloats, ntiles):
  """Per-device split-K scratch: partial tiles + per-tile arrival counters (zero between launches;
  launches on one stream are ordered, so one buffer per device and stream is enough)."""
  key = (device.index, torch.cuda.current_stream(device).cuda_stream)
  ws, counters = _SPLITK_WS.get(key, (None, None))
  if ws is None or ws.numel() < nfloats or counters.numel() < ntiles:
    ws = torch.empty((max(nfloats, ws.numel() if ws is not None else 0),), device=device,
                     dtype=torch.float32)
    counters = torch.zeros((max(ntiles, 256),), device=device, dtype=torch.int32)
    _SPLITK_WS[key] = (ws, counters)
  return ws, counters


def linear_tf32x3_grouped(x, w_hi, w_lo, bias, groups, relu=False):
  """Block-diagonal layer: out[:, g*N:(g+1)*N] = act(x[:, g*K:(g+1)*K] @ W_g^T + b_g) with
  x [M, groups*K], w_hi/w_lo [groups*N, K] (stacked), bias [groups*N]."""
  _need_cuda(x, w_hi, w_lo, bias)
  x = _f32c(x)
  M = x.shape[0]
  K = w_hi.shape[1]
  N = w_hi.shape[0] // groups
  assert x.shape[1] == groups * K and w_hi.shape[0] == groups * N
  out = torch.empty((M, groups * N), device=x.device, dtype=torch.float32)
  with torch.cuda.device(x.device):
    _lib.check(_lib.load().lnb_linear_tf32x3_grouped(_stream(x), _ptr(x), _ptr(w_hi), _ptr(w_lo),
                                                     _ptr(bias), M, groups, N, K,
                                                     int(bool(relu)), _ptr(out)),
               'lnb_linear_tf32x3_grouped')
  return out


class GraphPrep(tuple):
  """(ell_val, ell_idx, ell_max, gext, tiles) plus the compact Ritz row list of the same pass
  (attributes rowmap [B*K] int32, nrows [1] int32; see ritz_rowmap)."""
  rowmap = None
  nrows = None


def graph_prepare(L, Q, binarize=False):
  """Per-forward compression of the dense operators L [B,N,N,E1] (ELL rows), the real extents
  of every graph, the packed-tile assignment for the fused convolution kernel and the compact
  list of non-zero Ritz rows.  Returns GraphPrep(ell_val, ell_idx, ell_max, gext, tiles)."""
  _need_cuda(L, Q)
  L, Q = _f32c(L), _f32c(Q)
  B, N, _, E1 = L.shape
  K = Q.shape[2]
  dev = L.device
  ell_val = torch.empty((B, E1, N, N), device=dev, dtype=torch.float32)
  ell_idx = torch.empty((B, E1, N, N), device=dev, dtype=torch.uint8)
  ell_max = torch.empty((B, E1), device=dev, dtype=torch.int32)
  gext = torch.empty((B, 2), device=dev, dtype=torch.int32)
  tiles = torch.empty((4 * B + 2,), device=dev, dtype=torch.int32)   # tile table + scratch
  rowmap = torch.empty((B * K,), device=dev, dtype=torch.int32)
  nrows = torch.empty((1,), device=dev, dtype=torch.int32)
  with torch.cuda.device(dev):
    _lib.check(_lib.load().lnb_graph_prepare(_stream(L), _ptr(L), _ptr(Q), B, N, E1, K,
                                             _ptr(ell_val), _ptr(ell_idx), _ptr(ell_max),
                                             _ptr(gext), _ptr(tiles), _ptr(rowmap), _ptr(nrows),
                                             1 if binarize else 0),
               'lnb_graph_prepare')
  prep = GraphPrep((ell_val, ell_idx, ell_max, gext, tiles))
  prep.rowmap, prep.nrows = rowmap, nrows
  return prep


_INV_SQRT_DEG = {}


def _inv_sqrt_deg_table(device):
  """deg^-1/2 in fp64 for deg = 0..255 exactly as the reference's host code computes it
  (np.power(deg, -0.5) with inf -> 0, utils/data_helper.py:104-107), cached per device."""
  key = device.index if device.index is not None else torch.cuda.current_device()
  if key not in _INV_SQRT_DEG:
    import numpy as np
    deg = np.arange(256, dtype=np.float64)
    with np.errstate(divide='ignore'):
      t = np.power(deg, -0.5)
    t[np.isinf(t)] = 0.0
    _INV_SQRT_DEG[key] = torch.from_numpy(t).to(device)
  return _INV_SQRT_DEG[key]


def graph_prepare_sparse(sizes, node_ptr, node_feat, edge_ptr, edges, V_rows, N, E1, binarize=False,
                         want_dense=False):
  """GPU-side batch construction from sparse records (see lnb_graph_prepare_sparse).
  sizes [B] int32, node_ptr [B+1] int32, node_feat [>= node_ptr[B]] int32, edge_ptr [B+1] int32,
  edges [>= edge_ptr[B], 4] uint8, V_rows [>= node_ptr[B], K] fp32 -- all CUDA.
  Returns (GraphPrep, node_ids [B,N] int64, mask [B,N] uint8, V [B,N,K], L [B,N,N,E1] or None)."""
  _need_cuda(sizes, node_ptr, node_feat, edge_ptr, edges, V_rows)
  dev = sizes.device
  B = sizes.shape[0]
  K = V_rows.shape[1]
  assert sizes.dtype == torch.int32 and node_ptr.dtype == torch.int32 and node_feat.dtype == torch.int32
  assert edge_ptr.dtype == torch.int32 and edges.dtype == torch.uint8 and edges.shape[1] == 4
  assert V_rows.dtype == torch.float32 and V_rows.is_contiguous() and edges.is_contiguous()
  ell_val = torch.empty((B, E1, N, N), device=dev, dtype=torch.float32)
  ell_idx = torch.empty((B, E1, N, N), device=dev, dtype=torch.uint8)
  ell_max = torch.empty((B, E1), device=dev, dtype=torch.int32)
  gext = torch.empty((B, 2), device=dev, dtype=torch.int32)
  tiles = torch.empty((4 * B + 2,), device=dev, dtype=torch.int32)
  rowmap = torch.empty((B * K,), device=dev, dtype=torch.int32)
  nrows = torch.empty((1,), device=dev, dtype=torch.int32)
  node_ids = torch.empty((B, N), device=dev, dtype=torch.int64)
  mask = torch.empty((B, N), device=dev, dtype=torch.uint8)
  V = torch.empty((B, N, K), device=dev, dtype=torch.float32)
  L = torch.empty((B, N, N, E1), device=dev, dtype=torch.float32) if want_dense else None
  with torch.cuda.device(dev):
    _lib.check(_lib.load().lnb_graph_prepare_sparse(
        _stream(sizes), _ptr(sizes), _ptr(node_ptr), _ptr(node_feat), _ptr(edge_ptr), _ptr(edges),
        _ptr(V_rows), _ptr(_inv_sqrt_deg_table(dev)), B, int(N), int(E1), int(K),
        1 if binarize else 0, _ptr(ell_val), _ptr(ell_idx), _ptr(ell_max), _ptr(gext), _ptr(tiles),
        _ptr(rowmap), _ptr(nrows), _ptr(node_ids), _ptr(mask), _ptr(V), _ptr(L)),
               'lnb_graph_prepare_sparse')
  prep = GraphPrep((ell_val, ell_idx, ell_max, gext, tiles))
  prep.rowmap, prep.nrows = rowmap, nrows
  return prep, node_ids, mask, V, L


def graph_prepare_sparse_packed(blob, B, N, E1, K, binarize=False, want_dense=False, host_tiles=True):
  """graph_prepare_sparse on a packed batch (data.pack_sparse: one contiguous uint8 buffer, see
  lnb_graph_prepare_sparse_packed).  host_tiles: the batch carries the tile table and the Ritz-row
  prefix sums (data.pack_sparse always writes them), so no tile-assignment kernel is launched and the
  stack kernel reads the table in place.  Returns (GraphPrep, node_ids, mask, V, L or None)."""
  _need_cuda(blob)
  assert blob.dtype == torch.uint8 and blob.is_contiguous()
  dev = blob.device
  ell_val = torch.empty((B, E1, N, N), device=dev, dtype=torch.float32)
  ell_idx = torch.empty((B, E1, N, N), device=dev, dtype=torch.uint8)
  ell_max = torch.empty((B, E1), device=dev, dtype=torch.int32)
  gext = torch.empty((B, 2), device=dev, dtype=torch.int32)
  if host_tiles:
    from .data import packed_offsets
    off_tiles = packed_offsets(B, K)[5]
    tiles = blob[off_tiles:off_tiles + 4 * (B + 2)].view(torch.int32)
  else:
    tiles = torch.empty((4 * B + 2,), device=dev, dtype=torch.int32)
  rowmap = torch.empty((B * K,), device=dev, dtype=torch.int32)
  nrows = torch.empty((1,), device=dev, dtype=torch.int32)
  node_ids = torch.empty((B, N), device=dev, dtype=torch.int64)
  mask = torch.empty((B, N), device=dev, dtype=torch.uint8)
  V = torch.empty((B, N, K), device=dev, dtype=torch.float32)
  L = torch.empty((B, N, N, E1), device=dev, dtype=torch.float32) if want_dense else None
  with torch.cuda.device(dev):
    _lib.check(_lib.load().lnb_graph_prepare_sparse_packed(
        _stream(blob), _ptr(blob), _ptr(_inv_sqrt_deg_table(dev)), int(B), int(N), int(E1), int(K),
        (1 if binarize else 0) | (2 if host_tiles else 0), _ptr(ell_val), _ptr(ell_idx), _ptr(ell_max), _ptr(gext), _ptr(tiles),
        _ptr(rowmap), _ptr(nrows), _ptr(node_ids), _ptr(mask), _ptr(V), _ptr(L)),
               'lnb_graph_prepare_sparse_packed')
  prep = GraphPrep((ell_val, ell_idx, ell_max, gext, tiles))
  prep.rowmap, prep.nrows = rowmap, nrows
  return prep, node_ids, mask, V, L


def fused_conv_supported(N, Din, K, H, n_short, dense_filter, S=8, E1=7):
  """Shapes the fused tcgen05 convolution kernel handles (others use the unfused ops);
  mirrors the checks of lnb_spectral_conv_fused."""
  if (n_short or dense_filter or N > 128 or Din % 32 or K > 32 or K % 4 or H % 4 or H > 128 or
      E1 > 16):
    return False
  smem = 2 * 32768 + 256 + 1024 + 2 * 128 * (max(Din, H) + 4) * 4 + 128 * K * 4 + 4096
  return smem <= 227 * 1024


def spectral_conv_fused(X, Q, coeff, prep, w_hi, w_lo, bias, relu=True, write_pad=True):
  """One fused spectral convolution layer: X [B,N,Din], Q [B,N,K], coeff [B,K,S] diagonal
  filter coefficients (None when there are no long scales), prep = graph_prepare(L, Q),
  W [H, (S+E1)*Din] split -> [B,N,H].  write_pad=False leaves the rows of padded nodes
  unwritten (fine between layers: nothing reads them)."""
  _need_cuda(X, Q, coeff, w_hi, w_lo, bias)
  X, Q = _f32c(X), _f32c(Q)
  ell_val, ell_idx, ell_max, gext, tiles = prep
  B, N, Din = X.shape
  K = Q.shape[2]
  S = 0
  if coeff is not None:
    coeff = _f32c(coeff)
    S = coeff.shape[2]
  E1 = ell_val.shape[1]
  H = w_hi.shape[0]
  assert w_hi.shape[1] == (S + E1) * Din, (w_hi.shape, S, E1, Din)
  out = torch.empty((B, N, H), device=X.device, dtype=torch.float32)
  with torch.cuda.device(X.device):
    _lib.check(_lib.load().lnb_spectral_conv_fused(
        _stream(X), _ptr(X), _ptr(Q), _ptr(coeff), _ptr(ell_val), _ptr(ell_idx), _ptr(ell_max),
        _ptr(gext), _ptr(tiles), _ptr(w_hi), _ptr(w_lo), _ptr(bias), B, N, Din, E1, K, S, H,
        int(bool(relu)), int(bool(write_pad)), _ptr(out)), 'lnb_spectral_conv_fused')
  return out


def spectral_stack_forward(prep, Q, w_hi, w_lo, bias, dins, H, S, coeff=None, coeff_stride=0,
                           X=None, node_ids=None, emb=None, want_state=False, write_pad=True,
                           readout=None, mask=None, relu=True):
  """All convolution layers (+ optional embedding gather and readout) in one persistent kernel.

  prep = graph_prepare(L, Q); w_hi/w_lo [len(dins)*H, Kw] stacked split weights, bias
  [len(dins)*H]; dins = input width per layer; coeff = tensor whose layer l block starts at
  element l*coeff_stride (None when S == 0); X [B,N,dins[0]] or node_ids [B,N] + emb;
  readout = (W_out [P,H], b_out [P], w_att [H], b_att [1]) -> score [B,P].
  Returns (state or None, score or None)."""
  ell_val, ell_idx, ell_max, gext, tiles = prep
  _need_cuda(Q, w_hi, w_lo, bias, coeff, X, node_ids, emb, mask)
  Q = _f32c(Q)
  B, N, K = Q.shape
  E1 = ell_val.shape[1]
  dev = Q.device
  d = SpectralStack()
  if X is not None:
    X = _f32c(X)
    d.X = X.data_ptr()
  else:
    node_ids = node_ids.contiguous().long()
    emb = _f32c(emb)
    d.node_ids, d.emb_table, d.emb_rows = node_ids.data_ptr(), emb.data_ptr(), emb.shape[0]
  d.Q = Q.data_ptr()
  if coeff is not None:
    d.coeff, d.coeff_layer_stride = coeff.data_ptr(), int(coeff_stride)
  d.ell_val, d.ell_idx, d.ell_max = ell_val.data_ptr(), ell_idx.data_ptr(), ell_max.data_ptr()
  d.gext, d.tiles = gext.data_ptr(), tiles.data_ptr()
  d.W_hi, d.W_lo = w_hi.data_ptr(), w_lo.data_ptr()
  d.bias = bias.data_ptr() if bias is not None else None
  state = score = None
  if want_state or readout is None:
    state = torch.empty((B, N, H), device=dev, dtype=torch.float32)
    d.out_state = state.data_ptr()
  keep = []
  if readout is not None:
    W_out, b_out, w_att, b_att = [_f32c(t) for t in readout]
    keep += [W_out, b_out, w_att, b_att]
    score = torch.empty((B, W_out.shape[0]), device=dev, dtype=torch.float32)
    d.W_out, d.b_out, d.w_att, d.b_att = (W_out.data_ptr(), b_out.data_ptr(), w_att.data_ptr(),
                                          b_att.data_ptr())
    d.score, d.P = score.data_ptr(), W_out.shape[0]
    if mask is not None:
      if mask.dtype != torch.uint8:            # any non-zero byte counts as 'real node' in the kernel
        mask = (mask != 0).to(torch.uint8)
      mask = mask.contiguous()
      d.mask = mask.data_ptr()
  for i, v in enumerate(dins):
    d.Din[i] = int(v)
  d.num_layers, d.Kw, d.write_pad = len(dins), int(w_hi.shape[1]), int(bool(write_pad))
  d.B, d.N, d.E1, d.K, d.S, d.H, d.relu = B, N, E1, K, int(S), int(H), int(bool(relu))
  with torch.cuda.device(dev):
    _lib.check(_lib.load().lnb_spectral_stack_forward(_stream(Q), ctypes.byref(d)),
               'lnb_spectral_stack_forward')
  return state, score


def ritz_rowmap(gext, K):
  """Compact list of the (graph, k) rows with k < k_eff(graph): (rowmap [B*K] int32, nrows [1])."""
  _need_cuda(gext)
  B = gext.shape[0]
  rowmap = torch.empty((B * K,), device=gext.device, dtype=torch.int32)
  nrows = torch.empty((1,), device=gext.device, dtype=torch.int32)
  with torch.cuda.device(gext.device):
    _lib.check(_lib.load().lnb_ritz_rowmap(_stream(gext), _ptr(gext), B, int(K), _ptr(rowmap),
                                           _ptr(nrows)), 'lnb_ritz_rowmap')
  return rowmap, nrows


def ritz_filter_mlp(table, w_hi, w_lo, bias_all, num_layers, rowmap=None, nrows=None):
  """coeff[l, r, :] = MLP_l(table[r, :]) for all layers in one persistent kernel.
  table [R, S]; w_hi/w_lo [L*(3*Hd+S), Hd] stacked split weights; returns coeff [L, R, S]
  (rows not listed in rowmap are left unwritten)."""
  _need_cuda(table, w_hi, w_lo, bias_all, rowmap, nrows)
  table = _f32c(table)
  R, S = table.shape
  Hd = w_hi.shape[1]
  coeff = torch.empty((num_layers, R, S), device=table.device, dtype=torch.float32)
  with torch.cuda.device(table.device):
    _lib.check(_lib.load().lnb_ritz_filter_mlp(_stream(table), _ptr(table), _ptr(rowmap),
                                               _ptr(nrows), _ptr(w_hi), _ptr(w_lo),
                                               _ptr(bias_all), R, int(num_layers), S, Hd,
                                               _ptr(coeff)), 'lnb_ritz_filter_mlp')
  return coeff


def embedding_rows(idx, table):
  _need_cuda(idx, table)
  idx = idx.contiguous().long()
  table = _f32c(table)
  rows = idx.numel()
  out = torch.empty(tuple(idx.shape) + (table.shape[1],), device=table.device, dtype=torch.float32)
  with torch.cuda.device(table.device):
    _lib.check(_lib.load().lnb_embedding_rows(_stream(table), _ptr(idx), _ptr(table), rows,
                                              table.shape[0], table.shape[1], _ptr(out)),
               'lnb_embedding_rows')
  return out


def ritz_power_table(D, powers):
  """table[..., s] = D ** powers[s]  (model/lanczos_net.py:146-149)."""
  _need_cuda(D)
  D = _f32c(D)
  S = len(powers)
  out = torch.empty(tuple(D.shape) + (S,), device=D.device, dtype=torch.float32)
  with torch.cuda.device(D.device):
    _lib.check(_lib.load().lnb_ritz_power_table(_stream(D), _ptr(D), D.numel(), _ints(powers), S,
                                                _ptr(out)), 'lnb_ritz_power_table')
  return out


def readout(state, W_out, b_out, w_att, b_att, mask=None):
  _need_cuda(state, W_out, b_out, w_att, b_att, mask)
  state = _f32c(state)
  B, N, H = state.shape
  P = W_out.shape[0]
  if mask is not None:
    mask = (mask != 0).to(torch.uint8).contiguous()
  out = torch.empty((B, P), device=state.device, dtype=torch.float32)
  with torch.cuda.device(state.device):
    _lib.check(_lib.load().lnb_readout(_stream(state), _ptr(state), _ptr(_f32c(W_out)),
                                       _ptr(_f32c(b_out)), _ptr(_f32c(w_att)), _ptr(_f32c(b_att)),
                                       _ptr(mask), B, N, H, P, _ptr(out)), 'lnb_readout')
  return out


def operator_chain_supported(N, steps):
  return N <= 32 and steps <= 64


def operator_chain(L, X, steps, block_of_step, out, out_col0, chebyshev=False):
  """Power / Chebyshev chain of channel 0 of L [B,N,N,E1] applied to X [B,N,D]; result i goes to
  column block out_col0 + block_of_step[i] of out [B,N,C*D] (block < 0: not stored)."""
  _need_cuda(L, X, out)
  L, X = _f32c(L), _f32c(X)
  B, N, D = X.shape
  E1 = L.shape[3]
  assert out.dtype == torch.float32 and out.is_contiguous() and out.shape[:2] == (B, N)
  sel = (ctypes.c_int * steps)(*[int(v) for v in block_of_step])
  with torch.cuda.device(X.device):
    _lib.check(_lib.load().lnb_operator_chain(_stream(X), _ptr(L), _ptr(X), B, N, E1, D, int(steps),
                                              1 if chebyshev else 0, sel, _ptr(out),
                                              out.stride(0), out.stride(1), int(out_col0)),
               'lnb_operator_chain')
  return out


def graph_messages_supported(N, K, E1, S, max_short):
  return N <= 32 and (S == 0 or K <= 32) and E1 <= 16 and S <= 8 and max_short <= 64


def graph_messages(L, X, Q, filt, dense_filter, short_dist, out):
  """The whole message matrix [short walk | long scales | edge types] of a general-shape layer in one
  launch (see lnb_graph_messages).  filt: [B,S,K,K] dense blocks (dense_filter) or [B,K,S] diagonal
  coefficients, None when there are no long scales; out [B,N,>=C*D] contiguous."""
  _need_cuda(L, X, Q, filt, out)
  L, X = _f32c(L), _f32c(X)
  B, N, D = X.shape
  E1 = L.shape[3]
  S = K = 0
  if filt is not None:
    filt, Q = _f32c(filt), _f32c(Q)
    K = Q.shape[2]
    S = filt.shape[1] if dense_filter else filt.shape[2]
  steps = sorted(short_dist)
  max_short = max(steps) if steps else 0
  sel = [steps.index(s) if s in steps else -1 for s in range(1, max_short + 1)]
  arr = (ctypes.c_int * max(1, max_short))(*([int(v) for v in sel] or [0]))
  assert out.dtype == torch.float32 and out.is_contiguous()
  with torch.cuda.device(X.device):
    _lib.check(_lib.load().lnb_graph_messages(
        _stream(X), _ptr(L), _ptr(X), _ptr(Q), _ptr(filt), B, N, E1, D, K, S,
        1 if dense_filter else 0, max_short, arr, len(steps), _ptr(out), out.stride(0), out.stride(1)),
               'lnb_graph_messages')
  return out


def gaussian_laplacian(x, L):
  _need_cuda(x, L)
  x, L = _f32c(x), _f32c(L)
  B, N, Dx = x.shape
  E1 = L.shape[3]
  out = torch.empty((B, N, N), device=x.device, dtype=torch.float32)
  with torch.cuda.device(x.device):
    _lib.check(_lib.load().lnb_gaussian_laplacian(_stream(x), _ptr(x), _ptr(L), B, N, Dx, E1,
                                                  _ptr(out)), 'lnb_gaussian_laplacian')
  return out


def lanczos_tridiag(A, mask, q1, K):
  """Returns dict(T [B,K,K], Q [B,N,K], alpha [B,K], beta [B,K], idx [B] int32)."""
  _need_cuda(A, mask, q1)
  A = _f32c(A)
  B, N = A.shape[0], A.shape[1]
  q1 = _f32c(q1).reshape(B, N)
  if mask is not None:
    mask = (mask != 0).to(torch.uint8).contiguous()
  dev = A.device
  T = torch.empty((B, K, K), device=dev, dtype=torch.float32)
  Q = torch.empty((B, N, K), device=dev, dtype=torch.float32)
  alpha = torch.empty((B, K), device=dev, dtype=torch.float32)
  beta = torch.empty((B, K), device=dev, dtype=torch.float32)
  idx = torch.empty((B,), device=dev, dtype=torch.int32)
  with torch.cuda.device(dev):
    _lib.check(_lib.load().lnb_lanczos_tridiag(_stream(A), _ptr(A), _ptr(mask), _ptr(q1), B, N, K,
                                               _ptr(T), _ptr(Q), _ptr(alpha), _ptr(beta),
                                               _ptr(idx)), 'lnb_lanczos_tridiag')
  return {'T': T, 'Q': Q, 'alpha': alpha, 'beta': beta, 'idx': idx}


def tridiag_ritz(alpha, beta, Q):
  """Ritz values (descending |theta|) and vectors V = Q S.  Returns (theta, V, status)."""
  _need_cuda(alpha, beta, Q)
  alpha, beta, Q = _f32c(alpha), _f32c(beta), _f32c(Q)
  B, N, K = Q.shape
  theta = torch.empty((B, K), device=Q.device, dtype=torch.float32)
  V = torch.empty((B, N, K), device=Q.device, dtype=torch.float32)
  status = torch.empty((B,), device=Q.device, dtype=torch.int32)
  with torch.cuda.device(Q.device):
    _lib.check(_lib.load().lnb_tridiag_ritz(_stream(Q), _ptr(alpha), _ptr(beta), _ptr(Q), B, N, K,
                                            _ptr(theta), _ptr(V), _ptr(status)), 'lnb_tridiag_ritz')
  return theta, V, status


def lanczos_ritz(A, mask, q1, K, want_ritz=True, want_T=True, want_Q=True, proper=False):
  """adjacency operator -> Ritz pairs in ONE launch (Lanczos + QL + Ritz vectors fused).
  Returns dict(alpha, beta [B,K], idx [B] int32, T [B,K,K], Q [B,N,K] when asked for, and with
  want_ritz theta [B,K] by descending |theta|, V [B,N,K] = Q S, status [B] (bit 0: QL not
  converged, bit 1: operator streamed because its non-zeros did not fit on chip)).
  proper=False reproduces the reference's masking rules of _lanczos_layer; proper=True returns the
  textbook Krylov factorisation (LNB_LANCZOS_PROPER) whose Ritz values are eigenvalues of A."""
  _need_cuda(A, mask, q1)
  A = _f32c(A)
  B, N = A.shape[0], A.shape[1]
  q1 = _f32c(q1).reshape(B, N)
  if mask is not None:
    mask = (mask != 0).to(torch.uint8).contiguous()
  dev = A.device
  out = {'alpha': torch.empty((B, K), device=dev, dtype=torch.float32),
         'beta': torch.empty((B, K), device=dev, dtype=torch.float32),
         'idx': torch.empty((B,), device=dev, dtype=torch.int32)}
  if want_T:
    out['T'] = torch.empty((B, K, K), device=dev, dtype=torch.float32)
  if want_Q:
    out['Q'] = torch.empty((B, N, K), device=dev, dtype=torch.float32)
  if want_ritz:
    out['theta'] = torch.empty((B, K), device=dev, dtype=torch.float32)
    out['V'] = torch.empty((B, N, K), device=dev, dtype=torch.float32)
    out['status'] = torch.empty((B,), device=dev, dtype=torch.int32)
  with torch.cuda.device(dev):
    _lib.check(_lib.load().lnb_lanczos_ritz(
        _stream(A), _ptr(A), _ptr(mask), _ptr(q1), B, N, K, 1 if proper else 0,
        _ptr(out.get('T')), _ptr(out.get('Q')),
        _ptr(out['alpha']), _ptr(out['beta']), _ptr(out['idx']), _ptr(out.get('theta')),
        _ptr(out.get('V')), _ptr(out.get('status'))), 'lnb_lanczos_ritz')
  return out


def tridiag_powers(T, powers):
  """out[b, r, s, c] = (T_b ** powers[s])[r, c]  (MLP input layout of ada_lanczos_net.py:274)."""
  _need_cuda(T)
  T = _f32c(T)
  B, K = T.shape[0], T.shape[1]
  S = len(powers)
  out = torch.empty((B, K, S, K), device=T.device, dtype=torch.float32)
  with torch.cuda.device(T.device):
    _lib.check(_lib.load().lnb_tridiag_powers(_stream(T), _ptr(T), B, K, _ints(powers), S,
                                              _ptr(out)), 'lnb_tridiag_powers')
  return out


def symmetrize_filters(Y, K, S):
  """G[b,s,r,c] = (Y[b,r,c,s] + Y[b,c,r,s]) / 2 for Y viewed as [B,K,K,S]."""
  _need_cuda(Y)
  Y = _f32c(Y)
  B = Y.shape[0]
  G = torch.empty((B, S, K, K), device=Y.device, dtype=torch.float32)
  with torch.cuda.device(Y.device):
    _lib.check(_lib.load().lnb_symmetrize_filters(_stream(Y), _ptr(Y), B, K, S, _ptr(G)),
               'lnb_symmetrize_filters')
  return G


def segment_sum_forward(data, segment_index, num_segments, output=None):
  _need_cuda(data, segment_index, output)
  data = _f32c(data)
  seg = segment_index.contiguous().long()
  B, d1, d2 = data.shape
  if output is None:
    output = torch.zeros((B, num_segments, d2), device=data.device, dtype=torch.float32)
  with torch.cuda.device(data.device):
    _lib.check(_lib.load().lnb_unsorted_segment_sum_forward(
        _stream(data), _ptr(data), _ptr(seg), _ints([B, d1, d2]), int(num_segments), _ptr(output)),
               'lnb_unsorted_segment_sum_forward')
  return output


def segment_sum_backward(grad_output, segment_index, data_shape, grad_data=None):
  _need_cuda(grad_output, segment_index, grad_data)
  grad_output = _f32c(grad_output)
  seg = segment_index.contiguous().long()
  B, d1, d2 = [int(v) for v in data_shape]
  if grad_data is None:
    grad_data = torch.empty((B, d1, d2), device=grad_output.device, dtype=torch.float32)
  with torch.cuda.device(grad_output.device):
    _lib.check(_lib.load().lnb_unsorted_segment_sum_backward(
        _stream(grad_output), _ptr(grad_output), _ptr(seg), _ints([B, d1, d2]),
        int(grad_output.shape[1]), _ptr(grad_data)), 'lnb_unsorted_segment_sum_backward')
  return grad_data
