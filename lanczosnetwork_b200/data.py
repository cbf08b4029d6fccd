"""Host-side graph preparation for the hot path (numpy; no torch, no CUDA).

Mirrors the parts of the reference that *feed* the spectral-convolution forward:

  * ``get_laplacian(adj, 'L4')``      utils/data_helper.py:119-166 (L4 branch :155-156,
                                      normalisation :92-116)
  * ``get_graph_laplacian_eigs``      utils/data_helper.py:169-226 (dense eigh branch,
                                      -|lambda| stable ordering :217-223)
  * ``prepare_graph``                 dataset/get_qm8_data.py:60-90 / get_graph_data.py:52-92
  * ``collate``                       dataset/qm8.py:57-90,220-291 (default branch)
  * synthetic generators              dataset/get_graph_data.py:15-49 (G(n,p) regression set)
                                      and a QM8-shaped molecule sampler

All arithmetic follows the reference's fp64-then-cast-to-fp32 convention so the padded
batch tensors are bit-identical to what the reference loader would hand the model.
"""
import numpy as np

__all__ = [
    'check_dist', 'get_laplacian', 'get_graph_laplacian_eigs', 'prepare_graph',
    'collate', 'sparse_collate', 'pack_sparse', 'packed_offsets', 'synthetic_molecule', 'synthetic_qm8_samples', 'synthetic_qm8_batch',
    'synthetic_regression_graphs',
]

EPS = float(np.finfo(np.float32).eps)


def check_dist(dist):
  """utils/data_helper.py:9-14: diffusion distances must be ints or the string 'inf'."""
  for dd in dist:
    if not isinstance(dd, int) and dd != 'inf':
      raise ValueError("Non-supported value of diffusion distance")
  return dist


def _normalize_sym(mat, exponent=0.5):
  deg = mat.sum(axis=1)
  with np.errstate(divide='ignore'):
    scale = np.power(deg, -exponent)
  scale[np.isinf(scale)] = 0.0
  # same rounding as diag(s) @ M @ diag(s): fl(fl(s_i * m_ij) * s_j)
  return (scale[:, None] * mat) * scale[None, :]


def get_laplacian(adj, graph_laplacian_type='L4', alpha=0.5):
  """Dense graph operators of utils/data_helper.py:119-166.  The hot path only consumes
  'L4' (GCN renormalisation); L1/L2/L6 are provided because the same file defines them."""
  adj = np.asarray(adj)
  if not np.issubdtype(adj.dtype, np.floating):
    adj = adj.astype(np.float64)
  # NB: like the reference, arithmetic stays in the adjacency's dtype (the shipped
  # preprocessors feed float64; eye() promotes the L4 / L2 sums to float64 either way)
  if adj.ndim != 2 or adj.shape[0] != adj.shape[1]:
    raise ValueError('adjacency must be square')
  eye = np.eye(adj.shape[0])
  if graph_laplacian_type == 'L1':
    return np.diag(adj.sum(axis=1)) - adj
  if graph_laplacian_type == 'L2':
    return eye - _normalize_sym(adj)
  if graph_laplacian_type == 'L4':
    return _normalize_sym(eye + adj)
  if graph_laplacian_type == 'L6':
    return _normalize_sym(adj, exponent=alpha)
  raise ValueError('Unsupported Graph Laplacian!')


def get_graph_laplacian_eigs(adj, k=100, graph_laplacian_type='L4'):
  """(eigs[:k], V[:, :k], L) with eigenpairs ordered by descending |lambda| (stable), the
  dense-eigh branch of utils/data_helper.py:169-226."""
  lap = get_laplacian(adj, graph_laplacian_type)
  vals, vecs = np.linalg.eigh(lap)
  order = np.argsort(-np.abs(vals), kind='mergesort')[:k]
  return vals[order], vecs[:, order], lap


def prepare_graph(adjs, node_feat, label=None):
  """Per-graph record with the keys the collate step reads."""
  adjs = np.asarray(adjs, dtype=np.float64)
  if adjs.ndim == 2:
    adjs = adjs[:, :, None]
  D, V, L4 = get_graph_laplacian_eigs(adjs.sum(axis=2))
  L_multi = np.stack([get_laplacian(adjs[:, :, e]) for e in range(adjs.shape[2])], axis=2)
  # bond list {u, v, type}, u <= v, each undirected bond once: the sparse record sparse_collate ships
  edges = np.array([(u, v, c) for c in range(adjs.shape[2])
                    for u, v in zip(*np.nonzero(np.triu(adjs[:, :, c])))], dtype=np.uint8).reshape(-1, 3)
  rec = {'node_feat': np.asarray(node_feat), 'L_multi': L_multi, 'L_simple_4': L4,
         'D_simple': D, 'V_simple': V, 'edges': edges}
  if label is not None:
    rec['label'] = np.asarray(label)
  return rec


def collate(samples, num_eigs, num_nodes=None):
  """Zero-pad a list of ``prepare_graph`` records to the batch-max node count (or to ``num_nodes`` if
  that is larger: fixed shapes for a captured training step, train.GraphedStep).

  Returns numpy arrays: node_feat (B,N) int64 or (B,N,D) float32, node_mask (B,N) uint8,
  L (B,N,N,E+1) float32 with channel 0 the simple-graph operator, D (B,K), V (B,N,K)."""
  sizes = np.array([s['L_simple_4'].shape[0] for s in samples])
  B, N = len(samples), max(int(sizes.max()), int(num_nodes or 0))
  E = samples[0]['L_multi'].shape[2]
  nf0 = np.asarray(samples[0]['node_feat'])
  node_feat = (np.zeros((B, N), np.int64) if nf0.ndim == 1
               else np.zeros((B, N, nf0.shape[1]), np.float32))
  mask = (np.arange(N)[None, :] < sizes[:, None]).astype(np.uint8)
  L = np.zeros((B, N, N, E + 1), np.float32)
  D = np.zeros((B, num_eigs), np.float32)
  V = np.zeros((B, N, num_eigs), np.float32)
  for b, s in enumerate(samples):
    n = int(sizes[b])
    node_feat[b, :n] = s['node_feat']
    L[b, :n, :n, 0] = s['L_simple_4']
    L[b, :n, :n, 1:] = s['L_multi']
    kk = min(num_eigs, s['D_simple'].shape[0])
    D[b, :kk] = s['D_simple'][:kk]
    V[b, :n, :kk] = s['V_simple'][:, :kk]
  out = {'node_feat': node_feat, 'node_mask': mask, 'L': L, 'D': D, 'V': V}
  if 'label' in samples[0]:
    out['label'] = np.concatenate([np.asarray(s['label'], np.float32).reshape(1, -1)
                                   for s in samples], axis=0)
  return out


def sparse_collate(samples, num_eigs):
  """The same batch as ``collate`` in SPARSE form for the GPU-side batch construction
  (ops.graph_prepare_sparse / LanczosNet.forward_sparse): nothing is padded on the host and the
  dense operators are not shipped at all.

  Returns numpy arrays: sizes [B] int32, node_ptr [B+1] int32 (prefix sums), node_feat [sum n] int32,
  edge_ptr [B+1] int32, edges [sum E, 4] uint8 = {u, v, bond type, 0}, D [B,K] float32 (truncated /
  zero padded like dataset/qm8.py:268-287), V_rows [sum n, K] float32 (rows of real nodes only),
  N = batch-max node count (the reference's padding target), num_edgetype, label if present."""
  sizes = np.array([s['L_simple_4'].shape[0] for s in samples], np.int32)
  B = len(samples)
  node_ptr = np.zeros(B + 1, np.int32)
  node_ptr[1:] = np.cumsum(sizes)
  edge_ptr = np.zeros(B + 1, np.int32)
  edge_ptr[1:] = np.cumsum([len(s['edges']) for s in samples])
  node_feat = np.concatenate([np.asarray(s['node_feat']).astype(np.int32) for s in samples])
  edges = np.zeros((int(edge_ptr[-1]), 4), np.uint8)
  D = np.zeros((B, num_eigs), np.float32)
  V_rows = np.zeros((int(node_ptr[-1]), num_eigs), np.float32)
  for b, s in enumerate(samples):
    edges[edge_ptr[b]:edge_ptr[b + 1], :3] = s['edges']
    kk = min(num_eigs, s['D_simple'].shape[0])
    D[b, :kk] = s['D_simple'][:kk]
    V_rows[node_ptr[b]:node_ptr[b + 1], :kk] = s['V_simple'][:, :kk]
  out = {'sizes': sizes, 'node_ptr': node_ptr, 'node_feat': node_feat, 'edge_ptr': edge_ptr,
         'edges': edges, 'D': D, 'V_rows': V_rows, 'N': int(sizes.max()),
         'num_edgetype': int(samples[0]['L_multi'].shape[2])}
  if 'label' in samples[0]:
    out['label'] = np.concatenate([np.asarray(s['label'], np.float32).reshape(1, -1)
                                   for s in samples], axis=0)
  return out


PACK_MAGIC = 0x4c4e4231          # "LNB1"


def _align16(x):
  return (int(x) + 15) & ~15


def packed_offsets(B, K):
  """Byte offsets of the fixed-position segments of a packed batch (they depend on B and K only):
  (off_sizes, off_node_ptr, off_edge_ptr, off_D, off_variable)."""
  off_sizes = 64
  off_node_ptr = off_sizes + _align16(4 * B)
  off_edge_ptr = off_node_ptr + _align16(4 * (B + 1))
  off_D = off_edge_ptr + _align16(4 * (B + 1))
  off_tiles = off_D + _align16(4 * B * K)
  off_krow = off_tiles + _align16(4 * (B + 2))
  off_var = off_krow + _align16(4 * (B + 1))
  return off_sizes, off_node_ptr, off_edge_ptr, off_D, off_var, off_tiles, off_krow


def host_tile_table(sizes, k_eff, rows_per_tile=128, graphs_per_tile=32):
  """Packed-tile table of the fused convolution kernel, computed on the host with the rule of
  lnb_graph_prepare (csrc/spectral_conv_fused.cu: next-fit over consecutive graphs, sum n_eff <= 128,
  sum ceil4(k_eff) <= 128, <= 32 graphs per tile; the first graph of a tile always fits).
  Returns int32 [B+2]: [T, first graph of tile 0..T-1, B, 0 ...]."""
  B = len(sizes)
  tiles = np.zeros(B + 2, np.int32)
  if B == 0:
    return tiles
  cn = np.concatenate([[0], np.cumsum(np.asarray(sizes, np.int64))])
  ck = np.concatenate([[0], np.cumsum((np.asarray(k_eff, np.int64) + 3) // 4 * 4)])
  # jump table for EVERY start i at once (largest j with prefix[j] - prefix[i] <= limit: graphs i .. j-1
  # share a tile), then the tiles are the orbit of graph 0 -- the pointer-jumping form of tile_assign_kernel
  first = np.arange(B, dtype=np.int64)
  jn = np.searchsorted(cn, cn[:-1] + rows_per_tile, side='right') - 1
  jk = np.searchsorted(ck, ck[:-1] + rows_per_tile, side='right') - 1
  nxt = np.maximum(first + 1, np.minimum(np.minimum(jn, jk), np.minimum(first + graphs_per_tile, B))).tolist()
  starts, i = [], 0
  while i < B:
    starts.append(i)
    i = nxt[i]
  T = len(starts)
  tiles[1:1 + T] = starts
  tiles[0] = T
  tiles[1 + T] = B
  return tiles


def ritz_extents(V_rows, node_ptr):
  """k_eff per graph = last non-zero column of its Ritz rows + 1 (what the device measures in
  lnb_graph_prepare); vectorised over the batch (every graph has at least one node)."""
  B, K = len(node_ptr) - 1, V_rows.shape[1]
  if B == 0:
    return np.zeros(0, np.int64)
  any_col = np.logical_or.reduceat(V_rows != 0, np.asarray(node_ptr[:-1], np.int64), axis=0)   # [B, K]
  last = K - np.argmax(any_col[:, ::-1], axis=1)
  return np.where(any_col.any(axis=1), last, 0).astype(np.int64)


def pack_sparse(sp):
  """A ``sparse_collate`` batch as ONE contiguous uint8 buffer (layout: include/lanczosnet_b200.h,
  lnb_graph_prepare_sparse_packed): a 16-int header with the byte offsets of the segments, the
  fixed-size segments (sizes, node_ptr, edge_ptr, D), then node ids, Ritz rows and the bond list.
  One H2D copy per step ships the whole batch.  Returns dict(blob, B, N, K, num_edgetype[, label])."""
  B, K = sp['D'].shape
  off_sizes, off_node_ptr, off_edge_ptr, off_D, off, off_tiles, off_krow = packed_offsets(B, K)
  # extents the device would measure: k_eff = last non-zero column of the graph's Ritz rows + 1
  k_eff = ritz_extents(sp['V_rows'], sp['node_ptr'])
  tiles = host_tile_table(sp['sizes'], k_eff)
  krow = np.zeros(B + 1, np.int32)
  krow[1:] = np.cumsum(np.minimum(k_eff, K))
  off_nf = off
  off_v = off_nf + _align16(sp['node_feat'].nbytes)
  off_e = off_v + _align16(sp['V_rows'].nbytes)
  total = off_e + _align16(sp['edges'].nbytes)
  blob = np.zeros(total, np.uint8)
  hdr = blob[:64].view(np.int32)
  hdr[:13] = [PACK_MAGIC, B, K, off_sizes, off_node_ptr, off_edge_ptr, off_D, off_nf, off_v, off_e, total,
              off_tiles, off_krow]
  for off_, arr in ((off_tiles, tiles), (off_krow, krow), (off_sizes, sp['sizes']), (off_node_ptr, sp['node_ptr']), (off_edge_ptr, sp['edge_ptr']),
                    (off_D, sp['D']), (off_nf, sp['node_feat']), (off_v, sp['V_rows']), (off_e, sp['edges'])):
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    blob[off_:off_ + raw.size] = raw
  out = {'blob': blob, 'B': int(B), 'N': int(sp['N']), 'K': int(K), 'num_edgetype': int(sp['num_edgetype'])}
  if 'label' in sp:
    out['label'] = sp['label']
  return out


class PackedMolecules(object):
  """A whole split flattened ONCE into the segments of the packed batch format (node ids, bond lists,
  Ritz rows, Ritz values, extents): a batch over any index set is then a handful of vectorised gathers
  into one buffer -- the blob of ``pack_sparse(sparse_collate([samples[i] for i in idx], K))``, byte for
  byte -- instead of a Python loop over molecules per step (the reference pads and stacks per batch in
  DataLoader workers, dataset/qm8.py:220-291).  At 1024 molecules: ~1 ms per batch against 9 ms for
  sparse_collate + pack_sparse and 33 ms for the padded collate, i.e. one loader thread keeps up with
  a GPU step of 0.5 ms only with this path."""

  def __init__(self, samples, num_eigs):
    sp = sparse_collate(samples, num_eigs)
    self.K = int(num_eigs)
    self.num_edgetype = sp['num_edgetype']
    self.sizes, self.node_ptr, self.edge_ptr = sp['sizes'], sp['node_ptr'].astype(np.int64), sp['edge_ptr'].astype(np.int64)
    self.node_feat, self.edges, self.V_rows, self.D = sp['node_feat'], sp['edges'], sp['V_rows'], sp['D']
    self.k_eff = ritz_extents(self.V_rows, self.node_ptr)
    self.label = sp.get('label')

  def __len__(self):
    return len(self.sizes)

  @staticmethod
  def _ranges(starts, lens):
    """Concatenation of arange(starts[i], starts[i] + lens[i]) without a Python loop."""
    total = int(lens.sum())
    if total == 0:
      return np.zeros(0, np.int64)
    ends = np.cumsum(lens)
    return np.repeat(starts - (ends - lens), lens) + np.arange(total, dtype=np.int64)

  def max_bytes(self, B):
    """Upper bound of a B-molecule blob (for a reusable pinned staging buffer)."""
    n = int(np.sort(self.sizes)[-B:].sum())
    e = int(np.sort(np.diff(self.edge_ptr))[-B:].sum())
    return packed_offsets(B, self.K)[4] + _align16(4 * n) + _align16(4 * n * self.K) + _align16(4 * e)

  def batch(self, idx, out=None):
    """Packed batch of the molecules ``idx`` (order kept).  ``out``: optional uint8 buffer (e.g. the numpy
    view of a pinned tensor) of at least the blob's size; the returned blob is a view of it."""
    idx = np.asarray(idx, np.int64)
    B, K = len(idx), self.K
    sizes = self.sizes[idx]
    n_len = sizes.astype(np.int64)
    e_len = self.edge_ptr[idx + 1] - self.edge_ptr[idx]
    node_ptr = np.zeros(B + 1, np.int32)
    node_ptr[1:] = np.cumsum(n_len)
    edge_ptr = np.zeros(B + 1, np.int32)
    edge_ptr[1:] = np.cumsum(e_len)
    rows = self._ranges(self.node_ptr[idx], n_len)
    erow = self._ranges(self.edge_ptr[idx], e_len)
    k_eff = self.k_eff[idx]
    krow = np.zeros(B + 1, np.int32)
    krow[1:] = np.cumsum(np.minimum(k_eff, K))
    off_sizes, off_node_ptr, off_edge_ptr, off_D, off_nf, off_tiles, off_krow = packed_offsets(B, K)
    off_v = off_nf + _align16(4 * len(rows))
    off_e = off_v + _align16(4 * len(rows) * K)
    total = off_e + _align16(4 * len(erow))
    if out is None:
      blob = np.zeros(total, np.uint8)
    else:
      if out.dtype != np.uint8 or out.ndim != 1 or out.size < total:
        raise ValueError('PackedMolecules.batch: out must be a flat uint8 buffer of >= %d bytes' % total)
      blob = out[:total]
      blob[:off_nf] = 0                              # header + fixed segments (alignment gaps stay zero)
      for a, b in ((off_nf + 4 * len(rows), off_v), (off_v + 4 * len(rows) * K, off_e), (off_e + 4 * len(erow), total)):
        blob[a:b] = 0
    blob[:64].view(np.int32)[:13] = [PACK_MAGIC, B, K, off_sizes, off_node_ptr, off_edge_ptr, off_D, off_nf, off_v,
                                     off_e, total, off_tiles, off_krow]

    def put(off, arr):
      raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
      blob[off:off + raw.size] = raw

    put(off_tiles, host_tile_table(sizes, k_eff))
    put(off_krow, krow)
    put(off_sizes, sizes)
    put(off_node_ptr, node_ptr)
    put(off_edge_ptr, edge_ptr)
    put(off_D, self.D[idx])
    np.take(self.node_feat, rows, out=blob[off_nf:off_nf + 4 * len(rows)].view(np.int32))
    np.take(self.V_rows, rows, axis=0, out=blob[off_v:off_v + 4 * len(rows) * K].view(np.float32).reshape(len(rows), K))
    np.take(self.edges, erow, axis=0, out=blob[off_e:off_e + 4 * len(erow)].reshape(len(erow), 4))
    res = {'blob': blob, 'B': int(B), 'N': int(sizes.max()) if B else 0, 'K': K, 'num_edgetype': self.num_edgetype}
    if self.label is not None:
      res['label'] = self.label[idx]
    return res


# ----------------------------------------------------------------------------
# synthetic inputs (no dataset can be downloaded; shapes follow the reference)
# ----------------------------------------------------------------------------
def synthetic_molecule(rng, num_nodes, num_bond_type=6, num_atom=70, max_degree=4,
                       extra_edge_frac=0.3):
  """Random connected, degree-capped, molecule-like multigraph.

  A random spanning tree (each new atom bonds to a random earlier atom with free valence)
  plus ~extra_edge_frac*n ring-closing bonds; every bond gets one of num_bond_type channels.
  Returns (node_feat (n,) int64 in [0,num_atom), adjs (n,n,num_bond_type) float64)."""
  n = int(num_nodes)
  adjs = np.zeros((n, n, num_bond_type), np.float64)
  deg = np.zeros(n, np.int64)
  for v in range(1, n):
    free = np.flatnonzero(deg[:v] < max_degree)
    u = int(free[rng.randint(len(free))]) if len(free) else int(rng.randint(v))
    c = int(rng.randint(num_bond_type))
    adjs[u, v, c] = adjs[v, u, c] = 1.0
    deg[u] += 1
    deg[v] += 1
  for _ in range(int(round(extra_edge_frac * n))):
    u, v = int(rng.randint(n)), int(rng.randint(n))
    if u == v or adjs[u, v].sum() > 0 or deg[u] >= max_degree or deg[v] >= max_degree:
      continue
    c = int(rng.randint(num_bond_type))
    adjs[u, v, c] = adjs[v, u, c] = 1.0
    deg[u] += 1
    deg[v] += 1
  node_feat = rng.randint(0, num_atom, size=n).astype(np.int64)
  return node_feat, adjs


def synthetic_qm8_sizes(rng, batch_size, min_nodes=3, max_nodes=26, mean_nodes=16.0):
  sizes = np.clip(np.rint(rng.normal(mean_nodes, 4.5, size=batch_size)), min_nodes,
                  max_nodes).astype(np.int64)
  sizes[rng.randint(batch_size)] = max_nodes      # batch-max padding target N = max_nodes
  return sizes


def synthetic_qm8_samples(batch_size, seed=1234, num_bond_type=6, num_atom=70, num_label=16,
                          max_nodes=26):
  """Per-molecule records of a QM8-shaped batch: n_b in [3,26], mean ~16."""
  rng = np.random.RandomState(seed)
  sizes = synthetic_qm8_sizes(rng, batch_size, max_nodes=max_nodes)
  samples = []
  for n in sizes:
    nf, adjs = synthetic_molecule(rng, n, num_bond_type, num_atom)
    samples.append(prepare_graph(adjs, nf, label=rng.randn(1, num_label)))
  return samples


def synthetic_qm8_batch(batch_size, seed=1234, num_eigs=20, num_bond_type=6, num_atom=70,
                        num_label=16, max_nodes=26):
  """QM8-shaped padded batch: ``collate`` of ``synthetic_qm8_samples``."""
  return collate(synthetic_qm8_samples(batch_size, seed, num_bond_type, num_atom, num_label,
                                       max_nodes), num_eigs)


def synthetic_regression_graphs(num_graphs=16, seed=123, min_num_nodes=20, max_num_nodes=100,
                                node_emb_dim=10, graph_emb_dim=2, edge_prob=0.5):
  """The reference's synthetic graph-regression set (dataset/get_graph_data.py:15-49):
  X ~ randn(n,10), A = G(n, 0.5) with one edge type, Y ~ randn(1,2).  Same RNG call order
  as the reference so identical seeds give identical graphs."""
  import networkx as nx
  rng = np.random.RandomState(seed)
  sizes = rng.randint(min_num_nodes, high=max_num_nodes + 1, size=num_graphs)
  out = []
  for n in sizes:
    X = rng.randn(n, node_emb_dim)
    g = nx.fast_gnp_random_graph(int(n), edge_prob, seed=int(rng.randint(1000)))
    A = np.asarray(nx.to_numpy_array(g), dtype=np.float64)[:, :, None]
    Y = rng.randn(1, graph_emb_dim)
    out.append(prepare_graph(A, X, label=Y))
  return out
