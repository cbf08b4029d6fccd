"""Property tests (hypothesis, CPU) of the host-side logic around the hot path: the packed-tile rule
that the stack kernel relies on, the layout of a packed sparse batch, and invariants of the oracle's
Lanczos restatement that do not depend on any fixture (test infrastructure pinning itself)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from lanczosnetwork_b200 import data
from oracle import lanczos_oracle as orc


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 128), st.integers(0, 64)), min_size=0, max_size=300))
def test_host_tile_table_is_the_next_fit_packing(graphs):
  """Rule of lnb_graph_prepare / tile_assign (csrc/spectral_conv_fused.cu): consecutive graphs share a
  tile while sum n <= 128, sum ceil4(k_eff) <= 128 and <= 32 graphs; next-fit means a tile is closed
  ONLY when the next graph would break one of the limits."""
  sizes = np.array([g[0] for g in graphs], np.int32)
  k_eff = np.array([min(g[1], g[0]) for g in graphs], np.int64)
  tiles = data.host_tile_table(sizes, k_eff)
  B = len(graphs)
  T = int(tiles[0])
  assert tiles.shape == (B + 2,) and tiles.dtype == np.int32
  starts = tiles[1:2 + T]
  assert starts[-1] == B and (T == 0) == (B == 0)
  if B:
    assert starts[0] == 0 and np.all(np.diff(starts) >= 1)
  k4 = (k_eff + 3) // 4 * 4
  for t in range(T):
    a, b = int(starts[t]), int(starts[t + 1])
    n_sum, k_sum = int(sizes[a:b].sum()), int(k4[a:b].sum())
    assert b - a <= 32
    if b - a > 1:                              # a lone graph always fits (the kernel's tile is 128 rows)
      assert n_sum <= 128 and k_sum <= 128
    if b < B:                                  # closed for a reason
      assert b - a == 32 or n_sum + sizes[b] > 128 or k_sum + k4[b] > 128
  assert np.all(tiles[2 + T:] == 0)


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 2000), st.integers(1, 64))
def test_packed_offsets_are_aligned_disjoint_segments(B, K):
  off_sizes, off_node_ptr, off_edge_ptr, off_D, off_var, off_tiles, off_krow = data.packed_offsets(B, K)
  segs = sorted([(off_sizes, 4 * B), (off_node_ptr, 4 * (B + 1)), (off_edge_ptr, 4 * (B + 1)), (off_D, 4 * B * K),
                 (off_tiles, 4 * (B + 2)), (off_krow, 4 * (B + 1))])
  assert segs[0][0] == 64                      # the 16-int header
  for (o, n), (o2, _) in zip(segs, segs[1:] + [(off_var, 0)]):
    assert o % 16 == 0 and o + n <= o2


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 12), st.integers(0, 10 ** 6), st.sampled_from([4, 20]))
def test_pack_sparse_round_trips_every_segment(B, seed, K):
  samples = data.synthetic_qm8_samples(B, seed=seed)
  sp = data.sparse_collate(samples, K)
  pk = data.pack_sparse(sp)
  blob = pk['blob']
  hdr = blob[:64].view(np.int32)
  assert hdr[0] == data.PACK_MAGIC and (hdr[1], hdr[2]) == (B, K) and hdr[10] == blob.size and blob.size % 16 == 0

  def seg(off, dtype, count):
    return blob[off:off + count * np.dtype(dtype).itemsize].view(dtype)

  n_nodes, n_edges = int(sp['node_ptr'][-1]), int(sp['edge_ptr'][-1])
  assert np.array_equal(seg(hdr[3], np.int32, B), sp['sizes'])
  assert np.array_equal(seg(hdr[4], np.int32, B + 1), sp['node_ptr'])
  assert np.array_equal(seg(hdr[5], np.int32, B + 1), sp['edge_ptr'])
  assert np.array_equal(seg(hdr[6], np.float32, B * K).reshape(B, K), sp['D'])
  assert np.array_equal(seg(hdr[7], np.int32, n_nodes), sp['node_feat'])
  assert np.array_equal(seg(hdr[8], np.float32, n_nodes * K).reshape(n_nodes, K), sp['V_rows'])
  assert np.array_equal(seg(hdr[9], np.uint8, n_edges * 4).reshape(n_edges, 4), sp['edges'])
  # the tile table and the Ritz-row prefix that ride along (LNB_PACKED_HOST_TILES)
  tiles = seg(hdr[11], np.int32, B + 2)
  krow = seg(hdr[12], np.int32, B + 1)
  k_eff = np.diff(krow)
  assert krow[0] == 0 and np.all(k_eff >= 0) and np.all(k_eff <= K)
  assert np.array_equal(tiles, data.host_tile_table(sp['sizes'], k_eff))
  for b in range(B):
    rows = sp['V_rows'][sp['node_ptr'][b]:sp['node_ptr'][b + 1]]
    assert not rows[:, k_eff[b]:].any() and (k_eff[b] == 0 or rows[:, k_eff[b] - 1].any())


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 24), st.integers(1, 12), st.integers(0, 10 ** 6))
def test_oracle_lanczos_is_a_krylov_factorisation_where_it_says_valid(N, K, seed):
  """On the directions the reference's rules keep (columns < idx): Q has orthonormal columns, T is
  symmetric tridiagonal with T = Q^T A Q on that block, and everything past idx is exactly zero
  except the one sub-diagonal beta the reference keeps in front of a breakdown."""
  rng = np.random.RandomState(seed)
  n = int(rng.randint(2, N + 1))
  adj = np.triu((rng.rand(n, n) < 0.4).astype(np.float64), 1)
  adj = adj + adj.T
  A = np.zeros((1, N, N))
  A[0, :n, :n] = data.get_laplacian(adj)
  mask = np.zeros((1, N), np.uint8)
  mask[0, :n] = 1
  q1 = rng.randn(1, N)
  out = orc.lanczos_tridiagonalise(torch.from_numpy(A), torch.from_numpy(mask), torch.from_numpy(q1), K)
  T, Q, idx = out['T'][0].numpy(), out['Q'][0].numpy(), int(out['idx'][0])
  assert 0 <= idx <= min(n, K)
  assert np.array_equal(T, T.T) and not np.triu(T, 2).any()
  assert not Q[:, idx:].any() and not Q[idx:, :].any()
  # the alpha of the breakdown step is dropped, the beta in front of it stays (ada_lanczos_net.py:213-224)
  assert not T[idx + 1:, :].any() and (idx >= K or (T[idx, idx] == 0 and not T[idx, :max(idx - 1, 0)].any()))
  if idx == n and idx >= 1:
    # the row mask (rows >= idx zeroed, ada_lanczos_net.py:226-237) removes nothing: a true factorisation
    Qk = Q[:n, :idx]
    np.testing.assert_allclose(Qk.T @ Qk, np.eye(idx), atol=1e-8)
    np.testing.assert_allclose(np.triu(np.tril((Qk.T @ A[0, :n, :n] @ Qk), 1), -1)[:idx - 1, :idx - 1] if idx > 1 else np.zeros((0, 0)),
                               T[:idx - 1, :idx - 1] if idx > 1 else np.zeros((0, 0)), atol=1e-8)


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 10), st.integers(0, 10 ** 6))
def test_sparse_records_rebuild_the_padded_batch(B, seed):
  """The sparse records of a batch (what crosses PCIe on the forward_sparse path) carry everything the
  reference's padded batch holds: rebuilding the operators from the bond lists through the host mirror
  of the L4 normalisation gives ``collate``'s tensor bit for bit, for any batch."""
  samples = data.synthetic_qm8_samples(B, seed=seed)
  dense = data.collate(samples, 20)
  sp = data.sparse_collate(samples, 20)
  N = dense['L'].shape[1]
  assert sp['N'] == N and int(sp['node_ptr'][-1]) == int(dense['node_mask'].sum())
  rebuilt = np.zeros_like(dense['L'])
  for b in range(B):
    n = int(sp['sizes'][b])
    adjs = np.zeros((n, n, 6))
    for u, v, c, _ in sp['edges'][sp['edge_ptr'][b]:sp['edge_ptr'][b + 1]]:
      adjs[u, v, c] = adjs[v, u, c] = 1.0
    rebuilt[b, :n, :n, 0] = data.get_laplacian(adjs.sum(axis=2))
    for c in range(6):
      rebuilt[b, :n, :n, 1 + c] = data.get_laplacian(adjs[:, :, c])
  assert np.array_equal(rebuilt, dense['L'])
  # and 14x fewer bytes than the padded tensors at the bench shape is a property of the data, not of B
  sparse_bytes = sum(np.asarray(sp[k]).nbytes for k in ('sizes', 'node_ptr', 'node_feat', 'edge_ptr', 'edges', 'V_rows', 'D'))
  dense_bytes = sum(dense[k].nbytes for k in ('node_feat', 'node_mask', 'L', 'D', 'V'))
  assert sparse_bytes < dense_bytes


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 10 ** 6), st.lists(st.integers(0, 59), min_size=1, max_size=40), st.sampled_from([4, 20]))
def test_packed_molecules_batches_are_pack_sparse_byte_for_byte(seed, idx, K):
  """data.PackedMolecules flattens a split once and assembles a batch by vectorised gathers; the blob is
  the one pack_sparse(sparse_collate(...)) builds molecule by molecule -- same bytes, same header -- for
  any index list (repeats and any order included), also into a reused buffer full of stale bytes."""
  samples = data.synthetic_qm8_samples(60, seed=seed % 7)       # a few distinct pools, cached below
  pool = _POOLS.setdefault((seed % 7, K), data.PackedMolecules(samples, K))
  ref = data.pack_sparse(data.sparse_collate([samples[i] for i in idx], K))
  got = pool.batch(idx)
  assert np.array_equal(got['blob'], ref['blob'])
  assert (got['B'], got['N'], got['K'], got['num_edgetype']) == (ref['B'], ref['N'], ref['K'], ref['num_edgetype'])
  assert np.array_equal(got['label'], ref['label'])
  stale = np.full(pool.max_bytes(len(idx)) + 64, 0xAB, np.uint8)
  again = pool.batch(idx, out=stale)
  assert again['blob'].base is stale and np.array_equal(again['blob'], ref['blob'])
  assert pool.max_bytes(len(idx)) >= ref['blob'].size


_POOLS = {}
