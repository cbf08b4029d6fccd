"""Drop-in for the reference ``model.LanczosNet`` (model/lanczos_net.py:13-199): same
constructor, parameter names and ``forward(node_feat, L, D, V, label=None, mask=None)``;
the forward runs in hand-written sm_100a CUDA (no CPU path)."""
import torch
import torch.nn as nn

from .. import data as data_mod
from .. import ops
from ._common import Ragged, SpectralNetBase

__all__ = ['LanczosNet']


class LanczosNet(SpectralNetBase):

  def __init__(self, config):
    super(LanczosNet, self).__init__()
    self.num_atom = config.dataset.num_atom
    self._setup_common(config, config.dataset.num_bond_type,
                       len(config.model.long_diffusion_dist), 128)
    dims = self._build_layers()
    self.embedding = nn.Embedding(self.num_atom, self.input_dim)
    self._build_spectral_filter()
    self._build_head(dims)
    self._init_param()

  def forward(self, node_feat, L, D, V, label=None, mask=None):
    """
      node_feat: long B x N (atom ids); L: float B x N x N x (E+1); D: Ritz values B x K;
      V: Ritz vectors B x N x K; label: B x P; mask: B x N (uint8 / bool / float).
      Returns score (B x P) or (score, loss) when label is given.
    """
    dev = self._device()
    if self._check_mode():
      score = self._train_impl(*[self._to(dev, t) for t in (node_feat, L, D, V, mask)])
    else:
      score = self._graph_forward(self._forward_impl, (node_feat, L, D, V, mask))
    return self._finish(score, self._to(dev, label))

  def _train_impl(self, node_feat, L, D, V, mask):
    from ..train import ritz_stack_train
    return ritz_stack_train(self, None, node_feat, L, D, V, mask)

  def _forward_impl(self, node_feat, L, D, V, mask):
    return self._ritz_conv_stack(None, node_feat.long(), L.float().contiguous(),
                                 D.float().contiguous(), V.float().contiguous(), mask)

  def forward_sparse(self, batch, label=None):
    """Forward from a SPARSE batch (lanczosnetwork_b200.data.sparse_collate -> torch tensors, pinned
    host or device): per-molecule node ids, bond lists and the Ritz pairs of the real nodes.  The
    padded operators, mask, ELL rows and tile table are built on the device
    (lnb_graph_prepare_sparse); the dense B x N x N x (E+1) tensor of the reference's collate
    (dataset/qm8.py:220-262) is never materialised and never crosses PCIe.  Same scores as
    ``forward`` on the collated batch, bit for bit.  Returns score or (score, loss)."""
    if self._check_mode():
      raise NotImplementedError('forward_sparse is an inference path; train through forward()')
    dev = self._device()
    if 'blob' in batch:
      # packed batch (data.pack_sparse): ONE H2D copy of exactly the bytes present
      B, N, K = int(batch['B']), int(batch['N']), int(batch['K'])
      cap = data_mod.packed_offsets(B, K)[4] + 16 * 3 + 4 * B * N + 4 * B * N * K + 4 * B * N * 4
      blob = batch['blob']
      score = self._graph_forward(lambda b_: self._forward_packed_impl(B, N, K, b_),
                                  (Ragged(blob, max(cap, int(blob.shape[0]))),),
                                  extra_key=('packed', B, N, K))
      return self._finish(score, self._to(dev, label))
    N, B = int(batch['N']), int(batch['sizes'].shape[0])
    inputs = (batch['sizes'], batch['node_ptr'], Ragged(batch['node_feat'], B * N), batch['edge_ptr'],
              Ragged(batch['edges']), Ragged(batch['V_rows'], B * N), batch['D'])
    score = self._graph_forward(lambda *a: self._forward_sparse_impl(N, *a), inputs,
                                extra_key=('sparse', N))
    return self._finish(score, self._to(dev, label))

  def _forward_packed_impl(self, B, N, K, blob):
    E1 = self.num_edgetype + 1
    dense = not self._sparse_stack_ok(N, E1, K)
    off_D = data_mod.packed_offsets(B, K)[3]
    D = blob[off_D:off_D + 4 * B * K].view(torch.float32).reshape(B, K)     # fixed address in the buffer
    prep, node_ids, mask, V, L = ops.graph_prepare_sparse_packed(
        blob, B, N, E1, K, binarize=getattr(self, '_binarize_operators', False), want_dense=dense)
    return self._ritz_conv_stack(None, node_ids, L, D, V, mask, prep=prep, dims_hint=(N, E1))

  def _forward_sparse_impl(self, N, sizes, node_ptr, node_feat, edge_ptr, edges, V_rows, D):
    E1 = self.num_edgetype + 1
    K = V_rows.shape[1]
    dense = not self._sparse_stack_ok(N, E1, K)
    prep, node_ids, mask, V, L = ops.graph_prepare_sparse(
        sizes, node_ptr, node_feat, edge_ptr, edges, V_rows, N, E1,
        binarize=getattr(self, '_binarize_operators', False), want_dense=dense)
    return self._ritz_conv_stack(None, node_ids, L, D.float().contiguous(), V, mask, prep=prep,
                                 dims_hint=(N, E1))
