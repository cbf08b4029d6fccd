"""Drop-in for the reference ``model.GCN`` (model/gcn.py:8-119) -- SURVEY 8(f3): the sibling model
whose layer is the edge-type part of the LanczosNet layer (``msg = [L_e X]_e``, Linear, ReLU;
model/gcn.py:84-92) followed by the same gated readout (:95-110).  Same constructor, parameter
names and ``forward(node_feat, L, label=None, mask=None)``; the forward is ONE launch of the
fused convolution-stack kernel with no long scales (embedding gather + all layers + readout)."""
import torch
import torch.nn as nn

from ._common import SpectralNetBase, _opt
from ..spectral_conv import WeightCache

__all__ = ['GCN', 'GCNFP']


class GCN(SpectralNetBase):

  def __init__(self, config):
    super(GCN, self).__init__()
    m = config.model
    self.config = config
    self.input_dim = m.input_dim
    self.hidden_dim = m.hidden_dim
    self.output_dim = m.output_dim
    self.num_layer = m.num_layer
    self.num_atom = config.dataset.num_atom
    self.num_edgetype = config.dataset.num_bond_type
    self.dropout = _opt(m, 'dropout', 0.0)
    # no diffusion scales: the message is the E+1 edge-type products only
    self.short_diffusion_dist, self.long_diffusion_dist = [], []
    self.num_scale_short = self.num_scale_long = 0
    self.num_eig_vec = 0
    self.spectral_filter_kind = None
    self._wcache = WeightCache()
    dims = self._build_layers()
    self.embedding = nn.Embedding(self.num_atom, self.input_dim)
    self._build_head(dims)
    self._init_param()

  def forward(self, node_feat, L, label=None, mask=None):
    """
      node_feat: long B x N (atom ids); L: float B x N x N x (E+1); label: B x P;
      mask: B x N (uint8 / bool / float).  Returns score (B x P) or (score, loss).
    """
    dev = self._device()
    if self._check_mode():
      score = self._train_impl(*[self._to(dev, t) for t in (node_feat, L, mask)])
    else:
      score = self._graph_forward(self._forward_impl, (node_feat, L, mask))
    return self._finish(score, self._to(dev, label))

  def _train_impl(self, node_feat, L, mask):
    from ..train import ritz_stack_train
    if getattr(self, '_binarize_operators', False):
      L = (L != 0).to(torch.float32)                 # model/gcnfp.py:83
    return ritz_stack_train(self, None, node_feat, L, None, None, mask)

  def _forward_impl(self, node_feat, L, mask):
    L = L.float().contiguous()
    B, N = node_feat.shape
    # no Ritz vectors: an all-zero block makes lnb_graph_prepare take the extents from L alone
    V = torch.zeros((B, N, 4), device=L.device, dtype=torch.float32)
    return self._ritz_conv_stack(None, node_feat.long(), L, None, V, mask)


class GCNFP(GCN):
  """Drop-in for the reference ``model.GCNFP`` (model/gcnfp.py:8-125): GCN on the non-zero pattern
  of the operators (``L[L != 0] = 1.0``, :83).  The 0/1 values are produced while the operators are
  compressed (lnb_graph_prepare flag), so the dense tensor is never rewritten -- unlike the
  reference, the caller's ``L`` is left untouched."""
  _binarize_operators = True
