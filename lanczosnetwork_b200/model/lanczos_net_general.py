"""Drop-in for the reference ``model.LanczosNetGeneral`` (model/lanczos_net_general.py:13-201):
LanczosNet with float node features instead of an atom embedding (:156) and the edge-type
count taken from ``config.dataset.num_edge_type`` (:24)."""
import torch

from .. import ops
from ..spectral_conv import GraphContext, graph_conv_layer, ritz_filter_coefficients
from ._common import SpectralNetBase

__all__ = ['LanczosNetGeneral']


class LanczosNetGeneral(SpectralNetBase):

  def __init__(self, config):
    super(LanczosNetGeneral, self).__init__()
    self.node_emb_dim = config.dataset.node_emb_dim
    self.graph_emb_dim = config.dataset.graph_emb_dim
    self._setup_common(config, config.dataset.num_edge_type,
                       len(config.model.long_diffusion_dist), 128)
    dims = self._build_layers()
    assert self.input_dim == self.node_emb_dim      # lanczos_net_general.py:45-46
    assert self.output_dim == self.graph_emb_dim
    self._build_spectral_filter()
    self._build_head(dims)
    self._init_param()

  def forward(self, node_feat, L, D, V, label=None, mask=None):
    """
      node_feat: float B x N x D node features; L: B x N x N x (E+1); D: Ritz values B x K;
      V: Ritz vectors B x N x K; label: B x P; mask: B x N.
    """
    self._check_mode()
    dev = self._device()
    score = self._graph_forward(self._forward_impl, (node_feat, L, D, V, mask))
    return self._finish(score, self._to(dev, label))

  def _forward_impl(self, node_feat, L, D, V, mask):
    L = L.float().contiguous()
    D = D.float().contiguous()
    V = V.float().contiguous()
    state = node_feat.float().contiguous()

    ctx = GraphContext(L, V)
    coeffs = table = None
    if self.num_scale_long > 0:
      mlp = self._filter_mlp_params() if self.spectral_filter_kind == 'MLP' else None
      dims = [state.shape[2]] + list(self.hidden_dim)
      all_fused = all(ops.fused_conv_supported(L.shape[1], dims[t], V.shape[2], dims[t + 1], 0,
                                               False, self.num_scale_long, L.shape[3])
                      for t in range(self.num_layer)) and not self.short_diffusion_dist
      gext = ctx.prep()[3] if (mlp is not None and all_fused) else None
      coeffs, table = ritz_filter_coefficients(D, self.long_diffusion_dist, mlp, self._wcache,
                                               gext)

    for tt in range(self.num_layer):
      coeff = None
      if self.num_scale_long > 0:
        coeff = coeffs[tt] if coeffs is not None else table
      state = graph_conv_layer(state, ctx, coeff, False, self.short_diffusion_dist,
                               self.num_scale_long, self.filter[tt].weight, self.filter[tt].bias,
                               self._wcache, 'filter.%d' % tt, last=(tt == self.num_layer - 1))
    return self._readout(state, mask)
