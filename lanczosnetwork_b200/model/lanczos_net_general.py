"""Drop-in for the reference ``model.LanczosNetGeneral`` (model/lanczos_net_general.py:13-201):
LanczosNet with float node features instead of an atom embedding (:156) and the edge-type
count taken from ``config.dataset.num_edge_type`` (:24)."""

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
    dev = self._device()
    if self._check_mode():
      score = self._train_impl(*[self._to(dev, t) for t in (node_feat, L, D, V, mask)])
    else:
      score = self._graph_forward(self._forward_impl, (node_feat, L, D, V, mask))
    return self._finish(score, self._to(dev, label))

  def _train_impl(self, node_feat, L, D, V, mask):
    from ..train import ritz_stack_train
    return ritz_stack_train(self, node_feat, None, L, D, V, mask)

  def _forward_impl(self, node_feat, L, D, V, mask):
    return self._ritz_conv_stack(node_feat.float().contiguous(), None, L.float().contiguous(),
                                 D.float().contiguous(), V.float().contiguous(), mask)
