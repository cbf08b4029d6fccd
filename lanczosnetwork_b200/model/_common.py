"""Shared host-side scaffolding of the three drop-in modules.

Parameter containers, names and initialisation mirror the reference so its checkpoints load
unchanged (model/lanczos_net.py:15-93, utils/train_helper.py:14-32): ``embedding.weight``,
``filter.{i}.{weight,bias}``, ``spectral_filter.{l}.{0,2,4,6}.{weight,bias}``,
``att_func.0.{weight,bias}``.  The forward math lives in CUDA (lanczosnetwork_b200.ops).
"""
import os

import torch
import torch.nn as nn

from .. import _lib, ops
from ..data import check_dist
from ..spectral_conv import (GraphContext, WeightCache, graph_conv_layer,
                             ritz_filter_coefficients)


class Ragged(object):
  """A tensor whose leading dimension varies from batch to batch (bond lists, rows of real nodes).
  CUDA-graph replay keeps a static buffer of ``capacity`` rows and copies only the rows present; the
  consumer kernels read their extents from the pointer arrays that travel with the batch.  The
  default capacity rounds the row count up to a bucket so batches of similar size share a graph."""

  def __init__(self, tensor, capacity=None, bucket=4096):
    self.tensor = tensor
    rows = int(tensor.shape[0])
    self.capacity = int(capacity) if capacity is not None else max(bucket, -(-rows // bucket) * bucket)
    if rows > self.capacity:
      raise ValueError('Ragged: %d rows exceed the capacity %d' % (rows, self.capacity))

  @property
  def rows(self):
    return int(self.tensor.shape[0])

  def static_shape(self):
    return (self.capacity,) + tuple(self.tensor.shape[1:])


def _raw(t):
  return t.tensor if isinstance(t, Ragged) else t


def _opt(obj, name, default):
  return getattr(obj, name) if hasattr(obj, name) else default


class SpectralNetBase(nn.Module):
  """Common constructor pieces; subclasses define the embedding and the long-scale operator."""

  def _setup_common(self, config, num_edgetype, filter_mlp_in, filter_mlp_hidden):
    m = config.model
    self.config = config
    self.input_dim = m.input_dim
    self.hidden_dim = m.hidden_dim
    self.output_dim = m.output_dim
    self.num_layer = m.num_layer
    self.num_edgetype = num_edgetype
    self.dropout = _opt(m, 'dropout', 0.0)
    self.short_diffusion_dist = check_dist(m.short_diffusion_dist)
    self.long_diffusion_dist = check_dist(m.long_diffusion_dist)
    self.max_short_diffusion_dist = max(self.short_diffusion_dist) if self.short_diffusion_dist else None
    self.max_long_diffusion_dist = max(self.long_diffusion_dist) if self.long_diffusion_dist else None
    self.num_scale_short = len(self.short_diffusion_dist)
    self.num_scale_long = len(self.long_diffusion_dist)
    self.num_eig_vec = m.num_eig_vec
    self.spectral_filter_kind = m.spectral_filter_kind
    self._filter_mlp_dims = (filter_mlp_in, filter_mlp_hidden)
    self._wcache = WeightCache()

  def _build_layers(self):
    C = self.num_scale_short + self.num_scale_long + self.num_edgetype + 1
    dims = [self.input_dim] + list(self.hidden_dim) + [self.output_dim]
    self.filter = nn.ModuleList(
        [nn.Linear(dims[t] * C, dims[t + 1]) for t in range(self.num_layer)] +
        [nn.Linear(dims[-2], dims[-1])])
    return dims

  def _build_spectral_filter(self):
    if self.spectral_filter_kind == 'MLP' and self.num_scale_long > 0:
      n_in, hid = self._filter_mlp_dims
      self.spectral_filter = nn.ModuleList([
          nn.Sequential(nn.Linear(n_in, hid), nn.ReLU(), nn.Linear(hid, hid), nn.ReLU(),
                        nn.Linear(hid, hid), nn.ReLU(), nn.Linear(hid, n_in))
          for _ in range(self.num_layer)
      ])

  def _build_head(self, dims):
    self.att_func = nn.Sequential(nn.Linear(dims[-2], 1), nn.Sigmoid())
    loss = self.config.model.loss
    if loss == 'CrossEntropy':
      self.loss_func = torch.nn.CrossEntropyLoss()
    elif loss == 'MSE':
      self.loss_func = torch.nn.MSELoss()
    elif loss == 'L1':
      self.loss_func = torch.nn.L1Loss()
    else:
      raise ValueError("Non-supported loss function!")

  def _init_param(self):
    """Xavier-uniform weights, zero biases for filter / att_func / spectral_filter Linears
    (model/lanczos_net.py:74-93); the embedding keeps nn.Embedding's default N(0,1)."""
    groups = [self.filter, self.att_func]
    if hasattr(self, 'spectral_filter'):
      groups += list(self.spectral_filter)
    for grp in groups:
      for mod in grp:
        if isinstance(mod, nn.Linear):
          nn.init.xavier_uniform_(mod.weight.data)
          if mod.bias is not None:
            mod.bias.data.zero_()

  # ------------------------------------------------------------------------------------------
  def _device(self):
    dev = self.filter[0].weight.device
    if dev.type != 'cuda':
      raise RuntimeError(
          '%s runs on CUDA (sm_100a) only -- move the module with .cuda(); there is no CPU '
          'fallback (the CPU reference is the oracle).' % type(self).__name__)
    # DataParallel replicas share the master's WeightCache object: bypass it there (see WeightCache)
    if getattr(self, '_is_replica', False) and hasattr(self, '_wcache') and not self._wcache.bypass:
      self._wcache = type(self._wcache)()
      self._wcache.bypass = True
    return dev

  def invalidate_caches(self):
    """Drop the cached tf32 weight splits and the captured CUDA graphs (call after editing weights
    through ``p.data`` or any other route that does not bump the parameters' version counters)."""
    if hasattr(self, '_wcache'):
      self._wcache.invalidate()
    for name in ('_graphs', '_graphs_resident', '_resident_seen'):
      self.__dict__.pop(name, None)

  def graph_stats(self):
    """Counters of the CUDA-graph cache: captures (each costs a warm-up, two captures and two device
    synchronisations), replays, and live graphs -- a capture count that keeps growing means the input
    shapes thrash the cache (pad / bucket the batch shapes)."""
    st = self.__dict__.setdefault('_graph_stats', {'captures': 0, 'replays': 0})
    return dict(st, live=len(self.__dict__.get('_graphs', {})),
                live_resident=len(self.__dict__.get('_graphs_resident', {})))

  def _has_trainable_parameters(self):
    """nn.DataParallel replicas hold their parameter copies as plain attributes (``parameters()`` is
    empty there, the copies are listed in ``_former_parameters``): look at both."""
    for m in self.modules():
      for group in (m._parameters, getattr(m, '_former_parameters', None) or {}):
        for p in group.values():
          if p is not None and p.requires_grad:
            return True
    return False

  def _check_mode(self):
    """Returns True when this call has to be differentiable (autograd on, trainable parameters):
    the forward then runs the training path of lanczosnetwork_b200.train (unfused, every
    contraction and its adjoint in this library's kernels) instead of the fused inference kernels.
    Models without a training path raise."""
    if torch.is_grad_enabled() and self._has_trainable_parameters():
      if not hasattr(self, '_train_impl'):
        raise NotImplementedError(
            '%s has no training path in this build: call it under torch.no_grad() (as '
            'runner.test() / the validation loop do).' % type(self).__name__)
      return True
    if self.training and self.dropout > 0.0:
      raise NotImplementedError('dropout > 0 in training mode needs autograd enabled (the inference '
                                'kernels implement eval() semantics)')
    return False

  @staticmethod
  def _to(dev, t, dtype=None):
    if t is None:
      return None
    if t.device != dev:
      t = t.to(dev, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
      t = t.to(dtype)
    return t

  # ------------------------------------------------------------------------------------------
  # CUDA-graph replay of the inference forward: the forward is ~15 short kernel launches issued
  # through ctypes; capturing them once per input signature removes the per-launch host cost
  # (CUDA streams and graphs instead of a tracing compiler).  Inputs are copied into static
  # buffers (H2D straight from pinned host memory, or D2D), the graph is replayed, the small
  # score tensor is cloned out.  Recaptured when shapes or any parameter version change.
  use_cuda_graph = os.environ.get('LNB_NO_GRAPH', '0') != '1'   # LNB_NO_GRAPH=1: eager (profiling)

  def _param_signature(self):
    return tuple((p.data_ptr(), p._version) for p in self.parameters())

  def _graph_forward(self, impl, inputs, extra_key=()):
    """impl(*device_tensors) -> score; inputs: tuple of tensors / Ragged / None (CPU or CUDA).

    Two graph slots with their own static buffers alternate, and the input copies run on a
    dedicated copy stream: the H2D (or D2D) transfer of call i+1 overlaps the replay of call i
    (the forward returns without synchronising), ordered by events only."""
    dev = self._device()
    eligible = (self.use_cuda_graph and not self.training and not torch.is_grad_enabled() and
                not getattr(self, '_is_replica', False) and
                not torch.cuda.is_current_stream_capturing())
    if not eligible:
      return impl(*[self._to(dev, _raw(t)) for t in inputs])
    key = (dev.index,) + tuple(extra_key) + tuple(
        None if t is None else ((t.static_shape(), t.tensor.dtype, 'ragged') if isinstance(t, Ragged)
                                else (tuple(t.shape), t.dtype)) for t in inputs)
    cache = self.__dict__.setdefault('_graphs', {})
    stats = self.__dict__.setdefault('_graph_stats', {'captures': 0, 'replays': 0})
    entry = cache.get(key)
    if entry is not None:
      cache[key] = cache.pop(key)               # LRU: most recently used last
    sig = self._param_signature()
    cur = torch.cuda.current_stream(dev)
    # Inputs already resident on this device: a graph bound to their addresses needs no copy at
    # all.  Such a graph is captured the second time the same buffers show up (data loaders /
    # serving loops that recycle a few device buffers); any live tensor found at a captured
    # address with the captured shape and dtype is read correctly, so no reference is kept.
    raw = [_raw(t) for t in inputs]
    if all(t is None or (t.is_cuda and t.device == dev and t.is_contiguous()) for t in raw):
      pkey = key + tuple(None if t is None else t.data_ptr() for t in raw)
      zc = self.__dict__.setdefault('_graphs_resident', {})
      hit = zc.get(pkey)
      if hit is not None and hit['sig'] == sig:
        zc[pkey] = zc.pop(pkey)
        stats['replays'] += 1
        hit['graph'].replay()
        _lib.note_graph_replay(hit['kernels'])
        return hit['out'].clone()
      seen = self.__dict__.setdefault('_resident_seen', {})
      if len(seen) > 256:
        seen.clear()
      seen[pkey] = seen.get(pkey, 0) + 1
      if seen[pkey] >= 2 and entry is not None and entry['sig'] == sig:   # caches are warm
        if len(zc) >= 16:
          zc.pop(next(iter(zc)))
        if '_resident_pool' not in self.__dict__:
          self._resident_pool = torch.cuda.graph_pool_handle()
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        n0 = int(_lib.load().lnb_launch_count())
        with torch.cuda.graph(graph, pool=self._resident_pool):
          out = impl(*raw)
        hit = {'graph': graph, 'out': out, 'sig': sig,
               'kernels': int(_lib.load().lnb_launch_count()) - n0}
        zc[pkey] = hit
        stats['captures'] += 1
        graph.replay()
        _lib.note_graph_replay(hit['kernels'])
        return out.clone()
    if entry is None or entry['sig'] != sig:
      slots = []
      for _ in range(2):
        static_in = [None if t is None else
                     (torch.zeros(t.static_shape(), dtype=t.tensor.dtype, device=dev)
                      if isinstance(t, Ragged) else torch.empty(t.shape, dtype=t.dtype, device=dev))
                     for t in inputs]
        for s_, t in zip(static_in, inputs):
          if isinstance(t, Ragged):
            s_[:t.rows].copy_(t.tensor, non_blocking=True)
          elif s_ is not None:
            s_.copy_(t, non_blocking=True)
        if not slots:
          side = torch.cuda.Stream(device=dev)
          side.wait_stream(cur)
          with torch.cuda.stream(side):
            impl(*static_in)                   # warm-up: fills the weight caches
          cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        n0 = int(_lib.load().lnb_launch_count())
        with torch.cuda.graph(graph):
          static_out = impl(*static_in)
        slots.append({'graph': graph, 'in': static_in, 'out': static_out,
                      'kernels': int(_lib.load().lnb_launch_count()) - n0,
                      'free': torch.cuda.Event(), 'ready': torch.cuda.Event()})
        slots[-1]['free'].record(cur)
      entry = {'sig': sig, 'slots': slots, 'next': 0, 'copy': torch.cuda.Stream(device=dev)}
      stats['captures'] += 1
      cache.pop(key, None)
      if len(cache) >= 8:                      # bound the number of live graphs: evict the LRU entry
        cache.pop(next(iter(cache)))
      cache[key] = entry
    slot = entry['slots'][entry['next']]
    entry['next'] ^= 1
    copy = entry['copy']
    copy.wait_event(slot['free'])              # the previous replay of this slot has consumed its inputs
    if any(t is not None and t.is_cuda for t in raw):
      copy.wait_stream(cur)                    # device inputs produced on the caller's stream
    with torch.cuda.stream(copy):
      for s_, t in zip(slot['in'], inputs):
        if isinstance(t, Ragged):
          s_[:t.rows].copy_(t.tensor, non_blocking=True)     # only the rows present cross PCIe
        elif s_ is not None:
          s_.copy_(t, non_blocking=True)
      slot['ready'].record(copy)
    cur.wait_event(slot['ready'])
    stats['replays'] += 1
    slot['graph'].replay()
    slot['free'].record(cur)
    _lib.note_graph_replay(slot['kernels'])
    return slot['out'].clone()

  def _filter_mlp_params(self):
    if not hasattr(self, 'spectral_filter'):
      return None
    out = []
    for l, seq in enumerate(self.spectral_filter):
      out.append([('spectral_filter.%d.%d' % (l, i), seq[i].weight, seq[i].bias)
                  for i in (0, 2, 4, 6)])
    return out

  def _ritz_conv_stack(self, state, node_ids, L, D, V, mask, prep=None, dims_hint=None):
    """Convolution stack + readout of the Ritz-pair models (LanczosNet, LanczosNetGeneral).
    ``prep`` (ops.GraphPrep built on the device by ops.graph_prepare_sparse) replaces the pass over
    the dense operators; ``L`` may then be None when every layer runs in the fused stack
    (``dims_hint`` = (N, E1) of the absent tensor).

    Consecutive layers the fused kernel supports run as ONE persistent launch (embedding gather
    in front when every layer qualifies, readout behind); a leading layer with an unsupported
    input width (e.g. LanczosNetGeneral's 10 features) runs through the unfused ops first."""
    S, nl = self.num_scale_long, self.num_layer
    if L is not None:
      N, K, E1 = L.shape[1], V.shape[2], L.shape[3]
    else:
      (N, E1), K = dims_hint, V.shape[2]
    din0 = self.embedding.weight.shape[1] if node_ids is not None else state.shape[2]
    dims = [din0] + list(self.hidden_dim)
    ok = [ops.fused_conv_supported(N, dims[t], K, dims[t + 1], len(self.short_diffusion_dist),
                                   False, S, E1) for t in range(nl)]
    H = dims[1]
    uniform = all(d == H for d in dims[1:])
    first = 0
    while first < nl and not ok[first]:
      first += 1
    stack_ok = uniform and first < nl and all(ok[first:]) and nl - first <= 8
    binarize = getattr(self, '_binarize_operators', False)
    if binarize and not (stack_ok and first == 0):
      L = (L != 0).to(L.dtype)          # shapes off the fused path read the dense operators
      binarize = False
    if L is None and not (stack_ok and first == 0):
      raise RuntimeError('sparse batches without the dense operators need every layer on the fused '
                         'stack kernel; call ops.graph_prepare_sparse(..., want_dense=True)')
    ctx = GraphContext(L, V, binarize)
    if prep is not None:
      ctx._prep = prep
    coeffs = table = None
    if S > 0:
      mlp = self._filter_mlp_params() if self.spectral_filter_kind == 'MLP' else None
      # the power table does not depend on graph_prepare: fork it onto a side stream so the
      # two run concurrently (also inside the captured graph)
      cur = torch.cuda.current_stream(D.device)
      side = self.__dict__.setdefault('_side_streams', {}).get(D.device.index)
      if side is None:
        side = self._side_streams[D.device.index] = torch.cuda.Stream(device=D.device)
      side.wait_stream(cur)
      with torch.cuda.stream(side):
        table = ops.ritz_power_table(D, self.long_diffusion_dist)
      table.record_stream(cur)
      gext = ctx.prep() if (mlp is not None and stack_ok and first == 0) else None
      cur.wait_stream(side)
      coeffs, table = ritz_filter_coefficients(D, self.long_diffusion_dist, mlp, self._wcache, gext,
                                               table=table)

    def layer_coeff(t):
      if S == 0:
        return None
      return coeffs[t] if coeffs is not None else table

    if node_ids is not None and not (stack_ok and first == 0):
      state = ops.embedding_rows(node_ids, self.embedding.weight)
    for t in range(first if stack_ok else nl):           # unfused / single-layer prefix
      state = graph_conv_layer(state, ctx, layer_coeff(t), False, self.short_diffusion_dist, S,
                               self.filter[t].weight, self.filter[t].bias, self._wcache,
                               'filter.%d' % t, last=(t == nl - 1),
                               next_fused=(t + 1 < nl and ok[t + 1]))
    if not stack_ok:
      return self._readout(state, mask)
    layers = list(range(first, nl))
    kw = (S + E1) * max(dims[t] for t in layers)
    w_hi, w_lo, bias = self._wcache.split_conv_stack(
        'filter.stack.%d' % first, [self.filter[t].weight for t in layers],
        [self.filter[t].bias for t in layers], kw)
    if S == 0:
      coeff, stride = None, 0
    elif coeffs is not None:
      coeff, stride = coeffs[first], coeffs.stride(0)
    else:
      coeff, stride = table, 0
    head, att = self.filter[nl], self.att_func[0]
    if head.weight.shape[0] > 48:                        # fused readout holds <= 48 outputs
      state, _ = ops.spectral_stack_forward(
          ctx.prep(), V, w_hi, w_lo, bias, [dims[t] for t in layers], H, S, coeff=coeff,
          coeff_stride=stride, X=None if (node_ids is not None and first == 0) else state,
          node_ids=node_ids if first == 0 else None,
          emb=self.embedding.weight if (node_ids is not None and first == 0) else None,
          want_state=True)
      return self._readout(state, mask)
    _, score = ops.spectral_stack_forward(
        ctx.prep(), V, w_hi, w_lo, bias, [dims[t] for t in layers], H, S, coeff=coeff,
        coeff_stride=stride, X=None if (node_ids is not None and first == 0) else state,
        node_ids=node_ids if first == 0 else None,
        emb=self.embedding.weight if (node_ids is not None and first == 0) else None,
        readout=(head.weight, head.bias, att.weight.reshape(-1), att.bias), mask=mask)
    return score

  def _sparse_stack_ok(self, N, E1, K):
    """True when every layer of this model runs inside the one-launch stack kernel, so a sparse
    batch never needs the dense operator tensor."""
    S = self.num_scale_long
    din0 = self.embedding.weight.shape[1]
    dims = [din0] + list(self.hidden_dim)
    ok = all(ops.fused_conv_supported(N, dims[t], K, dims[t + 1], len(self.short_diffusion_dist), False, S, E1)
             for t in range(self.num_layer))
    return ok and all(d == dims[1] for d in dims[1:]) and self.num_layer <= 8

  def _readout(self, state, mask):
    head = self.filter[self.num_layer]
    att = self.att_func[0]
    return ops.readout(state, head.weight, head.bias, att.weight.reshape(-1), att.bias, mask)

  def _finish(self, score, label):
    if label is not None:
      return score, self.loss_func(score, label)
    return score
