"""Generate golden vectors by EXECUTING THE REFERENCE (lrjconan/LanczosNetwork @ 5ee5467,
mounted read-only at /root/reference) in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference cannot travel to the GPU box, so its inputs/outputs are committed as small
fixtures.  Weights are NOT stored: they are regenerated from numpy seeds by
tests/helpers.deterministic_state_dict and loaded into the reference modules with
load_state_dict, so the fixtures hold only inputs, outputs and the seed.

Shims (none touches /root/reference; all are outside the hot path):
  * operators._ext.segment_reduction is pre-registered as an empty stub module (the reference
    imports it at package import time and never calls it);
  * np.expand_dims(2-D, axis=3) in the reference collate (dataset/qm8.py:254-259) raises on
    numpy 2 -> the collate is run with axis 3 mapped to 2 (the evident intent);
  * configs are SimpleNamespace objects (easydict is absent).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, REF)

for name in ('operators._ext', 'operators._ext.segment_reduction'):
  sys.modules[name] = types.ModuleType(name)

import operators  # noqa: E402  (reference package)
operators._ext = sys.modules['operators._ext']
operators._ext.segment_reduction = sys.modules['operators._ext.segment_reduction']

from model import LanczosNet, AdaLanczosNet, LanczosNetGeneral  # noqa: E402  (reference)
from model import GCN, GCNFP, DCNN, ChebyNet  # noqa: E402  (reference)
from utils import data_helper as ref_dh  # noqa: E402
import dataset.qm8 as ref_qm8  # noqa: E402
import dataset.graph_data as ref_gd  # noqa: E402

from helpers import deterministic_state_dict  # noqa: E402
from lanczosnetwork_b200 import configs, data  # noqa: E402


class _NpShim(object):
  """numpy proxy whose expand_dims maps the out-of-range axis of the reference collate."""

  def __getattr__(self, k):
    return getattr(np, k)

  @staticmethod
  def expand_dims(a, axis):
    a = np.asarray(a)
    if axis > a.ndim:
      axis = a.ndim
    return np.expand_dims(a, axis)


ref_qm8.np = _NpShim()
ref_gd.np = _NpShim()


def ref_prepare(adjs, node_feat, label):
  """dataset/get_qm8_data.py:60-90 / get_graph_data.py:52-92 using the reference helpers."""
  adj_simple = np.sum(adjs, axis=2)
  D_list, V_list, L_list = ref_dh.get_multi_graph_laplacian_eigs(
      adjs, graph_laplacian_type='L4', use_eigen_decomp=True, is_sym=True)
  D, V, L4 = ref_dh.get_graph_laplacian_eigs(
      adj_simple, graph_laplacian_type='L4', use_eigen_decomp=True, is_sym=True)
  return {
      'node_feat': node_feat, 'L_multi': np.stack(L_list, axis=2), 'L_simple_4': L4,
      'L_simple_6': L4, 'L_simple_7': L4,
      'D_simple': D if D is not None else np.ones(adjs.shape[0]),
      'V_simple': V if V is not None else np.eye(adjs.shape[0]),
      'label': label,
  }


def ref_collate(cls, cfg, samples):
  ds = cls.__new__(cls)
  ds.config = cfg
  ds.num_edgetype = getattr(cfg.dataset, 'num_bond_type', getattr(cfg.dataset, 'num_edge_type', 1))
  ds.model_name = cfg.model.name
  ds.use_eigs = True
  ds.num_eigs = cfg.model.num_eig_vec
  return ds.collate_fn(samples)


def save(name, **arrays):
  path = os.path.join(HERE, name)
  np.savez_compressed(path, **arrays)
  print('%-32s %8.1f KB' % (name, os.path.getsize(path) / 1024.0))


# ---------------------------------------------------------------------------------------------
def golden_data_helper():
  """The reference's only fixture: the 6-node adjacency of utils/data_helper.py:297-299."""
  adj = np.array([[0, 1, 0, 0, 1, 0], [1, 0, 1, 0, 1, 0], [0, 1, 0, 1, 0, 0],
                  [0, 0, 1, 0, 1, 1], [1, 1, 0, 1, 0, 0], [0, 0, 0, 1, 0, 0]]).astype(np.float32)
  out = {'adj': adj}
  for kind in ('L1', 'L2', 'L4', 'L6'):
    out[kind] = ref_dh.get_laplacian(adj, graph_laplacian_type=kind)
  D, V, L = ref_dh.get_graph_laplacian_eigs(adj, k=100, graph_laplacian_type='L4',
                                             use_eigen_decomp=True)
  out.update(D=D, V=V)
  D3, V3, _ = ref_dh.get_graph_laplacian_eigs(adj, k=3, graph_laplacian_type='L4',
                                              use_eigen_decomp=True)
  out.update(D3=D3, V3=V3)
  save('data_helper_fixture.npz', **out)


def golden_lanczosnet_qm8(B=8, data_seed=7, weight_seed=1234):
  cfg = configs.qm8_lanczos_net()
  rng = np.random.RandomState(data_seed)
  sizes = [26, 3, 9, 20, 21, 14, 17, 25][:B]
  samples, adj_list, nf_list = [], [], []
  for n in sizes:
    nf, adjs = data.synthetic_molecule(rng, n)
    adj_list.append(adjs)
    nf_list.append(nf)
    samples.append(ref_prepare(adjs, nf, rng.randn(1, 16)))
  batch = ref_collate(ref_qm8.QM8Data, cfg, samples)
  torch.manual_seed(0)
  model = LanczosNet(cfg)
  model.load_state_dict(deterministic_state_dict(model, weight_seed))
  model.eval()
  with torch.no_grad():
    score, loss = model(batch['node_feat'], batch['L'], batch['D'], batch['V'],
                        label=batch['label'], mask=batch['node_mask'])
    D_pow = [torch.pow(batch['D'], p) for p in cfg.model.long_diffusion_dist]
    Lf0 = model._get_spectral_filters(D_pow, batch['V'], 0)
    # no-MLP branch of the spectral filter (model/lanczos_net.py:118-121)
    cfg2 = configs.qm8_lanczos_net(spectral_filter_kind='power', num_layer=2,
                                   hidden_dim=[32, 32])
    model2 = LanczosNet(cfg2)
    model2.load_state_dict(deterministic_state_dict(model2, weight_seed + 100))
    model2.eval()
    score2 = model2(batch['node_feat'], batch['L'], batch['D'], batch['V'],
                    mask=batch['node_mask'])
  pack = np.zeros((B, 26, 26, 6), np.uint8)
  for b, a in enumerate(adj_list):
    pack[b, :a.shape[0], :a.shape[0]] = a
  save('lanczosnet_qm8.npz', sizes=np.array(sizes), adjs=pack,
       node_feat=batch['node_feat'].numpy(), node_mask=batch['node_mask'].numpy(),
       L=batch['L'].numpy(), D=batch['D'].numpy(), V=batch['V'].numpy(),
       label=batch['label'].numpy(), score=score.numpy(), loss=np.array(float(loss)),
       Lf0=Lf0.numpy(), score_power=score2.numpy(), weight_seed=np.array(weight_seed))


def golden_gcn_qm8(weight_seed=2468):
  """Reference GCN (model/gcn.py) on the inputs of lanczosnet_qm8.npz (the QM8 collate feeds GCN the
  same L = [L_simple_4 | L_multi], dataset/qm8.py:220-262); only the outputs are stored."""
  g = np.load(os.path.join(HERE, 'lanczosnet_qm8.npz'))
  cfg = configs.qm8_gcn()
  model = GCN(cfg)
  model.load_state_dict(deterministic_state_dict(model, weight_seed))
  model.eval()
  nf, L = torch.from_numpy(g['node_feat']), torch.from_numpy(g['L'])
  mask = torch.from_numpy(g['node_mask'])
  with torch.no_grad():
    score, loss = model(nf, L, label=torch.from_numpy(g['label']), mask=mask)
    score_nomask = model(nf, L)
  # GCNFP (model/gcnfp.py): same parameters layout, operators binarised IN PLACE by the forward
  cfg_fp = configs.qm8_gcn(name='GCNFP')
  model_fp = GCNFP(cfg_fp)
  model_fp.load_state_dict(deterministic_state_dict(model_fp, weight_seed + 1))
  model_fp.eval()
  with torch.no_grad():
    score_fp = model_fp(nf, L.clone(), mask=mask)
  # DCNN (model/dcnn.py): edge types + six powers of the simple-graph operator per layer
  model_dc = DCNN(configs.qm8_dcnn())
  model_dc.load_state_dict(deterministic_state_dict(model_dc, weight_seed + 2))
  model_dc.eval()
  with torch.no_grad():
    score_dc = model_dc(nf, L, mask=mask)
  # ChebyNet (model/cheby_net.py): Chebyshev chain on channel 0 + bond-type channels
  model_ch = ChebyNet(configs.qm8_cheby_net())
  model_ch.load_state_dict(deterministic_state_dict(model_ch, weight_seed + 3))
  model_ch.eval()
  with torch.no_grad():
    score_ch = model_ch(nf, L, mask=mask)
  save('gcn_qm8.npz', score=score.numpy(), loss=np.array(float(loss)),
       score_nomask=score_nomask.numpy(), score_fp=score_fp.numpy(), score_dcnn=score_dc.numpy(),
       score_cheby=score_ch.numpy(), weight_seed=np.array(weight_seed))


def golden_general_synth(num_graphs=16, weight_seed=4321):
  cfg = configs.graph_lanczos_net()
  import networkx as nx
  # dataset/get_graph_data.py:15-49 verbatim recipe (nx.to_numpy_matrix -> to_numpy_array)
  npr = np.random.RandomState(123)
  N = npr.randint(20, high=101, size=num_graphs)
  samples = []
  for ii in range(num_graphs):
    X = npr.randn(N[ii], 10)
    A = np.expand_dims(np.asarray(nx.to_numpy_array(
        nx.fast_gnp_random_graph(int(N[ii]), 0.5, seed=int(npr.randint(1000))))), axis=2)
    Y = npr.randn(1, 2)
    samples.append(ref_prepare(A, X, Y))
  batch = ref_collate(ref_gd.GraphData, cfg, samples)
  # reference GraphData.collate pads node_feat as float64 -> the runner feeds .float()
  node_feat = batch['node_feat'].float()
  model = LanczosNetGeneral(cfg)
  model.load_state_dict(deterministic_state_dict(model, weight_seed))
  model.eval()
  with torch.no_grad():
    score = model(node_feat, batch['L'], batch['D'], batch['V'], mask=batch['node_mask'])
  save('lanczosnet_general_synth.npz', sizes=N, node_feat=node_feat.numpy(),
       node_mask=batch['node_mask'].numpy(), L=batch['L'].numpy(), D=batch['D'].numpy(),
       V=batch['V'].numpy(), score=score.numpy(), weight_seed=np.array(weight_seed))


def _qm8_like_operator(rng, sizes, N):
  """Padded L4 operators of random molecules + mask."""
  B = len(sizes)
  A = np.zeros((B, N, N), np.float32)
  mask = np.zeros((B, N), np.uint8)
  for b, n in enumerate(sizes):
    _, adjs = data.synthetic_molecule(rng, n)
    A[b, :n, :n] = ref_dh.get_laplacian(adjs.sum(axis=2), graph_laplacian_type='L4')
    mask[b, :n] = 1
  return A, mask


def golden_lanczos_layer():
  """AdaLanczosNet._lanczos_layer (model/ada_lanczos_net.py:139-247) on several regimes."""
  out = {}
  rng = np.random.RandomState(11)
  cases = {
      # name: (sizes, N, K)
      'qm8': ([26, 25, 22, 21, 20, 19, 12, 8, 5, 3, 16, 24], 26, 20),   # n_b >, ==, < K
      'small': ([12, 10, 7, 4], 12, 20),                                  # N < K -> zero pad
      'nomask': ([18] * 4, 18, 10),                                       # mask=None branch
      'cta64': ([64, 50, 33, 64], 64, 20),                                # CTA-per-graph kernel
      'cta100': ([100, 77], 100, 40),
  }
  for name, (sizes, N, K) in cases.items():
    A, mask = _qm8_like_operator(rng, sizes, N)
    if name.startswith('cta'):
      # denser G(n,p) operators for the larger graphs
      import networkx as nx
      for b, n in enumerate(sizes):
        g = nx.fast_gnp_random_graph(n, min(0.5, 8.0 / n), seed=int(rng.randint(1000)))
        A[b] = 0
        A[b, :n, :n] = ref_dh.get_laplacian(np.asarray(nx.to_numpy_array(g)),
                                            graph_laplacian_type='L4')
    stub = types.SimpleNamespace(num_eig_vec=K, use_reorthogonalization=True)
    seed = 100 + len(name)
    torch.manual_seed(seed)
    q1 = torch.randn(len(sizes), N, 1)
    torch.manual_seed(seed)
    At = torch.from_numpy(A)
    mt = None if name == 'nomask' else torch.from_numpy(mask)
    with torch.no_grad():
      T, Q = AdaLanczosNet._lanczos_layer(stub, At, mt)
    out[name + '_A'] = A
    out[name + '_mask'] = mask
    out[name + '_q1'] = q1.numpy()[:, :, 0]
    out[name + '_K'] = np.array(K)
    out[name + '_T'] = T.numpy()
    out[name + '_Q'] = Q.numpy()
  save('ada_lanczos_layer.npz', **out)


def golden_ada_forward(weight_seed=999):
  """AdaLanczosNet end to end on a reduced config (the class hard-codes the 4096-wide MLP,
  so the layer count / K / scales are reduced to keep weight regeneration cheap)."""
  cfg = configs.qm8_ada_lanczos_net(num_layer=2, hidden_dim=[32, 32], num_eig_vec=8,
                                    long_diffusion_dist=[2, 5], short_diffusion_dist=[1, 3])
  rng = np.random.RandomState(5)
  sizes = [14, 9, 12, 5, 11, 13]
  samples = []
  for n in sizes:
    nf, adjs = data.synthetic_molecule(rng, n)
    samples.append(ref_prepare(adjs, nf, rng.randn(1, 16)))
  cfg_l = configs.qm8_lanczos_net(num_eig_vec=8)
  batch = ref_collate(ref_qm8.QM8Data, cfg_l, samples)
  model = AdaLanczosNet(cfg)
  model.load_state_dict(deterministic_state_dict(model, weight_seed))
  model.eval()
  seed = 77
  torch.manual_seed(seed)
  q1 = torch.randn(len(sizes), max(sizes), 1)
  torch.manual_seed(seed)
  with torch.no_grad():
    state = model.embedding(batch['node_feat'])
    adj = torch.zeros_like(batch['L'][:, :, :, 0])
    adj[batch['L'][:, :, :, 0] != 0.0] = 1.0
    Le = model._get_graph_laplacian(state, adj)
    torch.manual_seed(seed)
    score = model(batch['node_feat'], batch['L'], mask=batch['node_mask'])
  save('ada_forward_small.npz', sizes=np.array(sizes), node_feat=batch['node_feat'].numpy(),
       node_mask=batch['node_mask'].numpy(), L=batch['L'].numpy(), q1=q1.numpy()[:, :, 0],
       Le=Le.numpy(), score=score.numpy(), weight_seed=np.array(weight_seed),
       torch_seed=np.array(seed))


def golden_train_grads():
  """Gradients of the REFERENCE classes' own autograd (runner/qm8_runner.py:226-247: train mode, MSE
  loss, ``train_loss.backward()``) on the inputs of the forward goldens: pins the oracle's differentiable
  use (tests/test_oracle_golden.py), which in turn is what the GPU training tests are checked against.
  Only the loss and a digest per parameter (sum, sum of squares, the first 8 entries) are stored."""
  out = {}

  def digest(prefix, model, loss):
    out[prefix + '_loss'] = np.array(float(loss))
    for name, p in model.named_parameters():
      g = p.grad.detach().double().numpy().reshape(-1)
      out['%s|%s' % (prefix, name)] = np.concatenate([[g.sum(), (g * g).sum()], g[:8]])

  g = np.load(os.path.join(HERE, 'lanczosnet_qm8.npz'))
  nf, L, D, V = [torch.from_numpy(g[k]) for k in ('node_feat', 'L', 'D', 'V')]
  label, mask = torch.from_numpy(g['label']), torch.from_numpy(g['node_mask'])
  cfg = configs.qm8_lanczos_net(num_layer=2, hidden_dim=[64, 64])
  model = LanczosNet(cfg)
  model.load_state_dict(deterministic_state_dict(model, 11))
  model.train()
  _, loss = model(nf, L, D, V, label=label, mask=mask)
  loss.backward()
  digest('lanczosnet', model, loss)

  gcn = GCN(configs.qm8_gcn(num_layer=2, hidden_dim=[64, 64]))
  gcn.load_state_dict(deterministic_state_dict(gcn, 9))
  gcn.train()
  _, loss = gcn(nf, L, label=label, mask=mask)
  loss.backward()
  digest('gcn', gcn, loss)

  a = np.load(os.path.join(HERE, 'ada_forward_small.npz'))
  cfg_a = configs.qm8_ada_lanczos_net(num_layer=2, hidden_dim=[32, 32], num_eig_vec=8,
                                      long_diffusion_dist=[2, 5], short_diffusion_dist=[1, 3])
  ada = AdaLanczosNet(cfg_a)
  ada.load_state_dict(deterministic_state_dict(ada, int(a['weight_seed'])))
  ada.train()
  lab = torch.from_numpy(np.random.RandomState(0).randn(a['score'].shape[0], a['score'].shape[1]).astype(np.float32))
  torch.manual_seed(int(a['torch_seed']))
  _, loss = ada(torch.from_numpy(a['node_feat']), torch.from_numpy(a['L']), label=lab,
                mask=torch.from_numpy(a['node_mask']))
  loss.backward()
  digest('ada', ada, loss)

  # the operator-chain models and the general (feature-input) model, configs of tests/test_gpu_train.py
  dc = DCNN(configs.qm8_dcnn(num_layer=2, hidden_dim=[32, 32], diffusion_dist=[2, 5]))
  dc.load_state_dict(deterministic_state_dict(dc, 3))
  dc.train()
  _, loss = dc(nf, L, label=label, mask=mask)
  loss.backward()
  digest('dcnn', dc, loss)
  ch = ChebyNet(configs.qm8_cheby_net(num_layer=2, hidden_dim=[32, 32], polynomial_order=4))
  ch.load_state_dict(deterministic_state_dict(ch, 3))
  ch.train()
  _, loss = ch(nf, L, label=label, mask=mask)
  loss.backward()
  digest('cheby', ch, loss)
  gg = np.load(os.path.join(HERE, 'lanczosnet_general_synth.npz'))
  gen = LanczosNetGeneral(configs.graph_lanczos_net())
  gen.load_state_dict(deterministic_state_dict(gen, 5))
  gen.train()
  _, loss = gen(torch.from_numpy(gg['node_feat']), torch.from_numpy(gg['L']), torch.from_numpy(gg['D']),
                torch.from_numpy(gg['V']), label=torch.zeros(gg['score'].shape), mask=torch.from_numpy(gg['node_mask']))
  loss.backward()
  digest('general', gen, loss)
  save('train_grads.npz', **out)


if __name__ == '__main__':
  if 'grads-only' in sys.argv:                     # needs the forward goldens on disk
    golden_train_grads()
    sys.exit(0)
  golden_data_helper()
  golden_lanczosnet_qm8()
  golden_gcn_qm8()
  golden_general_synth()
  golden_lanczos_layer()
  golden_ada_forward()
  golden_train_grads()
