"""Pin the oracle against outputs of the REFERENCE itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import deterministic_state_dict, load_golden, oracle_spec
from lanczosnetwork_b200 import configs
from lanczosnetwork_b200.model import (AdaLanczosNet, ChebyNet, DCNN, GCN, GCNFP, LanczosNet,
                                       LanczosNetGeneral)
from oracle import graph_prep
from oracle import lanczos_oracle as orc


def test_data_helper_fixture():
  g = load_golden('data_helper_fixture.npz')
  L4 = graph_prep.laplacian_L4(g['adj'])
  np.testing.assert_allclose(L4, g['L4'], rtol=0, atol=1e-15)
  # the survey's hand-checked known answers
  np.testing.assert_allclose(L4[0], [1 / 3, 0.288675, 0, 0, 0.288675, 0], atol=1e-6)
  D, V = graph_prep.eig_topk_by_magnitude(L4, k=100)
  np.testing.assert_allclose(D, g['D'], atol=1e-12)
  np.testing.assert_allclose(D, [1.0, 0.72039, 0.427016, -0.316134, 0.146376, -0.060981], atol=1e-6)
  # eigenvectors up to sign
  np.testing.assert_allclose(np.abs(V), np.abs(g['V']), atol=1e-10)
  D3, V3 = graph_prep.eig_topk_by_magnitude(L4, k=3)
  np.testing.assert_allclose(D3, g['D3'], atol=1e-12)


def test_collate_matches_reference():
  g = load_golden('lanczosnet_qm8.npz')
  samples = []
  for b, n in enumerate(g['sizes']):
    rec = graph_prep.prepare_molecule(g['adjs'][b, :n, :n].astype(np.float64))
    rec['node_feat'] = g['node_feat'][b, :n]
    samples.append(rec)
  out = graph_prep.collate(samples, 20)
  assert np.array_equal(out['node_feat'], g['node_feat'])
  assert np.array_equal(out['node_mask'], g['node_mask'])
  assert np.array_equal(out['L'], g['L'])          # bit-exact operator construction
  np.testing.assert_allclose(out['D'], g['D'], atol=1e-6)
  # Ritz vectors: compare the sign/rotation-invariant filter V diag(D) V^T
  rec_o = np.einsum('bnk,bk,bmk->bnm', out['V'], out['D'], out['V'])
  rec_g = np.einsum('bnk,bk,bmk->bnm', g['V'], g['D'], g['V'])
  np.testing.assert_allclose(rec_o, rec_g, atol=2e-6)


def test_lanczosnet_forward_matches_reference():
  g = load_golden('lanczosnet_qm8.npz')
  cfg = configs.qm8_lanczos_net()
  mod = LanczosNet(cfg)
  params = deterministic_state_dict(mod, int(g['weight_seed']))
  spec = oracle_spec(mod, 'LanczosNet')
  score = orc.lanczos_net_forward(params, spec, g['node_feat'], g['L'], g['D'], g['V'],
                                  g['node_mask'])
  np.testing.assert_allclose(score.numpy(), g['score'], rtol=1e-4, atol=2e-6)
  p32 = orc._cast(params, torch.float32)
  Lf0 = orc.spectral_filters_from_ritz(p32, spec, torch.from_numpy(g['D']),
                                       torch.from_numpy(g['V']), 0)
  np.testing.assert_allclose(Lf0.numpy(), g['Lf0'], rtol=1e-4, atol=1e-6)
  # fp64 oracle: the reference fp32 result sits within fp32 rounding of it
  s64 = orc.lanczos_net_forward(params, spec, g['node_feat'], g['L'], g['D'], g['V'],
                                g['node_mask'], dtype=torch.float64)
  assert np.abs(s64.numpy() - g['score']).max() < 2e-5


def test_gcn_forward_matches_reference():
  """the sibling GCN (model/gcn.py) on the inputs of the LanczosNet fixture."""
  g, gg = load_golden('lanczosnet_qm8.npz'), load_golden('gcn_qm8.npz')
  mod = GCN(configs.qm8_gcn())
  params = deterministic_state_dict(mod, int(gg['weight_seed']))
  spec = oracle_spec(mod, 'GCN')
  score = orc.gcn_forward(params, spec, g['node_feat'], g['L'], g['node_mask'])
  np.testing.assert_allclose(score.numpy(), gg['score'], rtol=1e-4, atol=2e-6)
  nomask = orc.gcn_forward(params, spec, g['node_feat'], g['L'], None)
  np.testing.assert_allclose(nomask.numpy(), gg['score_nomask'], rtol=1e-4, atol=2e-6)
  s64 = orc.gcn_forward(params, spec, g['node_feat'], g['L'], g['node_mask'], dtype=torch.float64)
  assert np.abs(s64.numpy() - gg['score']).max() < 2e-5
  # GCNFP: the same layer on the non-zero pattern of the operators
  mod_fp = GCNFP(configs.qm8_gcn(name='GCNFP'))
  params_fp = deterministic_state_dict(mod_fp, int(gg['weight_seed']) + 1)
  fp = orc.gcn_forward(params_fp, oracle_spec(mod_fp, 'GCNFP'), g['node_feat'], g['L'], g['node_mask'],
                       binarize=True)
  np.testing.assert_allclose(fp.numpy(), gg['score_fp'], rtol=1e-4, atol=1e-5)
  # DCNN: edge types + powers of the simple-graph operator
  cfg = configs.qm8_dcnn()
  mod_dc = DCNN(cfg)
  params_dc = deterministic_state_dict(mod_dc, int(gg['weight_seed']) + 2)
  dc = orc.dcnn_forward(params_dc, cfg.model.diffusion_dist, cfg.dataset.num_bond_type,
                        cfg.model.num_layer, g['node_feat'], g['L'], g['node_mask'])
  np.testing.assert_allclose(dc.numpy(), gg['score_dcnn'], rtol=1e-4, atol=2e-6)
  # ChebyNet: Chebyshev chain on channel 0 + bond-type channels
  cfg = configs.qm8_cheby_net()
  mod_ch = ChebyNet(cfg)
  params_ch = deterministic_state_dict(mod_ch, int(gg['weight_seed']) + 3)
  ch = orc.cheby_net_forward(params_ch, cfg.model.polynomial_order, cfg.dataset.num_bond_type,
                             cfg.model.num_layer, g['node_feat'], g['L'], g['node_mask'])
  np.testing.assert_allclose(ch.numpy(), gg['score_cheby'], rtol=1e-4, atol=2e-6)


def test_lanczosnet_power_filter_matches_reference():
  g = load_golden('lanczosnet_qm8.npz')
  cfg = configs.qm8_lanczos_net(spectral_filter_kind='power', num_layer=2, hidden_dim=[32, 32])
  mod = LanczosNet(cfg)
  params = deterministic_state_dict(mod, int(g['weight_seed']) + 100)
  spec = oracle_spec(mod, 'LanczosNet')
  score = orc.lanczos_net_forward(params, spec, g['node_feat'], g['L'], g['D'], g['V'],
                                  g['node_mask'])
  np.testing.assert_allclose(score.numpy(), g['score_power'], rtol=1e-4, atol=2e-6)


def test_general_forward_matches_reference():
  g = load_golden('lanczosnet_general_synth.npz')
  cfg = configs.graph_lanczos_net()
  mod = LanczosNetGeneral(cfg)
  params = deterministic_state_dict(mod, int(g['weight_seed']))
  spec = oracle_spec(mod, 'LanczosNetGeneral')
  score = orc.lanczos_net_forward(params, spec, g['node_feat'], g['L'], g['D'], g['V'],
                                  g['node_mask'])
  np.testing.assert_allclose(score.numpy(), g['score'], rtol=1e-4, atol=5e-6)


@pytest.mark.parametrize('case', ['qm8', 'small', 'nomask', 'cta64', 'cta100'])
def test_lanczos_layer_matches_reference(case):
  g = load_golden('ada_lanczos_layer.npz')
  A = torch.from_numpy(g[case + '_A'])
  mask = None if case == 'nomask' else torch.from_numpy(g[case + '_mask'])
  q1 = torch.from_numpy(g[case + '_q1'])
  K = int(g[case + '_K'])
  out = orc.lanczos_tridiagonalise(A, mask, q1, K)
  T_ref, Q_ref = g[case + '_T'], g[case + '_Q']
  # integer structure is exact: which Krylov directions / node rows survive
  assert np.array_equal(out['T'].numpy() != 0, T_ref != 0)
  assert np.array_equal(out['Q'].numpy() != 0, Q_ref != 0)
  np.testing.assert_allclose(out['T'].numpy(), T_ref, atol=5e-5)
  np.testing.assert_allclose(out['Q'].numpy(), Q_ref, atol=2e-3)


def test_ada_forward_matches_reference():
  g = load_golden('ada_forward_small.npz')
  cfg = configs.qm8_ada_lanczos_net(num_layer=2, hidden_dim=[32, 32], num_eig_vec=8,
                                    long_diffusion_dist=[2, 5], short_diffusion_dist=[1, 3])
  mod = AdaLanczosNet(cfg)
  params = deterministic_state_dict(mod, int(g['weight_seed']))
  spec = oracle_spec(mod, 'AdaLanczosNet')
  score, aux = orc.ada_lanczos_net_forward(params, spec, g['node_feat'], g['L'], g['node_mask'],
                                           g['q1'], return_aux=True)
  np.testing.assert_allclose(aux['Le'].numpy(), g['Le'], rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(score.numpy(), g['score'], rtol=1e-3, atol=2e-5)


def test_tridiag_ritz_oracle_reconstructs_T():
  g = load_golden('ada_lanczos_layer.npz')
  T = g['qm8_T']
  alpha = np.diagonal(T, axis1=1, axis2=2)
  beta = np.diagonal(T, offset=1, axis1=1, axis2=2)
  theta, S = orc.tridiag_ritz(alpha, beta)
  rec = np.einsum('bik,bk,bjk->bij', S, theta, S)
  np.testing.assert_allclose(rec, T, atol=1e-12)
  assert np.all(np.diff(np.abs(theta), axis=1) <= 1e-15)


def _oracle_grad_digests(forward, params, monkeypatch):
  p64 = {k: v.detach().double().requires_grad_(v.is_floating_point()) for k, v in params.items()}
  monkeypatch.setattr(orc, '_cast', lambda p, dtype: p)      # the oracle detaches its parameters; keep the tape
  loss = forward(p64)
  loss.backward()
  return float(loss.detach()), {k: v.grad.numpy().reshape(-1) for k, v in p64.items() if v.grad is not None}


_WORST = {}


def _check_digests(prefix, g, loss, grads, rel):
  """tests/golden/train_grads.npz holds, per parameter of the reference's own backward, [sum, sum of
  squares, first 8 entries] of the gradient; ``rel`` is relative to the gradient's largest entry."""
  assert abs(loss - float(g[prefix + '_loss'])) <= 1e-5 * max(1.0, abs(loss))
  seen = 0
  for key in g:
    if not key.startswith(prefix + '|'):
      continue
    name = key.split('|', 1)[1]
    ours, ref = grads[name], g[key]
    scale = float(np.abs(ours).max()) + 1e-12
    assert np.abs(ours[:8] - ref[2:2 + min(8, ours.size)]).max() <= rel * scale, name
    _WORST[prefix] = max(_WORST.get(prefix, 0.0), float(np.abs(ours[:8] - ref[2:2 + min(8, ours.size)]).max() / scale))
    assert abs(ours.sum() - ref[0]) <= rel * scale * np.sqrt(ours.size) + 1e-12, name
    assert abs((ours * ours).sum() - ref[1]) <= 4 * rel * max(ref[1], 1e-30), name
    seen += 1
  assert seen == len(grads) and seen > 0


def test_oracle_autograd_matches_the_references_own_backward(monkeypatch):
  """SURVEY 8(f1) pin: the GPU training tests compare against autograd over the oracle; here autograd
  over the oracle (fp64) is compared with ``loss.backward()`` of the REFERENCE classes (fp32, train mode,
  runner/qm8_runner.py:226-247) on the same inputs and weights."""
  g = load_golden('train_grads.npz')
  q = load_golden('lanczosnet_qm8.npz')
  label = torch.from_numpy(q['label']).double()

  mod = LanczosNet(configs.qm8_lanczos_net(num_layer=2, hidden_dim=[64, 64]))
  params = deterministic_state_dict(mod, 11)
  spec = oracle_spec(mod, 'LanczosNet')
  loss, grads = _oracle_grad_digests(
      lambda p: torch.nn.functional.mse_loss(
          orc.lanczos_net_forward(p, spec, q['node_feat'], q['L'], q['D'], q['V'], q['node_mask'],
                                  dtype=torch.float64), label), params, monkeypatch)
  _check_digests('lanczosnet', g, loss, grads, 2e-5)        # measured 2.4e-7

  gcn = GCN(configs.qm8_gcn(num_layer=2, hidden_dim=[64, 64]))
  gp = deterministic_state_dict(gcn, 9)
  gspec = oracle_spec(gcn, 'GCN')
  loss, grads = _oracle_grad_digests(
      lambda p: torch.nn.functional.mse_loss(
          orc.gcn_forward(p, gspec, q['node_feat'], q['L'], q['node_mask'], dtype=torch.float64), label),
      gp, monkeypatch)
  _check_digests('gcn', g, loss, grads, 2e-5)               # measured 1.1e-7

  a = load_golden('ada_forward_small.npz')
  cfg = configs.qm8_ada_lanczos_net(num_layer=2, hidden_dim=[32, 32], num_eig_vec=8,
                                    long_diffusion_dist=[2, 5], short_diffusion_dist=[1, 3])
  ada = AdaLanczosNet(cfg)
  ap = deterministic_state_dict(ada, int(a['weight_seed']))
  aspec = oracle_spec(ada, 'AdaLanczosNet')
  lab = torch.from_numpy(np.random.RandomState(0).randn(*a['score'].shape).astype(np.float32)).double()
  loss, grads = _oracle_grad_digests(
      lambda p: torch.nn.functional.mse_loss(
          orc.ada_lanczos_net_forward(p, aspec, a['node_feat'], a['L'], a['node_mask'],
                                      torch.from_numpy(a['q1']).double(), dtype=torch.float64), lab),
      ap, monkeypatch)
  _check_digests('ada', g, loss, grads, 5e-4)               # measured 5.6e-6 (the reference ran its Lanczos recurrence in fp32)
  dc_cfg = configs.qm8_dcnn(num_layer=2, hidden_dim=[32, 32], diffusion_dist=[2, 5])
  dp = deterministic_state_dict(DCNN(dc_cfg), 3)
  loss, grads = _oracle_grad_digests(
      lambda p: torch.nn.functional.mse_loss(
          orc.dcnn_forward(p, dc_cfg.model.diffusion_dist, 6, 2, q['node_feat'], q['L'], q['node_mask'],
                           dtype=torch.float64), label), dp, monkeypatch)
  _check_digests('dcnn', g, loss, grads, 2e-5)
  cp = deterministic_state_dict(ChebyNet(configs.qm8_cheby_net(num_layer=2, hidden_dim=[32, 32], polynomial_order=4)), 3)
  loss, grads = _oracle_grad_digests(
      lambda p: torch.nn.functional.mse_loss(
          orc.cheby_net_forward(p, 4, 6, 2, q['node_feat'], q['L'], q['node_mask'], dtype=torch.float64), label),
      cp, monkeypatch)
  _check_digests('cheby', g, loss, grads, 2e-5)
  gg = load_golden('lanczosnet_general_synth.npz')
  gen = LanczosNetGeneral(configs.graph_lanczos_net())
  gnp = deterministic_state_dict(gen, 5)
  gen_spec = oracle_spec(gen, 'LanczosNetGeneral')
  loss, grads = _oracle_grad_digests(
      lambda p: torch.nn.functional.mse_loss(
          orc.lanczos_net_forward(p, gen_spec, gg['node_feat'], gg['L'], gg['D'], gg['V'], gg['node_mask'],
                                  dtype=torch.float64), torch.zeros(gg['score'].shape, dtype=torch.float64)),
      gnp, monkeypatch)
  _check_digests('general', g, loss, grads, 2e-5)
  print('worst relative gradient-entry error vs the reference backward:', _WORST)
