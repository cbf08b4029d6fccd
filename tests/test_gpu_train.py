"""Training path: gradients of the drop-in modules against torch.autograd over the
fp64 CPU oracle, and the loop body of the reference's QM8Runner.train (runner/qm8_runner.py:188-259).
``pytest -m gpu``."""
import numpy as np
import pytest
import torch

from helpers import deterministic_state_dict, load_golden, oracle_spec
from lanczosnetwork_b200 import configs, data
from lanczosnetwork_b200.model import AdaLanczosNet, ChebyNet, DCNN, GCN, LanczosNet, LanczosNetGeneral
from oracle import lanczos_oracle as orc

pytestmark = pytest.mark.gpu


def dev():
  return torch.device('cuda:0')


def _t(a):
  return torch.from_numpy(np.ascontiguousarray(a))


def _oracle_grads(forward, params, monkeypatch):
  """d loss / d params by autograd over the oracle in fp64 (its _cast detaches: bypassed here)."""
  p64 = {k: v.detach().double().requires_grad_(v.is_floating_point()) for k, v in params.items()}
  monkeypatch.setattr(orc, '_cast', lambda p, dtype: p)
  loss = forward(p64)
  loss.backward()
  return float(loss.detach()), {k: v.grad for k, v in p64.items() if v.grad is not None}


def _compare(mod, grads_ref, rel=2e-3):
  worst = 0.0
  for name, p in mod.named_parameters():
    assert p.grad is not None, name
    g, r = p.grad.detach().cpu().double(), grads_ref[name]
    scale = float(r.abs().max()) + 1e-12
    err = float((g - r).abs().max()) / scale
    worst = max(worst, err)
    assert err <= rel, (name, err, scale)
  return worst


def test_lanczosnet_gradients_match_fp64_oracle_autograd(monkeypatch):
  """Every parameter gradient of a 2-layer LanczosNet (embedding, filter MLPs, conv Linears, head,
  gate) through the library's kernels (tcgen05 dense + strided batched GEMM + segment-sum scatter)
  against autograd over the fp64 oracle: relative to the largest entry of each gradient, 2e-3
  (3xTF32 dense products, fp32 accumulation)."""
  g = load_golden('lanczosnet_qm8.npz')
  cfg = configs.qm8_lanczos_net(num_layer=2, hidden_dim=[64, 64])
  mod = LanczosNet(cfg)
  params = deterministic_state_dict(mod, 11)
  mod.load_state_dict(params)
  mod = mod.to(dev()).train()
  label = _t(g['label'])
  spec = oracle_spec(mod, 'LanczosNet')

  def fwd(p64):
    s = orc.lanczos_net_forward(p64, spec, g['node_feat'], g['L'], g['D'], g['V'], g['node_mask'],
                                dtype=torch.float64)
    return torch.nn.functional.mse_loss(s, label.double())

  loss_ref, grads_ref = _oracle_grads(fwd, params, monkeypatch)
  score, loss = mod(_t(g['node_feat']).to(dev()), _t(g['L']).to(dev()), _t(g['D']).to(dev()),
                    _t(g['V']).to(dev()), label=label.to(dev()), mask=_t(g['node_mask']).to(dev()))
  assert score.requires_grad and abs(float(loss.detach()) - loss_ref) <= 1e-5 * max(1.0, abs(loss_ref))
  loss.backward()
  _compare(mod, grads_ref)
  # the differentiable forward agrees with the fused inference kernels
  mod.eval()
  with torch.no_grad():
    fused = mod(_t(g['node_feat']).to(dev()), _t(g['L']).to(dev()), _t(g['D']).to(dev()),
                _t(g['V']).to(dev()), mask=_t(g['node_mask']).to(dev()))
  np.testing.assert_allclose(score.detach().cpu().numpy(), fused.cpu().numpy(), rtol=1e-4, atol=2e-5)


def test_general_and_gcn_gradients_match_oracle(monkeypatch):
  gg = load_golden('lanczosnet_general_synth.npz')
  cfg = configs.graph_lanczos_net()
  mod = LanczosNetGeneral(cfg)
  params = deterministic_state_dict(mod, 5)
  mod.load_state_dict(params)
  mod = mod.to(dev()).train()
  spec = oracle_spec(mod, 'LanczosNetGeneral')
  label = _t(gg['label']) if 'label' in gg else torch.zeros(gg['score'].shape)

  def fwd(p64):
    s = orc.lanczos_net_forward(p64, spec, gg['node_feat'], gg['L'], gg['D'], gg['V'], gg['node_mask'],
                                dtype=torch.float64)
    return torch.nn.functional.mse_loss(s, label.double())

  _, grads_ref = _oracle_grads(fwd, params, monkeypatch)
  _, loss = mod(_t(gg['node_feat']).to(dev()), _t(gg['L']).to(dev()), _t(gg['D']).to(dev()),
                _t(gg['V']).to(dev()), label=label.to(dev()), mask=_t(gg['node_mask']).to(dev()))
  loss.backward()
  _compare(mod, grads_ref)

  g = load_golden('lanczosnet_qm8.npz')
  gcn = GCN(configs.qm8_gcn(num_layer=2, hidden_dim=[64, 64]))
  gp = deterministic_state_dict(gcn, 9)
  gcn.load_state_dict(gp)
  gcn = gcn.to(dev()).train()
  gspec = oracle_spec(gcn, 'GCN')

  def fwd_gcn(p64):
    s = orc.gcn_forward(p64, gspec, g['node_feat'], g['L'], g['node_mask'], dtype=torch.float64)
    return torch.nn.functional.mse_loss(s, _t(g['label']).double())

  _, gref = _oracle_grads(fwd_gcn, gp, monkeypatch)
  _, loss = gcn(_t(g['node_feat']).to(dev()), _t(g['L']).to(dev()), label=_t(g['label']).to(dev()),
                mask=_t(g['node_mask']).to(dev()))
  loss.backward()
  _compare(gcn, gref)


def test_reference_training_loop_body_runs_and_learns():
  """The loop body of QM8Runner.train (runner/qm8_runner.py:226-259): nn.DataParallel(model).cuda(),
  Adam(lr), model.train(), ``_, train_loss = model(..., label=, mask=)``, ``train_loss.backward()``,
  ``optimizer.step()`` -- on a fixed batch the loss goes down, the CUDA-graph inference forward picks
  up the updated weights, and eval under no_grad still uses the fused kernels."""
  batch = data.synthetic_qm8_batch(64, seed=4)
  model = LanczosNet(configs.qm8_lanczos_net())
  model.load_state_dict(deterministic_state_dict(model, 1234))
  model = torch.nn.DataParallel(model, device_ids=[0]).cuda()
  params = filter(lambda p: p.requires_grad, model.parameters())
  optimizer = torch.optim.Adam(params, lr=1.0e-3, weight_decay=0.0)
  t = {k: _t(v).cuda() for k, v in batch.items()}
  model.eval()
  with torch.no_grad():
    before = model(t['node_feat'], t['L'], t['D'], t['V'], label=t['label'], mask=t['node_mask'])[1]
  losses = []
  for it in range(25):
    model.train()
    optimizer.zero_grad()
    _, train_loss = model(t['node_feat'], t['L'], t['D'], t['V'], label=t['label'], mask=t['node_mask'])
    train_loss.backward()
    optimizer.step()
    losses.append(float(train_loss))
  assert abs(losses[0] - float(before)) <= 1e-4 * max(1.0, float(before))
  assert max(losses[-3:]) < 0.95 * losses[0], losses
  model.eval()
  with torch.no_grad():
    after = model(t['node_feat'], t['L'], t['D'], t['V'], label=t['label'], mask=t['node_mask'])[1]
    again = model(t['node_feat'], t['L'], t['D'], t['V'], label=t['label'], mask=t['node_mask'])[1]
  assert float(after) < losses[0] and float(after) == float(again)


def test_dcnn_and_cheby_gradients_match_oracle(monkeypatch):
  """The operator-chain models train through the same adjoint kernels (L_0^T g chains)."""
  g = load_golden('lanczosnet_qm8.npz')
  label = _t(g['label'])
  dc_cfg = configs.qm8_dcnn(num_layer=2, hidden_dim=[32, 32], diffusion_dist=[2, 5])
  ch_cfg = configs.qm8_cheby_net(num_layer=2, hidden_dim=[32, 32], polynomial_order=4)
  cases = (
      (DCNN, dc_cfg, lambda p: orc.dcnn_forward(p, dc_cfg.model.diffusion_dist, 6, 2, g['node_feat'], g['L'],
                                                g['node_mask'], dtype=torch.float64)),
      (ChebyNet, ch_cfg, lambda p: orc.cheby_net_forward(p, 4, 6, 2, g['node_feat'], g['L'], g['node_mask'],
                                                         dtype=torch.float64)))
  for cls, cfg, fwd in cases:
    mod = cls(cfg)
    params = deterministic_state_dict(mod, 3)
    mod.load_state_dict(params)
    mod = mod.to(dev()).train()
    _, grads_ref = _oracle_grads(lambda p: torch.nn.functional.mse_loss(fwd(p), label.double()), params, monkeypatch)
    _, loss = mod(_t(g['node_feat']).to(dev()), _t(g['L']).to(dev()), label=label.to(dev()),
                  mask=_t(g['node_mask']).to(dev()))
    loss.backward()
    _compare(mod, grads_ref)


def test_ada_lanczos_net_gradients_match_oracle(monkeypatch):
  """AdaLanczosNet end to end on the tape: embedding -> learned Gaussian Laplacian -> K-step Lanczos
  -> learned filter on the powers of T -> graph convolutions -> readout.  Gradients against autograd
  over the fp64 oracle (which runs the reference's sequential Gram-Schmidt): the Lanczos recurrence
  amplifies rounding, so the bound is 2e-2 of each gradient's largest entry; the forward agrees with
  the fused inference path."""
  g = load_golden('ada_forward_small.npz')
  cfg = configs.qm8_ada_lanczos_net(num_layer=2, hidden_dim=[32, 32], num_eig_vec=8,
                                    long_diffusion_dist=[2, 5], short_diffusion_dist=[1, 3])
  mod = AdaLanczosNet(cfg)
  params = deterministic_state_dict(mod, int(g['weight_seed']))
  mod.load_state_dict(params)
  mod = mod.to(dev()).train()
  spec = oracle_spec(mod, 'AdaLanczosNet')
  B, N = g['node_feat'].shape
  torch.manual_seed(int(g['torch_seed']))
  q1 = torch.randn(B, N, 1)
  label = torch.from_numpy(np.random.RandomState(0).randn(B, g['score'].shape[1]).astype(np.float32))

  def fwd(p64):
    s = orc.ada_lanczos_net_forward(p64, spec, g['node_feat'], g['L'], g['node_mask'], q1[:, :, 0].double(),
                                    dtype=torch.float64)
    return torch.nn.functional.mse_loss(s, label.double())

  _, grads_ref = _oracle_grads(fwd, params, monkeypatch)
  torch.manual_seed(int(g['torch_seed']))
  score, loss = mod(_t(g['node_feat']).to(dev()), _t(g['L']).to(dev()), label=label.to(dev()),
                    mask=_t(g['node_mask']).to(dev()))
  loss.backward()
  _compare(mod, grads_ref, rel=2e-2)
  np.testing.assert_allclose(score.detach().cpu().numpy(), g['score'], rtol=1e-3, atol=5e-5)


@pytest.mark.parametrize('opt_name', ['sgd', 'adam'])
def test_graphed_training_step_matches_eager_steps(opt_name):
  """The captured step (forward + loss + backward + optimizer in one CUDA graph) walks the same
  trajectory as the eager loop body of the reference's runner over 6 steps / 3 rotating batches, and
  building the object does not advance training.  Losses agree to 1e-5 for both optimizers; the
  weights are compared under momentum SGD (linear in the gradient) -- Adam's m / sqrt(v) turns the
  float reordering of the embedding-gradient atomics into +-lr moves where a gradient is ~ 0."""
  from lanczosnetwork_b200 import data
  from lanczosnetwork_b200.train import GraphedStep
  cfg = configs.qm8_lanczos_net(num_layer=3, hidden_dim=[64, 64, 64])
  batches = []
  for i in range(3):
    b = data.collate(data.synthetic_qm8_samples(32, seed=50 + i), 20, num_nodes=27)
    b['label'] = np.random.RandomState(i).randn(32, 16).astype(np.float32)
    batches.append({k: torch.from_numpy(v).to(dev()) for k, v in b.items() if isinstance(v, np.ndarray)})

  def make():
    m = LanczosNet(cfg)
    m.load_state_dict(deterministic_state_dict(m, 77))
    m = m.to(dev()).train()
    if opt_name == 'adam':
      return m, torch.optim.Adam(m.parameters(), lr=1e-3)
    return m, torch.optim.SGD(m.parameters(), lr=1e-2, momentum=0.9)

  def call_args(b):
    return (b['node_feat'], b['L'], b['D'], b['V']), {'label': b['label'], 'mask': b['node_mask']}

  eager, opt_e = make()
  losses_e = []
  for i in range(6):
    a, kw = call_args(batches[i % 3])
    opt_e.zero_grad()
    _, loss = eager(*a, **kw)
    loss.backward()
    opt_e.step()
    losses_e.append(float(loss.detach()))

  graphed, opt_g = make()
  a, kw = call_args(batches[0])
  step = GraphedStep(graphed, opt_g, a, kw)
  for (n, p), (_, q) in zip(graphed.named_parameters(), make()[0].named_parameters()):
    assert torch.equal(p, q), n                              # warm-up rolled back
  losses_g = []
  for i in range(6):
    a, kw = call_args(batches[i % 3])
    _, loss = step(*a, **kw)
    losses_g.append(float(loss.detach()))
  np.testing.assert_allclose(losses_g, losses_e, rtol=1e-5)
  if opt_name == 'sgd':
    for (n, p), (_, q) in zip(graphed.named_parameters(), eager.named_parameters()):
      np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().cpu().numpy(), rtol=2e-4, atol=2e-6, err_msg=n)
  assert step.replays == 6


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (the reference runner\'s gpus: [0, 1])')
def test_data_parallel_two_gpus_inference_and_training():
  """How the reference goes multi-GPU (runner/qm8_runner.py:64-66,291-292): nn.DataParallel over the module,
  scatter of the batch, per-replica forward in threads, gather of (score, loss).  Scores of the two halves
  equal the single-GPU forward; the gradients that flow back through the replicas to the master
  parameters equal the single-GPU gradients of the same batch (equal halves: mean of the two MSEs)."""
  g = load_golden('lanczosnet_qm8.npz')
  cfg = configs.qm8_lanczos_net()
  base = LanczosNet(cfg)
  params = deterministic_state_dict(base, 21)
  base.load_state_dict(params)
  base = base.cuda(0)
  args = [_t(g[k]).cuda(0) for k in ('node_feat', 'L', 'D', 'V')]
  label, mask = _t(g['label']).cuda(0), _t(g['node_mask']).cuda(0)
  base.eval()
  with torch.no_grad():
    ref = base(*args, mask=mask)
  base.train()
  _, loss1 = base(*args, label=label, mask=mask)
  loss1.backward()
  grads1 = {n: p.grad.detach().clone() for n, p in base.named_parameters()}
  base.zero_grad()

  dp = torch.nn.DataParallel(base, device_ids=[0, 1])
  dp.eval()
  with torch.no_grad():
    score, loss = dp(*args, label=label, mask=mask)
  assert score.shape == ref.shape and loss.numel() == 2           # one loss per replica
  torch.testing.assert_close(score, ref, rtol=1e-5, atol=1e-6)
  dp.train()
  _, loss2 = dp(*args, label=label, mask=mask)
  loss2.float().mean().backward()                                   # runner/qm8_runner.py:243-247
  assert abs(float(loss2.mean().detach()) - float(loss1.detach())) <= 1e-5 * max(1.0, abs(float(loss1.detach())))
  for n, p in base.named_parameters():
    assert p.grad is not None, n
    scale = float(grads1[n].abs().max()) + 1e-12
    assert float((p.grad - grads1[n]).abs().max()) / scale <= 1e-4, n
