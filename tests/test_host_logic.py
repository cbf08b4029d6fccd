"""CPU tests of the host-side logic: graph preparation mirror vs the oracle and the reference
fixtures, C-ABI symbol coverage, module surface, gloo sharding.  No GPU, no compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import ROOT, load_golden
from lanczosnetwork_b200 import configs, data
from oracle import graph_prep


def test_get_laplacian_matches_reference_fixture():
  g = load_golden('data_helper_fixture.npz')
  for kind in ('L1', 'L2', 'L4', 'L6'):
    np.testing.assert_allclose(data.get_laplacian(g['adj'], kind), g[kind], rtol=0, atol=1e-15)
  D, V, L = data.get_graph_laplacian_eigs(g['adj'], k=100)
  np.testing.assert_allclose(D, g['D'], atol=1e-12)
  np.testing.assert_allclose(np.abs(V), np.abs(g['V']), atol=1e-10)
  D3, V3, _ = data.get_graph_laplacian_eigs(g['adj'], k=3)
  np.testing.assert_allclose(D3, g['D3'], atol=1e-12)
  with pytest.raises(ValueError):
    data.get_laplacian(g['adj'], 'L9')
  with pytest.raises(ValueError):
    data.check_dist([1, 2.5])
  assert data.check_dist([1, 'inf']) == [1, 'inf']


def test_collate_bit_exact_vs_reference_batch():
  """The product's collate reproduces the reference loader's padded batch bit-for-bit for the
  index / operator tensors (node_feat, mask, L) and the Ritz values."""
  g = load_golden('lanczosnet_qm8.npz')
  samples = []
  for b, n in enumerate(g['sizes']):
    samples.append(data.prepare_graph(g['adjs'][b, :n, :n].astype(np.float64),
                                      g['node_feat'][b, :n], label=g['label'][b:b + 1]))
  out = data.collate(samples, 20)
  assert np.array_equal(out['node_feat'], g['node_feat'])
  assert np.array_equal(out['node_mask'], g['node_mask'])
  assert np.array_equal(out['L'], g['L'])
  assert np.array_equal(out['label'], g['label'])
  np.testing.assert_allclose(out['D'], g['D'], atol=1e-6)
  rec_o = np.einsum('bnk,bk,bmk->bnm', out['V'], out['D'], out['V'])
  rec_g = np.einsum('bnk,bk,bmk->bnm', g['V'], g['D'], g['V'])
  np.testing.assert_allclose(rec_o, rec_g, atol=2e-6)
  # and equals the oracle's literal restatement
  o = graph_prep.collate([dict(graph_prep.prepare_molecule(g['adjs'][b, :n, :n]),
                               node_feat=g['node_feat'][b, :n])
                          for b, n in enumerate(g['sizes'])], 20)
  assert np.array_equal(out['L'], o['L'])


def test_synthetic_regression_graphs_match_reference_recipe():
  g = load_golden('lanczosnet_general_synth.npz')
  graphs = data.synthetic_regression_graphs(num_graphs=16, seed=123)
  out = data.collate(graphs, 20)
  assert np.array_equal(np.array([s['L_simple_4'].shape[0] for s in graphs]), g['sizes'])
  assert np.array_equal(out['node_feat'], g['node_feat'])
  assert np.array_equal(out['L'], g['L'])
  np.testing.assert_allclose(out['D'], g['D'], atol=1e-6)


def test_synthetic_qm8_batch_shape_and_ragged_edges():
  b = data.synthetic_qm8_batch(64, seed=5)
  assert b['L'].shape == (64, 26, 26, 7) and b['V'].shape == (64, 26, 20)
  n = b['node_mask'].sum(axis=1)
  assert n.max() == 26 and n.min() >= 3
  for i in range(64):
    k = int(n[i])
    assert not b['L'][i, k:].any() and not b['L'][i, :, k:].any() and not b['V'][i, k:].any()
    np.testing.assert_allclose(b['L'][i, :k, :k, 0], b['L'][i, :k, :k, 0].T, atol=0)
    # sum of the bond-type adjacencies is the simple graph: channel 0 has an entry wherever any bond has
    assert ((b['L'][i, :, :, 1:] != 0).any(axis=2) <= (b['L'][i, :, :, 0] != 0)).all()
    if k < 20:
      assert not b['D'][i, k:].any()          # zero-padded Ritz values (dataset/qm8.py:268-287)


def test_c_abi_exports_every_declared_symbol():
  """The shared library loads and exports every function include/lanczosnet_b200.h declares."""
  from lanczosnetwork_b200 import _lib, build
  build.build()
  header = open(os.path.join(ROOT, 'include', 'lanczosnet_b200.h')).read()
  declared = set(re.findall(r'\b(lnb_[a-z0-9_]+)\s*\(', header))
  assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for name in declared:
    assert hasattr(lib, name), name
  assert _lib.load().lnb_abi_version() == 1
  assert ctypes.sizeof(_lib.GemmDesc) == 26 * 8 + 6 * 4 + 2 * 4   # + alpha, beta, addend + 4 strides


def test_modules_mirror_reference_surface():
  from lanczosnetwork_b200.model import AdaLanczosNet, LanczosNet, LanczosNetGeneral
  m = LanczosNet(configs.qm8_lanczos_net())
  assert sum(p.numel() for p in m.parameters()) == 1851465       # SURVEY.md 8a1 [probe]
  shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
  assert shapes['filter.0.weight'] == (128, 15 * 64) and shapes['filter.1.weight'] == (128, 1920)
  assert shapes['filter.7.weight'] == (16, 128) and shapes['embedding.weight'] == (70, 64)
  assert shapes['spectral_filter.0.0.weight'] == (128, 8) and shapes['att_func.0.weight'] == (1, 128)
  g = LanczosNetGeneral(configs.graph_lanczos_net())
  assert 'embedding.weight' not in g.state_dict() and g.state_dict()['filter.0.weight'].shape == (128, 100)
  a = AdaLanczosNet(configs.qm8_ada_lanczos_net(num_layer=1, hidden_dim=[8], num_eig_vec=4,
                                                long_diffusion_dist=[2, 3]))
  assert a.use_reorthogonalization is True       # top-level hasattr quirk (ada_lanczos_net.py:35-36)
  assert a.state_dict()['embedding.weight'].shape == (70, 70)
  assert a.state_dict()['spectral_filter.0.0.weight'].shape == (4096, 4 * 4 * 2)
  with pytest.raises(ValueError):
    LanczosNet(configs.qm8_lanczos_net(loss='hinge'))
  with pytest.raises(ValueError):
    LanczosNet(configs.qm8_lanczos_net(long_diffusion_dist=[1.5]))
  # CPU module fails loudly instead of falling back
  with pytest.raises(RuntimeError):
    with torch.no_grad():
      m(torch.zeros(1, 4, dtype=torch.long), torch.zeros(1, 4, 4, 7), torch.zeros(1, 20),
        torch.zeros(1, 4, 20), mask=torch.ones(1, 4, dtype=torch.uint8))
  # same construction + init order as the reference => same RNG stream consumption
  torch.manual_seed(1234)
  m1 = LanczosNet(configs.qm8_lanczos_net())
  torch.manual_seed(1234)
  m2 = LanczosNet(configs.qm8_lanczos_net())
  assert all(torch.equal(a_, b_) for a_, b_ in zip(m1.state_dict().values(), m2.state_dict().values()))
  assert float(m1.filter[0].bias.abs().sum()) == 0.0


def test_dropin_rebinds_names_in_a_runner_namespace(monkeypatch):
  import types
  from lanczosnetwork_b200 import dropin
  fake_runner = types.ModuleType('fake_runner')
  fake_runner.LanczosNet = object
  fake_runner.GCN = object
  fake_runner.MPNN = 'untouched'          # models off the path keep the reference class
  dropin.patch_namespace(fake_runner)
  from lanczosnetwork_b200.model import GCN, LanczosNet
  assert fake_runner.LanczosNet is LanczosNet and fake_runner.GCN is GCN
  assert fake_runner.MPNN == 'untouched'
  # a training run (no -t): classes without a differentiable path keep the reference's class
  train_ns = types.ModuleType('fake_train_runner')
  train_ns.LanczosNet, train_ns.AdaLanczosNet, train_ns.DCNN = 'ref', 'ref', 'ref'
  from lanczosnetwork_b200.model import AdaLanczosNet, DCNN
  monkeypatch.delattr(DCNN, '_train_impl')
  dropin.patch_namespace(train_ns, training=True)
  assert train_ns.LanczosNet is LanczosNet and train_ns.AdaLanczosNet is AdaLanczosNet and train_ns.DCNN == 'ref'
  monkeypatch.undo()
  dropin.register_native_op()
  import importlib
  sr = importlib.import_module('operators._ext.segment_reduction')
  assert callable(sr.unsorted_segment_sum_forward_gpu)


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from lanczosnetwork_b200 import sharded
rank, world, _ = sharded.init_from_env('gloo')
n = 37
full = {'x': torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)}
pred = sharded.sharded_predict(lambda b: b['x'] * 2.0 + 1.0, full, rank, world)
ref = full['x'] * 2.0 + 1.0
assert torch.equal(pred, ref), (rank, pred.shape)
lo, hi, per = sharded.shard_indices(n, rank, world)
assert per == 19 and (lo, hi) == ((0, 19) if rank == 0 else (19, 37))
# one gather for a whole shard of per-step predictions (what bench.py times)
steps = [torch.full((4, 3), float(10 * rank + i)) for i in range(5)]
allp = sharded.gather_once(steps, world)
assert allp.shape == (world, 20, 3)
for r in range(world):
  for i in range(5):
    assert torch.equal(allp[r, 4 * i:4 * i + 4], torch.full((4, 3), float(10 * r + i)))
dist.barrier()
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_sharded_predict_world_size_2_gloo(tmp_path):
  script = tmp_path / 'worker.py'
  script.write_text(WORKER)
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
         '--master-addr', '127.0.0.1', '--master-port', '29533', str(script), ROOT]
  proc = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
  assert proc.returncode == 0, proc.stdout + proc.stderr
  assert proc.stdout.count('ok') == 2


def test_sparse_collate_and_pack_describe_the_same_batch():
  """Host side of the GPU-side batch construction: the bond lists of ``sparse_collate``
  rebuild the reference's padded operators bit for bit through the host mirror, sizes / pointers /
  Ritz rows agree with ``collate``, and ``pack_sparse`` lays the same arrays out behind its header."""
  samples = data.synthetic_qm8_samples(40, seed=9)
  dense = data.collate(samples, 20)
  sp = data.sparse_collate(samples, 20)
  B, N = dense['node_feat'].shape
  assert sp['N'] == N and sp['num_edgetype'] == 6
  assert np.array_equal(sp['sizes'], dense['node_mask'].sum(axis=1).astype(np.int32))
  assert np.array_equal(sp['node_ptr'][1:], np.cumsum(sp['sizes'])) and sp['node_ptr'][0] == 0
  assert np.array_equal(sp['D'], dense['D'])
  for b in range(B):
    n, r0 = int(sp['sizes'][b]), int(sp['node_ptr'][b])
    assert np.array_equal(sp['node_feat'][r0:r0 + n], dense['node_feat'][b, :n])
    assert np.array_equal(sp['V_rows'][r0:r0 + n], dense['V'][b, :n])
    adjs = np.zeros((n, n, 6))
    for u, v, c, pad in sp['edges'][sp['edge_ptr'][b]:sp['edge_ptr'][b + 1]]:
      assert u <= v and pad == 0
      adjs[u, v, c] = adjs[v, u, c] = 1.0
    assert np.array_equal(data.get_laplacian(adjs.sum(axis=2)).astype(np.float32), dense['L'][b, :n, :n, 0])
    for c in range(6):
      assert np.array_equal(data.get_laplacian(adjs[:, :, c]).astype(np.float32), dense['L'][b, :n, :n, 1 + c])
  pk = data.pack_sparse(sp)
  hdr = pk['blob'][:64].view(np.int32)
  assert hdr[0] == data.PACK_MAGIC and hdr[1] == B and hdr[2] == 20 and hdr[10] == pk['blob'].size
  assert tuple(hdr[3:7]) == data.packed_offsets(B, 20)[:4] and all(int(o) % 16 == 0 for o in hdr[3:11])
  for off, key in ((3, 'sizes'), (4, 'node_ptr'), (5, 'edge_ptr'), (6, 'D'), (7, 'node_feat'), (8, 'V_rows'), (9, 'edges')):
    raw = np.ascontiguousarray(sp[key]).view(np.uint8).reshape(-1)
    assert np.array_equal(pk['blob'][hdr[off]:hdr[off] + raw.size], raw), key
  # the fp64 table the device kernel multiplies with is numpy's own deg ** -0.5
  deg = np.arange(1, 12, dtype=np.float64)
  assert np.array_equal(np.power(deg, -0.5), deg ** -0.5)


def test_graphed_step_refuses_what_it_cannot_capture():
  """train.GraphedStep argument checks (no GPU needed): a module without a host-free training forward,
  a missing label, a module that is not on CUDA."""
  import torch
  from lanczosnetwork_b200 import configs
  from lanczosnetwork_b200.model import AdaLanczosNet, LanczosNet
  from lanczosnetwork_b200.train import GraphedStep
  small = dict(num_layer=1, hidden_dim=[16])
  ln = LanczosNet(configs.qm8_lanczos_net(**small))
  opt = torch.optim.SGD(ln.parameters(), lr=0.1)
  x = (torch.zeros(2, 3, dtype=torch.long), torch.zeros(2, 3, 3, 7), torch.zeros(2, 20), torch.zeros(2, 3, 20))
  with pytest.raises(TypeError):
    GraphedStep(torch.nn.Linear(2, 2), opt, x, {'label': torch.zeros(2, 16)})
  ada = AdaLanczosNet(configs.qm8_ada_lanczos_net(num_eig_vec=4, long_diffusion_dist=[1], short_diffusion_dist=[], **small))
  with pytest.raises(TypeError):
    GraphedStep(ada, opt, x[:2], {'label': torch.zeros(2, 16)})
  with pytest.raises(ValueError):
    GraphedStep(ln, opt, x, {'mask': torch.ones(2, 3)})
  with pytest.raises(RuntimeError):            # CPU module: no CPU fallback anywhere
    GraphedStep(ln, opt, x, {'label': torch.zeros(2, 16)})


def test_trainable_parameter_detection_sees_data_parallel_replicas():
  """nn.DataParallel replicas keep their parameter copies in ``_former_parameters`` (``parameters()`` is
  empty there): the training path must still be selected for them (regression: two-GPU training through
  the reference runner's DataParallel took the inference path)."""
  import torch
  from lanczosnetwork_b200 import configs
  from lanczosnetwork_b200.model import LanczosNet
  mod = LanczosNet(configs.qm8_lanczos_net(num_layer=1, hidden_dim=[16]))
  assert mod._has_trainable_parameters()
  # build the replica tree the way torch.nn.parallel.replicate does: children replicated, parameters demoted
  def demote(m):
    r = m._replicate_for_data_parallel()                       # empties r._parameters
    r._former_parameters = {}
    for k, c in m._modules.items():
      r._modules[k] = None if c is None else demote(c)
    for k, p in m._parameters.items():
      if p is None:
        r._parameters[k] = None
      else:
        cp = p.detach().clone().requires_grad_(p.requires_grad)   # stand-in for the Broadcast output
        setattr(r, k, cp)
        r._former_parameters[k] = cp
    return r
  rep = demote(mod)
  assert list(rep.parameters()) == [] and rep._has_trainable_parameters()
  for p in mod.parameters():
    p.requires_grad_(False)
  assert not demote(mod)._has_trainable_parameters() and not mod._has_trainable_parameters()


def test_bench_cpu_training_port_steps_and_restores_threads():
  """bench.train_cpu_port (the CPU leg of the train_qm8 workload) on a tiny model: returns a positive
  time and a thread count, leaves torch's thread setting and the oracle's cast hook as they were."""
  import torch
  import bench
  from helpers import deterministic_state_dict, oracle_spec
  from lanczosnetwork_b200 import configs
  from lanczosnetwork_b200.model import LanczosNet
  from oracle import lanczos_oracle as orc
  mod = LanczosNet(configs.qm8_lanczos_net(num_layer=1, hidden_dim=[16]))
  params = deterministic_state_dict(mod, 1)
  batches = []
  for i in range(3):
    b = data.collate(data.synthetic_qm8_samples(4, seed=i), 20, num_nodes=27)
    b['label'] = np.random.RandomState(i).randn(4, 16).astype(np.float32)
    batches.append(b)
  nt, cast = torch.get_num_threads(), orc._cast
  ms, threads = bench.train_cpu_port(batches, params, oracle_spec(mod, 'LanczosNet'))
  assert ms > 0 and threads >= 1 and torch.get_num_threads() == nt and orc._cast is cast
  assert batches[0]['L'].shape[1] == 27                      # collate(num_nodes=) pads to the fixed node count
