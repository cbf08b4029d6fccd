"""The UNMODIFIED reference launcher driving the drop-in classes.

Runs in a subprocess (the reference's top-level package names ``model`` / ``runner`` / ``utils`` /
``dataset`` must not leak into this interpreter): synthetic molecules are written in the
reference's on-disk format (dataset/get_qm8_data.py:45-101: one pickle per molecule +
QM8_meta.p), a checkpoint in the format of utils/train_helper.py:14-25, then
``runner.qm8_runner.QM8Runner(config).test()`` -- config read by the reference's own
``utils.arg_helper.get_config`` from its own ``config/qm8_lanczos_net.yaml`` -- is executed after
``dropin.install(compat=True)``.

CPU container: asserts that ``eval(self.model_conf.name)`` inside the runner resolved to the B200
class, that the reference loader + collate fed it, and that the forward refuses to run without CUDA
(no CPU fallback).  With a GPU *and* /root/reference present the logged MAE is compared with the
oracle's.  /root/reference does not exist on the GPU box: the test skips there.
"""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

REF = '/root/reference'

SCRIPT = r'''
import os, pickle, sys
import numpy as np
import torch
repo, ref, work = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, repo); sys.path.insert(0, os.path.join(repo, 'tests'))
from lanczosnetwork_b200 import data, dropin
from lanczosnetwork_b200 import model as b200_models

# ---- synthetic molecules in the reference's on-disk format --------------------------------
rng = np.random.RandomState(7)
pre = os.path.join(work, 'data', 'QM8', 'preprocess'); os.makedirs(pre)
records = []
for i, n in enumerate([9, 14, 5, 20, 11, 17, 8, 13, 26, 3]):
  nf, adjs = data.synthetic_molecule(rng, n)
  rec = data.prepare_graph(adjs, nf, label=rng.randn(1, 16))
  simple = adjs.sum(axis=2)
  rec['L_simple_6'] = data.get_laplacian(simple, 'L6')
  rec['L_simple_7'] = rec['L_simple_6']
  rec['label_weight'] = np.ones((1, 16))
  records.append(rec)
  pickle.dump(rec, open(os.path.join(pre, 'QM8_preprocess_test_%07d.p' % i), 'wb'))
std = np.linspace(0.5, 2.0, 16)
pickle.dump({'mean': np.zeros(16), 'std': std}, open(os.path.join(work, 'data', 'QM8', 'QM8_meta.p'), 'wb'))

patched = dropin.install(ref, compat=True)
names = [m.__name__ for m in patched]
assert 'runner.qm8_runner' in names and 'runner.graph_runner' in names, names
import runner.qm8_runner as qr
assert qr.LanczosNet is b200_models.LanczosNet and qr.AdaLanczosNet is b200_models.AdaLanczosNet
assert qr.GCN is b200_models.GCN
import model as ref_model
assert ref_model.MPNN.__module__ == 'model.mpnn'          # off-path models keep the reference class

# ---- the reference's own config loader on its own yaml -------------------------------------
os.chdir(work)
from utils.arg_helper import get_config
config = get_config(os.path.join(ref, 'config', 'qm8_lanczos_net.yaml'), exp_dir=os.path.join(work, 'exp'))
config.use_gpu = config.use_gpu and torch.cuda.is_available()
config.test.batch_size = 4
from helpers import deterministic_state_dict, oracle_spec
proto = b200_models.LanczosNet(config)
params = deterministic_state_dict(proto, 77)
ckpt = os.path.join(work, 'model_snapshot_best.pth')
torch.save({'model': params, 'optimizer': {}, 'step': 0}, ckpt)
config.test.test_model = ckpt

seen = {}
orig_forward = b200_models.LanczosNet.forward
def spy(self, node_feat, L, D, V, label=None, mask=None):
  seen['cls'] = type(self)
  seen['shapes'] = (tuple(node_feat.shape), tuple(L.shape), tuple(D.shape), tuple(V.shape))
  return orig_forward(self, node_feat, L, D, V, label=label, mask=mask)
b200_models.LanczosNet.forward = spy

run = qr.QM8Runner(config)
if not torch.cuda.is_available():
  try:
    run.test()
  except RuntimeError as exc:
    assert 'CUDA' in str(exc) and 'no CPU' in str(exc), exc
  else:
    raise SystemExit('the forward ran without CUDA: there must be no CPU fallback')
  assert seen['cls'] is b200_models.LanczosNet
  B, N = seen['shapes'][0]
  assert B == 4 and seen['shapes'][1] == (4, N, N, 7) and seen['shapes'][3] == (4, N, 20)
  print('RUNNER_OK cpu')
else:
  mae = run.test()
  from oracle import lanczos_oracle as orc
  import glob
  test_files = glob.glob(os.path.join(pre, 'QM8_preprocess_test_*.p'))     # the loader's own order
  recs = [pickle.load(open(f, 'rb')) for f in test_files]
  errs = []
  spec = oracle_spec(proto, 'LanczosNet')
  for i in range(0, len(recs), 4):
    b = data.collate(recs[i:i + 4], 20)
    ref_score = orc.lanczos_net_forward(params, spec, b['node_feat'], b['L'], b['D'], b['V'], b['node_mask']).numpy()
    errs.append(np.abs(ref_score - b['label']) * std.reshape(1, -1))
  want = float(np.mean(np.concatenate(errs)))
  assert abs(mae - want) <= 1e-5, (mae, want)
  print('RUNNER_OK gpu mae=%.7f oracle=%.7f' % (mae, want))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present (GPU box)')
def test_reference_runner_calls_the_dropin_classes(tmp_path):
  script = tmp_path / 'drive_runner.py'
  script.write_text(SCRIPT)
  work = tmp_path / 'work'
  work.mkdir()
  proc = subprocess.run([sys.executable, str(script), ROOT, REF, str(work)], capture_output=True,
                        text=True, timeout=600)
  assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
  assert 'RUNNER_OK' in proc.stdout


def test_dropin_install_raises_when_no_runner_can_be_patched(tmp_path):
  """A reference tree whose runners cannot be imported must not be reported as patched."""
  root = tmp_path / 'fake_ref'
  (root / 'model').mkdir(parents=True)
  (root / 'model' / '__init__.py').write_text('class LanczosNet(object):\n  pass\n')
  (root / 'runner').mkdir()
  (root / 'runner' / '__init__.py').write_text('')
  (root / 'runner' / 'qm8_runner.py').write_text('import a_module_that_does_not_exist_anywhere\n')
  code = ('import sys; sys.path.insert(0, %r)\n'
          'from lanczosnetwork_b200 import dropin\n'
          'try:\n'
          '  dropin.install(%r, runner_modules=("runner.qm8_runner",))\n'
          'except ImportError as exc:\n'
          '  assert "no runner module" in str(exc); print("RAISED")\n' % (ROOT, str(root)))
  proc = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
  assert proc.returncode == 0 and 'RAISED' in proc.stdout, proc.stdout + proc.stderr
