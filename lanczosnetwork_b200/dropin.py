"""Install the B200 drop-ins into an UNMODIFIED checkout of lrjconan/LanczosNetwork.

    python -m lanczosnetwork_b200.dropin /path/to/LanczosNetwork -c config/qm8_lanczos_net.yaml -t

What it does:
  1. registers ``operators._ext`` / ``operators._ext.segment_reduction`` in ``sys.modules`` so
     ``from model import *`` of the reference works without building its THC-era extension
     (model/mpnn.py:6 -> operators/functions/unsorted_segment_sum.py:5);
  2. rebinds ``LanczosNet`` / ``AdaLanczosNet`` / ``LanczosNetGeneral`` / ``GCN`` inside the runner
     modules' globals, because the runners resolve the class with ``eval(name)`` in their own
     namespace (runner/qm8_runner.py:59,288; runner/graph_runner.py:57,285);
  3. runs the reference ``run_exp.main()`` unchanged.
"""
import importlib
import os
import sys

from . import model as _models
from .operators import _ext as _ext_pkg

DROPIN_CLASSES = ('LanczosNet', 'AdaLanczosNet', 'LanczosNetGeneral', 'GCN', 'GCNFP', 'DCNN', 'ChebyNet')


def register_native_op():
  """Make ``operators._ext.segment_reduction`` importable under the reference's module path."""
  sys.modules.setdefault('operators._ext', _ext_pkg)
  sys.modules.setdefault('operators._ext.segment_reduction', _ext_pkg.segment_reduction)
  ops_pkg = sys.modules.get('operators')
  if ops_pkg is not None:
    setattr(ops_pkg, '_ext', _ext_pkg)


def patch_namespace(module, training=False):
  """Rebind the class names in ``module``'s globals to the B200 drop-ins.  ``training=True`` (a
  run without ``-t``) rebinds only the classes that have a differentiable training path
  (currently every class); one without it would keep the reference's trainable class
  instead of failing on the first ``loss.backward()``."""
  for name in DROPIN_CLASSES:
    if hasattr(module, name):
      cls = getattr(_models, name)
      if training and not hasattr(cls, '_train_impl'):
        continue
      setattr(module, name, cls)
  return module


def install(reference_root=None, runner_modules=('runner.qm8_runner', 'runner.graph_runner'),
            compat=False, training=False):
  """Returns the list of patched modules.  ``reference_root`` is put on sys.path if given.
  ``compat=True`` first installs the shims of ``lanczosnetwork_b200.compat`` (missing easydict /
  tensorboardX, PyYAML >= 6, numpy >= 2) so the 2019 checkout imports under a current stack.
  Raises ImportError when NO runner module could be imported and patched: the runners resolve
  the model class by name in their own namespace, so a silent miss would run the reference's
  classes while claiming the drop-in."""
  if compat:
    from . import compat as _compat
    _compat.install()
  if reference_root is not None:
    reference_root = os.path.abspath(reference_root)
    if reference_root not in sys.path:
      sys.path.insert(0, reference_root)
  register_native_op()
  patched = []
  ref_model = importlib.import_module('model')
  patched.append(patch_namespace(ref_model, training))
  errors = []
  for name in runner_modules:
    try:
      mod = importlib.import_module(name)
    except ImportError as exc:      # e.g. tensorboardX absent: that runner cannot be used anyway
      errors.append('%s: %s' % (name, exc))
      continue
    patched.append(patch_namespace(mod, training))
  if runner_modules and len(patched) == 1:
    raise ImportError('dropin.install: no runner module could be imported, nothing would call the '
                      'B200 classes (%s); pass compat=True for the shims of '
                      'lanczosnetwork_b200.compat' % '; '.join(errors))
  return patched


def main(argv=None):
  argv = list(sys.argv[1:] if argv is None else argv)
  if not argv:
    raise SystemExit(__doc__)
  root = argv.pop(0)
  install(root, compat=True, training=('-t' not in argv and '--test' not in argv))
  os.chdir(root)
  sys.argv = ['run_exp.py'] + argv
  run_exp = importlib.import_module('run_exp')
  run_exp.main()


if __name__ == '__main__':
  main()
