"""Shims that let the UNMODIFIED 2019 reference checkout import and run under a current stack.

None of this is on the hot path; it only keeps the reference's own launcher (run_exp.py ->
runner/qm8_runner.py -> dataset/qm8.py) alive so the drop-in classes can be exercised through it:

  * ``easydict`` (utils/arg_helper.py:5) and ``tensorboardX`` (runner/qm8_runner.py:13) are not
    installed here and cannot be (no network): tiny stand-ins are registered in ``sys.modules``
    only when the real packages are absent;
  * ``yaml.load(f)`` without a Loader (utils/arg_helper.py:38) raises on PyYAML >= 6: default to
    FullLoader;
  * ``np.expand_dims(2-D, axis=3)`` in the collate (dataset/qm8.py:254-259) raises AxisError on
    numpy >= 2 (numpy 1.x clamped the axis): restore the clamp;
  * ``np.float`` / ``np.int`` aliases (utils/spectral_graph_partition.py:43).

``install()`` is idempotent and opt-in (``dropin.install(compat=True)``).
"""
import sys
import types

import numpy as np

_DONE = {'installed': False}


class _EasyDict(dict):
  """Attribute-access dict, recursive on nested dicts / lists (what the reference uses of easydict)."""

  def __init__(self, d=None, **kw):
    super(_EasyDict, self).__init__()
    d = dict(d or {}, **kw)
    for k, v in d.items():
      self[k] = v

  @classmethod
  def _wrap(cls, v):
    if isinstance(v, dict) and not isinstance(v, cls):
      return cls(v)
    if isinstance(v, (list, tuple)):
      return type(v)(cls._wrap(x) for x in v)
    return v

  def __setitem__(self, k, v):
    super(_EasyDict, self).__setitem__(k, self._wrap(v))

  def __setattr__(self, k, v):
    self[k] = v

  def __getattr__(self, k):
    try:
      return self[k]
    except KeyError:
      raise AttributeError(k)


class _SummaryWriter(object):
  """No-op stand-in for tensorboardX.SummaryWriter."""

  def __init__(self, *a, **kw):
    pass

  def __getattr__(self, name):
    return lambda *a, **kw: None


def install():
  if _DONE['installed']:
    return
  try:
    import easydict  # noqa: F401
  except ImportError:
    mod = types.ModuleType('easydict')
    mod.EasyDict = _EasyDict
    sys.modules['easydict'] = mod
  try:
    import tensorboardX  # noqa: F401
  except ImportError:
    mod = types.ModuleType('tensorboardX')
    mod.SummaryWriter = _SummaryWriter
    sys.modules['tensorboardX'] = mod

  import yaml
  if not getattr(yaml.load, '_lnb_compat', False):
    _load = yaml.load

    def load(stream, Loader=None, **kw):
      return _load(stream, Loader=Loader or yaml.FullLoader, **kw)

    load._lnb_compat = True
    yaml.load = load

  if not getattr(np.expand_dims, '_lnb_compat', False):
    _expand = np.expand_dims

    def expand_dims(a, axis):
      nd = np.ndim(a)
      if isinstance(axis, int) and axis > nd:      # numpy 1.x behaviour the reference relied on
        axis = nd
      return _expand(a, axis)

    expand_dims._lnb_compat = True
    np.expand_dims = expand_dims
  for name, typ in (('float', float), ('int', int), ('bool', bool)):
    if name not in np.__dict__:
      setattr(np, name, typ)
  _DONE['installed'] = True
