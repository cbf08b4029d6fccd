"""lanczosnetwork_b200 -- B200-native (sm_100a) LanczosNet spectral-convolution forward path.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/lanczosnet_b200.h)
  _lib.py, ops.py  ctypes loader and torch-tensor front end of the C ABI
  spectral_conv.py the convolution layer assembled from the kernels
  model/           drop-ins for model.LanczosNet / AdaLanczosNet / LanczosNetGeneral
  operators/       drop-in for operators/segment_reduction (unsorted_segment_sum)
  data.py          host-side graph preparation (L4 operators, Ritz-pair provider, collate)
  dropin.py        installs the drop-ins into the unmodified reference runner
  sharded.py       one-process-per-GPU sharded inference with a single NCCL gather
"""
from .model import LanczosNet, AdaLanczosNet, LanczosNetGeneral  # noqa: F401

__version__ = '0.1.0'
