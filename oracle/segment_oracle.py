"""Oracle for operators/segment_reduction (numpy loops).  TEST INFRASTRUCTURE -- see
oracle/__init__.py.

The reference ships three mutually inconsistent behaviours; the oracle states each:

  * ``intended``  out[b, seg[b,c], :] += data[b,c,:] with an output of B x S x D2 -- what
    operators/functions/unsorted_segment_sum.py:8-28 allocates and the op name promises.
  * ``ref_cuda``  operators/src/cuda/segment_reduction.cu:39-53: output batch stride is
    dim1*dim2 (not S*dim2), per-batch ids seg[b,c].  Equals ``intended`` iff S == dim1.
  * ``ref_cpu``   operators/src/segment_reduction.cpp:6-30: same stride quirk and ids of
    batch 0 used for every batch (segment_ids_ptr[jj]).

The product implements ``intended`` (documented in DESIGN.md); the tests show the three
agree on the domain where the reference is self-consistent (S == dim1, ids shared across
the batch).
"""
import numpy as np


def segment_sum_forward(data, seg, num_segments, flavour='intended'):
  data = np.asarray(data, dtype=np.float32)
  seg = np.asarray(seg, dtype=np.int64)
  B, C, X = data.shape
  out = np.zeros((B, num_segments, X), dtype=np.float32)
  flat = out.reshape(-1)
  for b in range(B):
    for c in range(C):
      if flavour == 'intended':
        out[b, seg[b, c], :] += data[b, c, :]
      elif flavour == 'ref_cuda':
        pos = b * C * X + seg[b, c] * X
        flat[pos:pos + X] += data[b, c, :]
      elif flavour == 'ref_cpu':
        pos = b * C * X + seg.reshape(-1)[c] * X
        flat[pos:pos + X] += data[b, c, :]
      else:
        raise ValueError(flavour)
  return out


def segment_sum_backward(grad_out, seg, data_shape, flavour='intended'):
  grad_out = np.asarray(grad_out, dtype=np.float32)
  seg = np.asarray(seg, dtype=np.int64)
  B, C, X = data_shape
  grad = np.zeros((B, C, X), dtype=np.float32)
  flat = grad_out.reshape(-1)
  for b in range(B):
    for c in range(C):
      if flavour == 'intended':
        grad[b, c, :] = grad_out[b, seg[b, c], :]
      elif flavour == 'ref_cuda':
        pos = b * C * X + seg[b, c] * X
        grad[b, c, :] = flat[pos:pos + X]
      elif flavour == 'ref_cpu':
        pos = b * C * X + seg.reshape(-1)[c] * X
        grad[b, c, :] = flat[pos:pos + X]
      else:
        raise ValueError(flavour)
  return grad
