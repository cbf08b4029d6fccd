"""``operators._ext.segment_reduction`` -- the native module the reference python side imports
(operators/functions/unsorted_segment_sum.py:5) and calls with
``(data, segment_index, data.size(), output)`` (:19-28,37-42).  Same four names, same argument
order, each returns int 1 like the reference C functions (operators/src/segment_reduction.cpp:29,56).

The ``*_gpu`` entries launch the sm_100a kernels through the C ABI.  The CPU-tensor entries of
the reference (operators/src/segment_reduction.cpp) are NOT re-implemented on the host: they
raise, because this build has no CPU compute path.
"""
from ... import ops

__all__ = ['unsorted_segment_sum_forward', 'unsorted_segment_sum_forward_gpu',
           'unsorted_segment_sum_backward', 'unsorted_segment_sum_backward_gpu']


def _no_cpu(name):
  raise RuntimeError('%s: CPU tensors are not supported by the B200 build (no CPU fallback); '
                     'use the *_gpu entry with CUDA tensors' % name)


def unsorted_segment_sum_forward(data, segment_ids, data_shape, output):
  if not data.is_cuda:
    _no_cpu('unsorted_segment_sum_forward')
  return unsorted_segment_sum_forward_gpu(data, segment_ids, data_shape, output)


def unsorted_segment_sum_forward_gpu(data, segment_ids, data_shape, output):
  """output[b, ids[b,c], :] += data[b, c, :]; output is pre-zeroed by the caller."""
  assert tuple(data_shape) == tuple(data.shape)
  ops.segment_sum_forward(data, segment_ids, output.shape[1], output=output)
  return 1


def unsorted_segment_sum_backward(grad_output, segment_ids, data_shape, grad_data):
  if not grad_output.is_cuda:
    _no_cpu('unsorted_segment_sum_backward')
  return unsorted_segment_sum_backward_gpu(grad_output, segment_ids, data_shape, grad_data)


def unsorted_segment_sum_backward_gpu(grad_output, segment_ids, data_shape, grad_data):
  """grad_data[b, c, :] = grad_output[b, ids[b,c], :]."""
  ops.segment_sum_backward(grad_output, segment_ids, tuple(data_shape), grad_data=grad_data)
  return 1
