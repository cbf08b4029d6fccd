"""autograd front end of the segment-sum op (reference: operators/functions/unsorted_segment_sum.py:8-44).
data [B, dim1, dim2] is reduced over dim1 into [B, num_segments, dim2]."""
import torch
from torch.autograd import Function

from .._ext import segment_reduction

__all__ = ['UnsortedSegmentSumFunction', 'unsorted_segment_sum']


class UnsortedSegmentSumFunction(Function):

  @staticmethod
  def forward(ctx, data, segment_index, num_segments):
    ctx.save_for_backward(segment_index)
    ctx.data_shape = tuple(data.shape)
    data = data.contiguous()
    segment_index = segment_index.contiguous()
    output = torch.zeros((data.size(0), num_segments, data.size(2)), dtype=torch.float32,
                         device=data.device)
    segment_reduction.unsorted_segment_sum_forward(data, segment_index, data.size(), output)
    return output

  @staticmethod
  def backward(ctx, grad_output):
    (segment_index,) = ctx.saved_tensors
    grad_data = torch.zeros(ctx.data_shape, dtype=torch.float32, device=grad_output.device)
    segment_reduction.unsorted_segment_sum_backward(grad_output.contiguous(), segment_index,
                                                    ctx.data_shape, grad_data)
    return grad_data, None, None


unsorted_segment_sum = UnsortedSegmentSumFunction.apply
