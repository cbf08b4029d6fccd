"""One-process-per-GPU sharded inference (replaces the reference's single-process
nn.DataParallel, runner/qm8_runner.py:291-292).

Molecules are independent end to end, so the batch shards by index with replicated weights
and NO data-path collective; the only exchange is one NCCL all-gather of the per-graph
predictions ([n_local, P] fp32, ~64 B per molecule -- latency bound over NVSwitch).
On CPU (tests) the same code runs over gloo with any callable standing in for the model.
"""
import os

import torch
import torch.distributed as dist

__all__ = ['init_from_env', 'shard_indices', 'shard_batch', 'gather_predictions',
           'sharded_predict', 'gather_once']


def init_from_env(backend=None):
  """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns
  (rank, world_size, local_rank).  No-op for a single process."""
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1 and not dist.is_initialized():
    if backend is None:
      backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
      torch.cuda.set_device(local)
      dist.init_process_group(backend, device_id=torch.device('cuda', local))
    else:
      dist.init_process_group(backend)
  return rank, world, local


def shard_indices(num_items, rank, world):
  """Contiguous block partition: rank r owns [r*per, min((r+1)*per, n)) with per = ceil(n/world)."""
  per = (num_items + world - 1) // world
  lo = min(rank * per, num_items)
  hi = min(lo + per, num_items)
  return lo, hi, per


def shard_batch(batch, rank, world):
  """Slice every tensor of a dict batch along dim 0 to this rank's block."""
  n = next(iter(batch.values())).shape[0]
  lo, hi, _ = shard_indices(n, rank, world)
  return {k: v[lo:hi] for k, v in batch.items()}, (lo, hi)


def gather_predictions(local_pred, num_items, rank, world):
  """All-gather equal-size padded shards, return the first num_items rows (identical on every
  rank).  One collective per call."""
  if world == 1:
    return local_pred
  _, _, per = shard_indices(num_items, rank, world)
  P = local_pred.shape[1]
  padded = local_pred.new_zeros((per, P))
  padded[:local_pred.shape[0]] = local_pred
  out = local_pred.new_empty((per * world, P))
  dist.all_gather_into_tensor(out, padded)
  return out[:num_items]


def sharded_predict(predict_fn, batch, rank, world):
  """predict_fn(local_batch_dict) -> [n_local, P]; returns the full [n, P] predictions."""
  n = next(iter(batch.values())).shape[0]
  local, _ = shard_batch(batch, rank, world)
  pred = predict_fn(local)
  return gather_predictions(pred, n, rank, world)


def gather_once(local_preds, world):
  """ONE collective for a whole shard: the per-step predictions a rank kept on its device
  ([n_step, P] each, the same shapes on every rank) are concatenated and all-gathered in a single
  ``all_gather_into_tensor``.  Returns [world, steps * n_step, P] (rank-major), or the local stack for
  a single process."""
  local = torch.cat(list(local_preds), dim=0)
  if world == 1:
    return local.unsqueeze(0)
  out = local.new_empty((world * local.shape[0], local.shape[1]))
  dist.all_gather_into_tensor(out, local)
  return out.reshape(world, local.shape[0], local.shape[1])
