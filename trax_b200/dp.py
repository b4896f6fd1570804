"""Data-parallel plumbing around the layer: the layer itself has no collective.

Trax's only parallelism is `pmap(axis_name='batch')` with `psum(grads)/n` (`trax/optimizers/trainer.py:172-199`):
every device owns B/n examples with all heads and replicated weights.  The analogue here is one process per
GPU (`torch.distributed`, NCCL over NVLink on the GPU box, gloo in CPU tests): `shard_batch` picks a rank's
examples, `allreduce_mean_` averages the weight gradients.  Units (example, head) never exchange data.
"""
import torch
import torch.distributed as dist


def shard_batch(x, rank, world_size):
  """Rows [rank*B/n, (rank+1)*B/n) of the batch axis (trax `reshape_by_device`, layers/acceleration.py:219)."""
  b = x.shape[0]
  if b % world_size != 0:
    raise ValueError('batch %d not divisible by %d devices' % (b, world_size))
  per = b // world_size
  return x[rank * per:(rank + 1) * per]


def allreduce_mean_(grads, group=None):
  """In-place mean of a tuple of gradient tensors over the process group, as ONE flat all-reduce
  (6.3 MB at config 2/3: latency-bound, so a single bucket)."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return grads
  if not grads[0].is_cuda and torch.cuda.is_available():
    torch.cuda.synchronize()        # host-side gradients may still be in flight (asynchronous host I/O)
  flat = torch.cat([g.reshape(-1) for g in grads])
  on_host = not flat.is_cuda and dist.get_backend(group) == 'nccl'
  if on_host:                       # host-buffer (e2e) path: NCCL reduces device memory only
    flat = flat.cuda(non_blocking=True)
  dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  flat /= dist.get_world_size(group)
  if on_host:
    flat = flat.cpu()
  off = 0
  for g in grads:
    n = g.numel()
    g.copy_(flat[off:off + n].view_as(g))
    off += n
  return grads
