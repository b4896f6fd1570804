"""world_size-2 gloo test (CPU) of the data-parallel plumbing used by bench.py --gpus N."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, world, port, out):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from trax_b200 import dp
  x = torch.arange(4 * 3 * 2, dtype=torch.float32).reshape(4, 3, 2)
  mine = dp.shard_batch(x, rank, world)
  assert mine.shape == (2, 3, 2) and torch.equal(mine, x[2 * rank:2 * rank + 2])
  # per-rank "weight gradients" = sum over the rank's examples; mean over ranks == (global sum) / world
  g = (mine.sum(dim=(0, 1)).clone(), torch.full((2, 5), float(rank + 1)), torch.ones(3) * (10 ** rank))
  dp.allreduce_mean_(g)
  want0 = x.sum(dim=(0, 1)) / world
  ok = torch.allclose(g[0], want0) and torch.allclose(g[1], torch.full((2, 5), 1.5)) and torch.allclose(g[2], torch.ones(3) * 5.5)
  out[rank] = bool(ok)
  dist.barrier()
  dist.destroy_process_group()


def test_shard_batch_and_allreduce_mean_world2():
  world = 2
  port = _free_port()
  with mp.Manager() as m:
    out = m.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_allreduce_is_identity_without_process_group():
  from trax_b200 import dp
  g = (torch.ones(3), torch.zeros(2, 2))
  assert dp.allreduce_mean_(g) is g and torch.equal(g[0], torch.ones(3))
