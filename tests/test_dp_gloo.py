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
  # the layer's default: ONE contiguous buffer (dw_q | dw_v | dw_o are views of it), reduced in place
  flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
  view = flat[4:]
  same = dp.allreduce_mean_flat_(flat)
  ok = ok and same is flat and torch.allclose(flat, torch.arange(10, dtype=torch.float32) * 1.5) and torch.allclose(view, flat[4:])
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
  f = torch.ones(4)
  assert dp.allreduce_mean_flat_(f) is f and torch.equal(f, torch.ones(4))


# ---- head sharding (BASELINE config 5): the head sums of EA:2426 / EA:2430 across ranks ------------------------------------
class _OracleLayer:
  """Per-rank stand-in with the layer's interface whose compute is the CPU oracle (tests may use the oracle): lets the
  head-sharding plumbing of trax_b200.dp run under gloo without a GPU."""

  def __init__(self, cfg):
    self.cfg, self.weights, self.state = cfg, (), ()

  def _run(self, x, weights, state, output_grad, compute_output):
    import numpy as np
    from oracle import lsh_oracle as O
    out, _, dx, dw = O.forward_and_or_backward(
        self.cfg, x.numpy(), tuple(w.numpy() for w in weights), buckets=state[0].numpy(), update_state=False,
        output_grad=None if output_grad is None else output_grad.numpy(), compute_output=compute_output)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a))
    return t(out), None, t(dx), None if dw is None else tuple(t(g) for g in dw)

  def forward(self, x):
    return self._run(x, self.weights, self.state, None, True)[0]

  def backward(self, x, output, grad, weights, state, new_state, rng=None):
    _, _, dx, dw = self._run(x, weights, new_state, grad, False)
    return dx, dw

  def forward_and_or_backward(self, x, weights, state, rng, output_grad=None, compute_output=True, update_state=True):
    return self._run(x, weights, state, output_grad, compute_output)


def _head_case():
  import numpy as np
  from oracle import lsh_oracle as O
  H, B, L, D = 4, 2, 64, 32
  cfg = O.LSHConfig(n_heads=H, d_qk=8, d_v=8, causal=True, chunk_len=16, n_hashes=2, n_buckets=4)
  rng = np.random.default_rng(5)
  x = rng.standard_normal((B, L, D))
  w = tuple(rng.standard_normal(s) * 0.2 for s in ((H, D, 8), (H, D, 8), (H, 8, D)))
  rot = rng.standard_normal((B * H,) + tuple(O.rotations_shape(cfg, L))).astype(np.float32)
  dout = rng.standard_normal((B, L, D))
  _, buckets, _, _ = O.forward_and_or_backward(cfg, x, w, rotations=rot, update_state=True)
  full = O.forward_and_or_backward(cfg, x, w, buckets=buckets, update_state=False, output_grad=dout)
  return cfg, x, w, buckets, dout, full


def _head_worker(rank, world, port, out, reduce):
  import dataclasses
  import numpy as np
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from trax_b200 import dp
  cfg, x, w, buckets, dout, full = _head_case()
  t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
  local = _OracleLayer(dataclasses.replace(cfg, n_heads=cfg.n_heads // world))
  layer = dp.HeadShardedLSHSelfAttention(local, cfg.n_heads, reduce=reduce)
  layer.load_full(tuple(t(a) for a in w), (t(buckets), torch.zeros((buckets.shape[0], 2), dtype=torch.int32)))
  h0, h1 = dp.head_range(cfg.n_heads, rank, world)
  ok = tuple(layer.weights[0].shape) == (h1 - h0,) + w[0].shape[1:] and layer.state[0].shape[0] == x.shape[0] * (h1 - h0)
  got_out = layer.forward(t(x))
  dx, dw = layer.backward(t(x), got_out, t(dout), layer.weights, None, layer.state, None)
  o4, _, dx4, dw4 = layer.forward_and_or_backward(t(x), layer.weights, layer.state, None, output_grad=t(dout))
  want_out, want_dx = t(full[0]), t(full[2])
  if reduce == 'scatter':
    per = x.shape[1] // world
    want_out, want_dx = want_out[:, rank * per:(rank + 1) * per], want_dx[:, rank * per:(rank + 1) * per]
  ok = ok and torch.allclose(got_out, want_out, atol=1e-10) and torch.allclose(dx, want_dx, atol=1e-10)
  ok = ok and torch.allclose(o4, want_out, atol=1e-10) and torch.allclose(dx4, want_dx, atol=1e-10)
  for g, g4, f in zip(dw, dw4, full[3]):
    ok = ok and torch.allclose(g, t(f[h0:h1]), atol=1e-10) and torch.allclose(g4, t(f[h0:h1]), atol=1e-10)
  ok = ok and layer.comm_bytes == 4 * x.size * 8      # two calls with out + dx each... counted per call below
  out[rank] = bool(ok)
  dist.barrier()
  dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize('reduce', ['all', 'scatter'])
def test_head_sharded_layer_sums_heads_across_ranks_world2(reduce):
  """Heads 0-1 on rank 0, 2-3 on rank 1: all-reduced (or reduce-scattered) output and input gradient equal the unsharded
  oracle's, weight gradients equal its head slices (EA:2426, 2430, 2431)."""
  world = 2
  port = _free_port()
  with mp.Manager() as m:
    out = m.dict()
    mp.spawn(_head_worker, args=(world, port, out, reduce), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_heads_picks_unit_rows():
  from trax_b200 import dp
  H, B = 4, 3
  w = tuple(torch.arange(H * 2 * 3, dtype=torch.float32).reshape(H, 2, 3) + i for i in range(3))
  buckets = torch.arange(B * H).reshape(B * H, 1).repeat(1, 5).to(torch.int32)
  rng = torch.arange(B * H * 2).reshape(B * H, 2).to(torch.int32)
  ws, (bk, rg) = dp.shard_heads(w, (buckets, rng), H, rank=1, world_size=2)
  assert torch.equal(ws[0], w[0][2:4])
  assert bk[:, 0].tolist() == [2, 3, 6, 7, 10, 11]            # units b*H + h for h in {2, 3}
  assert torch.equal(rg, rng[[2, 3, 6, 7, 10, 11]])
  with pytest.raises(ValueError):
    dp.head_range(6, 0, 4)
