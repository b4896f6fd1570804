"""Data-parallel plumbing around the layer: the layer itself has no collective.

Trax's only parallelism is `pmap(axis_name='batch')` with `psum(grads)/n` (`trax/optimizers/trainer.py:172-199`):
every device owns B/n examples with all heads and replicated weights.  The analogue here is one process per
GPU (`torch.distributed`, NCCL over NVLink on the GPU box, gloo in CPU tests): `shard_batch` picks a rank's
examples, `allreduce_mean_` averages the weight gradients.  Units (example, head) never exchange data.
"""
import os

import torch
import torch.distributed as dist


def shard_batch(x, rank, world_size):
  """Rows [rank*B/n, (rank+1)*B/n) of the batch axis (trax `reshape_by_device`, layers/acceleration.py:219)."""
  b = x.shape[0]
  if b % world_size != 0:
    raise ValueError('batch %d not divisible by %d devices' % (b, world_size))
  per = b // world_size
  return x[rank * per:(rank + 1) * per]


class GradOverlap:
  """Mean of the layer's weight gradients over the data-parallel group, overlapped with the rest of the backward call.

  `lsh_layer_bwd` records one CUDA event when dw_o is final (before the attention-gradient kernels) and one when
  dw_q | dw_v are final (before the dx GEMM).  `reduce` makes a per-device communication stream wait for those events,
  all-reduces the two slices of the caller's flat gradient buffer IN PLACE there (NCCL `AVG`: no concatenation, no
  copy back, no separate division) and lets the caller's stream wait for the result — which by then has travelled
  underneath the remaining kernels of the call."""
  _per_device = {}

  def __init__(self, dev):
    self.dev = dev
    # above the compute stream by default: collectives must not queue behind GEMMs (LSH_COMM_STREAM_PRIORITY=0: same priority)
    self.comm = torch.cuda.Stream(device=dev, priority=int(os.environ.get('LSH_COMM_STREAM_PRIORITY', '-1')))
    self._events = None

  @classmethod
  def get(cls, dev):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or dist.get_backend() != 'nccl':
      return None
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    inst = cls._per_device.get(key)
    if inst is None:
      inst = cls._per_device[key] = cls(dev)
    return inst

  def events(self):
    """Two events per call from a small ring (torch creates the CUDA event at its first record)."""
    if self._events is None:
      self._events, self._next = [], 0
      main = torch.cuda.current_stream(self.dev)
      for _ in range(8):
        ev = torch.cuda.Event()
        ev.record(main)
        self._events.append(ev)
    pair = self._events[self._next], self._events[self._next + 1]
    self._next = (self._next + 2) % len(self._events)
    return pair

  def reduce(self, flat, n_qv, ev_o, ev_qv):
    main = torch.cuda.current_stream(self.dev)
    flat.record_stream(self.comm)
    with torch.cuda.stream(self.comm):
      if ev_o is not None:
        self.comm.wait_event(ev_o)
        dist.all_reduce(flat[n_qv:], op=dist.ReduceOp.AVG)
      self.comm.wait_event(ev_qv)
      dist.all_reduce(flat[:n_qv], op=dist.ReduceOp.AVG)
      if ev_o is None:                       # dw_o was finalised on the caller's stream after the C call
        self.comm.wait_event(main.record_event())
        dist.all_reduce(flat[n_qv:], op=dist.ReduceOp.AVG)
      done = self.comm.record_event()
    main.wait_event(done)


def allreduce_mean_flat_(flat, group=None):
  """In-place mean of ONE contiguous gradient buffer over the process group, on the caller's stream, after the call's last
  kernel (no concatenation, no copy back; NCCL divides itself).  The robust data-parallel default: its cost is the
  collective's own duration (6.3 MB at config 2/3) and nothing it does can take SMs from the layer's persistent kernels."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return flat
  if dist.get_backend(group) == 'nccl':
    dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
  else:
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
  return flat


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


# ---------------------------------------------------------------------------------------------------------------------
# Head sharding (BASELINE config 5): few examples, many heads, one long sequence
# ---------------------------------------------------------------------------------------------------------------------
def head_range(n_heads, rank, world_size):
  """Heads [h0, h1) owned by `rank`: contiguous blocks, the unit order b*H + h of EA:2406-2407 restricted to them."""
  if n_heads % world_size != 0:
    raise ValueError('n_heads %d not divisible by %d ranks' % (n_heads, world_size))
  per = n_heads // world_size
  return rank * per, (rank + 1) * per


def shard_heads(weights, state, n_heads, rank, world_size):
  """A rank's slice of the full layer's `(w_q, w_v, w_o)` (head-major, EA:1829-1830) and `(buckets, rng)` state
  (rows b*H + h, EA:1831): the result is the weights / state of a layer with n_heads / world_size heads."""
  h0, h1 = head_range(n_heads, rank, world_size)
  w = tuple(t[h0:h1].contiguous() for t in weights)
  s = state
  if state is not None and len(state) == 2:
    buckets, rng = state
    bsz = buckets.shape[0] // n_heads
    rows = torch.arange(bsz).repeat_interleave(h1 - h0) * n_heads + torch.arange(h0, h1).repeat(bsz)
    rows = rows.to(buckets.device)
    s = (buckets[rows].contiguous(), rng[rows.to(rng.device)].contiguous())
  return w, s


class HeadShardedLSHSelfAttention:
  """`LSHSelfAttention` with its heads split over the ranks of a process group (one process per GPU).

  The reference sums heads into the output and into the input gradient (`index_add` at EA:2426 and EA:2430); units are
  otherwise independent (EA:2402-2432).  Here every rank owns n_heads / world_size heads — their weights, their bucket
  state — and sees the whole input; the two head sums become `all_reduce(SUM)` of the (B, L, D) partial output and
  partial input gradient over NVLink (NCCL on the GPU box, gloo in the CPU tests).  There is no weight-gradient collective:
  a head's weights live on one rank.  `reduce='scatter'` leaves each rank with its L / world_size rows of the sum instead
  (a sequence-parallel consumer — the position-wise residual / feed-forward half of the block — then all-gathers only
  what the next attention layer needs; half the traffic of the all-reduce per call).

  `local_layer` is the per-rank layer (built by the caller with n_heads / world_size heads), so the CPU tests can run
  the same plumbing around the oracle.  Collectives are issued on the caller's stream, after the call's last kernel;
  they are INSIDE whatever the caller times.
  """

  def __init__(self, local_layer, n_heads, group=None, reduce='all'):
    if reduce not in ('all', 'scatter'):
      raise ValueError("reduce must be 'all' or 'scatter'")
    if getattr(local_layer, '_incremental', False):
      raise NotImplementedError("head sharding is built for the training path; a mode='predict' layer is not sharded")
    self._local = local_layer
    self._n_heads = n_heads
    self._group = group
    self._reduce = reduce
    self._world = dist.get_world_size(group) if dist.is_initialized() else 1
    self._rank = dist.get_rank(group) if dist.is_initialized() else 0
    head_range(n_heads, self._rank, self._world)
    self.comm_bytes = 0
    self.n_calls = 0

  @property
  def local(self):
    return self._local

  @property
  def weights(self):
    return self._local.weights

  @property
  def state(self):
    return self._local.state

  def load_full(self, weights, state):
    """Takes this rank's heads out of the full layer's weights / state (same values on every rank)."""
    self._local.weights, self._local.state = shard_heads(weights, state, self._n_heads, self._rank, self._world)

  def _head_sum(self, t):
    """Sum of the per-rank partial results over the group (EA:2426 / EA:2430 across ranks), in place on the tensor's
    device (NCCL reduces device memory, gloo host memory)."""
    if self._world == 1 or t is None:
      return t
    self.comm_bytes += t.numel() * t.element_size()
    self.n_calls += 1
    if self._reduce == 'all':
      dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._group)
      return t
    bsz, seqlen = t.shape[0], t.shape[1]
    if seqlen % self._world != 0:
      raise ValueError('seqlen %d not divisible by %d ranks' % (seqlen, self._world))
    per = seqlen // self._world
    mine = torch.empty((bsz, per) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device)
    chunks = [t[:, r * per:(r + 1) * per].contiguous() for r in range(self._world)]
    dist.reduce_scatter(mine, chunks, op=dist.ReduceOp.SUM, group=self._group)
    return mine

  # Host tensors with an NCCL group: upload once, run the local layer and the head sum on the device, download the SUM —
  # never a partial result (the local layer's own host path would download its partial output first).
  def _io(self, t):
    if t is None or t.is_cuda or not (dist.is_initialized() and dist.get_backend(self._group) == 'nccl'):
      return None
    from trax_b200.lsh_attention import _HostIO
    return _HostIO.get(torch.device('cuda', torch.cuda.current_device()))

  def forward(self, inputs):
    x = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
    io = self._io(x)
    if io is None:
      return self._head_sum(self._local.forward(inputs))
    x_d = io.upload(x).contiguous()
    self._x_stash = (x, x._version, x_d)
    dev_inputs = x_d if not isinstance(inputs, (tuple, list)) else (x_d,) + tuple(io.upload(t) for t in inputs[1:])
    out = io.download(self._head_sum(self._local.forward(dev_inputs)))
    io.finish()
    return out

  def backward(self, inputs, output, grad, weights, state, new_state, rng=None, **kwargs):
    x = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
    io = self._io(x)
    if io is not None:
      stash, self._x_stash = getattr(self, '_x_stash', None), None
      x_d = stash[2] if stash is not None and stash[0] is x and stash[1] == x._version else io.upload(x).contiguous()
      inputs = x_d if not isinstance(inputs, (tuple, list)) else (x_d,) + tuple(io.upload(t) for t in inputs[1:])
      grad = io.upload(grad)
    dx, dw = self._local.backward(inputs, output, grad, weights, state, new_state, rng, **kwargs)
    rest = ()
    if isinstance(dx, tuple):
      dx, rest = dx[0], dx[1:]
    dx = self._head_sum(dx)
    if io is not None:
      dx = io.download(dx, 'dx')
      dw = tuple(io.download(g, 'dw%d' % i) for i, g in enumerate(dw))
      io.finish()
    return ((dx,) + rest if rest else dx), dw

  def forward_and_or_backward(self, inputs, weights, state, rng, output_grad=None, compute_output=True, update_state=True):
    if self._io(inputs[0] if isinstance(inputs, (tuple, list)) else inputs) is not None:
      raise NotImplementedError('HeadShardedLSHSelfAttention.forward_and_or_backward takes device tensors under NCCL '
                                '(forward / backward accept host tensors)')
    out, new_state, dx, dw = self._local.forward_and_or_backward(inputs, weights, state, rng, output_grad=output_grad,
                                                                 compute_output=compute_output, update_state=update_state)
    if isinstance(dx, tuple):
      dx = (self._head_sum(dx[0]),) + dx[1:]
    else:
      dx = self._head_sum(dx)
    return self._head_sum(out), new_state, dx, dw
