"""`LSHSelfAttention` — drop-in for `trax.layers.research.efficient_attention.LSHSelfAttention`
(EA:1729-2561) on the train path, hosted on torch tensors, computed by hand-written sm_100a CUDA
through the C ABI in include/lsh_attn.h.

Kept from the reference (names, argument meaning, error behaviour):
  * the 21 constructor keywords (EA:1732-1748); `share_qk` is forced on (EA:1782); dropouts are
    zeroed unless mode == 'train' (EA:1790-1795); `n_in = 2 if masked else 1` (EA:1750).
  * `weights = (w_q (H,D,dq), w_v (H,D,dv), w_o (H,dv,D))` fp32 (EA:1845-1868, stacked EA:1829-1830).
  * `state = (buckets int32 (B*H, nh*max(L, max_length_for_buckets)), rng uint32 (B*H, 2))`
    (EA:1870-1887, stacked EA:1831).
  * `init`, `init_weights_and_state`, `forward`, `has_backward`, `backward`,
    `forward_and_or_backward`, `pure_fn`, `__call__` with the semantics of `trax/layers/base.py:265`,
    `:541`, `:644` and EA:2111-2126, 2246-2561; only `weights`, `state`, `rng` are settable public
    attributes (base.py:675-706).
  * mode='predict' (fast inference, EA:1999-2109, 2174-2244): state `(mem_end, (mem,), (buckets, buckets_idx, rng))`
    (EA:1833-1841, 1883-1887), forward-only — trax_b200/predict.py.
Not provided (raise, never silently differ): `use_reference_code=True` (would be a CPU path), `bias=True` (broken in the
reference too: EA:1921 unpacks exactly three weights).  attention_dropout (the (chunk_len, window) keep matrix of EA:254-262, applied inside
the attention kernels) and output_dropout (a column scaling of w_o) are supported; their masks are functions of `rng`, not
jax.random's bits, and can be supplied explicitly.

torch plays the role JAX plays for the reference: device memory, streams, autograd glue
(`torch.autograd.Function` ≙ `fastmath.custom_vjp` in base.py:644-673).
"""
import ctypes
import weakref
from typing import NamedTuple

import numpy as np
import torch

from trax_b200 import _lib
from trax_b200 import ops


class ShapeDtype(NamedTuple):
  """Stand-in for `trax.shapes.ShapeDtype` (shapes.py:23): what `init` receives."""
  shape: tuple
  dtype: object = torch.float32


def _to_int32_bits(rng):
  """uint32 key material → int32 tensor with the same bits (torch has no general uint32 kernels)."""
  if isinstance(rng, torch.Tensor):
    if rng.dtype == torch.int32:
      return rng
    if hasattr(torch, 'uint32') and rng.dtype == torch.uint32:
      return rng.view(torch.int32)
    rng = rng.cpu().numpy()
  arr = np.asarray(rng).astype(np.uint32)
  return torch.from_numpy(arr.view(np.int32).copy())


def _split_host(key, n):
  """Host-side key derivation used at init time only (EA:1819, 1824 call fastmath.random.split).

  Not threefry-compatible with JAX (documented): a NumPy Philox stream keyed by the parent key.
  """
  key = np.asarray(key, dtype=np.uint32).reshape(-1)[:2]
  gen = np.random.Generator(np.random.Philox(key=int(key[0]) << 32 | int(key[1])))
  return gen.integers(0, 2 ** 32, size=(n, 2), dtype=np.uint64).astype(np.uint32)


def _split_rngs(rng, n):
  """`combinators._split_rngs` (trax/layers/combinators.py): one sub-key per sublayer, `(None,) * n` without a key.  The
  same function of `rng` wherever a combinator hands keys to its sublayers (reversible.py:297, 328; EA:3570), so a sublayer
  that draws from its key (output / attention dropout) draws the same mask in the forward and in the backward pass.
  Not jax.random's bits (see _split_host)."""
  if rng is None:
    return (None,) * n
  key = np.asarray(rng.cpu() if isinstance(rng, torch.Tensor) else rng).astype(np.uint32).reshape(-1)
  return tuple(_split_host(key, n))


class _HostIO:
  """Host <-> device plumbing of the layer's host-tensor path (what bench.py's `e2e` leg times).

  Two side streams per device keep both PCIe directions busy: uploads of the NEXT call's inputs and downloads of the
  PREVIOUS call's results run beside the kernels on the caller's stream.  With `set_async_host_io(True)` a call returns
  as soon as its work is enqueued: results are pinned host tensors that become valid after `trax_b200.synchronize()`
  (JAX-style asynchronous dispatch); the default is to wait before returning.  (Re-use of forward's device copy of x by
  the matching backward call is the LAYER's business — `LSHSelfAttention._x_stash` — not a cache in here.)
  """
  _per_device = {}
  async_mode = False

  def __init__(self, dev):
    self.dev = dev
    self.h2d = torch.cuda.Stream(device=dev)
    self.d2h = torch.cuda.Stream(device=dev)
    self.rings = {}
    self.h2d_bytes = 0
    self.d2h_bytes = 0

  @classmethod
  def get(cls, dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    io = cls._per_device.get(key)
    if io is None:
      io = cls._per_device[key] = cls(dev)
    return io

  def upload(self, t):
    """Host tensor -> device copy usable on the current stream (stream-ordered, no host wait)."""
    if t is None or t.is_cuda:
      return t
    main = torch.cuda.current_stream(self.dev)
    with torch.cuda.stream(self.h2d):
      d = t.to(self.dev, non_blocking=True)
      ev = self.h2d.record_event()
    main.wait_event(ev)
    d.record_stream(main)
    self.h2d_bytes += t.numel() * t.element_size()
    return d

  def _staging(self, d, role):
    """Pinned destination of a download.  Synchronous mode: a fresh tensor per call.  Asynchronous mode: a ring of
    three staging buffers per (role, shape, dtype) — re-acquiring a slot waits for its previous copy (back-pressure,
    bounded pinned memory); a returned tensor is overwritten by the third-next call of the same kind."""
    if not _HostIO.async_mode:
      return torch.empty(d.shape, dtype=d.dtype, device='cpu', pin_memory=True), None
    key = (role, tuple(d.shape), d.dtype)
    ring = self.rings.setdefault(key, {'bufs': [], 'events': [], 'next': 0})
    i = ring['next']
    ring['next'] = (i + 1) % 3
    if len(ring['bufs']) <= i:
      ring['bufs'].append(torch.empty(d.shape, dtype=d.dtype, device='cpu', pin_memory=True))
      ring['events'].append(None)
    elif ring['events'][i] is not None:
      ring['events'][i].synchronize()
    return ring['bufs'][i], (ring, i)

  def download(self, d, role='out'):
    """Device tensor -> pinned host tensor, copied on the download stream after the current stream's work so far."""
    main = torch.cuda.current_stream(self.dev)
    ev = main.record_event()
    h, slot = self._staging(d, role)
    with torch.cuda.stream(self.d2h):
      self.d2h.wait_event(ev)
      h.copy_(d, non_blocking=True)
      if slot is not None:
        slot[0]['events'][slot[1]] = self.d2h.record_event()
    d.record_stream(self.d2h)
    self.d2h_bytes += d.numel() * d.element_size()
    return h

  def finish(self):
    if not _HostIO.async_mode:
      self.d2h.synchronize()


def set_async_host_io(flag):
  """Host-tensor calls return without waiting for their device->host copies (see _HostIO)."""
  _HostIO.async_mode = bool(flag)


_REUSE_UPLOAD = {'on': True}


def set_reuse_forward_upload(flag):
  """`backward(x_host, ...)` right after `forward(x_host)` of the SAME host tensor object (weak reference identity and
  `_version`) re-uses the device copy forward made instead of crossing PCIe again (default on).  The stash lives on the
  layer, is consumed by the next `backward` and dropped by any other call.  Turn it off if host inputs are mutated in
  place between the two calls through a path torch cannot see (a NumPy view of a `torch.from_numpy` tensor does not bump
  `_version`)."""
  _REUSE_UPLOAD['on'] = bool(flag)


_GRAD_ALLREDUCE = {'on': False}


def set_weight_grad_allreduce(flag):
  """Data-parallel training: average (dw_q, dw_v, dw_o) over the default process group INSIDE the backward call, on
  the device, before they are returned / downloaded — the analogue of `psum(grads) / n` inside Trax's pmapped step
  (`trax/optimizers/trainer.py:172-199`).  Off by default (the caller then reduces the returned gradients itself).
  True / 'flat': one in-place all-reduce of the contiguous gradient buffer on the caller's stream after the call's last
  kernel.  'overlap': the two slices are reduced on a communication stream as soon as each is final
  (`dp.GradOverlap`) — underneath the remaining kernels; faster when the ranks run in step, but the collective's CTAs
  displace CTAs of the layer's persistent kernels for as long as they wait for the slowest rank."""
  if flag not in (False, True, None, 0, 1, 'flat', 'overlap'):
    raise ValueError("set_weight_grad_allreduce: False, True / 'flat' or 'overlap'")
  _GRAD_ALLREDUCE['on'] = 'overlap' if flag == 'overlap' else ('flat' if flag else False)


def synchronize():
  """Waits for every enqueued layer call, including the host copies of results (pairs with set_async_host_io)."""
  torch.cuda.synchronize()


def host_io_bytes(reset=False):
  """(h2d, d2h) bytes actually copied by host-tensor calls since the last reset."""
  h = sum(io.h2d_bytes for io in _HostIO._per_device.values())
  d = sum(io.d2h_bytes for io in _HostIO._per_device.values())
  if reset:
    for io in _HostIO._per_device.values():
      io.h2d_bytes = io.d2h_bytes = 0
  return h, d


class LSHSelfAttention:
  """LSH self-attention (see module docstring)."""

  def __init__(self,
               n_heads=2, d_qk=64, d_v=64, share_qk='unused',
               causal=False,
               masked=False,
               chunk_len=128, n_chunks_before=1, n_chunks_after=0,
               n_hashes=1,
               n_buckets=None,
               mode='train',
               predict_mem_len=2048, predict_drop_len=256,
               attention_dropout=0.0,
               output_dropout=0.0,
               max_length_for_buckets=None,
               bias=False,
               n_parallel_heads=1,
               use_python_loop=False,
               use_reference_code=False,
              ):
    del share_qk
    self._n_in = 2 if masked else 1                                 # EA:1750
    self._n_out = 1
    self._n_heads = n_heads
    if n_parallel_heads:                                            # EA:1753-1760
      if ((n_parallel_heads > n_heads and n_parallel_heads % n_heads != 0)
          or (n_parallel_heads < n_heads and n_heads % n_parallel_heads != 0)):
        raise ValueError('n_parallel_heads must be a multiple or fraction of n_heads')
    # All units run in one batched launch; n_parallel_heads / use_python_loop only shaped the
    # reference's memory use (EA:2297-2321) and do not change results.
    self._n_parallel_heads = n_parallel_heads or None
    self._use_python_loop = use_python_loop
    self._incremental = (mode == 'predict')                         # EA:1762-1766
    self._predict_hashes = True         # the predict state carries a bucket memory (SelfAttention: it does not)
    self._predict_mem_len, self._predict_drop_len = predict_mem_len, predict_drop_len
    if self._incremental:
      if masked or n_chunks_after:
        raise NotImplementedError("mode='predict' takes one input and no look-ahead (EA:1999-2001, 2085)")
      if not (isinstance(predict_mem_len, int) and isinstance(predict_drop_len, int) and 0 < predict_drop_len < predict_mem_len):
        raise ValueError("mode='predict' needs 0 < predict_drop_len < predict_mem_len")
    if use_reference_code:
      raise NotImplementedError('use_reference_code=True is a CPU loop in the reference; this build has no CPU path')
    if bias:
      raise NotImplementedError('bias=True is not supported (the reference unpacks 3 weights, EA:1921)')
    self._d_qk, self._d_v = d_qk, d_v
    self._share_qk = True                                           # EA:1782
    self._causal, self._masked = causal, masked
    self._chunk_len = chunk_len
    self._n_chunks_before, self._n_chunks_after = n_chunks_before, n_chunks_after
    self._bias = bias
    self._mode = mode
    if mode == 'train':                                             # EA:1790-1795
      self._attention_dropout, self._output_dropout = attention_dropout, output_dropout
    else:
      self._attention_dropout = self._output_dropout = 0.0
    if not 0.0 <= self._attention_dropout < 1.0:
      raise ValueError('attention_dropout must be in [0, 1)')
    if not 0.0 <= self._output_dropout < 1.0:
      raise ValueError('output_dropout must be in [0, 1)')
    self._n_hashes = n_hashes
    self._n_buckets = n_buckets
    self._max_length_for_buckets = max_length_for_buckets
    self._weights = ()
    self._state = ()
    self._rng = None
    self._separate_k = False            # SelfAttention(share_qk=False) sets it: weights (w_q, w_k, w_v, w_o), EA:1112-1128
    self._x_stash = None                # (weakref to a host x, its _version, device copy): forward -> matching backward
    self._rotations_override = None     # tests / a JAX host inject explicit rotations here
    self._out_keep_override = None      # likewise an explicit (d_model,) bool keep-mask for output dropout
    self._attn_keep_override = None     # and an explicit (chunk_len, window) bool keep-mask for attention dropout

  # ---- attribute discipline (base.py:675-706) -----------------------------------------------------
  def __setattr__(self, attr, value):
    if attr[0] != '_' and attr not in ('weights', 'state', 'rng'):
      raise ValueError("Trax layers only allow to set ('weights', 'state', 'rng') as public "
                       f'attribues, not {attr}.')
    super().__setattr__(attr, value)

  @property
  def n_in(self):
    return self._n_in

  @property
  def n_out(self):
    return self._n_out

  @property
  def weights(self):
    return self._weights

  @weights.setter
  def weights(self, w):
    self._weights = tuple(w) if isinstance(w, (list, tuple)) else w

  @property
  def state(self):
    return self._state

  @state.setter
  def state(self, s):
    self._state = tuple(s) if isinstance(s, (list, tuple)) else s

  @property
  def rng(self):
    if self._rng is None:
      self._rng = np.array([0, 0], dtype=np.uint32)                 # base.py: default key from seed 0
    return self._rng

  @rng.setter
  def rng(self, rng):
    self._rng = rng

  @property
  def has_backward(self):                                           # EA:2246-2249
    return True

  # ---- init (base.py:265-311, EA:1810-1887) -------------------------------------------------------
  def init(self, input_signature, rng=None, use_cache=False):
    del use_cache
    if rng is not None:
      self.rng = rng
    self.init_weights_and_state(input_signature)
    return self.weights, self.state

  def _kernel_initializer(self, shape, gen):                        # EA:1801-1808
    lim = np.sqrt(6.0 / (shape[0] + shape[1] * self._n_heads))
    return gen.uniform(-lim, lim, size=shape).astype(np.float32)

  def init_weights_and_state(self, input_signature, device=None):
    if not isinstance(input_signature, (tuple, list)) or isinstance(input_signature, ShapeDtype):
      input_signature = (input_signature,)
    shape = tuple(input_signature[0].shape)
    batch_size, seqlen, d_model = int(shape[0]), int(shape[1]), int(shape[2])
    device = device or ('cuda' if torch.cuda.is_available() else 'cpu')
    w_q, w_v, w_o = [], [], []
    weight_rngs = _split_host(self.rng, self._n_heads)              # EA:1819
    for i in range(self._n_heads):                                  # EA:1845-1868
      g = np.random.Generator(np.random.Philox(key=int(weight_rngs[i][0]) << 32 | int(weight_rngs[i][1])))
      w_q.append(self._kernel_initializer((d_model, self._d_qk), g))
      w_v.append(self._kernel_initializer((d_model, self._d_v), g))
      w_o.append(np.transpose(self._kernel_initializer((d_model, self._d_v), g)))
    state_rngs = _split_host(self.rng, self._n_heads * batch_size)  # EA:1824
    length = self._max_length_for_buckets or seqlen                 # EA:1880
    buckets = torch.zeros((self._n_heads * batch_size, self._n_hashes * length), dtype=torch.int32,
                          device=device)                            # EA:1881
    rng_state = _to_int32_bits(state_rngs).to(device)
    if hasattr(torch, 'uint32'):
      rng_state = rng_state.view(torch.uint32)
    self.weights = tuple(torch.from_numpy(np.ascontiguousarray(np.stack(w))).to(device) for w in (w_q, w_v, w_o))
    self.state = (buckets, rng_state)
    if self._incremental:                                           # EA:1833-1841, 1883-1887
      from trax_b200 import predict
      dtype = getattr(input_signature[0], 'dtype', torch.float32)
      self.state = predict.init_state(self, batch_size, d_model, dtype if isinstance(dtype, torch.dtype) else torch.float32,
                                      device, rng_state, self._predict_hashes)

  # ---- forward / backward (EA:2111-2126, 2251-2259) ------------------------------------------------
  def forward(self, inputs):
    weights, state, rng = self.weights, self.state, self.rng
    output, new_state, _, _ = self._forward_and_or_backward(
        inputs, weights, state, rng, compute_output=True, update_state=True, _stash='store')
    self.state = new_state
    return output

  def backward(self, inputs, output, grad, weights, state, new_state, rng=None, **kwargs):
    del output, state, kwargs
    _, _, inputs_grad, weights_grad = self._forward_and_or_backward(
        inputs, weights, new_state, rng, output_grad=grad, compute_output=False, update_state=False, _stash='consume')
    return inputs_grad, weights_grad

  def pure_fn(self, x, weights, state, rng, use_cache=False):       # base.py:541-600
    old_weights, old_state, old_rng = self.weights, self.state, self._rng
    self._rng = rng
    self.weights, self.state = weights, state
    try:
      outputs, s = self._do_custom_gradients(x)
    finally:
      self._rng = old_rng
      if not use_cache:
        self.weights, self.state = old_weights, old_state
    if use_cache:
      self.state = s
    return outputs, s

  def __call__(self, x, weights=None, state=None, rng=None):        # base.py:158-196
    weights = self.weights if weights is None else weights
    if state is not None:
      self.state = state
    outputs, new_state = self.pure_fn(x, weights, self.state, self.rng if rng is None else rng)
    self.state = new_state
    return outputs

  def _do_custom_gradients(self, x):
    """base.py:644-673 with torch.autograd.Function standing in for fastmath.custom_vjp."""
    layer = self
    have_single_input = not isinstance(x, (tuple, list))
    xs = (x,) if have_single_input else tuple(x)
    state, rng, weights = self.state, self._rng, self.weights

    holder = {}

    class _Fn(torch.autograd.Function):

      @staticmethod
      def forward(ctx, x0, *ws):
        inputs = x0 if have_single_input else (x0,) + xs[1:]
        out, new_state, _, _ = layer._forward_and_or_backward(
            inputs, tuple(ws), state, rng, compute_output=True, update_state=True, _stash='store')
        ctx.save_for_backward(x0, *ws)
        holder['new_state'] = new_state                             # residual (base.py:659)
        return out

      @staticmethod
      def backward(ctx, grad):
        x0, ws = ctx.saved_tensors[0], tuple(ctx.saved_tensors[1:])
        inputs = x0 if have_single_input else (x0,) + xs[1:]
        inputs_grad, weights_grad = layer.backward(
            inputs, None, grad.contiguous(), ws, state, holder['new_state'], rng)
        gx = inputs_grad if have_single_input else inputs_grad[0]
        return (gx,) + tuple(weights_grad)

    out = _Fn.apply(xs[0], *weights)
    return out, holder['new_state']

  # ---- the batched driver (EA:2261-2561) -----------------------------------------------------------
  def _output_multiplier(self, rng, d_model, dev):
    """keep / keep_prob of `apply_broadcasted_dropout` (EA:271-280) as a (d_model,) fp32 device tensor, or None.  The
    keep-mask is a deterministic function of `rng` (so the backward call, which gets the same rng, EA:2251-2259, re-draws
    the same mask) but not jax.random's bit stream — like the hash rotations, it can be supplied explicitly."""
    if not self._output_dropout:
      return None
    keep_prob = 1.0 - self._output_dropout
    if self._out_keep_override is not None:
      keep = torch.as_tensor(np.asarray(self._out_keep_override)).to(torch.bool).reshape(d_model)
    else:
      if rng is None:
        raise ValueError('output_dropout > 0 needs an rng (EA:274)')
      key = np.asarray(rng.cpu() if isinstance(rng, torch.Tensor) else rng).astype(np.uint32).reshape(-1)
      # (a Philox stream keyed by all 64 key bits: torch's CPU generator only looks at the low 32 bits of its seed)
      gen = np.random.Generator(np.random.Philox(key=((int(key[0]) << 32) | int(key[-1])) ^ 0x6f75745f64726f70))   # 'out_drop'
      keep = torch.from_numpy(gen.random(d_model) < keep_prob)
    return (keep.to(torch.float32) / keep_prob).to(dev)

  def _attention_multiplier(self, rng, dev):
    """keep / keep_prob of EA:254-262 as a (chunk_len, window) fp32 device tensor, or None: ONE matrix per call, shared by
    every chunk, head, example and hash round (`attend_rng` is the same for all units, EA:1920, 2334-2335).  Like the
    output-dropout mask it is a deterministic function of `rng` (the backward call re-draws the forward's mask) but not
    jax.random's bits; a JAX host supplies the matrix it drew itself (`_attn_keep_override`)."""
    if not self._attention_dropout:
      return None
    keep_prob = 1.0 - self._attention_dropout
    shape = (self._chunk_len, self._chunk_len * (1 + self._n_chunks_before + self._n_chunks_after))
    if self._attn_keep_override is not None:
      keep = torch.as_tensor(np.asarray(self._attn_keep_override)).to(torch.bool).reshape(shape)
    else:
      if rng is None:
        raise ValueError('attention_dropout > 0 needs an rng (EA:255-260)')
      key = np.asarray(rng.cpu() if isinstance(rng, torch.Tensor) else rng).astype(np.uint32).reshape(-1)
      gen = np.random.Generator(np.random.Philox(key=((int(key[0]) << 32) | int(key[-1])) ^ 0x6174746e5f647270))   # 'attn_drp'
      keep = torch.from_numpy(gen.random(shape) < keep_prob)
    return (keep.to(torch.float32) / keep_prob).contiguous().to(dev)

  def _dims(self, batch_size, seqlen, d_model, act_dtype):
    factors = ops.bucket_factors(self._n_buckets, seqlen, self._chunk_len)     # EA:1890-1902
    return _lib.make_dims(batch_size, self._n_heads, seqlen, d_model, self._d_qk, self._d_v,
                          self._chunk_len, self._n_chunks_before, self._n_chunks_after, self._n_hashes,
                          factors, self._causal, self._masked, act_dtype, separate_k=self._separate_k)

  def forward_and_or_backward(self, inputs, weights, state, rng, output_grad=None,
                              compute_output=True, update_state=True):
    """Performs batched forward and/or backward passes (EA:2261-2289); see _forward_and_or_backward."""
    return self._forward_and_or_backward(inputs, weights, state, rng, output_grad, compute_output, update_state)

  def _forward_and_or_backward(self, inputs, weights, state, rng, output_grad=None,
                               compute_output=True, update_state=True, _stash=None, _residual=None, _io_dtype=None,
                               _raw=False):
    """Performs batched forward and/or backward passes (EA:2261-2289).

    Returns (output, new_state, inputs_grad, weights_grad):
      output is not None iff compute_output; new_state iff update_state; grads iff output_grad given.
    Tensors may live on the GPU (no copies) or on the host (pinned or pageable): host inputs are
    copied to cuda:current, results copied back — the e2e path bench.py times.
    In predict mode the call goes to trax_b200/predict.py (EA:2336-2350); `_raw` is that module's way back to the
    training-path kernels for a prefix.
    """
    if self._incremental and not _raw:
      from trax_b200 import predict
      return predict.forward_and_or_backward(self, inputs, weights, state, rng, output_grad, compute_output, update_state)
    have_single_input = not isinstance(inputs, (tuple, list))
    if have_single_input:
      inputs = (inputs,)
    if len(inputs) != self._n_in:
      raise ValueError('LSHSelfAttention(masked=%s) takes %d inputs, got %d' % (self._masked, self._n_in, len(inputs)))
    x = inputs[0]
    compute_grad = output_grad is not None
    assert compute_output or compute_grad, 'No work to perform!'    # EA:2331
    if x.dim() != 3:
      raise ValueError('inputs[0] must have shape (batch, seqlen, d_model)')
    if not torch.cuda.is_available():
      raise _lib.LshAttnError('trax_b200.LSHSelfAttention needs a CUDA device (no CPU fallback)')
    if x.is_cuda and x.device.index != torch.cuda.current_device():
      # kernels, streams and the scratch buffer belong to the tensors' device, whatever the caller's current device is
      with torch.cuda.device(x.device):
        return self._forward_and_or_backward(inputs if not have_single_input else inputs[0], weights, state, rng, output_grad,
                                             compute_output, update_state, _stash, _residual, _io_dtype, _raw)
    lib = _lib.load()
    host_io = not x.is_cuda
    dev = torch.device('cuda', torch.cuda.current_device()) if host_io else x.device

    io = _HostIO.get(dev) if host_io else None

    def to_dev(t):
      if t is None or t.is_cuda:
        return t
      return _HostIO.get(dev).upload(t)
    # forward(x_host) -> backward(x_host, ...): the backward call may take over the device copy forward made of the very
    # same tensor object; every call drops whatever stash it finds (see set_reuse_forward_upload).
    stash, self._x_stash = self._x_stash, None
    if (host_io and _stash == 'consume' and stash is not None and _REUSE_UPLOAD['on'] and stash[0]() is x
        and stash[1] == x._version and stash[2].device == dev):
      x_d = stash[2]
    else:
      x_d = to_dev(x).contiguous()
    if host_io and _stash == 'store' and _REUSE_UPLOAD['on']:
      self._x_stash = (weakref.ref(x), x._version, x_d)
    mask_d = None
    if self._masked:
      mask_d = to_dev(inputs[1]).to(torch.uint8).contiguous()
    if len(weights) != (4 if self._separate_k else 3):
      raise ValueError('expected %d weight tensors, got %d' % (4 if self._separate_k else 3, len(weights)))
    w_k = None
    if self._separate_k:                                            # (w_q, w_k, w_v, w_o), EA:1126-1128
      w_q, w_k, w_v, w_o = (to_dev(w).to(torch.float32).contiguous() for w in weights)
    else:
      w_q, w_v, w_o = (to_dev(w).to(torch.float32).contiguous() for w in weights)
    out_mult = self._output_multiplier(rng, int(x_d.shape[2]), dev)
    attn_keep = self._attention_multiplier(rng, dev)
    if out_mult is not None:
      # EA:1995-1996: (o w_o) * m with m of shape (d_model,) shared by every position, head and example (EA:271-280, same
      # rng for all units EA:2334-2335) == o (w_o * m): the mask is folded into the packed weight in both passes and
      # into dw_o afterwards; the kernels are unchanged.
      w_o = w_o * out_mult
    buckets, hash_rng = state
    batch_size, seqlen, d_model = (int(s) for s in x_d.shape)
    if tuple(w_q.shape) != (self._n_heads, d_model, self._d_qk) or tuple(w_o.shape) != (self._n_heads, self._d_v, d_model):
      raise ValueError('weights do not match (n_heads, d_model, d_head) layout: %s %s %s'
                       % (tuple(w_q.shape), tuple(w_v.shape), tuple(w_o.shape)))
    # _io_dtype (ReversibleHalfResidual with f32 activations): the input arrives as bf16 straight from the LayerNorm kernel,
    # outputs and cotangents are f32 (dims.x_bf16, include/lsh_attn.h)
    io_dtype = x_d.dtype
    if _io_dtype is not None and _io_dtype != x_d.dtype:
      if not (x_d.dtype == torch.bfloat16 and _io_dtype == torch.float32):
        raise ValueError('_io_dtype: only bf16 input with f32 outputs is supported')
      io_dtype = torch.float32
    dims = self._dims(batch_size, seqlen, d_model, _lib.LSH_DTYPE_F32 if io_dtype == torch.float32 else _lib.LSH_DTYPE_BF16)
    dims.x_bf16 = 1 if io_dtype != x_d.dtype else 0
    _lib.check(lib.lsh_attn_check_dims(ctypes.byref(dims)), 'LSHSelfAttention')
    bh = batch_size * self._n_heads
    length = self._n_hashes * (self._max_length_for_buckets or seqlen)
    stream = ops._stream()

    new_state = None
    rotations = None
    if update_state:                                                # EA:1926-1937
      if self._rotations_override is not None:
        rotations = to_dev(self._rotations_override).to(torch.float32).contiguous()
        new_rng = hash_rng
      else:
        keys = to_dev(_to_int32_bits(hash_rng)).contiguous()
        rotations, new_keys = ops.make_rotations(dims, keys)        # split + normal (EA:1928, 92)
        new_rng = new_keys.view(torch.uint32) if hasattr(torch, 'uint32') else new_keys
      # the hash kernel writes every (round, position) entry; only the padding up to max_length_for_buckets needs zeros
      alloc = torch.zeros if length > self._n_hashes * seqlen else torch.empty
      buckets_d = alloc((bh, max(length, self._n_hashes * seqlen)), dtype=torch.int32, device=dev)
    else:                                                           # EA:1939-1941
      buckets_d = to_dev(buckets)
      if buckets_d.dtype != torch.int32 or buckets_d.dim() != 2 or buckets_d.shape[0] != bh \
          or buckets_d.shape[1] < self._n_hashes * seqlen or buckets_d.stride(1) != 1:
        raise ValueError('state buckets must be int32 of shape (B*H, >= n_hashes*seqlen), got %s %s'
                         % (tuple(buckets_d.shape), buckets_d.dtype))

    nbytes = lib.lsh_layer_workspace_bytes(ctypes.byref(dims), 1 if compute_grad else 0)
    if nbytes == 0:
      _lib.check(1, 'lsh_layer_workspace_bytes')
    ws = ops.workspace(dev, nbytes)

    out_d = None
    if compute_output:
      out_d = torch.empty(x_d.shape, dtype=io_dtype, device=dev)    # dtype of inputs[0], EA:2529-2530
    # residual of the enclosing reversible block, fused into the output projection's epilogue (reversible.py:318, 400):
    # out = residual + sign * attention_output
    res_d, res_sign = None, 1.0
    if _residual is not None and compute_output:
      res_d, res_sign = _residual
      if not res_d.is_cuda or res_d.shape != x_d.shape or res_d.dtype != io_dtype or not res_d.is_contiguous():
        raise ValueError('fused residual must be a contiguous device tensor shaped and typed like the input')
    inputs_grad = weights_grad = None
    if not compute_grad:
      _lib.check(lib.lsh_layer_fwd_res(
          ctypes.byref(dims), ops._ptr(x_d), ops._ptr(w_q), ops._ptr(w_v), ops._ptr(w_o), ops._ptr(w_k), ops._ptr(rotations),
          ops._ptr(mask_d), ops._ptr(attn_keep), ops._ptr(buckets_d), buckets_d.stride(0), ops._ptr(out_d), ops._ptr(res_d),
          ctypes.c_float(res_sign), ops._ptr(ws), ws.numel(), stream), 'lsh_layer_fwd')
    else:
      if update_state:
        # EA allows update_state together with output_grad; the hash must then run first.
        _lib.check(lib.lsh_layer_fwd(
            ctypes.byref(dims), ops._ptr(x_d), ops._ptr(w_q), ops._ptr(w_v), ops._ptr(w_o), ops._ptr(w_k), ops._ptr(rotations),
            ops._ptr(mask_d), ops._ptr(attn_keep), ops._ptr(buckets_d), buckets_d.stride(0),
            ops._ptr(out_d if out_d is not None else torch.empty(x_d.shape, dtype=io_dtype, device=dev)), ops._ptr(ws), ws.numel(), stream),
            'lsh_layer_fwd')
      g_d = to_dev(output_grad).to(io_dtype).contiguous()
      if g_d.shape != x_d.shape:
        raise ValueError('output_grad shape %s != input shape %s' % (tuple(g_d.shape), tuple(x_d.shape)))
      dx = torch.empty(x_d.shape, dtype=io_dtype, device=dev)
      # one contiguous gradient buffer (dw_q | dw_v | dw_o): the data-parallel mean below reduces slices of it in place
      n_q, n_v, n_o = w_q.numel() + (w_k.numel() if w_k is not None else 0), w_v.numel(), w_o.numel()
      dw_flat = torch.empty(n_q + n_v + n_o, dtype=torch.float32, device=dev)
      dw_q, dw_v = dw_flat[:w_q.numel()].view_as(w_q), dw_flat[n_q:n_q + n_v].view_as(w_v)
      dw_k = dw_flat[w_q.numel():n_q].view_as(w_k) if w_k is not None else None
      dw_o = dw_flat[n_q + n_v:].view_as(w_o)
      overlap = None
      if _GRAD_ALLREDUCE['on']:
        from trax_b200 import dp
      if _GRAD_ALLREDUCE['on'] == 'overlap':
        overlap = dp.GradOverlap.get(dev)         # None without an initialised NCCL group of more than one rank
      ev_o, ev_qv = (overlap.events() if overlap is not None else (None, None))
      _lib.check(lib.lsh_layer_bwd_res(
          ctypes.byref(dims), ops._ptr(x_d), ops._ptr(w_q), ops._ptr(w_v), ops._ptr(w_o), ops._ptr(w_k), ops._ptr(mask_d),
          ops._ptr(attn_keep), ops._ptr(buckets_d), buckets_d.stride(0), ops._ptr(g_d), ops._ptr(out_d), ops._ptr(dx), ops._ptr(dw_q),
          ops._ptr(dw_v), ops._ptr(dw_o), ops._ptr(dw_k), ops._ptr(ws), ws.numel(),
          ev_o.cuda_event if ev_o is not None and out_mult is None else None,
          ev_qv.cuda_event if ev_qv is not None else None, ops._ptr(res_d), ctypes.c_float(res_sign), stream), 'lsh_layer_bwd')
      if out_mult is not None:
        dw_o.mul_(out_mult)
      if overlap is not None:
        # psum(grads) / n (trainer.py:197-199) on the communication stream, started by the events lsh_layer_bwd recorded
        # when each gradient became final: dw_o travels under the attention-gradient kernels and the dw_q|dw_v / dx GEMMs,
        # dw_q|dw_v under the dx GEMM.  (With output dropout dw_o is rescaled after the call, so it goes last.)
        overlap.reduce(dw_flat, n_q + n_v, ev_o if out_mult is None else None, ev_qv)
      elif _GRAD_ALLREDUCE['on']:
        dp.allreduce_mean_flat_(dw_flat)          # (dw_q | dw_k | dw_v | dw_o are views of this buffer)
      if host_io:
        dx, dw_q, dw_v, dw_o = (io.download(t, r) for t, r in ((dx, 'dx'), (dw_q, 'dw_q'), (dw_v, 'dw_v'), (dw_o, 'dw_o')))
        if dw_k is not None:
          dw_k = io.download(dw_k, 'dw_k')
      inputs_grad = dx if have_single_input else (dx,) + (None,) * (len(inputs) - 1)
      weights_grad = (dw_q, dw_k, dw_v, dw_o) if dw_k is not None else (dw_q, dw_v, dw_o)
    if update_state:
      new_state = (buckets_d, new_rng)    # state stays on the device, like a jitted Trax layer's
    if compute_output and host_io:
      out_d = io.download(out_d)
    if host_io:
      io.finish()      # default: results are in pinned host memory when the call returns (see set_async_host_io)
    return out_d, new_state, inputs_grad, weights_grad
