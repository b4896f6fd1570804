"""Fast inference (`mode='predict'`) of `LSHSelfAttention` (EA:1999-2109, 2174-2244) and `SelfAttention` (EA:1200-1268).

The reference keeps, per layer, an input memory of `predict_mem_len` positions that new tokens are appended to (rolled by
`predict_drop_len` when full, `_use_predict_mem` EA:2174-2244) and — for the LSH layer — a bucket memory of the same length
per (example, head, hash round) (EA:1883-1887, rolled at EA:2036-2053).  State layout, as the reference stacks it
(EA:1829-1841):

    LSHSelfAttention:  (mem_end int32 (), (mem (B, M, D),), (buckets int32 (B*H, nh*M), buckets_idx int32 (B*H,), rng (B*H, 2)))
    SelfAttention:     (mem_end int32 (), (mem (B, M, D),), ())

`mem_end` and `buckets_idx` decide host control flow (roll or not, which branch), so they live on the HOST (0-d / 1-d CPU
int32 tensors); `mem`, `buckets`, `rng` live on the device.  The bookkeeping below is slicing and copies of device memory
(torch as the allocator / memcpy engine); the arithmetic is CUDA behind the C ABI:

  * a single new token (q_len == 1, EA:2032-2109): `lsh_predict_step` — projection of the memory (own tcgen05 GEMM), the
    training path's bit-exact hash for the new token's bucket ids, `predict_attend_kernel` (bucket-memory update, slot
    ranking, online softmax over the attended slots), `predict_out_kernel`;
  * a prefix (q_len > 1, only at the start of a sequence, EA:2004-2030): the training path's forward (`lsh_layer_fwd`) on the
    chunk-padded input, whose buckets are stored into the bucket memory.

Behaviour kept from the reference: states are values (a call returns NEW tensors and leaves the ones it was given alone);
`n_buckets=None` resolves per call from the number of rows hashed (EA:1893-1902: the padded prefix length, but 2 at a
single-token step); the hash rotations are a function of the state's key itself — the same at every call (EA:2014, 2066).
Where the reference "signals an error by introducing NaNs" (a long prefix that is not at the start, EA:2233-2243) or
silently computes on the wrong slots (several tokens appended after the start, EA:2005-2006 asserts only for Python ints),
this build raises.  No backward pass in predict mode (EA:2002-2003).
"""
import ctypes

import numpy as np
import torch

from trax_b200 import _lib
from trax_b200 import ops


# ---- memory bookkeeping (device-agnostic torch; checked on CPU against oracle/predict_oracle.py) --------------------------
def use_predict_mem(x, mem_end, mem, mem_len, drop_len):
  """`_use_predict_mem` (EA:2174-2244).  x (B, seqlen, D); mem (B, M, D); mem_end int.
  Returns (inputs, q_start, new_mem, new_mem_end): `inputs` is what the units attend over."""
  seqlen = int(x.shape[1])
  if seqlen <= drop_len and seqlen < mem_len:                        # EA:2179: a few tokens appended
    if mem_end + seqlen > mem_len:                                   # EA:2189-2198 roll_mem
      new_mem = torch.zeros_like(mem)
      new_mem[:, :mem_len - drop_len] = mem[:, drop_len:]
      mem_end -= drop_len
    else:
      new_mem = mem.clone()
    # EA:2199-2206; dynamic_update_slice clamps its start (jax.lax), and so does index_update's wrap for the one-token
    # case never trigger: mem_end + seqlen <= mem_len holds after the roll as long as drop_len >= seqlen
    start = max(0, min(mem_end, mem_len - seqlen))
    new_mem[:, start:start + seqlen] = x.to(new_mem.dtype)
    return new_mem, mem_end, new_mem, mem_end + seqlen
  if not (seqlen > drop_len or seqlen == mem_len):                   # EA:2209
    raise ValueError('predict mode: %d tokens at once need predict_drop_len < %d or == predict_mem_len' % (seqlen, seqlen))
  if mem_end != 0:                                                   # EA:2233-2243 (the reference returns NaNs)
    raise ValueError('predict mode: a prefix of %d tokens (> predict_drop_len = %d) is only valid at the start of a '
                     'sequence, but the memory already holds %d' % (seqlen, drop_len, mem_end))
  if seqlen >= mem_len:                                              # EA:2218-2222
    new_mem = x[:, seqlen - mem_len:].to(mem.dtype).clone()
  else:                                                              # EA:2223-2230
    new_mem = torch.zeros_like(mem)
    new_mem[:, :seqlen] = x.to(mem.dtype)
  return x, 0, new_mem, min(seqlen, mem_len)


def roll_buckets(buckets, buckets_idx, q_start, n_hashes, mem_len, drop_len):
  """EA:2036-2053: after the input memory rolled, shift the bucket memory by the same amount (a NEW tensor either way).
  buckets (BH, nh * M); buckets_idx, q_start ints."""
  if buckets_idx <= q_start:
    return buckets.clone()
  shift = min(buckets_idx - q_start, drop_len)                       # dynamic_slice_in_dim clamps so that the slice fits
  b3 = buckets.view(buckets.shape[0], n_hashes, mem_len)
  out = torch.zeros_like(b3)
  out[:, :, :mem_len - shift] = b3[:, :, shift:]
  return out.view(buckets.shape[0], n_hashes * mem_len)


def store_prefix_buckets(buckets, buckets_update, q_len, n_hashes, mem_len):
  """EA:2021-2028: the first q_len bucket ids of every round (the last `mem_len` of them if there are more) go to the start
  of the bucket memory.  buckets (BH, nh * M); buckets_update (BH, nh * padded_len)."""
  bh = buckets.shape[0]
  upd = buckets_update.view(bh, n_hashes, -1)[:, :, :q_len]
  if q_len > mem_len:
    upd = upd[:, :, q_len - mem_len:]
  out = buckets.clone().view(bh, n_hashes, mem_len)
  out[:, :, :upd.shape[2]] = upd
  return out.view(bh, n_hashes * mem_len)


def init_state(layer, batch_size, d_model, dtype, device, rng_state, with_buckets):
  """EA:1833-1841, 1883-1887."""
  m = layer._predict_mem_len
  mem = torch.zeros((batch_size, m, d_model), dtype=dtype, device=device)
  mem_end = torch.zeros((), dtype=torch.int32)
  if not with_buckets:
    return (mem_end, (mem,), ())
  bh = batch_size * layer._n_heads
  buckets = torch.zeros((bh, layer._n_hashes * m), dtype=torch.int32, device=device)
  return (mem_end, (mem,), (buckets, torch.zeros((bh,), dtype=torch.int32), rng_state))


def _unpack_state(layer, state, with_buckets):
  if not isinstance(state, (tuple, list)) or len(state) != 3:
    raise ValueError("predict mode: state must be (mem_end, mem, layer_state) (EA:1841); initialise the layer with mode='predict'")
  mem_end, mem, inner = state
  mem = mem[0] if isinstance(mem, (tuple, list)) else mem
  if mem.dim() != 3 or int(mem.shape[1]) != layer._predict_mem_len:
    raise ValueError('predict mode: memory of shape %s does not match predict_mem_len=%d' % (tuple(mem.shape), layer._predict_mem_len))
  if with_buckets:
    if len(inner) != 3:
      raise ValueError('predict mode: the LSH state is (buckets, buckets_idx, rng) (EA:1883-1887)')
    return int(mem_end), mem, inner
  return int(mem_end), mem, ()


# ---- the CUDA step -------------------------------------------------------------------------------------------------------
def _step(layer, mem, weights, q_start, buckets, rotations, causal):
  """One `lsh_predict_step` call: mem (B, M, D) with the new token stored at q_start; buckets (BH, nh*M) is updated IN
  PLACE (the caller passes a fresh tensor); rotations None = no hashing (SelfAttention).  Returns out (B, 1, D)."""
  lib = _lib.load()
  dev = mem.device
  batch_size, m, d_model = (int(s) for s in mem.shape)
  factors = ops.bucket_factors(layer._n_buckets, 2, layer._chunk_len) if rotations is not None else [2]   # EA:2064-2066: 2 rows
  dims = _lib.make_dims(batch_size, layer._n_heads, m, d_model, layer._d_qk, layer._d_v, layer._chunk_len,
                        layer._n_chunks_before, 0, layer._n_hashes, factors, causal, False, ops._act_dtype(mem),
                        separate_k=layer._separate_k)
  w_k = None
  if layer._separate_k:
    w_q, w_k, w_v, w_o = (w.to(dev).to(torch.float32).contiguous() for w in weights)
  else:
    w_q, w_v, w_o = (w.to(dev).to(torch.float32).contiguous() for w in weights)
  nbytes = lib.lsh_predict_workspace_bytes(ctypes.byref(dims))
  if nbytes == 0:
    _lib.check(1, 'lsh_predict_workspace_bytes')
  ws = ops.workspace(dev, nbytes)
  out = torch.empty((batch_size, 1, d_model), dtype=mem.dtype, device=dev)
  _lib.check(lib.lsh_predict_step(
      ctypes.byref(dims), ops._ptr(mem), ops._ptr(w_q), ops._ptr(w_v), ops._ptr(w_o), ops._ptr(w_k), ops._ptr(rotations),
      ops._ptr(buckets), buckets.stride(0) if buckets is not None else 0, ctypes.c_int32(q_start), ops._ptr(out), ops._ptr(ws),
      ws.numel(), ops._stream()), 'lsh_predict_step')
  return out


def _step_rotations(layer, batch_size, device, hash_rng):
  """The rotations `hash_vectors(q, hash_rng)` draws at a single-token step (EA:2066): a function of the state's key and of
  the shape (d_qk, n_hashes, R) only — the same at every step, and the same as the prefix call's when `n_buckets` is given."""
  if layer._rotations_override is not None:
    return layer._rotations_override.to(device).to(torch.float32).contiguous()
  from trax_b200.lsh_attention import _to_int32_bits
  factors = ops.bucket_factors(layer._n_buckets, 2, layer._chunk_len)
  dims = _lib.make_dims(batch_size, layer._n_heads, layer._predict_mem_len, 64, layer._d_qk, layer._d_v, layer._chunk_len,
                        layer._n_chunks_before, 0, layer._n_hashes, factors, True, False, _lib.LSH_DTYPE_BF16)
  keys = _to_int32_bits(hash_rng).to(device).contiguous()
  return ops.make_rotations(dims, keys)[0]


def _pure_step(layer, qk_mem, v_mem, q_start, buckets, rotations):
  """One `lsh_predict_attend` call for the weight-less core (EA:2858-2932): qk_mem / v_mem (B*H, M, d_head) with the new
  token stored at q_start; buckets (BH, nh*M) updated IN PLACE.  Returns out (B*H, 1, d_v) in the memory's dtype."""
  lib = _lib.load()
  dev = qk_mem.device
  bh, m = int(qk_mem.shape[0]), int(qk_mem.shape[1])
  io_dtype = qk_mem.dtype if qk_mem.dtype in (torch.float32, torch.bfloat16) else torch.float32
  # (B*H, M, d) x 2 -> (B, M, H, [qk | v]) bf16, the row layout the kernels read (one launch)
  qv = ops.pack_heads(qk_mem.to(io_dtype).contiguous(), v_mem.to(io_dtype).contiguous(), layer._n_heads)
  factors = ops.bucket_factors(layer._n_buckets, 2, layer._chunk_len)                                # EA:2886-2890: 2 rows
  dims = _lib.make_dims(bh // layer._n_heads, layer._n_heads, m, 64, layer._d_qk, layer._d_v, layer._chunk_len,
                        layer._n_chunks_before, 0, layer._n_hashes, factors, True, False, _lib.LSH_DTYPE_BF16)
  nbytes = lib.lsh_predict_attend_workspace_bytes(ctypes.byref(dims))
  if nbytes == 0:
    _lib.check(1, 'lsh_predict_attend_workspace_bytes')
  ws = ops.workspace(dev, nbytes)
  o = torch.empty((bh, layer._d_v), dtype=torch.float32, device=dev)
  _lib.check(lib.lsh_predict_attend(ctypes.byref(dims), ops._ptr(qv), ops._ptr(rotations), ops._ptr(buckets), buckets.stride(0),
                                    ctypes.c_int32(q_start), ops._ptr(o), ops._ptr(ws), ws.numel(), ops._stream()),
             'lsh_predict_attend')
  return o.view(bh, 1, layer._d_v).to(qk_mem.dtype)


def _run(layer, x_d, weights, mem_end, mem, inner, rng, step=None, train=None):
  """The call's control flow on tensors of ONE device (EA:2336-2350 + 1999-2109 / 1200-1268).  `step` / `train` default to
  the CUDA step and the layer's training-path forward; tests/test_predict_host.py swaps them for the oracle to check this
  function on the CPU.  Returns (output, (new_mem_end, new_mem, new_inner))."""
  step = step or _step
  train = train or (lambda x, w, st: layer._forward_and_or_backward(x, w, st, rng, compute_output=True, update_state=True,
                                                                    _raw=True))
  with_buckets = layer._predict_hashes
  seqlen = int(x_d.shape[1])
  m, drop = layer._predict_mem_len, layer._predict_drop_len
  att_in, q_start, new_mem, new_mem_end = use_predict_mem(x_d, mem_end, mem, m, drop)
  if with_buckets:
    buckets, buckets_idx, hash_rng = inner
    idx0 = int(buckets_idx.reshape(-1)[0]) if buckets_idx.numel() else 0
  if seqlen == 1:
    if with_buckets:
      new_buckets = roll_buckets(buckets, idx0, q_start, layer._n_hashes, m, drop)
      out = step(layer, new_mem, weights, q_start, new_buckets, _step_rotations(layer, int(new_mem.shape[0]), new_mem.device, hash_rng), True)
      new_inner = (new_buckets, torch.full_like(buckets_idx, q_start + 1), hash_rng)                 # EA:2108
    else:
      out = step(layer, new_mem, weights, q_start, None, None, layer._causal)
      new_inner = ()
  else:
    if q_start != 0:                                               # EA:2005-2006
      raise ValueError('predict mode: more than one token at a time only works at the start of a sequence '
                       '(the memory already holds %d)' % q_start)
    att_in = att_in.contiguous()
    rows = int(att_in.shape[1])
    if with_buckets:
      pad = (-rows) % layer._chunk_len                             # EA:2007-2011
      if pad:
        att_in = torch.cat([att_in, att_in.new_zeros((att_in.shape[0], pad, att_in.shape[2]))], dim=1)
      out, (buckets_update, _), _, _ = train(att_in, weights, (None, hash_rng))                      # EA:2012-2018
      out = out[:, :seqlen].contiguous()
      # (a training-path bucket row may be longer than n_hashes * rows: `max_length_for_buckets`, EA:1880)
      buckets_update = buckets_update[:, :layer._n_hashes * int(att_in.shape[1])].contiguous()
      new_buckets = store_prefix_buckets(buckets, buckets_update, seqlen, layer._n_hashes, m)
      new_inner = (new_buckets, buckets_idx + seqlen, hash_rng)                                       # EA:2030
    elif seqlen > layer._chunk_len and not getattr(layer, '_dense', False):   # EA:1250-1261: whole chunks of the prefix itself
      if seqlen % layer._chunk_len or rows != seqlen:
        raise ValueError('predict mode: a prefix longer than chunk_len must be a multiple of it and longer than '
                         'predict_drop_len (EA:1250-1252, 2209)')
      out, _, _, _ = train(att_in, weights, ())
      new_inner = ()
    else:                                                          # EA:1262-1267: each new token against every slot
      if not layer._causal:
        raise NotImplementedError('predict mode: a non-causal prefix of at most chunk_len tokens is not built')
      out = torch.cat([step(layer, new_mem, weights, j, None, None, True) for j in range(seqlen)], dim=1)
      new_inner = ()
  return out, (new_mem_end, new_mem, new_inner)


def forward_and_or_backward(layer, inputs, weights, state, rng, output_grad=None, compute_output=True, update_state=True):
  """`forward_and_or_backward` in predict mode (EA:2333-2350, 2554-2555).  Returns (output, new_state, None, None)."""
  if output_grad is not None or not update_state:
    raise NotImplementedError('predict mode is forward-only with update_state=True (EA:2002-2003)')
  x = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
  if x.dim() != 3:
    raise ValueError('inputs[0] must have shape (batch, seqlen, d_model)')
  if not torch.cuda.is_available():
    raise _lib.LshAttnError('trax_b200 predict mode needs a CUDA device (no CPU fallback)')
  mem_end, mem, inner = _unpack_state(layer, state, layer._predict_hashes)
  host_io = not x.is_cuda
  dev = mem.device if mem.is_cuda else (x.device if x.is_cuda else torch.device('cuda', torch.cuda.current_device()))
  with torch.cuda.device(dev):
    if layer._predict_hashes:
      inner = (inner[0].to(dev), inner[1], inner[2])
    out, (new_mem_end, new_mem, new_inner) = _run(layer, x.to(dev), weights, mem_end, mem.to(dev), inner, rng)
    new_state = (torch.tensor(new_mem_end, dtype=torch.int32), (new_mem,), new_inner)
    if host_io:
      out = out.cpu()
  return (out if compute_output else None), new_state, None, None


def _run_pure(layer, qk, v, mem_end, mems, inner, rng, step=None, train=None):
  """`_run` for the weight-less core (EA:3122-3146 + 2823-2932): the memory is the pair (qk_mem, v_mem), both rolled and
  updated alike (EA:2955-3033).  Returns (output (B*H, seqlen, d_v), (new_mem_end, (qk_mem, v_mem), new_inner))."""
  step = step or _pure_step
  train = train or (lambda inputs, st: layer.forward_and_or_backward(inputs, st, rng, compute_output=True, update_state=True,
                                                                     _raw=True))
  seqlen = int(qk.shape[1])
  m, drop = layer._predict_mem_len, layer._predict_drop_len
  att_qk, q_start, new_qk_mem, new_mem_end = use_predict_mem(qk, mem_end, mems[0], m, drop)
  att_v, _, new_v_mem, _ = use_predict_mem(v, mem_end, mems[1], m, drop)
  buckets, buckets_idx, hash_rng = inner
  idx0 = int(buckets_idx.reshape(-1)[0]) if buckets_idx.numel() else 0
  if seqlen == 1:
    new_buckets = roll_buckets(buckets, idx0, q_start, layer._n_hashes, m, drop)
    rotations = _step_rotations(layer, int(qk.shape[0]) // layer._n_heads, new_qk_mem.device, hash_rng)
    out = step(layer, new_qk_mem, new_v_mem, q_start, new_buckets, rotations)
    new_inner = (new_buckets, torch.full_like(buckets_idx, q_start + 1), hash_rng)                   # EA:2931
  else:
    if q_start != 0:                                               # EA:2831-2832
      raise ValueError('predict mode: more than one token at a time only works at the start of a sequence '
                       '(the memory already holds %d)' % q_start)
    rows = int(att_qk.shape[1])
    pad = (-rows) % layer._chunk_len                               # EA:2833-2838
    if pad:
      att_qk = torch.cat([att_qk, att_qk.new_zeros((att_qk.shape[0], pad, att_qk.shape[2]))], dim=1)
      att_v = torch.cat([att_v, att_v.new_zeros((att_v.shape[0], pad, att_v.shape[2]))], dim=1)
    out, (buckets_update, _), _ = train((att_qk.contiguous(), att_v.contiguous()), (None, hash_rng))  # EA:2839-2845
    out = out[:, :seqlen].contiguous()
    buckets_update = buckets_update[:, :layer._n_hashes * int(att_qk.shape[1])].contiguous()
    new_buckets = store_prefix_buckets(buckets, buckets_update, seqlen, layer._n_hashes, m)
    new_inner = (new_buckets, buckets_idx + seqlen, hash_rng)                                         # EA:2857
  return out.to(qk.dtype), (new_mem_end, (new_qk_mem, new_v_mem), new_inner)


def pure_init_state(layer, batch_x_heads, d_qk, d_v, dtype, device, rng_state):
  """EA:2677-2686, 2700-2708: memories for qk and v, bucket memory, counters."""
  m = layer._predict_mem_len
  mems = (torch.zeros((batch_x_heads, m, d_qk), dtype=dtype, device=device),
          torch.zeros((batch_x_heads, m, d_v), dtype=dtype, device=device))
  buckets = torch.zeros((batch_x_heads, layer._n_hashes * m), dtype=torch.int32, device=device)
  return (torch.zeros((), dtype=torch.int32), mems, (buckets, torch.zeros((batch_x_heads,), dtype=torch.int32), rng_state))


def pure_forward_and_or_backward(layer, inputs, state, rng, output_grad=None, compute_output=True, update_state=True):
  """`PureLSHSelfAttention.forward_and_or_backward` in predict mode (EA:3122-3146).  Returns (output, new_state, None)."""
  if output_grad is not None or not update_state:
    raise NotImplementedError('predict mode is forward-only with update_state=True (EA:2828-2829, 3125)')
  qk, v = inputs[0], inputs[1]
  if not isinstance(state, (tuple, list)) or len(state) != 3 or len(state[1]) != 2 or len(state[2]) != 3:
    raise ValueError("predict mode: state must be (mem_end, (qk_mem, v_mem), (buckets, buckets_idx, rng)) (EA:2677-2686)")
  if not qk.is_cuda:
    raise ValueError('PureLSHSelfAttention takes device tensors (its caller holds the projections on the device)')
  mem_end, mems, inner = state
  dev = qk.device
  with torch.cuda.device(dev):
    out, (new_mem_end, new_mems, new_inner) = _run_pure(
        layer, qk, v.to(dev), int(mem_end), (mems[0].to(dev), mems[1].to(dev)), (inner[0].to(dev), inner[1], inner[2]), rng)
  new_state = (torch.tensor(new_mem_end, dtype=torch.int32), new_mems, new_inner)
  return (out if compute_output else None), new_state, None


def rotations_shape(layer, n_rows):
  """(d_qk, n_hashes, R) of the rotations a call that hashes `n_rows` rows draws (EA:79-91, 1893-1902)."""
  return (layer._d_qk, layer._n_hashes, sum(ops.bucket_factors(layer._n_buckets, n_rows, layer._chunk_len)) // 2)


__all__ = ['use_predict_mem', 'roll_buckets', 'store_prefix_buckets', 'init_state', 'pure_init_state', 'forward_and_or_backward',
           'pure_forward_and_or_backward', 'rotations_shape']
del np
