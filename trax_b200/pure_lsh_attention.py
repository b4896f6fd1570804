"""`trax.layers.research.efficient_attention.PureLSHSelfAttention` (EA:2564-3265) on the same sm_100a kernels.

The weight-less core of the LSH layer: inputs `(qk, v)` — or `(qk, v, mask)` when `masked` — of
shape `(batch * n_heads, seqlen, d_head)`, no weights, state `(buckets, rng)` like `LSHSelfAttention`, output
`(batch * n_heads, seqlen, d_v)`.  `forward_unbatched` (EA:2739-2826) is `LSHSelfAttention.forward_unbatched` without the
three projections and without output dropout (EA:2822-2824): hash `qk`, sort, chunked shared-QK attention, un-sort,
multi-round combine.  Everything runs through the stage entry points of the C ABI (`lsh_hash`, `lsh_sort`,
`lsh_attend_fwd`, `lsh_combine_fwd`, `lsh_attend_bwd`); torch only re-packs the two inputs into the kernels'
`(B, L, H, [q | v])` row layout and back.  There is no CPU fallback.

Interface kept from the reference: constructor keywords (EA:2567-2586), `n_in = 3 if masked else 2` (EA:2589),
`init_weights_and_state` (EA:2648-2689: weights `()`, per-unit state), `forward` (EA:2935-2953), `backward`
(EA:3035-3050: returns `(inputs_grad, weights_grad = ())`), and
`forward_and_or_backward(inputs, state, rng, output_grad=None, compute_output=True, update_state=True)`
→ `(output, new_state, inputs_grad)` (EA:3052-3265; note: three results, no weights).
"""
import ctypes
import math

import numpy as np
import torch

from trax_b200 import _lib, ops
from trax_b200.lsh_attention import LSHSelfAttention, ShapeDtype, _split_host, _split_rngs, _to_int32_bits


class PureLSHSelfAttention(LSHSelfAttention):
  """LSH self-attention without weights (EA:2564)."""

  def __init__(self, *args, **kwargs):
    super().__init__(*args, **kwargs)
    self._n_in = 3 if self._masked else 2                           # EA:2587-2589

  # ---- init (EA:2648-2689) -------------------------------------------------------------------------
  def init_weights_and_state(self, input_signature, device=None):
    expected = 3 if self._masked else 2
    if not isinstance(input_signature, (tuple, list)) or len(input_signature) != expected \
        or isinstance(input_signature, ShapeDtype):
      raise ValueError(f'input_signature should be {expected}-tuple, but is: {input_signature}')   # EA:2651-2655
    shape = tuple(input_signature[0].shape)
    batch_x_heads, seqlen = int(shape[0]), int(shape[1])
    if batch_x_heads % self._n_heads != 0:                          # EA:2664
      raise ValueError('leading dimension %d is not a multiple of n_heads=%d' % (batch_x_heads, self._n_heads))
    device = device or ('cuda' if torch.cuda.is_available() else 'cpu')
    state_rngs = _split_host(self.rng, batch_x_heads)               # EA:2671
    length = self._max_length_for_buckets or seqlen                 # EA:2701
    buckets = torch.zeros((batch_x_heads, self._n_hashes * length), dtype=torch.int32, device=device)
    rng_state = _to_int32_bits(state_rngs).to(device)
    if hasattr(torch, 'uint32'):
      rng_state = rng_state.view(torch.uint32)
    self.weights = ()                                               # EA:2689
    self.state = (buckets, rng_state)
    if self._incremental:                                           # EA:2677-2686
      from trax_b200 import predict
      dtype = getattr(input_signature[0], 'dtype', torch.float32)
      self.state = predict.pure_init_state(self, batch_x_heads, int(shape[2]), int(tuple(input_signature[1].shape)[2]),
                                           dtype if isinstance(dtype, torch.dtype) else torch.float32, device, rng_state)

  # ---- forward / backward (EA:2935-2953, 3035-3050) --------------------------------------------------
  def forward(self, inputs):
    output, new_state, _ = self.forward_and_or_backward(inputs, self.state, self.rng, compute_output=True,
                                                        update_state=True)
    self.state = new_state
    return output

  def backward(self, inputs, output, grad, weights, state, new_state, rng=None, **kwargs):
    del output, state, kwargs
    _, _, inputs_grad = self.forward_and_or_backward(inputs, new_state, rng, output_grad=grad, compute_output=False,
                                                     update_state=False)
    return inputs_grad, ()                                          # zeros_like(()) == ()

  def pure_fn(self, inputs, weights, state, rng, use_cache=False):
    del weights, use_cache
    out, new_state, _ = self.forward_and_or_backward(inputs, state, rng, compute_output=True, update_state=True)
    return out, new_state

  # ---- the batched driver (EA:3052-3265) -------------------------------------------------------------
  def forward_and_or_backward(self, inputs, state, rng, output_grad=None, compute_output=True, update_state=True, _raw=False):
    """Returns (output, new_state, inputs_grad): output iff compute_output, new_state iff update_state,
    inputs_grad = (dqk, dv[, None for the mask]) iff output_grad is given.  In predict mode (EA:3122-3146) the call goes
    to trax_b200/predict.py; `_raw` is that module's way back to the training-path kernels for a prefix."""
    if not isinstance(inputs, (tuple, list)) or len(inputs) != self._n_in:
      raise ValueError('PureLSHSelfAttention(masked=%s) takes %d inputs' % (self._masked, self._n_in))
    if self._incremental and not _raw:
      if not torch.cuda.is_available():
        raise _lib.LshAttnError('trax_b200.PureLSHSelfAttention needs a CUDA device (no CPU fallback)')
      from trax_b200 import predict
      return predict.pure_forward_and_or_backward(self, inputs, state, rng, output_grad, compute_output, update_state)
    compute_grad = output_grad is not None
    assert compute_output or compute_grad, 'No work to perform!'
    if not torch.cuda.is_available():
      raise _lib.LshAttnError('trax_b200.PureLSHSelfAttention needs a CUDA device (no CPU fallback)')
    qk, v = inputs[0], inputs[1]
    if qk.dim() != 3 or v.shape != qk.shape[:2] + (self._d_v,) or qk.shape[2] != self._d_qk:
      raise ValueError('qk / v must have shape (batch*heads, seqlen, d_head); got %s %s' % (tuple(qk.shape), tuple(v.shape)))
    if not qk.is_cuda:
      raise ValueError('PureLSHSelfAttention takes device tensors (its caller holds the projections on the device)')
    bh, seqlen = int(qk.shape[0]), int(qk.shape[1])
    if bh % self._n_heads != 0:
      raise ValueError('leading dimension %d is not a multiple of n_heads=%d' % (bh, self._n_heads))
    batch = bh // self._n_heads
    dev = qk.device
    dims = self._dims(batch, seqlen, 64, _lib.LSH_DTYPE_BF16)
    _lib.check(_lib.load().lsh_attn_check_dims(ctypes.byref(dims)), 'PureLSHSelfAttention')
    attn_keep = self._attention_multiplier(rng, dev)                # EA:254-262 keep matrix (None without dropout)
    # (B*H, L, d) x 2  ->  (B, L, H, [q | v]) bf16: the row layout every kernel gathers from
    io_dtype = qk.dtype if qk.dtype in (torch.float32, torch.bfloat16) else torch.float32
    qv = ops.pack_heads(qk.to(io_dtype).contiguous(), v.to(device=dev, dtype=io_dtype).contiguous(), self._n_heads)
    mask_d = inputs[2].to(device=dev, dtype=torch.uint8).contiguous() if self._masked else None
    buckets, hash_rng = state
    length = self._n_hashes * (self._max_length_for_buckets or seqlen)

    new_state = None
    if update_state:                                                # EA:2749-2761
      if self._rotations_override is not None:
        rotations = self._rotations_override.to(device=dev, dtype=torch.float32).contiguous()
        new_rng = hash_rng
      else:
        keys = _to_int32_bits(hash_rng).to(dev).contiguous()
        rotations, new_keys = ops.make_rotations(dims, keys)
        new_rng = new_keys.view(torch.uint32) if hasattr(torch, 'uint32') else new_keys
      buckets_d = torch.zeros((bh, max(length, self._n_hashes * seqlen)), dtype=torch.int32, device=dev)
      ops.hash_qv(dims, qv, rotations, mask=mask_d, buckets=buckets_d)
      new_state = (buckets_d, new_rng)
    else:                                                           # EA:2762-2764
      buckets_d = buckets.to(dev)
      if buckets_d.dtype != torch.int32 or buckets_d.dim() != 2 or buckets_d.shape[0] != bh \
          or buckets_d.shape[1] < self._n_hashes * seqlen or buckets_d.stride(1) != 1:
        raise ValueError('state buckets must be int32 of shape (B*H, >= n_hashes*seqlen), got %s %s'
                         % (tuple(buckets_d.shape), buckets_d.dtype))

    sticker, _ = ops.sort(dims, buckets_d, want_undo=False)         # EA:2766-2778
    o_rounds, logits = ops.attend_fwd(dims, qv, sticker, mask=mask_d, attn_keep=attn_keep)   # EA:2780-2808 (un-sorted rows)
    o_comb, lse_tot = ops.combine_fwd(dims, o_rounds, logits)       # EA:2810-2814

    def unpack(t, col0, d):                                         # (B, L, H, :) columns -> (B*H, L, d) in the input dtype
      return ops.unpack_heads(t, col0, d, io_dtype).to(qk.dtype)
    output = unpack(o_comb, 0, self._d_v) if compute_output else None
    inputs_grad = None
    if compute_grad:
      do = ops.pack_heads(output_grad.to(device=dev, dtype=io_dtype).contiguous(), None, self._n_heads)
      dqv = ops.attend_bwd(dims, qv, sticker, o_comb, lse_tot, do, mask=mask_d, attn_keep=attn_keep)
      inputs_grad = (unpack(dqv, 0, self._d_qk), unpack(dqv, self._d_qk, self._d_v))
      if self._masked:
        inputs_grad = inputs_grad + (None,)
    return output, new_state, inputs_grad


def _rotary_tables(seqlen, d, device):
  """cos / sin of research/rotary_positional_embedding.py:29-35, (L, d) fp32."""
  inv_freq = torch.exp(torch.arange(0, d, 2, device=device, dtype=torch.float32) * -(math.log(10000.0) / d))
  freqs = torch.arange(seqlen, device=device, dtype=torch.float32)[:, None] * inv_freq[None, :]
  emb = torch.cat((freqs, freqs), dim=-1)
  return torch.cos(emb), torch.sin(emb)


def _rotate(x, cos, sin):
  """rotary_positional_embedding.py:37-44 on (B, L, d)."""
  half = x.shape[-1] // 2
  rot_half = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
  return (x.float() * cos + rot_half.float() * sin).to(x.dtype)


def _rotate_vjp(g, cos, sin):
  """Transpose of `_rotate` (it is linear): g cos + (u2, -u1) with u = g sin."""
  half = g.shape[-1] // 2
  u = g.float() * sin
  return (g.float() * cos + torch.cat((u[..., half:], -u[..., :half]), dim=-1)).to(g.dtype)


class PureLSHSelfAttentionWrapper:
  """`Serial(_ProjectAndSplitHeads, PureLSHSelfAttention, MergeHeads, Dense)` (EA:3493-3540) with the reference's
  hand-scheduled `forward_and_or_backward` (EA:3542-3620): Dense projections of x (B, L, d_model = n_heads * d_qk) to
  qk and v — two weights (EA:3360-3372) or three with qk = (q + k) / 2 (EA:3294-3312) — optional rotary embedding of
  q / k, head split to (B * n_heads, L, d_head), the weight-less LSH core above, head merge and the output Dense.

  Only `weights_format='model'` is built; `'sparse'` (EA:3314-3358) needs `sparsity.FactoredDense` / `LocallyConvDense`,
  which are outside the path.  The projections are plain GEMMs (cuBLAS through torch); everything between them is the
  CUDA path of `PureLSHSelfAttention`, which refuses host tensors and a missing GPU — there is no CPU fallback here
  either (`pure_lsh_implementation` is the reference's own injection point, EA:3504; tests use it to check this host
  logic against the oracle without a GPU).

  Weights / state keep Serial's one-slot-per-sublayer layout that `forward_and_or_backward` indexes (EA:3571-3586):
  `weights = (qkv, (), (), dense)`, `state = ((), (buckets, rng), (), ())`; `qkv` is a tuple of `num_weights` Dense
  weights, each `(w, b)` (`bias=True`) or a bare `w` of shape (d_model, d_model) (core.py:79-98, 100-118).
  """

  def __init__(self, n_heads=1, d_qk=64, d_v=64, causal=False, masked=False, output_dropout=0.0, attention_dropout=0.0,
               pure_lsh_implementation=None, bias=True, mode='train', num_weights=3, sparsity=16, weights_format='model',
               rotary_position_emb=False, **pure_lsh_implementation_kwargs):
    del sparsity
    assert weights_format in ('heads', 'model', 'sparse')            # EA:3290
    if weights_format != 'model':
      raise NotImplementedError("PureLSHSelfAttentionWrapper: only weights_format='model' is built (EA:3294-3312, "
                                "3360-3372); 'sparse' needs the sparsity layers, 'heads' is a TODO in the reference (EA:3376)")
    if num_weights not in (2, 3):
      raise ValueError('num_weights must be 2 (qk, v) or 3 (q, k, v)')
    if d_v != d_qk:
      raise ValueError('the v projection is Dense(d_model) split into n_heads (EA:3368-3371): d_v must equal d_qk')
    impl = pure_lsh_implementation or PureLSHSelfAttention
    self._attn = impl(n_heads=n_heads, d_qk=d_qk, d_v=d_v, causal=causal, masked=masked, mode=mode,
                      output_dropout=output_dropout, attention_dropout=attention_dropout,
                      **pure_lsh_implementation_kwargs)                # EA:3521-3530
    self._n_heads, self._d_model = n_heads, d_qk * n_heads            # EA:3511
    self._bias, self._num_weights, self._rotary, self._masked = bias, num_weights, rotary_position_emb, masked
    self._n_in = 2 if masked else 1
    self.weights, self.state, self.rng = (), (), None

  @property
  def n_in(self):
    return self._n_in

  @property
  def n_out(self):
    return 1

  @property
  def has_backward(self):
    return True

  @property
  def sublayers(self):
    return (None, self._attn, None, None)

  # ---- init: Dense weights (core.py:100-118: Glorot-uniform kernel, N(0, 1e-6) bias), core state -------------------
  def init(self, input_signature, rng=None, use_cache=False):
    del use_cache
    if rng is not None:
      self.rng = rng
    self.init_weights_and_state(input_signature)
    return self.weights, self.state

  def init_weights_and_state(self, input_signature, device=None):
    sig = input_signature[0] if isinstance(input_signature, (tuple, list)) and not isinstance(input_signature, ShapeDtype) \
        else input_signature
    batch, seqlen, d_model = (int(s) for s in sig.shape)
    if d_model != self._d_model:
      raise ValueError('input depth %d != n_heads * d_qk = %d' % (d_model, self._d_model))
    device = device or ('cuda' if torch.cuda.is_available() else 'cpu')
    key = np.array([0, 0], dtype=np.uint32) if self.rng is None else self.rng
    sub = _split_host(key, 4)                                         # one key per sublayer (combinators.py Serial init)
    def dense(k):
      g = np.random.Generator(np.random.Philox(key=int(k[0]) << 32 | int(k[1])))
      lim = np.sqrt(6.0 / (2 * d_model))
      w = torch.from_numpy(g.uniform(-lim, lim, size=(d_model, d_model)).astype(np.float32)).to(device)
      if not self._bias:
        return w
      return (w, torch.from_numpy((1e-6 * g.standard_normal(d_model)).astype(np.float32)).to(device))
    qkv = tuple(dense(k) for k in _split_host(sub[0], self._num_weights))
    self._attn.rng = sub[1]
    d_head = d_model // self._n_heads
    core_sig = (ShapeDtype((batch * self._n_heads, seqlen, d_head)),) * 2
    if self._masked:
      core_sig = core_sig + (ShapeDtype((batch, seqlen)),)
    self._attn.init_weights_and_state(core_sig, device=device)
    self.weights = (qkv, (), (), dense(sub[3]))
    self.state = ((), self._attn.state, (), ())

  # ---- pieces ---------------------------------------------------------------------------------------------------
  def _split(self, t):                                                # attention.py:347-364
    b, l, f = t.shape
    return t.view(b, l, self._n_heads, f // self._n_heads).permute(0, 2, 1, 3).reshape(b * self._n_heads, l, -1)

  def _merge(self, t):                                                # attention.py:369-388
    bh, l, d = t.shape
    return t.view(bh // self._n_heads, self._n_heads, l, d).permute(0, 2, 1, 3).reshape(-1, l, self._n_heads * d)

  @staticmethod
  def _kernel_bias(w):
    return (w[0], w[1]) if isinstance(w, (tuple, list)) else (w, None)

  def _run(self, inputs, weights, state, output_grad, update_state, rng=None):
    x, mask = (inputs[0], inputs[1]) if isinstance(inputs, (tuple, list)) else (inputs, None)
    if (mask is not None) != self._masked:
      raise ValueError('PureLSHSelfAttentionWrapper(masked=%s) takes %d inputs' % (self._masked, self._n_in))
    if x.dim() != 3 or x.shape[2] != self._d_model:
      raise ValueError('x must have shape (batch, seqlen, %d); got %s' % (self._d_model, tuple(x.shape)))
    B, L, D = x.shape
    n = self._num_weights
    kb = [self._kernel_bias(w) for w in weights[0]]
    w_cat = torch.cat([k for k, _ in kb], dim=1).to(x.dtype)          # (D, n D): the n projections as one GEMM
    proj = torch.matmul(x.reshape(B * L, D), w_cat).view(B, L, n * D)
    if self._bias:
      proj = proj + torch.cat([b for _, b in kb]).to(x.dtype)
    parts = list(proj.split(D, dim=2))                                # q, k, v  or  qk, v
    if self._rotary:                                                  # EA:3303-3304, 3363-3365: q (and k), never v
      cos, sin = _rotary_tables(L, D, x.device)
      parts = [_rotate(p, cos, sin) if i < n - 1 else p for i, p in enumerate(parts)]
    qk = (parts[0] + parts[1]) / 2.0 if n == 3 else parts[0]          # EA:3306
    core_in = (self._split(qk).contiguous(), self._split(parts[-1]).contiguous()) + ((mask,) if self._masked else ())
    compute_grad = output_grad is not None
    d_merged = None
    w_o, b_o = self._kernel_bias(weights[3])
    if compute_grad:                                                  # EA:3590-3597: Dense and MergeHeads cotangents
      dy = output_grad.to(device=x.device, dtype=x.dtype).reshape(B * L, D)
      d_merged = torch.matmul(dy, w_o.to(x.dtype).t()).view(B, L, D)
    core_out, new_core_state, core_grads = self._attn.forward_and_or_backward(
        core_in, state[1], _split_rngs(rng, 4)[1], output_grad=self._split(d_merged).contiguous() if compute_grad else None,
        compute_output=True, update_state=update_state)               # EA:3575-3577 + 3600-3602 in one call
    merged = self._merge(core_out)                                    # EA:3579-3581
    out = torch.matmul(merged.reshape(B * L, D), w_o.to(x.dtype)).view(B, L, D)      # EA:3584-3586
    if b_o is not None:
      out = out + b_o.to(x.dtype)
    new_state = ((), new_core_state, (), ()) if update_state else None
    if not compute_grad:
      return out, new_state, None, None
    d_wo = torch.matmul(merged.reshape(B * L, D).t(), dy).float()
    d_dense = (d_wo, dy.float().sum(dim=0)) if b_o is not None else d_wo
    d_qk, d_v = self._merge(core_grads[0]), self._merge(core_grads[1])
    d_parts = [d_qk / 2.0, d_qk / 2.0, d_v] if n == 3 else [d_qk, d_v]
    if self._rotary:
      d_parts = [_rotate_vjp(d, cos, sin) if i < n - 1 else d for i, d in enumerate(d_parts)]
    d_proj = torch.cat(d_parts, dim=2).reshape(B * L, n * D)          # EA:3605: the Branch's vjp, again one GEMM each way
    dx = torch.matmul(d_proj, w_cat.t()).view(B, L, D)
    d_wcat = torch.matmul(x.reshape(B * L, D).t(), d_proj).float()
    d_bcat = d_proj.float().sum(dim=0) if self._bias else None
    d_qkv = tuple((d_wcat[:, i * D:(i + 1) * D].contiguous(), d_bcat[i * D:(i + 1) * D].contiguous()) if self._bias
                  else d_wcat[:, i * D:(i + 1) * D].contiguous() for i in range(n))
    inputs_grad = (dx, None) if self._masked else dx
    return out, new_state, inputs_grad, (d_qkv, (), (), d_dense)     # EA:3611-3614

  # ---- Layer interface ------------------------------------------------------------------------------------------
  def forward(self, inputs):
    """Serial.forward: hashes (update_state) and stores the core's new state."""
    out, new_state, _, _ = self._run(inputs, self.weights, self.state, None, True, self.rng)
    self.state = new_state
    self._attn.state = new_state[1]
    return out

  def pure_fn(self, inputs, weights, state, rng, use_cache=False):
    del use_cache
    out, new_state, _, _ = self._run(inputs, weights, state, None, True, rng)
    return out, new_state

  def forward_and_or_backward(self, inputs, weights, state, rng, output_grad=None, compute_output=True,
                              update_state=True):
    """EA:3542-3620 → (output, None, inputs_grad, weights_grad); like the reference it only serves the reversible
    backward pass: `compute_output`, not `update_state`, and an `output_grad` (EA:3566-3568)."""
    assert compute_output
    assert not update_state
    assert output_grad is not None
    return self._run(inputs, weights, state, output_grad, False, rng)

  def backward(self, inputs, output, grad, weights, state, new_state, rng=None, **kwargs):
    del output, state, kwargs
    _, _, inputs_grad, weights_grad = self.forward_and_or_backward(inputs, weights, new_state, rng, output_grad=grad,
                                                                   compute_output=True, update_state=False)
    return inputs_grad, weights_grad
