"""`trax.layers.reversible.ReversibleHalfResidual(LayerNorm(), attention_layer=<LSH attention>)` — the immediate caller of
the hot path on both sides (`trax/layers/reversible.py:244-412`, `models/reformer/reformer.py`
builds exactly this pair per attention half-block).

  forward           (y1, x2) = (x1 + Attn(LN(x2)), x2)                                   reversible.py:296-321
  reverse_and_grad  z = LN(x2);  (res, _, dz, dw) = Attn.forward_and_or_backward(z, w, new_state, rng,
                    output_grad=ct_y1, compute_output=True, update_state=False)          reversible.py:371-378
                    ct_x2 += LN_vjp(dz);  x1 = y1 - res                                  reversible.py:384-400

The attention call is ONE fused forward+backward of the layer (the recompute is shared between inverting the block and
its gradient).  LayerNorm forward / VJP and the residual add / subtract are the memory-bound kernels of
`csrc/residual.cu` behind `lsh_layernorm_fwd`, `lsh_layernorm_bwd`, `lsh_residual_add`, `lsh_residual_sub`.
Weights `((scale, bias), attention_weights)` and state `((), attention_state)` follow the sublayer order
`(compute_residual, attention_layer)` of reversible.py:283.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from trax_b200 import _lib, ops
from trax_b200.lsh_attention import _split_rngs


def _rows(x):
  return int(x.numel() // x.shape[-1]), int(x.shape[-1])


def layernorm_fwd(x, scale, bias, epsilon=1e-6, z_bf16=False):
  """trax/layers/normalization.py:129-136.  Returns (z, stats (rows, 2) f32 = {mean, rstd}).  z_bf16 (f32 activations
  only): z is written as bf16 — the rounding the attention layer applies to its input anyway — for layer calls with
  `_io_dtype=torch.float32` (no f32 z round trip, no conversion pass)."""
  lib = _lib.load()
  x = x.contiguous()
  rows, d = _rows(x)
  stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
  if z_bf16 and x.dtype == torch.float32:
    z = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.lsh_layernorm_fwd_bf16(rows, d, ops._ptr(x), ops._ptr(scale), ops._ptr(bias), ops._ptr(z), ops._ptr(stats),
                                          ctypes.c_float(epsilon), ops._stream()), 'lsh_layernorm_fwd_bf16')
    return z, stats
  z = torch.empty_like(x)
  _lib.check(lib.lsh_layernorm_fwd(rows, d, ops._act_dtype(x), ops._ptr(x), ops._ptr(scale), ops._ptr(bias), ops._ptr(z),
                                   ops._ptr(stats), ctypes.c_float(epsilon), ops._stream()), 'lsh_layernorm_fwd')
  return z, stats


def layernorm_bwd(x, dz, ct_in, stats, scale):
  """VJP of layernorm_fwd at x: returns (ct_in + dx, d_scale, d_bias)."""
  lib = _lib.load()
  x, dz = x.contiguous(), dz.contiguous().to(x.dtype)
  rows, d = _rows(x)
  ct_out = torch.empty_like(x)
  d_scale = torch.empty(d, dtype=torch.float32, device=x.device)
  d_bias = torch.empty(d, dtype=torch.float32, device=x.device)
  ct_in = None if ct_in is None else ct_in.contiguous().to(x.dtype)
  _lib.check(lib.lsh_layernorm_bwd(rows, d, ops._act_dtype(x), ops._ptr(x), ops._ptr(dz), ops._ptr(ct_in), ops._ptr(stats),
                                   ops._ptr(scale), ops._ptr(ct_out), ops._ptr(d_scale), ops._ptr(d_bias), ops._stream()),
             'lsh_layernorm_bwd')
  return ct_out, d_scale, d_bias


def _residual(a, b, sign):
  lib = _lib.load()
  a, b = a.contiguous(), b.contiguous().to(a.dtype)
  out = torch.empty_like(a)
  fn = lib.lsh_residual_add if sign > 0 else lib.lsh_residual_sub
  _lib.check(fn(a.numel(), ops._act_dtype(a), ops._ptr(a), ops._ptr(b), ops._ptr(out), ops._stream()), 'lsh_residual')
  return out


class ReversibleHalfResidual:
  """ReversibleHalfResidual(LayerNorm(epsilon), attention_layer=attention_layer): inputs / outputs (accumulator, context)."""

  def __init__(self, attention_layer, epsilon=1e-6):
    if not hasattr(attention_layer, 'forward_and_or_backward'):           # reversible.py:281
      raise ValueError('attention_layer must provide forward_and_or_backward')
    if attention_layer.n_in != 1:
      raise NotImplementedError('masked attention inside the reversible block is not supported')
    self._attention_layer = attention_layer
    self._epsilon = float(epsilon)
    self._ln_weights = ()
    self._rng = None

  @property
  def rng(self):
    if self._rng is None:
      self._rng = np.array([0, 0], dtype=np.uint32)                       # base.py: default key from seed 0
    return self._rng

  @rng.setter
  def rng(self, rng):
    self._rng = rng

  n_in = n_out = 2                                                          # reversible.py:288-294 (1 + 1 context)

  @property
  def sublayers(self):
    return ('LayerNorm', self._attention_layer)

  @property
  def weights(self):
    return (self._ln_weights, self._attention_layer.weights)

  @weights.setter
  def weights(self, w):
    self._ln_weights, self._attention_layer.weights = tuple(w[0]), w[1]

  @property
  def state(self):
    return ((), self._attention_layer.state)

  @state.setter
  def state(self, s):
    self._attention_layer.state = s[1]

  def init(self, input_signature, rng=None):
    """input_signature: (accumulator signature, context signature), both (batch, seqlen, d_model)."""
    ctx = input_signature[1]
    d_model = int(ctx.shape[-1])
    dev = 'cuda' if torch.cuda.is_available() else 'cpu'
    self._ln_weights = (torch.ones(d_model, dtype=torch.float32, device=dev),      # normalization.py:138-142
                        torch.zeros(d_model, dtype=torch.float32, device=dev))
    if rng is not None:
      self.rng = rng
    self._attention_layer.init(ctx, rng=rng)
    return self.weights, self.state

  def forward(self, xs):
    accumulator, context = xs
    scale, bias = self._ln_weights
    rngs = _split_rngs(self.rng, 2)                                        # reversible.py:297: (LayerNorm, attention)
    fused = self._fused(accumulator, context)
    z, _ = layernorm_fwd(context, scale, bias, self._epsilon, z_bf16=fused)
    attn = self._attention_layer
    with torch.no_grad():
      if fused:
        # output = accumulator + residual (reversible.py:318) leaves the output projection's epilogue: no separate pass;
        # with f32 activations the LayerNorm hands its result over as bf16 (what the layer makes of its input anyway)
        out, new_state, _, _ = attn._forward_and_or_backward(z, attn.weights, attn.state, rngs[1], compute_output=True,
                                                             update_state=True, _residual=(accumulator.contiguous(), +1.0),
                                                             _io_dtype=context.dtype)
        attn.state = new_state
        return out, context
      residual, new_state = attn.pure_fn(z, attn.weights, attn.state, rngs[1])     # reversible.py:308
    attn.state = new_state                                                 # buckets of this step, read back by reverse_and_grad
    return _residual(accumulator, residual, +1.0), context

  def _fused(self, accumulator, z):
    """The residual add / subtract can ride in the attention layer's output-projection epilogue (device tensors only)."""
    from trax_b200.lsh_attention import LSHSelfAttention
    return (isinstance(self._attention_layer, LSHSelfAttention) and not self._attention_layer._incremental   # (predict mode:
            and accumulator.is_cuda and z.is_cuda                                           # the decode step has no epilogue)
            and accumulator.shape == z.shape and accumulator.dtype == z.dtype)

  def reverse(self, output, weights=(), state=(), new_state=(), rng=None):
    raise NotImplementedError('Only reverse_and_grad is actually used.')     # reversible.py:323-324

  def reverse_and_grad(self, output, ct, weights=(), state=(), new_state=(), rng=None):
    """Returns (inputs, (inputs_ct, weights_ct)) like reversible.py:326-412."""
    del state
    accumulator_output, context = output
    accumulator_output_ct, context_ct = ct
    (scale, bias), attn_weights = weights if weights else self.weights
    attn_state = (new_state if new_state else self.state)[1]
    rngs = _split_rngs(rng, 2)                                             # reversible.py:328: same split as forward
    fused = self._fused(accumulator_output, context)
    z, stats = layernorm_fwd(context, scale, bias, self._epsilon, z_bf16=fused)
    if fused:
      # reconstructed_x = accumulator_output - residual (reversible.py:400), again in the epilogue of the same GEMM
      reconstructed_x, _, dz, attn_weights_ct = self._attention_layer._forward_and_or_backward(
          z, attn_weights, attn_state, rngs[1], output_grad=accumulator_output_ct, compute_output=True, update_state=False,
          _residual=(accumulator_output.contiguous(), -1.0), _io_dtype=context.dtype)
      context_ct_new, d_scale, d_bias = layernorm_bwd(context, dz, context_ct, stats, scale)
    else:
      residual, _, dz, attn_weights_ct = self._attention_layer.forward_and_or_backward(
          z, attn_weights, attn_state, rngs[1], output_grad=accumulator_output_ct, compute_output=True, update_state=False)
      context_ct_new, d_scale, d_bias = layernorm_bwd(context, dz, context_ct, stats, scale)
      reconstructed_x = _residual(accumulator_output, residual, -1.0)
    return (reconstructed_x, context), ((accumulator_output_ct, context_ct_new), ((d_scale, d_bias), attn_weights_ct))
