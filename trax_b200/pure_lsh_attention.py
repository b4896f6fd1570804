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

import torch

from trax_b200 import _lib, ops
from trax_b200.lsh_attention import LSHSelfAttention, ShapeDtype, _split_host, _to_int32_bits


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
  def forward_and_or_backward(self, inputs, state, rng, output_grad=None, compute_output=True, update_state=True):
    """Returns (output, new_state, inputs_grad): output iff compute_output, new_state iff update_state,
    inputs_grad = (dqk, dv[, None for the mask]) iff output_grad is given."""
    del rng
    if not isinstance(inputs, (tuple, list)) or len(inputs) != self._n_in:
      raise ValueError('PureLSHSelfAttention(masked=%s) takes %d inputs' % (self._masked, self._n_in))
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
    # (B*H, L, d) x 2  ->  (B, L, H, [q | v]) bf16: the row layout every kernel gathers from
    qv = torch.cat([qk.view(batch, self._n_heads, seqlen, self._d_qk), v.view(batch, self._n_heads, seqlen, self._d_v)],
                   dim=3).permute(0, 2, 1, 3).to(torch.bfloat16).contiguous()
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
    o_rounds, logits = ops.attend_fwd(dims, qv, sticker, mask=mask_d)   # EA:2780-2808 (un-sorted rows)
    o_comb, lse_tot = ops.combine_fwd(dims, o_rounds, logits)       # EA:2810-2814

    def unpack(t, d):                                               # (B, L, H, d) -> (B*H, L, d) in the input dtype
      return t.permute(0, 2, 1, 3).reshape(bh, seqlen, d).to(qk.dtype)
    output = unpack(o_comb, self._d_v) if compute_output else None
    inputs_grad = None
    if compute_grad:
      do = output_grad.to(dev).view(batch, self._n_heads, seqlen, self._d_v).permute(0, 2, 1, 3).to(torch.bfloat16).contiguous()
      dqv = ops.attend_bwd(dims, qv, sticker, o_comb, lse_tot, do, mask=mask_d)
      inputs_grad = (unpack(dqv[..., :self._d_qk], self._d_qk), unpack(dqv[..., self._d_qk:], self._d_v))
      if self._masked:
        inputs_grad = inputs_grad + (None,)
    return output, new_state, inputs_grad
