"""`trax.layers.research.efficient_attention.SelfAttention` (EA:936-1726) — the chunked local attention ReformerLM
interleaves with the LSH layer — for the configuration the existing kernels cover: `share_qk=True` with a `chunk_len`.

With shared queries and keys, `SelfAttention.forward_unbatched` (EA:1136-1197) is `LSHSelfAttention.forward_unbatched`
with ONE hash round and the identity permutation: `attend` gets `q_info = arange(seqlen)`, keys are the length-normalised
queries, a token does not attend to itself (EA:1175-1178), look-back / look-ahead chunks wrap around (EA:122-142), the
padding mask flips the sign of `kv_info` (EA:1185-1186).  Sorting all-zero bucket ids by `seqlen * bucket + position`
(EA:1946-1947) IS the identity, so the layer runs the LSH core — same CUDA kernels, no hashing — on constant buckets.
The default `share_qk=False` (separate key projection, no key normalisation, self-attention allowed) needs a separate-K
variant of the attend kernels and raises; its oracle (`oracle/self_attention_oracle.py`) is already pinned to the reference.

Interface kept: the constructor keywords (EA:939-953), weights `(w_q, w_v, w_o)` for `share_qk` (EA:1126), state `()`
(EA:1130-1131), `forward`, `backward`, `forward_and_or_backward` → `(output, new_state, inputs_grad, weights_grad)`.
"""
import torch

from trax_b200.lsh_attention import LSHSelfAttention


class SelfAttention(LSHSelfAttention):
  """Chunked local self-attention with shared query/key projections (EA:936)."""

  def __init__(self, n_heads=2, d_qk=64, d_v=64, share_qk=False, causal=False, masked=False, chunk_len=None,
               n_chunks_before=0, n_chunks_after=0, bias=False, mode='train', predict_mem_len=None, predict_drop_len=None,
               attention_dropout=0.0, output_dropout=0.0, n_parallel_heads=None, use_python_loop=False,
               use_reference_code=False):
    del predict_mem_len, predict_drop_len
    if not share_qk:
      raise NotImplementedError('SelfAttention(share_qk=False) needs attend kernels with a separate key projection, without '
                                'key normalisation and self-exclusion (EA:1160-1162, 230, 1175-1178); only share_qk=True is built')
    if chunk_len is None:
      raise NotImplementedError('SelfAttention(chunk_len=None) is dense attention over the whole sequence; the kernels are '
                                'chunked (chunk_len 32 / 64 / 128 / 256)')
    super().__init__(n_heads=n_heads, d_qk=d_qk, d_v=d_v, causal=causal, masked=masked, chunk_len=chunk_len,
                     n_chunks_before=n_chunks_before, n_chunks_after=n_chunks_after, n_hashes=1, n_buckets=2, mode=mode,
                     attention_dropout=attention_dropout, output_dropout=output_dropout, bias=bias,
                     n_parallel_heads=n_parallel_heads, use_python_loop=use_python_loop,
                     use_reference_code=use_reference_code)

  def init_weights_and_state(self, input_signature, device=None):
    super().init_weights_and_state(input_signature, device=device)   # same (w_q, w_v, w_o) shapes and init (EA:1061-1066)
    self.state = ()                                                  # EA:1130-1131

  def forward(self, inputs):
    output, _, _, _ = self.forward_and_or_backward(inputs, self.weights, self.state, self.rng, compute_output=True,
                                                   update_state=True)
    return output

  def _forward_and_or_backward(self, inputs, weights, state, rng, output_grad=None, compute_output=True, update_state=True,
                               _stash=None):
    x = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
    if not torch.cuda.is_available():
      from trax_b200 import _lib
      raise _lib.LshAttnError('trax_b200.SelfAttention needs a CUDA device (no CPU fallback)')
    dev = x.device if x.is_cuda else torch.device('cuda', torch.cuda.current_device())
    # identity permutation == stable sort of constant bucket ids (EA:1946-1947)
    buckets = torch.zeros((int(x.shape[0]) * self._n_heads, int(x.shape[1])), dtype=torch.int32, device=dev)
    out, _, inputs_grad, weights_grad = super()._forward_and_or_backward(
        inputs, weights, (buckets, None), rng, output_grad=output_grad, compute_output=compute_output, update_state=False,
        _stash=_stash)
    return out, (state if update_state else None), inputs_grad, weights_grad
