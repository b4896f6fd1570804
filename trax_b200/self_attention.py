"""`trax.layers.research.efficient_attention.SelfAttention` (EA:936-1726) — the chunked local attention ReformerLM
interleaves with the LSH layer (`reformer_enwik8.gin:23-28`: 3 of 4 layers) — for every `share_qk` with a `chunk_len`.

With shared queries and keys, `SelfAttention.forward_unbatched` (EA:1136-1197) is `LSHSelfAttention.forward_unbatched`
with ONE hash round and the identity permutation: `attend` gets `q_info = arange(seqlen)`, keys are the length-normalised
queries, a token does not attend to itself (EA:1175-1178), look-back / look-ahead chunks wrap around (EA:122-142), the
padding mask flips the sign of `kv_info` (EA:1185-1186).  Sorting all-zero bucket ids by `seqlen * bucket + position`
(EA:1946-1947) IS the identity, so the layer runs the LSH core — same CUDA kernels, no hashing — on constant buckets.
The default `share_qk=False` (EA:1133-1197) has a key projection of its own, `k = x w_k` (EA:1160-1162), which is NOT
length-normalised (EA:229-231 apply to shared-QK only; the keys are only divided by sqrt(d_qk), EA:232), and a token may
attend to itself (`exclude_self=self._share_qk`, EA:1175-1178).  The kernels take it as `LshAttnDims.separate_k`: the
projections are ONE GEMM onto q | v | k rows, the attention kernels read their key tiles from the k columns with a constant
scale and without the self mask, and the key-side cotangent goes to `dw_k` instead of `dw_q` (no normalisation VJP).

Interface kept: the constructor keywords (EA:939-953), weights `(w_q, w_v, w_o)` for `share_qk`, `(w_q, w_k, w_v, w_o)`
otherwise (EA:1112-1128), state `()` (EA:1130-1131), `forward`, `backward`, `forward_and_or_backward` →
`(output, new_state, inputs_grad, weights_grad)`.  `mode='predict'` (EA:1200-1268; state `(mem_end, (mem,), ())`) runs
through trax_b200/predict.py.  `chunk_len=None` (one dense window over the whole sequence) is built for predict mode only
(a decode step attends over the whole memory whatever the chunk length, EA:1262-1267); in the other modes it raises.
"""
import torch

from trax_b200.lsh_attention import LSHSelfAttention


class SelfAttention(LSHSelfAttention):
  """Chunked local self-attention with shared query/key projections (EA:936)."""

  def __init__(self, n_heads=2, d_qk=64, d_v=64, share_qk=False, causal=False, masked=False, chunk_len=None,
               n_chunks_before=0, n_chunks_after=0, bias=False, mode='train', predict_mem_len=None, predict_drop_len=None,
               attention_dropout=0.0, output_dropout=0.0, n_parallel_heads=None, use_python_loop=False,
               use_reference_code=False):
    dense = chunk_len is None
    if dense and mode != 'predict':
      raise NotImplementedError('SelfAttention(chunk_len=None) is dense attention over the whole sequence; the training '
                                'kernels are chunked (chunk_len 32 / 64 / 128 / 256).  In predict mode it is built: every '
                                'decode step attends over the whole memory anyway (EA:1262-1267)')
    super().__init__(n_heads=n_heads, d_qk=d_qk, d_v=d_v, causal=causal, masked=masked, chunk_len=64 if dense else chunk_len,
                     n_chunks_before=n_chunks_before, n_chunks_after=n_chunks_after, n_hashes=1, n_buckets=2, mode=mode,
                     predict_mem_len=predict_mem_len, predict_drop_len=predict_drop_len, attention_dropout=attention_dropout,
                     output_dropout=output_dropout, bias=bias, n_parallel_heads=n_parallel_heads,
                     use_python_loop=use_python_loop, use_reference_code=use_reference_code)
    self._share_qk = bool(share_qk)
    self._separate_k = not share_qk
    self._predict_hashes = False
    self._dense = dense                 # chunk_len=None (predict mode only): prefixes are never chunked (EA:1250)

  def init_weights_and_state(self, input_signature, device=None):
    super().init_weights_and_state(input_signature, device=device)   # same (w_q, w_v, w_o) shapes and init (EA:1061-1066)
    if self._separate_k:                                             # (w_q, w_k, w_v, w_o), EA:1112-1128
      import numpy as np
      from trax_b200.lsh_attention import _split_host
      w_q, w_v, w_o = self.weights
      d_model = int(w_q.shape[1])
      keys = _split_host(np.asarray(self.rng, np.uint32) ^ np.uint32(0x6b657973), self._n_heads)     # 'keys'
      w_k = np.stack([self._kernel_initializer((d_model, self._d_qk), np.random.Generator(np.random.Philox(
          key=int(k[0]) << 32 | int(k[1])))) for k in keys])
      self.weights = (w_q, torch.from_numpy(np.ascontiguousarray(w_k)).to(w_q.device), w_v, w_o)
    if not self._incremental:
      self.state = ()                                                # EA:1130-1131 (predict mode: (mem_end, (mem,), ()), EA:1093-1101)

  def forward(self, inputs):
    output, new_state, _, _ = self.forward_and_or_backward(inputs, self.weights, self.state, self.rng, compute_output=True,
                                                           update_state=True)
    if self._incremental:
      self.state = new_state
    return output

  def _forward_and_or_backward(self, inputs, weights, state, rng, output_grad=None, compute_output=True, update_state=True,
                               _stash=None, _residual=None, _io_dtype=None, _raw=False):
    if self._incremental and not _raw:
      from trax_b200 import predict
      return predict.forward_and_or_backward(self, inputs, weights, state, rng, output_grad, compute_output, update_state)
    x = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
    if not torch.cuda.is_available():
      from trax_b200 import _lib
      raise _lib.LshAttnError('trax_b200.SelfAttention needs a CUDA device (no CPU fallback)')
    dev = x.device if x.is_cuda else torch.device('cuda', torch.cuda.current_device())
    # identity permutation == stable sort of constant bucket ids (EA:1946-1947)
    buckets = torch.zeros((int(x.shape[0]) * self._n_heads, int(x.shape[1])), dtype=torch.int32, device=dev)
    out, _, inputs_grad, weights_grad = super()._forward_and_or_backward(
        inputs, weights, (buckets, None), rng, output_grad=output_grad, compute_output=compute_output, update_state=False,
        _stash=_stash, _residual=_residual, _io_dtype=_io_dtype, _raw=True)
    return out, (state if update_state else None), inputs_grad, weights_grad
