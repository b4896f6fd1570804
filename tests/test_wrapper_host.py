"""CPU tests of `trax_b200.PureLSHSelfAttentionWrapper`'s host logic (EA:3493-3620): projections, QK averaging, rotary
embedding and its transpose, head split / merge, output Dense and every weight gradient, with the attention core
replaced — through the reference's own `pure_lsh_implementation` argument (EA:3504) — by the oracle.  The CUDA core under
the same wrapper is covered by tests/test_gpu_pure_lsh.py."""
import numpy as np
import pytest
import torch

from oracle import lsh_oracle as O
from tests import util


class OracleCore:
  """Stands in for PureLSHSelfAttention: same constructor keywords and `forward_and_or_backward` contract (EA:3052-3265),
  computed per unit by the oracle in fp64."""

  def __init__(self, n_heads, d_qk, d_v, causal, masked, mode, output_dropout, attention_dropout, chunk_len, n_hashes,
               n_buckets):
    del mode, output_dropout, attention_dropout
    self.cfg = O.LSHConfig(n_heads=n_heads, d_qk=d_qk, d_v=d_v, causal=causal, masked=masked, chunk_len=chunk_len,
                           n_chunks_before=1, n_chunks_after=0, n_hashes=n_hashes, n_buckets=n_buckets)
    self.state, self.rng, self.rotations = (), None, None

  def init_weights_and_state(self, sig, device=None):
    bh, seqlen = sig[0].shape[0], sig[0].shape[1]
    self.state = (torch.zeros((bh, self.cfg.n_hashes * seqlen), dtype=torch.int32), torch.zeros((bh, 2), dtype=torch.int32))

  def forward_and_or_backward(self, inputs, state, rng, output_grad=None, compute_output=True, update_state=True):
    qk, v = inputs[0].double().numpy(), inputs[1].double().numpy()
    mask = inputs[2].numpy().astype(bool) if self.cfg.masked else None
    w_q, w_v, w_o = util.core_identity_weights()
    outs, grads, buckets = [], [], []
    for u in range(qk.shape[0]):
      r = O.forward_unit(self.cfg, np.concatenate([qk[u], v[u]], axis=1), w_q, w_v, w_o,
                         buckets=None if update_state else state[0][u].numpy(),
                         rotations=self.rotations[u] if update_state else None,
                         mask=None if mask is None else mask[u // self.cfg.n_heads])
      outs.append(r.out)
      buckets.append(r.buckets)
      if output_grad is not None:
        grads.append(O.backward_unit(self.cfg, r, output_grad[u].double().numpy())[0])
    out = torch.from_numpy(np.stack(outs)).to(inputs[0].dtype)
    new_state = (torch.from_numpy(np.stack(buckets)), state[1]) if update_state else None
    g = None
    if output_grad is not None:
      g = torch.from_numpy(np.stack(grads)).to(inputs[0].dtype)
      g = (g[..., :64].contiguous(), g[..., 64:].contiguous()) + ((None,) if self.cfg.masked else ())
    return out, new_state, g


def _leaves(w):
  return list(w) if isinstance(w, (tuple, list)) else [w]


@pytest.mark.parametrize('num_weights,bias,rotary,masked', [(3, True, False, False), (2, False, True, False),
                                                            (3, False, True, True), (2, True, False, True)])
def test_wrapper_host_logic_matches_oracle(num_weights, bias, rotary, masked):
  import trax_b200
  B, L, H = 2, 64, 2
  D = 64 * H
  wrap = trax_b200.PureLSHSelfAttentionWrapper(
      n_heads=H, d_qk=64, d_v=64, causal=True, masked=masked, pure_lsh_implementation=OracleCore, bias=bias,
      num_weights=num_weights, weights_format='model', rotary_position_emb=rotary, chunk_len=16, n_hashes=2, n_buckets=4)
  sig = trax_b200.ShapeDtype((B, L, D))
  weights, state = wrap.init((sig, trax_b200.ShapeDtype((B, L))) if masked else sig, rng=np.array([7, 9], np.uint32))
  assert len(weights) == 4 and weights[1] == () and weights[2] == () and len(weights[0]) == num_weights
  assert state[0] == () and state[1][0].shape == (B * H, 2 * L)
  rng = np.random.default_rng(17)
  if bias:                                                          # the 1e-6 initial biases would not exercise the bias path
    weights = (tuple((w, torch.from_numpy(rng.standard_normal(D).astype(np.float32) * 0.1)) for w, _ in weights[0]), (), (),
               (weights[3][0], torch.from_numpy(rng.standard_normal(D).astype(np.float32) * 0.1)))
    wrap.weights = weights
  core = wrap.sublayers[1]
  core.rotations = rng.standard_normal((B * H, 64, 2, 2)).astype(np.float32)
  x = torch.from_numpy(rng.standard_normal((B, L, D)).astype(np.float32))
  mask = rng.random((B, L)) > 0.2 if masked else None
  dout = rng.standard_normal((B, L, D)).astype(np.float32)
  if masked:
    dout = dout * mask[:, :, None]
  inputs = (x, torch.from_numpy(mask)) if masked else x

  out = wrap.forward(inputs)                                        # hashes, stores the buckets
  buckets = wrap.state[1][0].numpy()
  np_w = lambda w: tuple(l.numpy().astype(np.float64) for l in w) if isinstance(w, tuple) else w.numpy().astype(np.float64)
  qkv_w, dense_w = tuple(np_w(w) for w in weights[0]), np_w(weights[3])
  want_out, want_b, _, _ = O.pure_lsh_wrapper(core.cfg, x.numpy(), qkv_w, dense_w, rotations=core.rotations, mask=mask,
                                              rotary_position_emb=rotary)
  np.testing.assert_array_equal(buckets, want_b)
  np.testing.assert_allclose(out.numpy(), want_out, rtol=2e-4, atol=2e-5)

  with pytest.raises(AssertionError):                               # EA:3566-3568
    wrap.forward_and_or_backward(inputs, weights, wrap.state, None, output_grad=None, update_state=False)
  out2, new_state, dx, dw = wrap.forward_and_or_backward(inputs, weights, wrap.state, None,
                                                         output_grad=torch.from_numpy(dout), update_state=False)
  assert new_state is None and dw[1] == () and dw[2] == ()
  _, _, want_dx, (want_dqkv, want_ddense) = O.pure_lsh_wrapper(core.cfg, x.numpy(), qkv_w, dense_w, buckets=buckets, mask=mask,
                                                               output_grad=dout, rotary_position_emb=rotary)
  np.testing.assert_allclose(out2.numpy(), want_out, rtol=2e-4, atol=2e-5)
  if masked:
    assert dx[1] is None
    dx = dx[0]
  np.testing.assert_allclose(dx.numpy(), want_dx, rtol=2e-3, atol=2e-4)
  for i in range(num_weights):
    for got, want in zip(_leaves(dw[0][i]), _leaves(want_dqkv[i])):
      np.testing.assert_allclose(got.numpy(), want, rtol=2e-3, atol=2e-3)
  for got, want in zip(_leaves(dw[3]), _leaves(want_ddense)):
    np.testing.assert_allclose(got.numpy(), want, rtol=2e-3, atol=2e-3)
  # Layer.backward signature (base.py:541-673 custom-gradient path)
  dx_b, dw_b = wrap.backward(inputs, out, torch.from_numpy(dout), weights, None, wrap.state, None)
  np.testing.assert_array_equal((dx_b[0] if masked else dx_b).numpy(), dx.numpy())


def test_wrapper_rejects_what_is_not_built():
  import trax_b200
  with pytest.raises(NotImplementedError):
    trax_b200.PureLSHSelfAttentionWrapper(n_heads=2, weights_format='sparse')
  with pytest.raises(ValueError):
    trax_b200.PureLSHSelfAttentionWrapper(n_heads=2, num_weights=4)
  wrap = trax_b200.PureLSHSelfAttentionWrapper(n_heads=2, causal=True, chunk_len=64, n_hashes=1)
  wrap.init(trax_b200.ShapeDtype((1, 128, 128)))
  with pytest.raises(ValueError):
    wrap.init(trax_b200.ShapeDtype((1, 128, 256)))                  # depth != n_heads * d_qk
  if not torch.cuda.is_available():                                 # default core = the CUDA one: no CPU fallback
    with pytest.raises(Exception):
      wrap.forward(torch.zeros(1, 128, 128))
