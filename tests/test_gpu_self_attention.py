"""`trax_b200.SelfAttention(chunk_len=...)` (EA:936-1726), both `share_qk` settings, vs `oracle/self_attention_oracle.py`,
which is pinned against the reference's SelfAttention by the live sweep (tests/test_reference_pin.py)."""
import numpy as np
import pytest
import torch

from oracle import self_attention_oracle as S
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,L,D,H,C,nb,na,causal,masked,dtype', [
    (2, 512, 128, 2, 128, 1, 0, True, False, torch.float32),       # the enwik8 SelfAttention shape; tcgen05 kernels
    (1, 1024, 256, 4, 128, 1, 0, True, False, torch.bfloat16),
    (2, 256, 64, 2, 64, 1, 1, False, True, torch.float32),         # bidirectional, padding mask, look-ahead chunk
    (1, 256, 64, 2, 64, 0, 0, True, False, torch.float32),         # own chunk only
])
def test_self_attention_share_qk_matches_oracle(B, L, D, H, C, nb, na, causal, masked, dtype):
  import trax_b200
  rng = np.random.default_rng(41)
  rnd = util.bf16_round if dtype == torch.bfloat16 else (lambda a: np.asarray(a, np.float32))
  x, dout = rnd(rng.standard_normal((B, L, D))), rnd(rng.standard_normal((B, L, D)))
  mask = (rng.random((B, L)) > 0.2) if masked else None
  if masked:
    dout = dout * mask[:, :, None]
  layer = trax_b200.SelfAttention(n_heads=H, d_qk=64, d_v=64, share_qk=True, causal=causal, masked=masked, chunk_len=C,
                                  n_chunks_before=nb, n_chunks_after=na)
  sig = trax_b200.ShapeDtype((B, L, D))
  weights, state = layer.init((sig, trax_b200.ShapeDtype((B, L))) if masked else sig)
  assert state == () and [tuple(w.shape) for w in weights] == [(H, D, 64), (H, D, 64), (H, 64, D)]
  cfg = S.SelfAttentionConfig(n_heads=H, d_qk=64, d_v=64, share_qk=True, causal=causal, masked=masked, chunk_len=C,
                              n_chunks_before=nb, n_chunks_after=na)
  w_np = tuple(w.cpu().numpy().astype(np.float64) for w in weights)
  want_out, want_dx, want_dw = S.forward_and_or_backward(cfg, x, w_np, mask=mask, output_grad=dout)
  x_d = torch.from_numpy(x).cuda().to(dtype)
  inputs = (x_d, torch.from_numpy(mask).cuda()) if masked else x_d
  out = layer.forward(inputs)
  util.assert_close_layer(out.float().cpu().numpy(), want_out, 'out')
  dx, dw = layer.backward(inputs, out, torch.from_numpy(dout).cuda().to(dtype), weights, (), (), None)
  util.assert_close_layer((dx[0] if masked else dx).float().cpu().numpy(), want_dx, 'dx')
  for name, g, w in zip(('dw_q', 'dw_v', 'dw_o'), dw, want_dw):
    util.assert_close_layer(g.cpu().numpy(), w, name)


@pytest.mark.parametrize('B,L,D,H,C,nb,na,causal,masked,dtype', [
    (2, 512, 128, 2, 128, 1, 0, True, False, torch.float32),       # reformer_enwik8.gin:23-28, 40: 3 of 4 layers use this
    (1, 1024, 256, 4, 128, 1, 0, True, False, torch.bfloat16),
    (2, 256, 64, 2, 64, 1, 1, False, True, torch.float32),         # bidirectional, padding mask, look-ahead chunk
    (1, 256, 64, 2, 64, 0, 0, True, False, torch.float32),         # own chunk only
    (1, 512, 128, 2, 256, 1, 0, True, False, torch.float32),       # chunk 256
])
def test_self_attention_separate_keys_matches_oracle(B, L, D, H, C, nb, na, causal, masked, dtype):
  """The default share_qk=False (EA:1133-1197): own key projection, keys not normalised, self-attention allowed; weights
  (w_q, w_k, w_v, w_o); output, dx and all four weight gradients against the oracle (fp64)."""
  import trax_b200
  rng = np.random.default_rng(43)
  rnd = util.bf16_round if dtype == torch.bfloat16 else (lambda a: np.asarray(a, np.float32))
  x, dout = rnd(rng.standard_normal((B, L, D))), rnd(rng.standard_normal((B, L, D)))
  mask = (rng.random((B, L)) > 0.2) if masked else None
  if masked:
    dout = dout * mask[:, :, None]
  layer = trax_b200.SelfAttention(n_heads=H, d_qk=64, d_v=64, causal=causal, masked=masked, chunk_len=C,
                                  n_chunks_before=nb, n_chunks_after=na)
  sig = trax_b200.ShapeDtype((B, L, D))
  weights, state = layer.init((sig, trax_b200.ShapeDtype((B, L))) if masked else sig)
  assert state == () and [tuple(w.shape) for w in weights] == [(H, D, 64), (H, D, 64), (H, D, 64), (H, 64, D)]
  assert not torch.equal(weights[0], weights[1])                   # w_k is drawn from its own key
  cfg = S.SelfAttentionConfig(n_heads=H, d_qk=64, d_v=64, share_qk=False, causal=causal, masked=masked, chunk_len=C,
                              n_chunks_before=nb, n_chunks_after=na)
  w_np = tuple(w.cpu().numpy().astype(np.float64) for w in weights)
  want_out, want_dx, want_dw = S.forward_and_or_backward(cfg, x, w_np, mask=mask, output_grad=dout)
  x_d = torch.from_numpy(x).cuda().to(dtype)
  inputs = (x_d, torch.from_numpy(mask).cuda()) if masked else x_d
  out = layer.forward(inputs)
  util.assert_close_layer(out.float().cpu().numpy(), want_out, 'out')
  dx, dw = layer.backward(inputs, out, torch.from_numpy(dout).cuda().to(dtype), weights, (), (), None)
  util.assert_close_layer((dx[0] if masked else dx).float().cpu().numpy(), want_dx, 'dx')
  assert len(dw) == 4
  for name, g, w in zip(('dw_q', 'dw_k', 'dw_v', 'dw_o'), dw, want_dw):
    util.assert_close_layer(g.cpu().numpy(), w, name)
  # the fused call ReversibleHalfResidual makes (reversible.py:374-378) returns the same output and gradients
  o2, _, dx2, dw2 = layer.forward_and_or_backward(inputs, weights, (), None, output_grad=torch.from_numpy(dout).cuda().to(dtype),
                                                  compute_output=True, update_state=False)
  assert torch.equal(o2, out) and torch.equal(dx2[0] if masked else dx2, dx[0] if masked else dx)
  for a, b in zip(dw, dw2):
    assert torch.equal(a, b)


def test_self_attention_autograd_with_separate_keys():
  import trax_b200
  layer = trax_b200.SelfAttention(n_heads=2, causal=True, chunk_len=64, n_chunks_before=1)
  layer.init(trax_b200.ShapeDtype((1, 256, 64)))
  x = torch.randn((1, 256, 64), device='cuda', requires_grad=True)
  ws = tuple(w.clone().requires_grad_(True) for w in layer.weights)
  out, _ = layer.pure_fn(x, ws, (), None)
  out.square().sum().backward()
  assert x.grad is not None and all(w.grad is not None and bool(torch.isfinite(w.grad).all()) for w in ws)
  assert float(ws[1].grad.abs().max()) > 0                          # the key projection receives a gradient


def test_self_attention_rejects_what_is_not_built():
  import trax_b200
  with pytest.raises(NotImplementedError):
    trax_b200.SelfAttention(n_heads=2, share_qk=True, causal=True)                              # chunk_len=None outside predict mode
  with pytest.raises(ValueError):                                   # predict mode (tests/test_zgpu_predict.py) needs its memory sizes
    trax_b200.SelfAttention(n_heads=2, causal=True, chunk_len=128, mode='predict')
  with pytest.raises(NotImplementedError):                          # ... and is forward-only (EA:2002-2003)
    sa = trax_b200.SelfAttention(n_heads=2, causal=True, chunk_len=64, mode='predict', predict_mem_len=128, predict_drop_len=64)
    _, st = sa.init(trax_b200.ShapeDtype((1, 1, 64)))
    sa.forward_and_or_backward(torch.zeros((1, 1, 64), device='cuda'), sa.weights, st, None,
                               output_grad=torch.zeros((1, 1, 64), device='cuda'))
  layer = trax_b200.SelfAttention(n_heads=2, causal=True, chunk_len=64)
  layer.init(trax_b200.ShapeDtype((1, 128, 64)))
  with pytest.raises(ValueError):                                   # share_qk=False takes four weights
    layer.forward_and_or_backward(torch.zeros((1, 128, 64), device='cuda'), layer.weights[:3], (), None)
