"""`trax_b200.SelfAttention(share_qk=True, chunk_len=...)` (EA:936-1726) vs `oracle/self_attention_oracle.py`, which is pinned
against the reference's SelfAttention by the live sweep (tests/test_reference_pin.py)."""
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


def test_self_attention_rejects_what_is_not_built():
  import trax_b200
  with pytest.raises(NotImplementedError):
    trax_b200.SelfAttention(n_heads=2, causal=True, chunk_len=128, n_chunks_before=1)          # share_qk=False (default)
  with pytest.raises(NotImplementedError):
    trax_b200.SelfAttention(n_heads=2, share_qk=True, causal=True)                              # chunk_len=None
