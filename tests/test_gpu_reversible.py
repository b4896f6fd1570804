"""GPU parity tests of the reversible block around the layer (`trax_b200.ReversibleHalfResidual`)
against the oracle's restatement of `trax/layers/reversible.py:296-412` + `layers/normalization.py:129-136`."""
import numpy as np
import pytest
import torch

from oracle import lsh_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('rows,D', [(37, 256), (1000, 1024), (8, 2048)])
def test_layernorm_fwd_bwd(rows, D, dtype):
  from trax_b200 import reversible as R
  rng = np.random.default_rng(rows + D)
  rnd = util.bf16_round if dtype == torch.bfloat16 else (lambda a: np.asarray(a, np.float32))
  x = rnd(rng.standard_normal((rows, D)) * 2 + 0.5)
  dz = rnd(rng.standard_normal((rows, D)))
  ct = rnd(rng.standard_normal((rows, D)))
  scale = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
  bias = (0.1 * rng.standard_normal(D)).astype(np.float32)
  cu = lambda a, dt=torch.float32: torch.from_numpy(np.asarray(a, np.float32)).cuda().to(dt)
  z, stats = R.layernorm_fwd(cu(x, dtype), cu(scale), cu(bias))
  util.assert_close_layer(z.float().cpu().numpy(), O.layernorm(x, scale, bias), 'z')
  ct_out, d_scale, d_bias = R.layernorm_bwd(cu(x, dtype), cu(dz, dtype), cu(ct, dtype), stats, cu(scale))
  dx, ws, wb = O.layernorm_vjp(x, scale, dz)
  util.assert_close_layer(ct_out.float().cpu().numpy(), ct + dx, 'ct + dx')
  util.assert_close_layer(d_scale.cpu().numpy(), ws, 'd_scale')
  util.assert_close_layer(d_bias.cpu().numpy(), wb, 'd_bias')
  only_dx, _, _ = R.layernorm_bwd(cu(x, dtype), cu(dz, dtype), None, stats, cu(scale))
  util.assert_close_layer(only_dx.float().cpu().numpy(), dx, 'dx')


@pytest.mark.parametrize('B,L,D,dtype,cfg', [
    (1, 512, 256, torch.float32, util.make_cfg(H=2, C=128, nh=2, n_buckets=8)),
    (2, 512, 256, torch.bfloat16, util.make_cfg(H=4, C=128, nh=4, n_buckets=None)),
    (1, 1024, 256, torch.float32, util.make_cfg(H=2, C=64, nh=1, n_buckets=32)),
])
def test_reversible_half_forward_and_reverse_and_grad(B, L, D, dtype, cfg):
  import trax_b200
  rng = np.random.default_rng(17)
  rnd = util.bf16_round if dtype == torch.bfloat16 else (lambda a: np.asarray(a, np.float32))
  x1, x2 = rnd(rng.standard_normal((B, L, D))), rnd(rng.standard_normal((B, L, D)))
  ct_y1, ct_x2 = rnd(rng.standard_normal((B, L, D))), rnd(rng.standard_normal((B, L, D)))
  scale = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
  bias = (0.1 * rng.standard_normal(D)).astype(np.float32)
  attn_w = O.init_weights(cfg.n_heads, D, 64, 64, seed=3)
  factors = O.bucket_factors(cfg.n_buckets, L, cfg.chunk_len)
  rot = rng.standard_normal((B * cfg.n_heads, 64, cfg.n_hashes, sum(factors) // 2)).astype(np.float32)

  attn = trax_b200.LSHSelfAttention(n_heads=cfg.n_heads, d_qk=64, d_v=64, causal=True, chunk_len=cfg.chunk_len,
                                    n_hashes=cfg.n_hashes, n_buckets=cfg.n_buckets)
  block = trax_b200.ReversibleHalfResidual(attn)
  sig = trax_b200.ShapeDtype((B, L, D))
  block.init((sig, sig))
  cu = lambda a, dt=torch.float32: torch.from_numpy(np.asarray(a, np.float32)).cuda().to(dt)
  block.weights = ((cu(scale), cu(bias)), tuple(cu(w) for w in attn_w))
  attn._rotations_override = torch.from_numpy(rot)

  # forward: (x1, x2) -> (x1 + Attn(LN(x2)), x2); bucket ids bit-exact would need identical z bits, so they are taken from
  # the GPU state for the reverse pass (the reference stores them in new_state for the same reason, reversible.py:263-265)
  y1, ctx = block.forward((cu(x1, dtype), cu(x2, dtype)))
  buckets = block.state[1][0].cpu().numpy()
  z = O.layernorm(x2, scale, bias)
  want_res, _, _, _ = O.forward_and_or_backward(cfg, z, attn_w, buckets=buckets, update_state=False)
  util.assert_close_layer(y1.float().cpu().numpy(), x1 + want_res, 'y1')
  assert ctx.data_ptr() == ctx.data_ptr() and torch.equal(ctx.float().cpu(), torch.from_numpy(x2))

  # reverse_and_grad from (y1, x2) and cotangents
  y1_np = rnd(y1.float().cpu().numpy())
  (rx1, rx2), ((g_y1, g_x2), ((d_scale, d_bias), dw)) = block.reverse_and_grad(
      (y1, ctx), (cu(ct_y1, dtype), cu(ct_x2, dtype)), block.weights, None, block.state, None)
  (wx1, _), ((_, w_ct_x2), ((w_ds, w_db), w_dw)) = O.reversible_half_reverse_and_grad(
      cfg, y1_np, x2, ct_y1, ct_x2, (scale, bias), attn_w, buckets)
  util.assert_close_layer(rx1.float().cpu().numpy(), wx1, 'reconstructed x1')
  util.assert_close_layer(rx1.float().cpu().numpy(), x1, 'reconstructed x1 vs the original input', rtol=3e-2)
  assert torch.equal(rx2, ctx) and torch.equal(g_y1.float().cpu(), torch.from_numpy(ct_y1))
  util.assert_close_layer(g_x2.float().cpu().numpy(), w_ct_x2, 'ct_x2')
  util.assert_close_layer(d_scale.cpu().numpy(), w_ds, 'd_scale')
  util.assert_close_layer(d_bias.cpu().numpy(), w_db, 'd_bias')
  for n, g, w in zip(('dw_q', 'dw_v', 'dw_o'), dw, w_dw):
    util.assert_close_layer(g.float().cpu().numpy(), w, n)


def test_reversible_half_around_the_pure_lsh_wrapper():
  """The Terraformer-style block: ReversibleHalfResidual(LayerNorm, attention_layer=PureLSHSelfAttentionWrapper)
  (reversible.py:281-286 accepts any layer with `forward_and_or_backward`; EA:3542-3620 serves exactly this call)."""
  import trax_b200
  B, L, H = 1, 512, 4
  D = 64 * H
  cfg = util.make_cfg(H=H, C=128, nh=2, n_buckets=8)
  rng = np.random.default_rng(29)
  x1, x2 = (rng.standard_normal((B, L, D)).astype(np.float32) for _ in range(2))
  ct_y1, ct_x2 = (rng.standard_normal((B, L, D)).astype(np.float32) for _ in range(2))
  scale = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
  bias = (0.1 * rng.standard_normal(D)).astype(np.float32)
  wrap = trax_b200.PureLSHSelfAttentionWrapper(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=128, n_hashes=2,
                                               n_buckets=8, bias=True, num_weights=3)
  block = trax_b200.ReversibleHalfResidual(wrap)
  sig = trax_b200.ShapeDtype((B, L, D))
  (_, attn_w), _ = block.init((sig, sig), rng=np.array([5, 6], np.uint32))
  cu = lambda a: torch.from_numpy(np.asarray(a, np.float32)).cuda()
  block.weights = ((cu(scale), cu(bias)), attn_w)
  y1, ctx = block.forward((cu(x1), cu(x2)))
  buckets = block.state[1][1][0].cpu().numpy()                      # wrapper state -> core state -> buckets
  np_w = lambda w: tuple(l.cpu().numpy().astype(np.float64) for l in w)
  qkv_w, dense_w = tuple(np_w(w) for w in attn_w[0]), np_w(attn_w[3])
  z = O.layernorm(x2, scale, bias)
  want_res, _, want_dz, (want_dqkv, want_ddense) = O.pure_lsh_wrapper(cfg, z, qkv_w, dense_w, buckets=buckets,
                                                                     output_grad=ct_y1)
  util.assert_close_layer(y1.cpu().numpy(), x1 + want_res, 'y1')
  (rx1, rx2), ((g_y1, g_x2), ((d_scale, d_bias), dw)) = block.reverse_and_grad(
      (y1, ctx), (cu(ct_y1), cu(ct_x2)), block.weights, None, block.state, None)
  util.assert_close_layer(rx1.cpu().numpy(), x1, 'reconstructed x1', rtol=3e-2)
  assert torch.equal(rx2, ctx) and torch.equal(g_y1.cpu(), torch.from_numpy(ct_y1))
  w_dx2, w_ds, w_db = O.layernorm_vjp(x2, scale, want_dz)
  util.assert_close_layer(g_x2.cpu().numpy(), ct_x2 + w_dx2, 'ct_x2')
  util.assert_close_layer(d_scale.cpu().numpy(), w_ds, 'd_scale')
  util.assert_close_layer(d_bias.cpu().numpy(), w_db, 'd_bias')
  assert dw[1] == () and dw[2] == ()
  for i in range(3):
    for got, want, nm in zip(dw[0][i], want_dqkv[i], ('kernel', 'bias')):
      util.assert_close_layer(got.cpu().numpy(), want, 'd_qkv[%d] %s' % (i, nm))
  for got, want, nm in zip(dw[3], want_ddense, ('kernel', 'bias')):
    util.assert_close_layer(got.cpu().numpy(), want, 'd_dense %s' % nm)


def test_reversible_half_with_output_dropout_uses_one_mask_in_both_passes():
  """The attention sub-key is `_split_rngs(rng, 2)[1]` in `forward` AND in `reverse_and_grad` (reversible.py:297, 328): with
  output dropout the mask of the backward pass must be the forward's, otherwise x1 = y1 - residual is reconstructed
  from a different residual."""
  import trax_b200
  B, L, D = 1, 512, 256
  rng = np.random.default_rng(5)
  x1, x2 = (rng.standard_normal((B, L, D)).astype(np.float32) for _ in range(2))
  ct_y1, ct_x2 = (rng.standard_normal((B, L, D)).astype(np.float32) for _ in range(2))
  attn = trax_b200.LSHSelfAttention(n_heads=2, d_qk=64, d_v=64, causal=True, chunk_len=128, n_hashes=2, n_buckets=8,
                                    output_dropout=0.5)
  block = trax_b200.ReversibleHalfResidual(attn)
  sig = trax_b200.ShapeDtype((B, L, D))
  key = np.array([11, 22], np.uint32)
  block.init((sig, sig), rng=key)
  cu = lambda a: torch.from_numpy(a).cuda()
  y1, ctx = block.forward((cu(x1), cu(x2)))
  res = (y1 - cu(x1)).cpu().numpy()
  dropped = np.all(res == 0, axis=(0, 1))
  assert 0.2 < dropped.mean() < 0.8, 'output dropout 0.5 should zero about half of the d_model columns'
  (rx1, _), ((_, g_x2), _) = block.reverse_and_grad((y1, ctx), (cu(ct_y1), cu(ct_x2)), block.weights, None, block.state,
                                                    key)
  util.assert_close_layer(rx1.cpu().numpy(), x1, 'x1 reconstructed through the same dropout mask', rtol=3e-2)
  # a different key draws a different mask: the reconstruction must then differ (the test has teeth)
  (bad, _), _ = block.reverse_and_grad((y1, ctx), (cu(ct_y1), cu(ct_x2)), block.weights, None, block.state,
                                       np.array([33, 44], np.uint32))
  assert np.abs(bad.cpu().numpy() - x1).max() > 1e-2
  with pytest.raises(ValueError):
    block.reverse_and_grad((y1, ctx), (cu(ct_y1), cu(ct_x2)), block.weights, None, block.state, None)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_residual_fused_into_the_output_projection_matches_the_separate_pass(dtype):
  """`lsh_layer_fwd_res` / `lsh_layer_bwd_res`: out = residual + sign * attention_output from the GEMM epilogue
  (reversible.py:318, 400) against the plain call followed by the add / subtract; the gradients must not change."""
  import trax_b200
  layer = trax_b200.LSHSelfAttention(n_heads=4, causal=True, chunk_len=128, n_hashes=2, n_buckets=None)
  layer.init(trax_b200.ShapeDtype((2, 1024, 256)))
  g = torch.Generator('cuda').manual_seed(21)
  x = torch.randn(2, 1024, 256, device='cuda', generator=g).to(dtype)
  acc = torch.randn(2, 1024, 256, device='cuda', generator=g).to(dtype)
  ct = torch.randn(2, 1024, 256, device='cuda', generator=g).to(dtype)
  plain, state, _, _ = layer.forward_and_or_backward(x, layer.weights, layer.state, None)
  fused, _, _, _ = layer._forward_and_or_backward(x, layer.weights, state, None, update_state=False, _residual=(acc, 1.0))
  want = acc.float() + plain.float()
  tol = 1e-5 if dtype == torch.float32 else 2 ** -7 * float(want.abs().max())
  assert float((fused.float() - want).abs().max()) <= tol
  out0, _, dx0, dw0 = layer.forward_and_or_backward(x, layer.weights, state, None, output_grad=ct, update_state=False)
  out1, _, dx1, dw1 = layer._forward_and_or_backward(x, layer.weights, state, None, output_grad=ct, update_state=False,
                                                     _residual=(acc, -1.0))
  assert float((out1.float() - (acc.float() - out0.float())).abs().max()) <= tol
  assert torch.equal(dx0, dx1) and all(torch.equal(a, b) for a, b in zip(dw0, dw1))


def test_layernorm_bf16_output_feeds_the_layer_without_a_conversion_pass():
  """f32 activations: `lsh_layernorm_fwd_bf16` + a layer call with dims.x_bf16 must give exactly what the f32 LayerNorm output
  followed by the layer's own f32 -> bf16 conversion gives (one rounding of the same fp32 number either way)."""
  import trax_b200
  from trax_b200 import reversible as R
  layer = trax_b200.LSHSelfAttention(n_heads=4, causal=True, chunk_len=128, n_hashes=2, n_buckets=None)
  layer.init(trax_b200.ShapeDtype((2, 1024, 256)))
  g = torch.Generator('cuda').manual_seed(33)
  ctx = torch.randn(2, 1024, 256, device='cuda', generator=g)
  ct = torch.randn(2, 1024, 256, device='cuda', generator=g)
  scale = torch.rand(256, device='cuda', generator=g) + 0.5
  bias = torch.randn(256, device='cuda', generator=g)
  z32, st32 = R.layernorm_fwd(ctx, scale, bias)
  z16, st16 = R.layernorm_fwd(ctx, scale, bias, z_bf16=True)
  assert z16.dtype == torch.bfloat16 and torch.equal(z16, z32.to(torch.bfloat16)) and torch.equal(st16, st32)
  out0, state, _, _ = layer.forward_and_or_backward(z32, layer.weights, layer.state, None)
  out1, _, dx1, dw1 = layer._forward_and_or_backward(z16, layer.weights, state, None, output_grad=ct, update_state=False,
                                                     _io_dtype=torch.float32)
  out2, _, dx2, dw2 = layer.forward_and_or_backward(z32, layer.weights, state, None, output_grad=ct, update_state=False)
  assert out1.dtype == torch.float32 and dx1.dtype == torch.float32
  assert torch.equal(out1, out2) and torch.equal(out0, out2) and torch.equal(dx1, dx2)
  assert all(torch.equal(a, b) for a, b in zip(dw1, dw2))
