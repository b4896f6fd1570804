"""GPU parity tests of `trax_b200.PureLSHSelfAttention` (EA:2564-3265) against the CPU oracle and
against the full layer, modelled on `efficient_attention_test.py:376-440` (LSH == projections + weight-less core + w_o).
"""
import numpy as np
import pytest
import torch

from oracle import lsh_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def _pure(cfg, **kw):
  import trax_b200
  return trax_b200.PureLSHSelfAttention(
      n_heads=cfg.n_heads, d_qk=64, d_v=64, causal=cfg.causal, masked=cfg.masked, chunk_len=cfg.chunk_len,
      n_chunks_before=cfg.n_chunks_before, n_chunks_after=cfg.n_chunks_after, n_hashes=cfg.n_hashes,
      n_buckets=cfg.n_buckets, **kw)


PURE_CASES = [
    # (B, L, cfg)
    (2, 512, util.make_cfg(H=2, C=128, nh=4, n_buckets=8)),                                     # tcgen05 path
    (1, 1024, util.make_cfg(H=2, C=64, nh=1, n_buckets=32)),                                    # BASELINE config 1 core
    (1, 512, util.make_cfg(H=2, C=64, nb=1, na=1, nh=2, n_buckets=16, causal=False, masked=True)),
]


@pytest.mark.parametrize('B,L,cfg', PURE_CASES)
def test_pure_lsh_matches_oracle(B, L, cfg):
  """Output and (dqk, dv) vs the oracle's unit with identity projections; buckets bit-exact with the oracle's hash."""
  H = cfg.n_heads
  rng = np.random.default_rng(5)
  qk = util.bf16_round(rng.standard_normal((B * H, L, 64)))
  v = util.bf16_round(rng.standard_normal((B * H, L, 64)))
  dout = util.bf16_round(rng.standard_normal((B * H, L, 64)))
  mask = (rng.random((B, L)) > 0.2) if cfg.masked else None
  if mask is not None:
    dout = dout * np.repeat(mask, H, axis=0)[:, :, None]
  factors = O.bucket_factors(cfg.n_buckets, L, cfg.chunk_len)
  rot = rng.standard_normal((B * H, 64, cfg.n_hashes, sum(factors) // 2)).astype(np.float32)
  layer = _pure(cfg)
  sig = [trax_sig((B * H, L, 64)), trax_sig((B * H, L, 64))] + ([trax_sig((B, L))] if cfg.masked else [])
  weights, state = layer.init(tuple(sig))
  assert weights == () and state[0].shape == (B * H, cfg.n_hashes * L) and state[1].shape == (B * H, 2)
  layer._rotations_override = torch.from_numpy(rot)
  inputs = (torch.from_numpy(qk).cuda(), torch.from_numpy(v).cuda())
  if cfg.masked:
    inputs = inputs + (torch.from_numpy(mask).cuda(),)
  out = layer.forward(inputs)
  dq_dv, dw = layer.backward(inputs, out, torch.from_numpy(dout).cuda(), (), None, layer.state, None)
  assert dw == () and out.shape == (B * H, L, 64)
  got_buckets = layer.state[0].cpu().numpy()
  w_q, w_v, w_o = util.core_identity_weights()
  for u in range(B * H):
    b = u // H
    x = np.concatenate([qk[u], v[u]], axis=1).astype(np.float64)
    r = O.forward_unit(cfg, x, w_q, w_v, w_o, rotations=rot[u], mask=None if mask is None else mask[b])
    np.testing.assert_array_equal(got_buckets[u], r.buckets)
    util.assert_close(out[u].float().cpu().numpy(), r.out, 'out[%d]' % u)
    g = O.backward_unit(cfg, r, dout[u].astype(np.float64))[0]
    util.assert_close(dq_dv[0][u].float().cpu().numpy(), g[:, :64], 'dqk[%d]' % u)
    util.assert_close(dq_dv[1][u].float().cpu().numpy(), g[:, 64:], 'dv[%d]' % u)


def trax_sig(shape):
  import trax_b200
  return trax_b200.ShapeDtype(shape)


def test_lsh_equals_projections_plus_pure_core():
  """efficient_attention_test.py:376-440: LSHSelfAttention(x) == (PureLSH(x w_q, x w_v) per head) w_o summed over heads,
  with the same hash rotations."""
  import trax_b200
  B, L, D, H = 1, 512, 128, 2
  cfg = util.make_cfg(H=H, C=128, nh=2, n_buckets=8)
  rng = np.random.default_rng(11)
  x = torch.from_numpy(util.bf16_round(rng.standard_normal((B, L, D)))).cuda()
  rot = torch.from_numpy(rng.standard_normal((B * H, 64, cfg.n_hashes, 4)).astype(np.float32))
  full = trax_b200.LSHSelfAttention(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=128, n_hashes=2, n_buckets=8)
  full.init(trax_b200.ShapeDtype((B, L, D)))
  full._rotations_override = rot
  want = full.forward(x.to(torch.bfloat16)).float()
  w_q, w_v, w_o = (w.float() for w in full.weights)                 # (H, D, 64), (H, D, 64), (H, 64, D)
  xb = x.to(torch.bfloat16).float()
  q = torch.einsum('bld,hdk->bhlk', xb, w_q.to(torch.bfloat16).float()).reshape(B * H, L, 64)
  v = torch.einsum('bld,hdk->bhlk', xb, w_v.to(torch.bfloat16).float()).reshape(B * H, L, 64)
  pure = _pure(cfg)
  pure.init((trax_b200.ShapeDtype((B * H, L, 64)), trax_b200.ShapeDtype((B * H, L, 64))))
  pure._rotations_override = rot
  o = pure.forward((q, v)).float()                                   # (B*H, L, 64)
  np.testing.assert_array_equal(pure.state[0].cpu().numpy(), full.state[0].cpu().numpy())   # same buckets
  got = torch.einsum('bhlk,hkd->bld', o.reshape(B, H, L, 64).to(torch.bfloat16).float(), w_o.to(torch.bfloat16).float())
  util.assert_close(got.cpu().numpy(), want.cpu().numpy(), 'LSH vs projections + PureLSH + w_o')


@pytest.mark.parametrize('num_weights,bias,rotary,dtype', [(3, True, False, torch.float32), (2, False, True, torch.bfloat16)])
def test_pure_lsh_wrapper_matches_oracle(num_weights, bias, rotary, dtype):
  """`PureLSHSelfAttentionWrapper` (EA:3493-3620) on the CUDA core vs the oracle's restatement.  The buckets the device
  hashed (from its bf16 qk) are handed to the oracle, so near-tie hash flips do not enter the comparison; they must be
  valid bucket ids and survive the backward call unchanged."""
  import trax_b200
  B, L, H = 2, 512, 2
  D = 64 * H
  cfg = util.make_cfg(H=H, C=128, nh=2, n_buckets=8)
  wrap = trax_b200.PureLSHSelfAttentionWrapper(
      n_heads=H, d_qk=64, d_v=64, causal=True, bias=bias, num_weights=num_weights, weights_format='model',
      rotary_position_emb=rotary, chunk_len=128, n_hashes=2, n_buckets=8)
  weights, _ = wrap.init(trax_b200.ShapeDtype((B, L, D)), rng=np.array([3, 4], np.uint32))
  rng = np.random.default_rng(23)
  x = util.bf16_round(rng.standard_normal((B, L, D)))
  dout = util.bf16_round(rng.standard_normal((B, L, D)))
  xd = torch.from_numpy(x).cuda().to(dtype)
  out = wrap.forward(xd)
  buckets = wrap.state[1][0].cpu().numpy()
  assert buckets.shape == (B * H, 2 * L) and buckets.min() >= 0 and buckets.max() < 2 * 8
  assert (buckets[:, :L] < 8).all() and (buckets[:, L:] >= 8).all()                # per-round offsets (EA:1913-1915)
  np_w = lambda w: tuple(l.cpu().numpy().astype(np.float64) for l in w) if isinstance(w, tuple) \
      else w.cpu().numpy().astype(np.float64)
  qkv_w, dense_w = tuple(np_w(w) for w in weights[0]), np_w(weights[3])
  want_out, _, want_dx, (want_dqkv, want_ddense) = O.pure_lsh_wrapper(
      cfg, x, qkv_w, dense_w, buckets=buckets, output_grad=dout, rotary_position_emb=rotary)
  util.assert_close_layer(out.float().cpu().numpy(), want_out, 'wrapper out')
  out2, new_state, dx, dw = wrap.forward_and_or_backward(xd, weights, wrap.state, None,
                                                         output_grad=torch.from_numpy(dout).cuda(), update_state=False)
  assert new_state is None
  np.testing.assert_array_equal(wrap.state[1][0].cpu().numpy(), buckets)
  util.assert_close_layer(out2.float().cpu().numpy(), want_out, 'wrapper out (backward call)')
  util.assert_close_layer(dx.float().cpu().numpy(), want_dx, 'wrapper dx')
  leaves = lambda w: list(w) if isinstance(w, tuple) else [w]
  for i in range(num_weights):
    for got, want in zip(leaves(dw[0][i]), leaves(want_dqkv[i])):
      util.assert_close_layer(got.float().cpu().numpy(), want, 'wrapper d_qkv[%d]' % i)
  for got, want in zip(leaves(dw[3]), leaves(want_ddense)):
    util.assert_close_layer(got.float().cpu().numpy(), want, 'wrapper d_dense')


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_head_pack_and_unpack_kernels_match_the_torch_layout_ops(dtype):
  """`lsh_pack_heads` / `lsh_unpack_heads` against the cat / permute / cast they replace on the core's path (EA:3052-3070,
  3245-3265): bit-exact in both directions, both dtypes, ragged widths."""
  from trax_b200 import ops
  B, H, L = 2, 3, 257
  g = torch.Generator('cuda').manual_seed(3)
  qk = torch.randn(B * H, L, 64, device='cuda', generator=g).to(dtype)
  v = torch.randn(B * H, L, 40, device='cuda', generator=g).to(dtype)
  got = ops.pack_heads(qk, v, H)
  want = torch.cat([qk.view(B, H, L, 64), v.view(B, H, L, 40)], dim=3).permute(0, 2, 1, 3).to(torch.bfloat16).contiguous()
  assert got.shape == want.shape and torch.equal(got, want)
  single = ops.pack_heads(v, None, H)
  assert torch.equal(single, v.view(B, H, L, 40).permute(0, 2, 1, 3).to(torch.bfloat16).contiguous())
  back_q = ops.unpack_heads(got, 0, 64, dtype)
  back_v = ops.unpack_heads(got, 64, 40, dtype)
  assert torch.equal(back_q, qk.to(torch.bfloat16).to(dtype)) and torch.equal(back_v, v.to(torch.bfloat16).to(dtype))
