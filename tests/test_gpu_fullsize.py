"""Size-independent properties at BASELINE config 2's full size (B1, L65536, H8, chunk 128, 4 hashes), where the
float64 oracle is too slow: permutation validity and sortedness, hash determinism, invariance of the attention
kernels to the CTA work partition (bit-exact), combine weights, and linearity of the backward in the cotangent."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, H, L, D, C, NH = 1, 8, 65536, 1024, 128, 4


@pytest.fixture(scope='module')
def c2():
  from trax_b200 import ops, _lib
  dims = _lib.make_dims(B, H, L, D, 64, 64, C, 1, 0, NH, ops.bucket_factors(None, L, C), True, False, 1)
  g = torch.Generator('cuda').manual_seed(0)
  qv = torch.randn((B, L, H, 128), device='cuda', generator=g).bfloat16()
  keys = torch.arange(2 * B * H, dtype=torch.int32, device='cuda').reshape(B * H, 2) + 5
  rot, _ = ops.make_rotations(dims, keys)
  buckets = ops.hash_qv(dims, qv, rot)
  sticker, undo = ops.sort(dims, buckets)
  return dict(dims=dims, qv=qv, rot=rot, buckets=buckets, sticker=sticker, undo=undo)


def test_buckets_in_range_and_deterministic(c2):
  from trax_b200 import ops
  b = c2['buckets'].view(B * H, NH, L)
  nb = 32 * 32
  lo = (torch.arange(NH, device='cuda') * nb).view(1, NH, 1)
  assert bool(((b >= lo) & (b < lo + nb)).all())
  again = ops.hash_qv(c2['dims'], c2['qv'], c2['rot'])
  assert torch.equal(again, c2['buckets'])
  # every one of the 1024 buckets is used somewhere (random rotations on N(0,1) vectors)
  assert int(torch.unique(b[0, 0]).numel()) > 900


def test_permutation_is_valid_sorted_and_stable(c2):
  sticker, undo, buckets = c2['sticker'].long(), c2['undo'].long(), c2['buckets'].long()
  ar = torch.arange(NH * L, device='cuda')
  for u in range(B * H):
    assert torch.equal(undo[u][sticker[u]], ar)                       # inverse permutation (EA:1953)
    key = L * buckets[u][sticker[u]] + sticker[u] % L                 # sorted key (EA:1947), strictly increasing
    assert bool((key[1:] > key[:-1]).all())
    assert torch.equal(sticker[u] // L, ar // L)                      # hash rounds stay contiguous


def test_attention_is_invariant_to_the_cta_partition(c2, monkeypatch):
  """The persistent kernels split the chunk list over CTAs; any split must give bit-identical rows."""
  from trax_b200 import ops
  o1, l1 = ops.attend_fwd(c2['dims'], c2['qv'], c2['sticker'])
  oc, lse = ops.combine_fwd(c2['dims'], o1, l1)
  do = torch.randn(oc.shape, device='cuda', generator=torch.Generator('cuda').manual_seed(1)).bfloat16()
  g1 = ops.attend_bwd(c2['dims'], c2['qv'], c2['sticker'], oc, lse, do)
  monkeypatch.setenv('LSH_ATTN_MAX_CTAS', '37')
  o2, l2 = ops.attend_fwd(c2['dims'], c2['qv'], c2['sticker'])
  g2 = ops.attend_bwd(c2['dims'], c2['qv'], c2['sticker'], oc, lse, do)
  assert torch.equal(o1, o2) and torch.equal(l1, l2) and torch.equal(g1, g2)
  assert bool(torch.isfinite(o1.float()).all()) and bool(torch.isfinite(g1.float()).all())


def test_combine_is_a_convex_combination_and_backward_is_linear(c2):
  from trax_b200 import ops
  o_r, logits = ops.attend_fwd(c2['dims'], c2['qv'], c2['sticker'])
  oc, lse = ops.combine_fwd(c2['dims'], o_r, logits)
  w = torch.exp(logits.view(B * H, NH, L) - lse.view(B * H, 1, L))
  wsum = w.sum(1)
  normal = lse > -1e4              # rows at the -1e5 level (only their own key class visible) carry fp32 ulp(1e5) = 0.008
  assert float((wsum[normal] - 1).abs().max()) < 1e-4 and float((wsum[~normal] - 1).abs().max()) < 2e-2
  lo = o_r.float().view(B * H, NH, L, 64).amin(1).view(B, H, L, 64).permute(0, 2, 1, 3)
  hi = o_r.float().view(B * H, NH, L, 64).amax(1).view(B, H, L, 64).permute(0, 2, 1, 3)
  assert bool(((oc.float() >= lo - 2e-2) & (oc.float() <= hi + 2e-2)).all())
  # position 0 of every unit attends only to itself: o == v_0 (EA quirk list)
  torch.testing.assert_close(oc[0, 0].float(), c2['qv'][0, 0, :, 64:].float(), rtol=1e-2, atol=1e-2)
  gen = torch.Generator('cuda').manual_seed(2)
  do = torch.randn(oc.shape, device='cuda', generator=gen).bfloat16()
  g1 = ops.attend_bwd(c2['dims'], c2['qv'], c2['sticker'], oc, lse, do).float()
  g2 = ops.attend_bwd(c2['dims'], c2['qv'], c2['sticker'], oc, lse, (2 * do.float()).bfloat16()).float()   # exact doubling in bf16
  err = (g2 - 2 * g1).norm() / (2 * g1).norm()
  assert float(err) < 1e-2, float(err)


def test_layer_step_runs_at_full_size_and_backward_matches_reduced_dout():
  """Full layer at config 2: out finite, state shapes, and dW linear in dout (dout -> 0 gives 0)."""
  import trax_b200
  layer = trax_b200.LSHSelfAttention(n_heads=H, causal=True, chunk_len=C, n_hashes=NH)
  layer.init(trax_b200.ShapeDtype((B, L, D)))
  x = torch.randn((B, L, D), device='cuda', generator=torch.Generator('cuda').manual_seed(3)).bfloat16()
  out = layer.forward(x)
  assert out.shape == x.shape and out.dtype == torch.bfloat16 and bool(torch.isfinite(out.float()).all())
  assert tuple(layer.state[0].shape) == (B * H, NH * L)
  dx, dw = layer.backward(x, out, torch.zeros_like(x), layer.weights, None, layer.state, None)
  assert float(dx.float().abs().max()) == 0.0 and all(float(g.abs().max()) == 0.0 for g in dw)


# ---- numeric parity AT SIZE: one (example, head) unit of every BASELINE config against the fp32 oracle ---------------------
AT_SIZE = {
    # name: (L, C, nh, n_buckets)
    'c2': (65536, 128, 4, None),            # BASELINE config 2 (headline): [32, 32] buckets
    'c3': (12288, 128, 2, 192),             # config 3: imagenet64, int n_buckets = 192 (R = 96)
    'c4-nh1': (16384, 128, 1, None),        # config 4: n_hashes sweep at seq 16384, [16, 16] buckets
    'c4-nh2': (16384, 128, 2, None),
    'c4-nh4': (16384, 128, 4, None),
    'c4-nh8': (16384, 128, 8, None),
}


@pytest.mark.parametrize('name', sorted(AT_SIZE))
def test_one_unit_at_size_against_the_oracle(name):
  """VERDICT r1 #4a: bucket ids, sticker and undo_sort bit-exact, per-round logits, combined output and the attention
  gradients within tolerance — at the sequence length, chunking, hash rounds and bucket factorisation each BASELINE config
  names, one unit (units are independent, EA:2402-2432).  The oracle runs in fp32 (its BLAS path; ~4 s at config 2)."""
  from oracle import lsh_oracle as O
  from tests import util
  from trax_b200 import ops, _lib
  L_, C_, nh, nbk = AT_SIZE[name]
  rng = np.random.default_rng(len(name) + L_ + nh)
  factors = O.bucket_factors(nbk, L_, C_)
  cfg = util.make_cfg(H=1, C=C_, nh=nh, n_buckets=nbk)
  qv = util.bf16_round(rng.standard_normal((1, L_, 1, 128)))
  rot = rng.standard_normal((1, 64, nh, sum(factors) // 2)).astype(np.float32)
  do = util.bf16_round(rng.standard_normal((1, L_, 1, 64)))
  dims = _lib.make_dims(1, 1, L_, 128, 64, 64, C_, 1, 0, nh, factors, True, False, _lib.LSH_DTYPE_BF16)
  cu = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).cuda().to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a)).cuda()
  qv_d = cu(qv, torch.bfloat16)
  buckets = ops.hash_qv(dims, qv_d, cu(rot))
  sticker, undo = ops.sort(dims, buckets)
  o_r, logits = ops.attend_fwd(dims, qv_d, sticker)
  o_c, lse = ops.combine_fwd(dims, o_r, logits)
  dqv = ops.attend_bwd(dims, qv_d, sticker, o_c, lse, cu(do, torch.bfloat16))
  torch.cuda.synchronize()

  w_q, w_v, w_o = (w.astype(np.float32) for w in util.core_identity_weights())
  res = O.forward_unit(cfg, qv[0, :, 0, :], w_q, w_v, w_o, rotations=rot[0], dtype=np.float32)
  np.testing.assert_array_equal(buckets[0].cpu().numpy(), res.buckets)                 # bit-exact (north_star 1)
  np.testing.assert_array_equal(sticker[0].cpu().numpy(), res.sticker)                 # bit-exact (north_star 2)
  np.testing.assert_array_equal(undo[0].cpu().numpy(), res.undo_sort)
  util.assert_close(logits[0].cpu().numpy(), res.logits, 'logits')
  util.assert_close(o_c[0, :, 0, :].float().cpu().numpy(), res.o, 'o_comb')
  grads = O.backward_unit(cfg, res, do[0, :, 0, :].astype(np.float32))[0]
  got = dqv[0, :, 0, :].float().cpu().numpy()
  util.assert_close(got[:, :64], grads[:, :64], 'dq', frac_bad_max=0.02)
  util.assert_close(got[:, 64:], grads[:, 64:], 'dv', frac_bad_max=0.02)


def test_config5_share_permutation_and_invariants():
  """BASELINE config 5, one GPU's share (2 of 16 heads, 2^20 tokens, n_buckets [32, 32] so the int32 key of EA:1947 does
  not wrap): bucket range, permutation validity / sortedness, finite outputs, partition invariance of the forward."""
  from trax_b200 import ops, _lib
  L5, H5 = 1 << 20, 2
  dims = _lib.make_dims(1, H5, L5, 1024, 64, 64, 128, 1, 0, 1, [32, 32], True, False, 1)
  g = torch.Generator('cuda').manual_seed(5)
  qv = torch.randn((1, L5, H5, 128), device='cuda', generator=g).bfloat16()
  keys = torch.arange(2 * H5, dtype=torch.int32, device='cuda').reshape(H5, 2) + 11
  rot, _ = ops.make_rotations(dims, keys)
  buckets = ops.hash_qv(dims, qv, rot)
  assert int(buckets.min()) >= 0 and int(buckets.max()) < 1024
  sticker, undo = ops.sort(dims, buckets)
  ar = torch.arange(L5, device='cuda')
  for u in range(H5):
    s, b = sticker[u].long(), buckets[u].long()
    assert torch.equal(undo[u].long()[s], ar)
    key = L5 * b[s] + s
    assert bool((key[1:] > key[:-1]).all())
  o1, l1 = ops.attend_fwd(dims, qv, sticker)
  assert bool(torch.isfinite(o1.float()).all()) and bool(torch.isfinite(l1).all())
  torch.testing.assert_close(o1[:, 0].float(), qv[0, 0, :, 64:].float(), rtol=1e-2, atol=1e-2)   # position 0 sees itself only
  do = torch.randn((1, L5, H5, 64), device='cuda', generator=g).bfloat16()
  oc = o1.view(1, H5, L5, 64).permute(0, 2, 1, 3).contiguous()                                 # nh = 1: o_rounds == o_comb rows
  g1 = ops.attend_bwd(dims, qv, sticker, oc, l1, do)
  assert bool(torch.isfinite(g1.float()).all())
