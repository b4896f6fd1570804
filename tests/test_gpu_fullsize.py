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
