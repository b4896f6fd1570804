"""GPU parity tests of the layer (`trax_b200.LSHSelfAttention`) against the CPU oracle, modelled on
`trax/layers/research/efficient_attention_test.py` (:63-73, :136-156, :210-284, :322-328, :376-440).
"""
import numpy as np
import pytest
import torch

from oracle import lsh_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def _layer(cfg, **kw):
  import trax_b200
  return trax_b200.LSHSelfAttention(
      n_heads=cfg.n_heads, d_qk=64, d_v=64, causal=cfg.causal, masked=cfg.masked, chunk_len=cfg.chunk_len,
      n_chunks_before=cfg.n_chunks_before, n_chunks_after=cfg.n_chunks_after, n_hashes=cfg.n_hashes,
      n_buckets=cfg.n_buckets, max_length_for_buckets=cfg.max_length_for_buckets, **kw)


def _case(seed, B, L, D, cfg, dtype):
  rng = np.random.default_rng(seed)
  x = rng.standard_normal((B, L, D)).astype(np.float32)
  if dtype == torch.bfloat16:
    x = util.bf16_round(x)
  weights = O.init_weights(cfg.n_heads, D, 64, 64, seed=seed + 1)
  factors = O.bucket_factors(cfg.n_buckets, L, cfg.chunk_len)
  rot = rng.standard_normal((B * cfg.n_heads, 64, cfg.n_hashes, sum(factors) // 2)).astype(np.float32)
  dout = rng.standard_normal((B, L, D)).astype(np.float32)
  if dtype == torch.bfloat16:
    dout = util.bf16_round(dout)
  mask = (rng.random((B, L)) > 0.2) if cfg.masked else None
  if mask is not None:
    dout = dout * mask[:, :, None]     # no gradient flows into padding outputs (see test_attend_bwd)
  return x, weights, rot, dout, mask


LAYER_CASES = [
    # (B, L, D, dtype, cfg)
    (1, 1024, 256, torch.float32, util.make_cfg(H=2, C=64, nh=1, n_buckets=32)),                  # BASELINE config 1
    (2, 512, 128, torch.bfloat16, util.make_cfg(H=4, C=128, nh=4, n_buckets=None)),               # config-2 style
    (1, 768, 256, torch.bfloat16, util.make_cfg(H=2, C=128, nh=2, n_buckets=12)),                 # int n_buckets, nh=2
    (2, 256, 64, torch.float32, util.make_cfg(H=2, C=64, nb=1, na=1, nh=2, n_buckets=8, causal=False, masked=True)),
    (2, 512, 128, torch.bfloat16, util.make_cfg(H=2, C=128, nh=1, n_buckets=8)),    # single round on the tcgen05 path (rows go straight to o_comb)
]


@pytest.mark.parametrize('B,L,D,dtype,cfg', LAYER_CASES)
def test_layer_fwd_bwd_shared_buckets(B, L, D, dtype, cfg):
  """T3 + T4: update_state=False with the oracle's buckets in `state` (reversible.py:374-378)."""
  x, weights, rot, dout, mask = _case(21, B, L, D, cfg, dtype)
  want_out, buckets, _, _ = O.forward_and_or_backward(cfg, x, weights, rotations=rot, mask=mask, update_state=True)
  _, _, want_dx, want_dw = O.forward_and_or_backward(cfg, x, weights, buckets=buckets, mask=mask, output_grad=dout,
                                                     update_state=False)
  layer = _layer(cfg)
  w_d = tuple(torch.from_numpy(w).cuda() for w in weights)
  x_d = torch.from_numpy(x).cuda().to(dtype)
  inputs = x_d if mask is None else (x_d, torch.from_numpy(mask).cuda())
  state = (torch.from_numpy(buckets).cuda(), torch.zeros((B * cfg.n_heads, 2), dtype=torch.int32, device='cuda'))
  out, new_state, dx, dw = layer.forward_and_or_backward(
      inputs, w_d, state, None, output_grad=torch.from_numpy(dout).cuda().to(dtype), compute_output=True,
      update_state=False)
  assert new_state is None and out.dtype == dtype and out.shape == x_d.shape
  util.assert_close_layer(out.float().cpu().numpy(), want_out, 'out')
  dx0 = dx if mask is None else dx[0]
  util.assert_close_layer(dx0.float().cpu().numpy(), want_dx, 'dx')
  for name, g, w in zip(('dw_q', 'dw_v', 'dw_o'), dw, want_dw):
    assert g.dtype == torch.float32
    util.assert_close_layer(g.cpu().numpy(), w, name)
  # cotangent = ones, as the reference test does (efficient_attention_test.py:81)
  ones = np.ones_like(dout)
  _, _, want_dx1, want_dw1 = O.forward_and_or_backward(cfg, x, weights, buckets=buckets, mask=mask, output_grad=ones,
                                                       update_state=False)
  out2, _, dx1, dw1 = layer.forward_and_or_backward(
      inputs, w_d, state, None, output_grad=torch.ones_like(x_d), compute_output=False, update_state=False)
  assert out2 is None
  util.assert_close_layer((dx1 if mask is None else dx1[0]).float().cpu().numpy(), want_dx1, 'dx(ones)')
  for name, g, w in zip(('dw_q', 'dw_v', 'dw_o'), dw1, want_dw1):
    util.assert_close_layer(g.cpu().numpy(), w, name + '(ones)')


@pytest.mark.parametrize('B,L,D,dtype,cfg', LAYER_CASES[:3])
def test_layer_forward_update_state(B, L, D, dtype, cfg):
  """update_state=True: buckets bit-exact vs the oracle hashing the DEVICE q (hash_vecs granularity), state shapes
  as efficient_attention_test.py:322-328, output within tolerance of the oracle run on those buckets."""
  from trax_b200 import ops, _lib
  x, weights, rot, _, mask = _case(22, B, L, D, cfg, dtype)
  layer = _layer(cfg)
  layer._rotations_override = torch.from_numpy(rot).cuda()
  w_d = tuple(torch.from_numpy(w).cuda() for w in weights)
  x_d = torch.from_numpy(x).cuda().to(dtype)
  layer.init(__import__('trax_b200').ShapeDtype((B, L, D)))
  layer.weights = w_d
  out = layer.forward(x_d)
  buckets, rng_state = layer.state
  assert tuple(buckets.shape) == (B * cfg.n_heads, cfg.n_hashes * L) and buckets.dtype == torch.int32
  assert tuple(rng_state.shape) == (B * cfg.n_heads, 2)
  # device q (bf16) -> oracle hash
  factors = O.bucket_factors(cfg.n_buckets, L, cfg.chunk_len)
  dims = _lib.make_dims(B, cfg.n_heads, L, D, 64, 64, cfg.chunk_len, 1, 0, cfg.n_hashes, factors, True, False, 1)
  wqv, _ = ops.pack_weights(dims, *w_d)
  qv = ops.project_qv(dims, x_d.to(torch.bfloat16).contiguous(), wqv).float().cpu().numpy()
  got_b = buckets.cpu().numpy()
  for u in range(B * cfg.n_heads):
    b, h = divmod(u, cfg.n_heads)
    want_b = O.hash_vectors(cfg, np.ascontiguousarray(qv[b, :, h, :64]), rot[u])
    np.testing.assert_array_equal(got_b[u], want_b)
  want_out, _, _, _ = O.forward_and_or_backward(cfg, x, weights, buckets=got_b, update_state=False)
  util.assert_close_layer(out.float().cpu().numpy(), want_out, 'out')
  # how far the device q is from the fp32 q: fraction of bucket ids that differ from hashing the oracle's own q
  b_ref = O.forward_and_or_backward(cfg, x, weights, rotations=rot, update_state=True)[1]
  print('bucket-id mismatch vs oracle-q hashing: %.4f' % float((b_ref != got_b).mean()))


def test_determinism_and_rng_advance():
  """efficient_attention_test.py:210-236: same seeds -> same output; the state's rng advances (EA:1928)."""
  import trax_b200
  cfg = util.make_cfg(H=2, C=64, nh=2, n_buckets=16)
  x = torch.randn(2, 256, 64, device='cuda', generator=torch.Generator('cuda').manual_seed(0))
  outs = []
  for _ in range(3):
    layer = _layer(cfg)
    layer.init(trax_b200.ShapeDtype((2, 256, 64)), rng=np.array([1, 2], np.uint32))
    s0 = layer.state[1].clone()
    outs.append(layer.forward(x))
    assert not torch.equal(layer.state[1].view(torch.int32), s0.view(torch.int32))
  assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
  out_b = layer.forward(x)                       # second step: new rotations from the advanced key
  assert not torch.equal(out_b, outs[0])


def test_unit_batching_invariance():
  """Analogue of n_parallel_heads ∈ {1,3,6,12} (efficient_attention_test.py:136-156): examples processed together
  equal examples processed one at a time."""
  cfg = util.make_cfg(H=2, C=64, nh=2, n_buckets=8)
  B, L, D = 3, 256, 64
  x, weights, rot, dout, _ = _case(23, B, L, D, cfg, torch.bfloat16)
  buckets = O.forward_and_or_backward(cfg, x, weights, rotations=rot, update_state=True)[1]
  layer = _layer(cfg)
  w_d = tuple(torch.from_numpy(w).cuda() for w in weights)
  x_d, g_d = torch.from_numpy(x).cuda().bfloat16(), torch.from_numpy(dout).cuda().bfloat16()
  b_d = torch.from_numpy(buckets).cuda()
  rng0 = torch.zeros((B * 2, 2), dtype=torch.int32, device='cuda')
  out, _, dx, dw = layer.forward_and_or_backward(x_d, w_d, (b_d, rng0), None, output_grad=g_d, update_state=False)
  dw_sum = [torch.zeros_like(g) for g in dw]
  for b in range(B):
    o1, _, dx1, dw1 = layer.forward_and_or_backward(
        x_d[b:b + 1].contiguous(), w_d, (b_d[2 * b:2 * b + 2].contiguous(), rng0[:2]), None,
        output_grad=g_d[b:b + 1].contiguous(), update_state=False)
    assert torch.equal(o1[0], out[b]) and torch.equal(dx1[0], dx[b])
    for acc, g in zip(dw_sum, dw1):
      acc += g
  for a, g in zip(dw_sum, dw):
    torch.testing.assert_close(a, g, rtol=1e-3, atol=1e-3)


def test_masked_invariance():
  """efficient_attention_test.py:238-284: changing masked-out inputs leaves unmasked outputs unchanged."""
  cfg = util.make_cfg(H=2, C=64, nb=1, na=1, nh=2, n_buckets=8, causal=False, masked=True)
  B, L, D = 1, 256, 64
  x, weights, rot, _, _ = _case(24, B, L, D, cfg, torch.float32)
  mask = np.ones((B, L), bool)
  mask[:, L // 2:] = False
  layer = _layer(cfg)
  layer._rotations_override = torch.from_numpy(rot).cuda()
  w_d = tuple(torch.from_numpy(w).cuda() for w in weights)
  st = (torch.zeros((2, 2 * L), dtype=torch.int32, device='cuda'), torch.zeros((2, 2), dtype=torch.int32, device='cuda'))
  x1 = torch.from_numpy(x).cuda()
  x2 = x1.clone()
  x2[:, L // 2:] = torch.randn_like(x2[:, L // 2:])
  m = torch.from_numpy(mask).cuda()
  o1 = layer.forward_and_or_backward((x1, m), w_d, st, None)[0]
  o2 = layer.forward_and_or_backward((x2, m), w_d, st, None)[0]
  torch.testing.assert_close(o1[:, :L // 2], o2[:, :L // 2], rtol=1e-5, atol=1e-5)


def test_autograd_through_pure_fn():
  """base.py:585-590, 644-673: pure_fn routes through the custom backward; torch autograd sees it."""
  import trax_b200
  cfg = util.make_cfg(H=2, C=64, nh=2, n_buckets=8)
  B, L, D = 1, 256, 64
  x, weights, rot, dout, _ = _case(25, B, L, D, cfg, torch.float32)
  layer = _layer(cfg)
  layer._rotations_override = torch.from_numpy(rot).cuda()
  layer.init(trax_b200.ShapeDtype((B, L, D)))
  w_d = tuple(torch.from_numpy(w).cuda().requires_grad_() for w in weights)
  x_d = torch.from_numpy(x).cuda().requires_grad_()
  out, new_state = layer.pure_fn(x_d, w_d, layer.state, None)
  (out * torch.from_numpy(dout).cuda()).sum().backward()
  buckets = new_state[0].cpu().numpy()
  _, _, want_dx, want_dw = O.forward_and_or_backward(cfg, x, weights, buckets=buckets, output_grad=dout, update_state=False)
  util.assert_close_layer(x_d.grad.cpu().numpy(), want_dx, 'dx')
  for name, w, g in zip(('dw_q', 'dw_v', 'dw_o'), w_d, want_dw):
    util.assert_close_layer(w.grad.cpu().numpy(), g, name)


def test_max_length_for_buckets_state_layout():
  """EA:1880, 1930-1941: state is padded to n_hashes*max_length_for_buckets and only the prefix is used."""
  import trax_b200
  cfg = util.make_cfg(H=2, C=64, nh=2, n_buckets=8, max_len=512)
  layer = _layer(cfg)
  layer.init(trax_b200.ShapeDtype((1, 256, 64)))
  assert tuple(layer.state[0].shape) == (2, 2 * 512)
  x = torch.randn(1, 256, 64, device='cuda')
  out = layer.forward(x)
  assert tuple(layer.state[0].shape) == (2, 2 * 512)
  assert (layer.state[0][:, 2 * 256:] == 0).all()
  out2 = layer.forward_and_or_backward(x, layer.weights, layer.state, None, update_state=False)[0]
  assert torch.equal(out, out2)


def test_unsupported_shapes_are_rejected():
  """no fallback — the reference's odd test shapes (d_qk=7, d_v=17, chunk_len=5) raise."""
  import trax_b200
  from trax_b200 import _lib
  layer = trax_b200.LSHSelfAttention(n_heads=6, d_qk=7, d_v=17, chunk_len=5, n_hashes=2, n_buckets=4, causal=True)
  layer.init(trax_b200.ShapeDtype((2, 10, 13)))
  with pytest.raises(_lib.LshAttnError):
    layer.forward(torch.zeros(2, 10, 13, device='cuda'))
  with pytest.raises(NotImplementedError):                          # predict mode is built (tests/test_zgpu_predict.py) but takes
    trax_b200.LSHSelfAttention(mode='predict', causal=True, masked=True)   # one input (EA:1999-2001)
  with pytest.raises(ValueError):
    trax_b200.LSHSelfAttention(attention_dropout=1.0)
  with pytest.raises(ValueError):
    trax_b200.LSHSelfAttention(n_heads=6, n_parallel_heads=4)


def test_host_buffers_roundtrip():
  """Host (pinned) tensors in -> host tensors out: the e2e path bench.py times."""
  import trax_b200
  cfg = util.make_cfg(H=2, C=64, nh=1, n_buckets=8)
  layer = _layer(cfg)
  layer.init(trax_b200.ShapeDtype((1, 256, 64)))
  x = torch.randn(1, 256, 64).pin_memory()
  out = layer.forward(x)
  assert not out.is_cuda and out.shape == x.shape
  out_d = layer.forward_and_or_backward(x.cuda(), layer.weights, layer.state, None, update_state=False)[0]
  torch.testing.assert_close(out, out_d.cpu(), rtol=0, atol=0)


def test_async_host_io_matches_sync():
  """Asynchronous dispatch (side-stream copies, staging rings, one-shot reuse of forward's upload in backward)
  returns the same bits as the default synchronous host path and as device-resident inputs."""
  import trax_b200
  cfg = util.make_cfg(H=2, C=128, nh=2, n_buckets=8)
  layer = _layer(cfg)
  layer.init(trax_b200.ShapeDtype((1, 512, 64)))
  g = torch.Generator().manual_seed(3)
  xs = [torch.randn(1, 512, 64, generator=g).pin_memory() for _ in range(4)]
  gs = [torch.randn(1, 512, 64, generator=g).pin_memory() for _ in range(4)]
  layer.forward(xs[0])                               # fixes the buckets in the state; every call below re-uses them
  state = layer.state

  def run(x, gr):
    # forward call then backward call on the same host tensor (the second upload of x is served from the first)
    out = layer.forward_and_or_backward(x, layer.weights, state, None, update_state=False)[0]
    _, _, dx, dw = layer.forward_and_or_backward(x, layer.weights, state, None, output_grad=gr, compute_output=False,
                                                 update_state=False)
    return out, dx, dw
  ref = []
  for x, gr in zip(xs, gs):
    out, dx, dw = run(x, gr)
    ref.append([t.clone() for t in (out, dx) + tuple(dw)])
  trax_b200.set_async_host_io(True)
  try:
    got = []
    for x, gr in zip(xs, gs):
      out, dx, dw = run(x, gr)
      trax_b200.synchronize()                      # staging buffers are recycled after three calls: copy out now
      got.append([t.clone() for t in (out, dx) + tuple(dw)])
  finally:
    trax_b200.set_async_host_io(False)
  for r, o in zip(ref, got):
    for a, b in zip(r, o):
      torch.testing.assert_close(a, b, rtol=0, atol=0)
  h2d, d2h = trax_b200.host_io_bytes(reset=True)
  assert h2d > 0 and d2h > 0


def test_hashing_call_and_recompute_agree_bitwise():
  """A forward call that hashes takes the per-token scales / normalised keys from the hash kernel; a call that re-uses
  the buckets (the backward's recompute) takes them from qscale_kernel.  Same operation order => the same output bits."""
  import trax_b200
  cfg = util.make_cfg(H=4, C=128, nh=4, n_buckets=None)
  layer = _layer(cfg)
  layer.init(trax_b200.ShapeDtype((2, 1024, 256)))
  x = torch.randn(2, 1024, 256, device='cuda', generator=torch.Generator('cuda').manual_seed(5)).bfloat16()
  out1 = layer.forward(x)                                                               # update_state=True
  out2 = layer.forward_and_or_backward(x, layer.weights, layer.state, None, update_state=False)[0]
  assert torch.equal(out1, out2)


def test_output_dropout_is_a_shared_column_mask():
  """EA:1995-1996, 271-280: out = (o w_o) * keep / keep_prob with ONE (d_model,) keep-mask for all positions, heads and
  examples.  Explicit mask vs the oracle (out, dx, dW); rng-derived mask: same rng -> same mask in the backward call,
  dropped columns exactly zero, other rng -> other mask."""
  import trax_b200
  B, L, D = 2, 512, 128
  cfg = util.make_cfg(H=2, C=128, nh=2, n_buckets=8)
  x, weights, rot, dout, _ = _case(31, B, L, D, cfg, torch.float32)
  rate = 0.25
  keep = np.random.default_rng(5).random(D) < 1 - rate
  mult = keep / (1 - rate)
  want_out, buckets, _, _ = O.forward_and_or_backward(cfg, x, weights, rotations=rot, out_keep=mult)
  _, _, want_dx, want_dw = O.forward_and_or_backward(cfg, x, weights, buckets=buckets, output_grad=dout, out_keep=mult,
                                                     update_state=False)
  layer = trax_b200.LSHSelfAttention(n_heads=2, d_qk=64, d_v=64, causal=True, chunk_len=128, n_hashes=2, n_buckets=8,
                                     output_dropout=rate)
  layer.init(trax_b200.ShapeDtype((B, L, D)))
  w_d = tuple(torch.from_numpy(w).cuda() for w in weights)
  x_d, g_d = torch.from_numpy(x).cuda(), torch.from_numpy(dout).cuda()
  state = (torch.from_numpy(buckets).cuda(), layer.state[1])
  layer._out_keep_override = keep
  out, _, dx, dw = layer.forward_and_or_backward(x_d, w_d, state, None, output_grad=g_d, update_state=False)
  util.assert_close_layer(out.cpu().numpy(), want_out, 'out')
  assert (out[..., torch.from_numpy(~keep).cuda()] == 0).all()
  util.assert_close_layer(dx.cpu().numpy(), want_dx, 'dx')
  for name, g, w in zip(('dw_q', 'dw_v', 'dw_o'), dw, want_dw):
    util.assert_close_layer(g.cpu().numpy(), w, name)
  assert (dw[2][..., torch.from_numpy(~keep).cuda()] == 0).all()
  # rng-derived masks
  layer._out_keep_override = None
  with pytest.raises(ValueError):
    layer.forward_and_or_backward(x_d, w_d, state, None, update_state=False)
  rng_a, rng_b = np.array([1, 2], np.uint32), np.array([1, 3], np.uint32)
  out_a, _, _, _ = layer.forward_and_or_backward(x_d, w_d, state, rng_a, update_state=False)
  out_a2, _, _, dw_a = layer.forward_and_or_backward(x_d, w_d, state, rng_a, output_grad=g_d, update_state=False)
  out_b, _, _, _ = layer.forward_and_or_backward(x_d, w_d, state, rng_b, update_state=False)
  assert torch.equal(out_a, out_a2)
  dropped_a, dropped_b = (out_a == 0).all(dim=0).all(dim=0), (out_b == 0).all(dim=0).all(dim=0)
  assert 0 < int(dropped_a.sum()) < D // 2 and not torch.equal(dropped_a, dropped_b)
  assert (dw_a[2][..., dropped_a] == 0).all() and (dw_a[2][..., ~dropped_a] != 0).any()
  # eval mode switches dropout off (EA:1790-1795)
  ev = trax_b200.LSHSelfAttention(n_heads=2, d_qk=64, d_v=64, causal=True, chunk_len=128, n_hashes=2, n_buckets=8,
                                  output_dropout=rate, mode='eval')
  ev.init(trax_b200.ShapeDtype((B, L, D)))
  out_e, _, _, _ = ev.forward_and_or_backward(x_d, w_d, state, None, update_state=False)
  assert not (out_e == 0).all(dim=0).all(dim=0).any()


def test_fresh_host_tensors_are_never_served_from_a_stale_device_copy():
  """forward-only loops and fused calls on FRESH host tensors of one shape (the allocator hands the freed
  address out again) must see their own data; only forward(x) -> backward(x) of the same tensor object shares the upload."""
  import trax_b200
  B, L, D = 1, 512, 128
  cfg = util.make_cfg(H=2, C=128, nh=2, n_buckets=8)
  layer = _layer(cfg)
  layer.init(trax_b200.ShapeDtype((B, L, D)), rng=np.array([3, 4], np.uint32))
  rng = np.random.default_rng(0)
  factors = O.bucket_factors(cfg.n_buckets, L, cfg.chunk_len)
  layer._rotations_override = torch.from_numpy(
      rng.standard_normal((B * cfg.n_heads, 64, cfg.n_hashes, sum(factors) // 2)).astype(np.float32))
  def fresh(i):
    return torch.from_numpy(np.random.default_rng(100 + i).standard_normal((B, L, D)).astype(np.float32))
  for i in range(6):
    want = layer.forward(fresh(i).cuda()).cpu()
    got = layer.forward(fresh(i))                       # fresh host tensor, same shape, likely the same address
    assert torch.equal(got, want), 'host call %d computed on another tensor\'s data' % i
  # fused call on host tensors after a forward of a DIFFERENT tensor: must not take the stash
  x0, x1, g = fresh(20), fresh(21), fresh(22)
  layer.forward(x0)
  state = layer.state
  _, _, dx_h, _ = layer.forward_and_or_backward(x1, layer.weights, state, None, output_grad=g, compute_output=False,
                                                update_state=False)
  _, _, dx_d, _ = layer.forward_and_or_backward(x1.cuda(), layer.weights, state, None, output_grad=g.cuda(),
                                                compute_output=False, update_state=False)
  assert torch.equal(dx_h, dx_d.cpu())
  # the legitimate pairing moves x once
  trax_b200.host_io_bytes(reset=True)
  out = layer.forward(x0)
  layer.backward(x0, out, g, layer.weights, None, layer.state, None)
  h2d, _ = trax_b200.host_io_bytes(reset=True)
  assert h2d == (x0.numel() + g.numel() + layer._rotations_override.numel()) * 4, h2d     # x once, the cotangent, the rotations
  # ... and an in-place change torch can see invalidates it
  out = layer.forward(x0)
  x0.add_(1.0)
  dx_h, _ = layer.backward(x0, out, g, layer.weights, None, layer.state, None)
  dx_d, _ = layer.backward(x0.cuda(), out, g.cuda(), layer.weights, None, layer.state, None)
  assert torch.equal(dx_h, dx_d.cpu())


@pytest.mark.parametrize('B,L,D,dtype,cfg', [
    (2, 512, 128, torch.bfloat16, util.make_cfg(H=4, C=128, nh=4, n_buckets=None)),               # config-2 style (tcgen05 forward)
    (1, 1024, 256, torch.float32, util.make_cfg(H=2, C=64, nh=1, n_buckets=32)),                  # BASELINE config 1
])
def test_layer_with_attention_and_output_dropout(B, L, D, dtype, cfg):
  """Both reference training configs use attention_dropout = 0.2 (reformer_enwik8.gin:94, reformer_imagenet64.gin:69):
  forward, fused forward+backward and backward against the oracle with the SAME keep matrices (EA:254-262, 271-280)."""
  x, weights, rot, dout, _ = _case(61, B, L, D, cfg, dtype)
  layer = _layer(cfg, attention_dropout=0.2, output_dropout=0.1)
  key = np.array([7, 9], np.uint32)
  W = cfg.chunk_len * (1 + cfg.n_chunks_before + cfg.n_chunks_after)
  attn_keep = layer._attention_multiplier(key, 'cpu').numpy().astype(np.float64)
  out_keep = layer._output_multiplier(key, D, 'cpu').numpy().astype(np.float64)
  assert attn_keep.shape == (cfg.chunk_len, W) and 0.1 < (attn_keep == 0).mean() < 0.3
  want_out, buckets, _, _ = O.forward_and_or_backward(cfg, x, weights, rotations=rot, update_state=True, attn_keep=attn_keep,
                                                      out_keep=out_keep)
  _, _, want_dx, want_dw = O.forward_and_or_backward(cfg, x, weights, buckets=buckets, output_grad=dout, update_state=False,
                                                     attn_keep=attn_keep, out_keep=out_keep)
  w_d = tuple(torch.from_numpy(w).cuda() for w in weights)
  x_d = torch.from_numpy(x).cuda().to(dtype)
  state = (torch.from_numpy(buckets).cuda(), torch.zeros((B * cfg.n_heads, 2), dtype=torch.int32, device='cuda'))
  out, _, dx, dw = layer.forward_and_or_backward(x_d, w_d, state, key, output_grad=torch.from_numpy(dout).cuda().to(dtype),
                                                 compute_output=True, update_state=False)
  util.assert_close_layer(out.float().cpu().numpy(), want_out, 'out')
  util.assert_close_layer(dx.float().cpu().numpy(), want_dx, 'dx')
  for name, g, w in zip(('dw_q', 'dw_v', 'dw_o'), dw, want_dw):
    util.assert_close_layer(g.cpu().numpy(), w, name)
  # the masks are functions of rng: another key changes the result, the same key reproduces it bit for bit
  out2, _, _, _ = layer.forward_and_or_backward(x_d, w_d, state, key, compute_output=True, update_state=False)
  out3, _, _, _ = layer.forward_and_or_backward(x_d, w_d, state, np.array([8, 9], np.uint32), compute_output=True,
                                                update_state=False)
  assert torch.equal(out, out2) and not torch.equal(out, out3)
  # eval mode switches both off (EA:1790-1795)
  ev = _layer(cfg, attention_dropout=0.2, output_dropout=0.1, mode='eval')
  want_plain, _, _, _ = O.forward_and_or_backward(cfg, x, weights, buckets=buckets, update_state=False)
  got_plain, _, _, _ = ev.forward_and_or_backward(x_d, w_d, state, None, compute_output=True, update_state=False)
  util.assert_close_layer(got_plain.float().cpu().numpy(), want_plain, 'eval-mode out')


def test_concurrent_streams_get_their_own_scratch_and_agree():
  """Two calls on different streams must not share the library's scratch buffer (ops.workspace is keyed by device AND
  stream): run the same forward on two side streams at once and compare with the default-stream result."""
  import trax_b200
  from trax_b200 import ops
  cfg = util.make_cfg(H=4, C=128, nh=2, n_buckets=None)
  layer = _layer(cfg)
  layer.init(trax_b200.ShapeDtype((1, 2048, 256)))
  g = torch.Generator('cuda').manual_seed(11)
  xs = [torch.randn(1, 2048, 256, device='cuda', generator=g).bfloat16() for _ in range(2)]
  layer.forward(xs[0])                       # fills the bucket state; any valid permutation serves both inputs below
  state = layer.state
  want = [layer.forward_and_or_backward(x, layer.weights, state, None, update_state=False)[0] for x in xs]
  torch.cuda.synchronize()
  streams = [torch.cuda.Stream(), torch.cuda.Stream()]
  got, bufs = [], []
  for x, st in zip(xs, streams):
    with torch.cuda.stream(st):
      got.append(layer.forward_and_or_backward(x, layer.weights, state, None, update_state=False)[0])
      bufs.append(ops.workspace(x.device, 1).data_ptr())
  torch.cuda.synchronize()
  assert bufs[0] != bufs[1]
  for a, b in zip(got, want):
    assert torch.equal(a, b)
