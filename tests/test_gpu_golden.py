"""The CUDA path against the COMMITTED golden fixtures: tests/golden/reference_live.npz holds outputs of the reference's own
code (tests/golden/make_reference_golden.py); lsh_small.npz / reversible_c128.npz hold oracle outputs (make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _cu(a, dt=torch.float32):
  return torch.from_numpy(np.asarray(a, np.float32)).cuda().to(dt)


def test_layer_against_lsh_small_fixture():
  """lsh_small.npz: B1 L128 D32 H2 chunk 64 2 hashes n_buckets [4, 2] (mma.sync path): the stable permutation of the
  fixture's buckets bit-exact, out / dx / dW within tolerance."""
  import trax_b200
  from trax_b200 import _lib, ops
  g = np.load(os.path.join(HERE, 'lsh_small.npz'))
  layer = trax_b200.LSHSelfAttention(n_heads=2, d_qk=64, d_v=64, causal=True, chunk_len=64, n_hashes=2, n_buckets=[4, 2])
  layer.init(trax_b200.ShapeDtype(g['x'].shape))
  layer.weights = tuple(_cu(g[k]) for k in ('w_q', 'w_v', 'w_o'))
  x = _cu(g['x'])
  # the fixture's buckets go in through the state (update_state=False, the reversible-layer call): the layer's own hash sees
  # bf16 tensor-core projections, where a near-tie of the argmax may legitimately fall the other way (the hash kernel's
  # bit-exactness on identical inputs is what tests/test_gpu_stages.py checks)
  state = (_cu(g['buckets']).to(torch.int32), layer.state[1])
  out, _, dx, dw = layer.forward_and_or_backward(x, layer.weights, state, None, output_grad=_cu(g['dout']),
                                                 compute_output=True, update_state=False)
  util.assert_close_layer(out.cpu().numpy(), g['out'], 'out')
  util.assert_close_layer(dx.cpu().numpy(), g['dx'], 'dx')
  for k, w in zip(('dw_q', 'dw_v', 'dw_o'), dw):
    util.assert_close_layer(w.cpu().numpy(), g[k], k)
  dims = _lib.make_dims(1, 2, 128, 32, 64, 64, 64, 1, 0, 2, [4, 2], True, False, _lib.LSH_DTYPE_F32)
  sticker, undo = ops.sort(dims, state[0])
  np.testing.assert_array_equal(sticker[0].cpu().numpy(), g['sticker0'])
  np.testing.assert_array_equal(undo[0].cpu().numpy(), g['undo0'])


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_reversible_block_against_reversible_c128_fixture(dtype):
  """reversible_c128.npz: B1 L256 D256 H2 chunk 128 2 hashes (tcgen05 path) inside ReversibleHalfResidual."""
  import trax_b200
  from tests.golden import make_golden as G
  g = np.load(os.path.join(HERE, 'reversible_c128.npz'))
  cfg, x1, x2, ct1, ct2, scale, bias, aw, rot = G.reversible_case()
  attn = trax_b200.LSHSelfAttention(n_heads=cfg.n_heads, d_qk=64, d_v=64, causal=True, chunk_len=cfg.chunk_len,
                                    n_hashes=cfg.n_hashes, n_buckets=cfg.n_buckets)
  block = trax_b200.ReversibleHalfResidual(attn)
  sig = trax_b200.ShapeDtype(x1.shape)
  block.init((sig, sig))
  block.weights = ((_cu(scale), _cu(bias)), tuple(_cu(w) for w in aw))
  # the fixture's buckets go in through new_state (see the note in the test above); reverse_and_grad recomputes the forward
  state = ((), (_cu(g['buckets']).to(torch.int32), block.state[1][1]))
  ctx = _cu(x2, dtype)
  (rx1, _), ((_, g2), ((ds, db), dw)) = block.reverse_and_grad((_cu(g['y1'], dtype), ctx), (_cu(ct1, dtype), _cu(ct2, dtype)),
                                                               block.weights, None, state, None)
  util.assert_close_layer(rx1.float().cpu().numpy(), x1, 'reconstructed x1', rtol=3e-2)
  util.assert_close_layer(g2.float().cpu().numpy(), g['ct2_out'], 'ct_x2')
  util.assert_close_layer(ds.cpu().numpy(), g['d_scale'], 'd_scale')
  util.assert_close_layer(db.cpu().numpy(), g['d_bias'], 'd_bias')
  for k, w in zip(('dw_q', 'dw_v', 'dw_o'), dw):
    util.assert_close_layer(w.float().cpu().numpy(), g[k], k)


# ---- reference_live.npz: what google/trax's own code returned for these inputs (tests/golden/make_reference_golden.py) ----
def _directional_ok(grad, direction, want, name):
  """<grad, dir> against the reference's directional derivative.  An elementwise relative error e on grad moves the inner
  product by about e * |grad| |dir| / sqrt(n); allow four times that at e = RTOL."""
  grad, direction = np.asarray(grad, np.float64), np.asarray(direction, np.float64)
  got = float((grad * direction).sum())
  tol = 4 * util.RTOL * np.linalg.norm(grad) * np.linalg.norm(direction) / np.sqrt(grad.size) + util.ATOL
  assert abs(got - want) <= tol, '%s: <grad, dir> = %.6g, the reference forward differentiates to %.6g (tol %.3g)' % (
      name, got, want, tol)


@pytest.mark.parametrize('name', ['lsh_c128', 'lsh_c64_auto', 'lsh_masked_factored'])
def test_layer_against_the_live_reference(name):
  """Output vs the reference layer's own output (buckets: the reference's, fed through the state as the reversible block
  does), gradients vs derivatives of the reference's forward and vs the oracle's VJP (pinned to the same numbers on CPU)."""
  import trax_b200
  from oracle import lsh_oracle as O
  from tests.golden import reference_cases as RC
  c, d, g = RC.CASES[name], RC.inputs(name), np.load(os.path.join(HERE, 'reference_live.npz'))
  B, H, L, D = c['B'], c['H'], c['L'], c['D']
  layer = trax_b200.LSHSelfAttention(n_heads=H, d_qk=64, d_v=64, causal=c['causal'], masked=c['masked'], chunk_len=c['C'],
                                     n_chunks_before=c['nb'], n_chunks_after=c['na'], n_hashes=c['nh'],
                                     n_buckets=c['n_buckets'])
  sig = trax_b200.ShapeDtype((B, L, D))
  layer.init((sig, trax_b200.ShapeDtype((B, L))) if c['masked'] else sig)
  weights = tuple(_cu(d[k]) for k in ('w_q', 'w_v', 'w_o'))
  inputs = (_cu(d['x']), torch.from_numpy(d['mask']).cuda()) if c['masked'] else _cu(d['x'])
  state = (torch.from_numpy(g[name + '/buckets']).cuda(), layer.state[1])
  out, _, dx, dw = layer.forward_and_or_backward(inputs, weights, state, None, output_grad=_cu(d['dout']),
                                                 compute_output=True, update_state=False)
  util.assert_close_layer(out.cpu().numpy(), g[name + '/out'], 'out vs the reference')
  dx = dx[0] if c['masked'] else dx
  cfg = O.LSHConfig(n_heads=H, d_qk=64, d_v=64, causal=c['causal'], masked=c['masked'], chunk_len=c['C'],
                    n_chunks_before=c['nb'], n_chunks_after=c['na'], n_hashes=c['nh'], n_buckets=c['n_buckets'])
  _, _, want_dx, want_dw = O.forward_and_or_backward(cfg, d['x'], (d['w_q'], d['w_v'], d['w_o']), buckets=g[name + '/buckets'],
                                                     mask=d['mask'], output_grad=d['dout'], update_state=False)
  for key, got, want in zip(('x', 'w_q', 'w_v', 'w_o'), (dx,) + tuple(dw), (want_dx,) + tuple(want_dw)):
    util.assert_close_layer(got.float().cpu().numpy(), want, 'd' + key)
    _directional_ok(got.float().cpu().numpy(), d['dir_' + key], float(g[name + '/ddir_' + key]), 'd' + key)


def test_pure_core_against_the_live_reference():
  """PureLSHSelfAttention with update_state=True on bf16-exact inputs and the reference's rotations: the device's buckets
  equal the reference's bit for bit; output and (dqk, dv) follow."""
  import trax_b200
  from tests.golden import reference_cases as RC
  name = 'pure_c128'
  c, d, g = RC.CASES[name], RC.inputs(name), np.load(os.path.join(HERE, 'reference_live.npz'))
  BH, L = c['B'] * c['H'], c['L']
  layer = trax_b200.PureLSHSelfAttention(n_heads=c['H'], d_qk=64, d_v=64, causal=True, chunk_len=c['C'], n_hashes=c['nh'],
                                         n_buckets=c['n_buckets'])
  sig = trax_b200.ShapeDtype((BH, L, 64))
  layer.init((sig, sig))
  layer._rotations_override = torch.from_numpy(g[name + '/rot'])
  inputs = (_cu(d['qk']), _cu(d['v']))
  out = layer.forward(inputs)
  np.testing.assert_array_equal(layer.state[0].cpu().numpy(), g[name + '/buckets'])
  util.assert_close(out.cpu().numpy(), g[name + '/out'], 'out vs the reference')
  (dqk, dv), _ = layer.backward(inputs, out, _cu(d['dout']), (), None, layer.state, None)
  _directional_ok(dqk.cpu().numpy(), d['dir_qk'], float(g[name + '/ddir_qk']), 'dqk')
  _directional_ok(dv.cpu().numpy(), d['dir_v'], float(g[name + '/ddir_v']), 'dv')


@pytest.mark.parametrize('name', ['wrapper_c128', 'wrapper_rotary'])
def test_wrapper_against_the_live_reference(name):
  """PureLSHSelfAttentionWrapper vs the reference's Serial run (reference buckets through the state, as in the reversible
  backward pass that `forward_and_or_backward` serves, EA:3566-3568)."""
  import trax_b200
  from oracle import lsh_oracle as O
  from tests.golden import reference_cases as RC
  c, d, g = RC.CASES[name], RC.inputs(name), np.load(os.path.join(HERE, 'reference_live.npz'))
  wrap = trax_b200.PureLSHSelfAttentionWrapper(
      n_heads=c['H'], d_qk=64, d_v=64, causal=True, bias=c['bias'], num_weights=c['num_weights'], weights_format='model',
      rotary_position_emb=c['rotary'], chunk_len=c['C'], n_hashes=c['nh'], n_buckets=c['n_buckets'])
  wrap.init(trax_b200.ShapeDtype((c['B'], c['L'], c['D'])))
  dev = lambda w: tuple(_cu(l) for l in w) if isinstance(w, tuple) else _cu(w)
  weights = (tuple(dev(w) for w in d['qkv']), (), (), dev(d['dense']))
  state = ((), (torch.from_numpy(g[name + '/buckets']).cuda(), wrap.state[1][1]), (), ())
  out, _, dx, dw = wrap.forward_and_or_backward(_cu(d['x']), weights, state, None, output_grad=_cu(d['dout']),
                                                update_state=False)
  util.assert_close_layer(out.cpu().numpy(), g[name + '/out'], 'out vs the reference')
  cfg = O.LSHConfig(n_heads=c['H'], d_qk=64, d_v=64, causal=True, masked=False, chunk_len=c['C'], n_chunks_before=1,
                    n_chunks_after=0, n_hashes=c['nh'], n_buckets=c['n_buckets'])
  _, _, want_dx, (want_qkv, want_dense) = O.pure_lsh_wrapper(cfg, d['x'], d['qkv'], d['dense'], buckets=g[name + '/buckets'],
                                                             output_grad=d['dout'], rotary_position_emb=c['rotary'])
  leaves = lambda w: list(w) if isinstance(w, tuple) else [w]
  util.assert_close_layer(dx.cpu().numpy(), want_dx, 'dx')
  _directional_ok(dx.cpu().numpy(), d['dir_x'], float(g[name + '/ddir_x']), 'dx')
  for i in range(c['num_weights']):
    for got, want in zip(leaves(dw[0][i]), leaves(want_qkv[i])):
      util.assert_close_layer(got.cpu().numpy(), want, 'd_qkv[%d]' % i)
    got = np.concatenate([l.cpu().numpy().ravel() for l in leaves(dw[0][i])])
    direction = np.concatenate([l.ravel() for l in leaves(d['dir_qkv'][i])])
    _directional_ok(got, direction, float(g[name + '/ddir_qkv%d' % i]), 'd_qkv[%d]' % i)
  for got, want in zip(leaves(dw[3]), leaves(want_dense)):
    util.assert_close_layer(got.cpu().numpy(), want, 'd_dense')


def test_reversible_block_against_the_live_reference():
  """ReversibleHalfResidual(LayerNorm, LSHSelfAttention) vs the reference block's forward output and the derivatives of
  that forward (tcgen05 shape)."""
  import trax_b200
  from tests.golden import reference_cases as RC
  name = 'reversible_c128'
  c, d, g = RC.CASES[name], RC.inputs(name), np.load(os.path.join(HERE, 'reference_live.npz'))
  attn = trax_b200.LSHSelfAttention(n_heads=c['H'], d_qk=64, d_v=64, causal=True, chunk_len=c['C'], n_hashes=c['nh'],
                                    n_buckets=c['n_buckets'])
  block = trax_b200.ReversibleHalfResidual(attn)
  sig = trax_b200.ShapeDtype((c['B'], c['L'], c['D']))
  block.init((sig, sig))
  block.weights = ((_cu(d['scale']), _cu(d['bias'])), tuple(_cu(d[k]) for k in ('w_q', 'w_v', 'w_o')))
  attn._rotations_override = torch.from_numpy(g[name + '/rot'])
  y1, ctx = block.forward((_cu(d['x1']), _cu(d['x2'])))              # own hash (bf16 projections): output still close
  mismatch = float((block.state[1][0].cpu().numpy() != g[name + '/buckets']).mean())
  assert mismatch <= 0.05, 'device buckets differ from the reference at %.2f%% of the tokens' % (100 * mismatch)
  state = ((), (torch.from_numpy(g[name + '/buckets']).cuda(), block.state[1][1]))
  y1_ref = _cu(g[name + '/y1'])
  (rx1, _), ((_, g2), ((ds, db), dw)) = block.reverse_and_grad(
      (y1_ref, ctx), (_cu(d['ct_y1']), torch.zeros_like(ctx)), block.weights, None, state, None)
  util.assert_close_layer(rx1.cpu().numpy(), d['x1'], 'x1 reconstructed from the reference y1', rtol=3e-2)
  for key, got in zip(('x2', 'scale', 'bias', 'w_q', 'w_v', 'w_o'), (g2, ds, db) + tuple(dw)):
    _directional_ok(got.float().cpu().numpy(), d['dir_' + key], float(g[name + '/ddir_' + key]), 'd' + key)
  if mismatch == 0.0:
    util.assert_close_layer(y1.cpu().numpy(), g[name + '/y1'], 'y1 vs the reference')
