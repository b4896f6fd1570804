"""The oracle against the LIVE REFERENCE.  tests/golden/reference_live.npz holds what google/trax's own code returned
(written by tests/golden/make_reference_golden.py through oracle/ref_live.py: the reference's files executed from
/root/reference under the reference's own NumPy backend).  Here:

  * the oracle must reproduce the reference's buckets bit for bit and its float64 outputs to 1e-12;
  * the oracle's analytic VJP must agree with derivatives of the reference's forward function (the reference's backward
    is `jax.vjp` of that function, EA:2399-2421);
  * where /root/reference is present (the build container), the fixture is regenerated from it and must come out equal —
    the committed numbers are the reference's, not ours.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import lsh_oracle as O
from oracle import ref_live
from tests import util
from tests.golden import reference_cases as RC

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
FIXTURE = os.path.join(GOLDEN, 'reference_live.npz')


def _cfg(c):
  return O.LSHConfig(n_heads=c['H'], d_qk=RC.D_HEAD, d_v=RC.D_HEAD, causal=c['causal'], masked=c['masked'],
                     chunk_len=c['C'], n_chunks_before=c['nb'], n_chunks_after=c['na'], n_hashes=c['nh'],
                     n_buckets=c['n_buckets'])


@pytest.mark.parametrize('name', [n for n, c in RC.CASES.items() if c['kind'] == 'lsh'])
def test_oracle_layer_matches_live_reference(name):
  c, d, g = RC.CASES[name], RC.inputs(name), np.load(FIXTURE)
  cfg, weights = _cfg(c), (d['w_q'], d['w_v'], d['w_o'])
  out, buckets, _, _ = O.forward_and_or_backward(cfg, d['x'], weights, rotations=g[name + '/rot'], mask=d['mask'])
  np.testing.assert_array_equal(buckets, g[name + '/buckets'])      # fp32 sequential-FMA hash == the reference's hash
  np.testing.assert_allclose(out, g[name + '/out'], rtol=1e-12, atol=1e-12)
  _, _, dx, dw = O.forward_and_or_backward(cfg, d['x'], weights, buckets=g[name + '/buckets'], mask=d['mask'],
                                           output_grad=d['dout'], update_state=False)
  for key, grad in zip(('x', 'w_q', 'w_v', 'w_o'), (dx,) + tuple(dw)):
    want = float(g[name + '/ddir_' + key])
    np.testing.assert_allclose((grad * d['dir_' + key]).sum(), want, rtol=1e-6, atol=1e-7, err_msg=key)


def test_oracle_pure_core_matches_live_reference():
  name = 'pure_c128'
  c, d, g = RC.CASES[name], RC.inputs(name), np.load(FIXTURE)
  cfg = _cfg(c)
  w_q, w_v, w_o = util.core_identity_weights()
  an = dict(qk=0.0, v=0.0)
  for u in range(c['B'] * c['H']):
    x = np.concatenate([d['qk'][u], d['v'][u]], axis=1)
    r = O.forward_unit(cfg, x, w_q, w_v, w_o, rotations=g[name + '/rot'][u])
    np.testing.assert_array_equal(r.buckets, g[name + '/buckets'][u])
    np.testing.assert_allclose(r.out, g[name + '/out'][u], rtol=1e-12, atol=1e-12)
    grad = O.backward_unit(cfg, r, d['dout'][u])[0]
    an['qk'] += (grad[:, :RC.D_HEAD] * d['dir_qk'][u]).sum()
    an['v'] += (grad[:, RC.D_HEAD:] * d['dir_v'][u]).sum()
  for key in ('qk', 'v'):
    np.testing.assert_allclose(an[key], float(g[name + '/ddir_' + key]), rtol=1e-6, atol=1e-7, err_msg=key)


def test_oracle_hash_auto_factor_list_matches_live_reference():
  """n_buckets=None with 2 L / C > 128: the reference picks the factor list itself (EA:1896-1902)."""
  name = 'hash_auto_factors'
  c, d, g = RC.CASES[name], RC.inputs(name), np.load(FIXTURE)
  assert O.bucket_factors(None, c['L'], c['C']) == [32, 8]
  q = d['x'][0] @ d['w_q'][0]
  for fn in (O.hash_vectors, O.hash_vectors_c):
    np.testing.assert_array_equal(fn(_cfg(c), q.astype(np.float32), g[name + '/rot'][0]), g[name + '/buckets'])


@pytest.mark.skipif(not ref_live.available(), reason='the reference checkout exists only in the build container')
def test_fixture_is_what_the_reference_returns(tmp_path):
  """Re-runs the reference (own process: the import stubs are process-global) and compares with the committed file."""
  out = str(tmp_path / 'regenerated.npz')
  repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  subprocess.run([sys.executable, os.path.join(GOLDEN, 'make_reference_golden.py'), '--out', out], check=True, cwd=repo,
                 timeout=900, stdout=subprocess.DEVNULL)
  new, old = np.load(out), np.load(FIXTURE)
  assert sorted(new.files) == sorted(old.files)
  for k in old.files:
    if old[k].dtype.kind in 'iuU':
      np.testing.assert_array_equal(new[k], old[k], err_msg=k)
    else:
      np.testing.assert_allclose(new[k], old[k], rtol=1e-7 if '/ddir_' in k else 1e-12, atol=1e-12, err_msg=k)


def _leaves(w):
  return list(w) if isinstance(w, tuple) else [w]


@pytest.mark.parametrize('name', [n for n, c in RC.CASES.items() if c['kind'] == 'wrapper'])
def test_oracle_wrapper_matches_live_reference(name):
  """`PureLSHSelfAttentionWrapper` run as the reference's Serial (Dense / Branch / rotary / SplitIntoHeads / core driver /
  MergeHeads / Dense are all the reference's code) vs `oracle.pure_lsh_wrapper`."""
  c, d, g = RC.CASES[name], RC.inputs(name), np.load(FIXTURE)
  cfg = _cfg(c)
  out, buckets, _, _ = O.pure_lsh_wrapper(cfg, d['x'], d['qkv'], d['dense'], rotations=g[name + '/rot'],
                                          rotary_position_emb=c['rotary'])
  np.testing.assert_array_equal(buckets, g[name + '/buckets'])
  np.testing.assert_allclose(out, g[name + '/out'], rtol=1e-12, atol=1e-12)
  _, _, dx, (d_qkv, d_dense) = O.pure_lsh_wrapper(cfg, d['x'], d['qkv'], d['dense'], buckets=g[name + '/buckets'],
                                                  output_grad=d['dout'], rotary_position_emb=c['rotary'])
  inner = lambda grads, dirs: sum((a * b).sum() for a, b in zip(_leaves(grads), _leaves(dirs)))
  np.testing.assert_allclose((dx * d['dir_x']).sum(), float(g[name + '/ddir_x']), rtol=1e-6, atol=1e-7)
  for i in range(c['num_weights']):
    np.testing.assert_allclose(inner(d_qkv[i], d['dir_qkv'][i]), float(g[name + '/ddir_qkv%d' % i]), rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(inner(d_dense, d['dir_dense']), float(g[name + '/ddir_dense']), rtol=1e-6, atol=1e-7)


def test_oracle_reversible_block_matches_live_reference():
  """`ReversibleHalfResidual(LayerNorm(), attention_layer=LSHSelfAttention)`: forward vs the reference's forward, and the
  cotangents `reverse_and_grad` must return vs derivatives of that forward (reversible.py:326-412 computes them with
  jax.vjp of the same two sublayers)."""
  name = 'reversible_c128'
  c, d, g = RC.CASES[name], RC.inputs(name), np.load(FIXTURE)
  cfg, ln, aw = _cfg(c), (d['scale'], d['bias']), (d['w_q'], d['w_v'], d['w_o'])
  (y1, x2), buckets = O.reversible_half_forward(cfg, d['x1'], d['x2'], ln, aw, g[name + '/rot'])
  np.testing.assert_array_equal(buckets, g[name + '/buckets'])
  np.testing.assert_allclose(y1, g[name + '/y1'], rtol=1e-12, atol=1e-12)
  (x1, _), ((ct1, ct2), ((d_scale, d_bias), dw)) = O.reversible_half_reverse_and_grad(
      cfg, y1, d['x2'], d['ct_y1'], np.zeros_like(d['x2']), ln, aw, buckets)
  np.testing.assert_allclose(x1, d['x1'], rtol=1e-12, atol=1e-12)   # the block inverts
  np.testing.assert_array_equal(ct1, d['ct_y1'])
  for key, grad in zip(('x2', 'scale', 'bias', 'w_q', 'w_v', 'w_o'), (ct2, d_scale, d_bias) + tuple(dw)):
    np.testing.assert_allclose((grad * d['dir_' + key]).sum(), float(g[name + '/ddir_' + key]), rtol=1e-6, atol=1e-7,
                               err_msg=key)


def test_reference_batched_driver_equals_its_reference_loop():
  """Recorded by the generator: the reference's batched driver (Python-loop mode, EA:2261-2561) and its
  `use_reference_code` loop (EA:2127-2170) return the same output — the oracle restates the latter."""
  assert float(np.load(FIXTURE)['lsh_c128/batched_driver_max_abs_diff']) <= 1e-12


@pytest.mark.skipif(not ref_live.available(), reason='the reference checkout exists only in the build container')
def test_randomised_sweep_against_the_live_reference():
  """oracle/ref_live_sweep.py: 32 random option combinations (incl. the reference tests' own d_qk 7 / d_v 17 / chunk 5
  shape, look-back 0..2, look-ahead, masks, factor lists, max_length_for_buckets): buckets equal, outputs to 1e-11,
  VJP vs a central difference of the reference's forward."""
  repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([sys.executable, os.path.join(repo, 'oracle', 'ref_live_sweep.py'), '32', '7'], cwd=repo, timeout=900,
                     capture_output=True, text=True)
  assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
  assert r.stdout.count('\nok') + r.stdout.startswith('ok') == 32


@pytest.mark.skipif(not ref_live.available(), reason='the reference checkout exists only in the build container')
def test_predict_mode_sweep_against_the_live_reference():
  """oracle/ref_live_predict.py: `oracle/predict_oracle.py` vs the reference's own `mode='predict'` code (LSHSelfAttention
  EA:1999-2109, 2174-2244 and SelfAttention EA:1200-1268), call by call: prefixes shorter / equal / longer than the memory,
  then single tokens until the memory has rolled twice; outputs to 1e-11, memory / bucket memory / counters exact."""
  repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([sys.executable, os.path.join(repo, 'oracle', 'ref_live_predict.py'), '16', '3'], cwd=repo, timeout=900,
                     capture_output=True, text=True)
  assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
  assert r.stdout.count('\nok') + r.stdout.startswith('ok') == 16


PREDICT_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_predict.npz')


@pytest.mark.parametrize('name', ['lsh', 'self'])
def test_predict_oracle_matches_the_committed_reference_outputs(name):
  """tests/golden/reference_predict.npz holds what the reference's own `mode='predict'` code returned (a prefix, then single
  tokens until the memory rolled twice; tests/golden/make_predict_golden.py): the oracle must reproduce the outputs (stored
  as float32) and the final state exactly — this runs wherever the fixture travels, /root/reference or not."""
  from oracle import predict_oracle as P
  from oracle import self_attention_oracle as SA
  from tests.golden import make_predict_golden as G
  g, c = np.load(PREDICT_FIXTURE), G.CASES[name]
  w, xs = G.inputs(name)
  pcfg = P.PredictConfig(predict_mem_len=G.M, predict_drop_len=G.DROP)
  outs, t0 = [], 0
  if name == 'lsh':
    cfg = O.LSHConfig(**c['kw'])
    state = P.init_state(cfg, pcfg, G.B, G.D)
  else:
    cfg = SA.SelfAttentionConfig(**c['kw'])
    state = (0, np.zeros((G.B, G.M, G.D)))
  for n in G.calls(c):
    x = xs[:, t0:t0 + n]
    t0 += n
    if name == 'lsh':
      out, state = P.predict_forward(cfg, pcfg, x, w, state, lambda u, n_rows: g['lsh/rot'][u])
    else:
      out, state = P.self_attention_predict_forward(cfg, pcfg, x, w, state)
    outs.append(out)
  np.testing.assert_allclose(np.concatenate(outs, axis=1), g[name + '/out'], rtol=1e-5, atol=1e-6)
  assert state[0] == int(g[name + '/mem_end'])
  np.testing.assert_array_equal(state[1].astype(np.float32), g[name + '/mem'])
  if name == 'lsh':
    np.testing.assert_array_equal(state[2][0], g['lsh/buckets'])
    np.testing.assert_array_equal(state[2][1], g['lsh/buckets_idx'])


@pytest.mark.skipif(not ref_live.available(), reason='the reference checkout exists only in the build container')
def test_predict_fixture_regenerates_from_the_live_reference(tmp_path):
  repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  out = str(tmp_path / 'regen.npz')
  subprocess.run([sys.executable, os.path.join(repo, 'tests', 'golden', 'make_predict_golden.py'), '--out', out], cwd=repo,
                 timeout=900, check=True, capture_output=True)
  a, b = np.load(PREDICT_FIXTURE), np.load(out)
  assert sorted(a.files) == sorted(b.files)
  for k in a.files:
    np.testing.assert_array_equal(a[k], b[k], err_msg=k)
