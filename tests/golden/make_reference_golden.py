"""Writes tests/golden/reference_live.npz BY RUNNING THE REFERENCE'S OWN CODE (google/trax at /root/reference) on CPU.

    python tests/golden/make_reference_golden.py [--out PATH]        # build container only (/root/reference must exist)

`oracle/ref_live.py` explains how the reference is made importable without JAX (its own NumPy backend; stubs only for
absent third-party packages; no reference source is copied).  For every case of `reference_cases.py` this script stores
what the reference returned:

  <case>/rot       hash rotations the reference drew (NumPy global generator, seeded; fastmath/numpy.py:37-40), per unit
  <case>/buckets   `LSHSelfAttention(use_reference_code=True).forward` state after the call (EA:2111-2170, 1926-1937)
  <case>/out       that call's output, float64
                   (`forward_unbatched(..., update_state=False)` with those buckets, EA:1939-1941, is checked here to
                   return the same numbers and is the function differentiated below)
  wrapper_* / reversible_*: the same for `PureLSHSelfAttentionWrapper` (Serial of Dense projections, core, Dense) and for
                   `ReversibleHalfResidual(LayerNorm, attention_layer=LSHSelfAttention)`, run through the batched drivers'
                   Python loop; every evaluation re-seeds NumPy so the same rotations (and, checked, buckets) recur.
  <case>/ddir_*    d/de <out(theta + e dir), dout> at e = 0 with the buckets held, by central differences of the reference's forward in
                   float64 — the reference's backward IS `jax.vjp` of this function (EA:2399-2421), so these numbers pin a
                   VJP without running JAX.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_live  # noqa: E402
from tests.golden import reference_cases as RC  # noqa: E402

EPS = 1e-6


def make_layer(R, c):
  kw = dict(n_heads=c['H'], d_qk=RC.D_HEAD, d_v=RC.D_HEAD, causal=c['causal'], masked=c['masked'], chunk_len=c['C'],
            n_chunks_before=c['nb'], n_chunks_after=c['na'], n_hashes=c['nh'], n_buckets=c['n_buckets'],
            use_reference_code=True)
  return (R.EA.PureLSHSelfAttention if c['kind'] == 'pure' else R.EA.LSHSelfAttention)(**kw)


def run_lsh(R, name, c, d, out):
  B, H, L = c['B'], c['H'], c['L']
  layer = make_layer(R, c)
  sig = R.shapes.ShapeDtype((B, L, c['D']), np.float64)
  layer.init((sig, R.shapes.ShapeDtype((B, L), np.bool_)) if c['masked'] else sig)
  weights = (d['w_q'], d['w_v'], d['w_o'])
  layer.weights = weights
  # the rotations the call below draws: same seed, same draw order (unit by unit, EA:2142-2158 -> EA:91-93)
  np.random.seed(d['rot_seed'])
  n_rot = sum(c['n_buckets']) // 2 if isinstance(c['n_buckets'], list) else \
      (c['n_buckets'] or 2 * max(1, L // c['C'])) // 2
  rot = np.stack([np.random.normal(size=(RC.D_HEAD, c['nh'], n_rot)).astype(np.float64).astype(np.float32)
                  for _ in range(B * H)])
  np.random.seed(d['rot_seed'])
  y = layer((d['x'], d['mask'])) if c['masked'] else layer(d['x'])
  buckets = np.asarray(layer.state[0])
  out[name + '/rot'], out[name + '/buckets'], out[name + '/out'] = rot, buckets.astype(np.int32), np.asarray(y, np.float64)

  def fixed(x, w_q, w_v, w_o):                                      # update_state=False: buckets from the state
    res = np.zeros((B, L, c['D']))
    for b in range(B):
      for h in range(H):
        args = (x[b], d['mask'][b]) if c['masked'] else (x[b],)
        o, _ = layer.forward_unbatched(*args, weights=(w_q[h], w_v[h], w_o[h]), state=(buckets[b * H + h], None),
                                       rng=None, update_state=False)
        res[b] += o
    return res
  base = (d['x'],) + weights
  np.testing.assert_allclose(fixed(*base), out[name + '/out'], rtol=1e-12, atol=1e-12)
  for i, key in enumerate(('x', 'w_q', 'w_v', 'w_o')):
    hi = [a + (EPS * d['dir_' + key] if j == i else 0) for j, a in enumerate(base)]
    lo = [a - (EPS * d['dir_' + key] if j == i else 0) for j, a in enumerate(base)]
    out[name + '/ddir_' + key] = np.float64(((fixed(*hi) - fixed(*lo)) * d['dout']).sum() / (2 * EPS))


def run_pure(R, name, c, d, out):
  BH, L = c['B'] * c['H'], c['L']
  layer = make_layer(R, c)
  np.random.seed(d['rot_seed'])
  rot = np.stack([np.random.normal(size=(RC.D_HEAD, c['nh'], c['n_buckets'] // 2)).astype(np.float64).astype(np.float32)
                  for _ in range(BH)])
  np.random.seed(d['rot_seed'])
  outs, buckets = [], []
  for u in range(BH):                                               # EA:2739-2826, update_state=True: hashes qk
    o, (b, _) = layer.forward_unbatched(d['qk'][u], d['v'][u], state=(np.zeros(c['nh'] * L, np.int32), None), rng=None,
                                        update_state=True)
    outs.append(o)
    buckets.append(b)
  buckets = np.stack(buckets).astype(np.int32)
  out[name + '/rot'], out[name + '/buckets'], out[name + '/out'] = rot, buckets, np.stack(outs).astype(np.float64)

  def fixed(qk, v):
    return np.stack([layer.forward_unbatched(qk[u], v[u], state=(buckets[u], None), rng=None, update_state=False)[0]
                     for u in range(BH)])
  np.testing.assert_allclose(fixed(d['qk'], d['v']), out[name + '/out'], rtol=1e-12, atol=1e-12)
  for key, hi, lo in (('qk', (d['qk'] + EPS * d['dir_qk'], d['v']), (d['qk'] - EPS * d['dir_qk'], d['v'])),
                      ('v', (d['qk'], d['v'] + EPS * d['dir_v']), (d['qk'], d['v'] - EPS * d['dir_v']))):
    out[name + '/ddir_' + key] = np.float64(((fixed(*hi) - fixed(*lo)) * d['dout']).sum() / (2 * EPS))


def run_hash(R, name, c, d, out):
  layer = make_layer(R, c)
  q = d['x'][0] @ d['w_q'][0]
  np.random.seed(d['rot_seed'])
  rot = np.random.normal(size=(RC.D_HEAD, c['nh'], (32 + 8) // 2)).astype(np.float64).astype(np.float32)
  np.random.seed(d['rot_seed'])
  out[name + '/rot'] = rot[None]
  out[name + '/buckets'] = np.asarray(layer.hash_vectors(q, None), np.int32)      # EA:1889-1916


def leaves(w):
  return list(w) if isinstance(w, tuple) else [w]


def set_leaves(R, layer, new_leaves):
  """Replaces the layer's weights, leaf by leaf in tree order, keeping the reference's own nesting."""
  old, _ = R.fastmath.tree_flatten(layer.weights), None
  assert [np.shape(a) for a in old] == [np.shape(a) for a in new_leaves], ([np.shape(a) for a in old],)
  tree, rest = R.fastmath.tree_unflatten(list(new_leaves), layer.weights)
  assert not rest
  layer.weights = tree


def run_wrapper(R, name, c, d, out):
  """PureLSHSelfAttentionWrapper as a Serial (EA:3512-3540), its core driven by the batched driver's Python loop."""
  B, H, L, D = c['B'], c['H'], c['L'], c['D']
  layer = R.EA.PureLSHSelfAttentionWrapper(
      n_heads=H, d_qk=RC.D_HEAD, d_v=RC.D_HEAD, causal=c['causal'], pure_lsh_implementation=R.EA.PureLSHSelfAttention,
      bias=c['bias'], num_weights=c['num_weights'], weights_format='model', rotary_position_emb=c['rotary'],
      chunk_len=c['C'], n_hashes=c['nh'], n_buckets=c['n_buckets'], use_python_loop=True, n_parallel_heads=1)
  layer.init(R.shapes.ShapeDtype((B, L, D), np.float64))
  np.random.seed(d['rot_seed'])
  out[name + '/rot'] = np.stack([np.random.normal(size=(RC.D_HEAD, c['nh'], c['n_buckets'] // 2)).astype(np.float64)
                                 .astype(np.float32) for _ in range(B * H)])

  def call(x, qkv, dense):                                          # re-seeded: every call draws the same rotations
    set_leaves(R, layer, [l for w in qkv for l in leaves(w)] + leaves(dense))
    np.random.seed(d['rot_seed'])
    y = layer(x)
    return np.asarray(y, np.float64), np.asarray(layer.state[1][0], np.int32)
  y, buckets = call(d['x'], d['qkv'], d['dense'])
  out[name + '/out'], out[name + '/buckets'] = y, buckets

  def ddir(hi, lo):
    (yh, bh), (yl, bl) = call(*hi), call(*lo)
    assert np.array_equal(bh, buckets) and np.array_equal(bl, buckets)   # the differentiated function holds the buckets
    return np.float64(((yh - yl) * d['dout']).sum() / (2 * EPS))
  shift = lambda w, dw, e: tuple(a + e * b for a, b in zip(w, dw)) if isinstance(w, tuple) else w + e * dw
  out[name + '/ddir_x'] = ddir((d['x'] + EPS * d['dir_x'], d['qkv'], d['dense']), (d['x'] - EPS * d['dir_x'], d['qkv'], d['dense']))
  for i in range(c['num_weights']):
    mv = lambda e: d['qkv'][:i] + (shift(d['qkv'][i], d['dir_qkv'][i], e),) + d['qkv'][i + 1:]
    out[name + '/ddir_qkv%d' % i] = ddir((d['x'], mv(EPS), d['dense']), (d['x'], mv(-EPS), d['dense']))
  out[name + '/ddir_dense'] = ddir((d['x'], d['qkv'], shift(d['dense'], d['dir_dense'], EPS)),
                                   (d['x'], d['qkv'], shift(d['dense'], d['dir_dense'], -EPS)))


def run_reversible(R, name, c, d, out):
  """ReversibleHalfResidual(LayerNorm(), attention_layer=LSHSelfAttention(...)).forward (reversible.py:296-321)."""
  B, H, L, D = c['B'], c['H'], c['L'], c['D']
  attn = R.EA.LSHSelfAttention(n_heads=H, d_qk=RC.D_HEAD, d_v=RC.D_HEAD, causal=c['causal'], chunk_len=c['C'],
                               n_hashes=c['nh'], n_buckets=c['n_buckets'], use_python_loop=True, n_parallel_heads=1)
  block = R.reversible.ReversibleHalfResidual(R.normalization.LayerNorm(), attention_layer=attn)
  sig = R.shapes.ShapeDtype((B, L, D), np.float64)
  block.init((sig, sig))
  np.random.seed(d['rot_seed'])
  out[name + '/rot'] = np.stack([np.random.normal(size=(RC.D_HEAD, c['nh'], c['n_buckets'] // 2)).astype(np.float64)
                                 .astype(np.float32) for _ in range(B * H)])

  def call(x2, scale, bias, w_q, w_v, w_o):
    set_leaves(R, block, [scale, bias, w_q, w_v, w_o])
    np.random.seed(d['rot_seed'])
    y1, ctx = block((d['x1'], x2))
    assert np.array_equal(ctx, x2)
    return np.asarray(y1, np.float64), np.asarray(block.state[1][0], np.int32)
  base = [d[k] for k in ('x2', 'scale', 'bias', 'w_q', 'w_v', 'w_o')]
  y1, buckets = call(*base)
  out[name + '/y1'], out[name + '/buckets'] = y1, buckets
  for i, key in enumerate(('x2', 'scale', 'bias', 'w_q', 'w_v', 'w_o')):
    hi = [a + (EPS * d['dir_' + key] if j == i else 0) for j, a in enumerate(base)]
    lo = [a - (EPS * d['dir_' + key] if j == i else 0) for j, a in enumerate(base)]
    (yh, bh), (yl, bl) = call(*hi), call(*lo)
    assert np.array_equal(bh, buckets) and np.array_equal(bl, buckets)
    out[name + '/ddir_' + key] = np.float64(((yh - yl) * d['ct_y1']).sum() / (2 * EPS))


def check_batched_driver(R, out):
  """The batched driver in Python-loop mode (EA:2261-2561) returns what the `use_reference_code` loop returns."""
  c, d = RC.CASES['lsh_c128'], RC.inputs('lsh_c128')
  res = []
  for kw in (dict(use_reference_code=True), dict(use_python_loop=True, n_parallel_heads=1)):
    layer = R.EA.LSHSelfAttention(n_heads=c['H'], d_qk=RC.D_HEAD, d_v=RC.D_HEAD, causal=True, chunk_len=c['C'],
                                  n_hashes=c['nh'], n_buckets=c['n_buckets'], **kw)
    layer.init(R.shapes.ShapeDtype((c['B'], c['L'], c['D']), np.float64))
    layer.weights = (d['w_q'], d['w_v'], d['w_o'])
    np.random.seed(d['rot_seed'])
    res.append((np.asarray(layer(d['x'])), np.asarray(layer.state[0])))
  assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][1], out['lsh_c128/buckets'])
  out['lsh_c128/batched_driver_max_abs_diff'] = np.float64(np.abs(res[0][0] - res[1][0]).max())


def generate():
  R = ref_live.load()
  out = {}
  for name, c in RC.CASES.items():
    {'lsh': run_lsh, 'pure': run_pure, 'hash': run_hash, 'wrapper': run_wrapper, 'reversible': run_reversible}[c['kind']](R, name, c, RC.inputs(name), out)
  check_batched_driver(R, out)
  out['stubbed_third_party'] = np.array(','.join(R.stubbed))
  return out


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--out', default=os.path.join(HERE, 'reference_live.npz'))
  args = ap.parse_args()
  arrays = generate()
  np.savez_compressed(args.out, **arrays)
  print('wrote %s: %d arrays, %d bytes' % (args.out, len(arrays), os.path.getsize(args.out)))
