"""ORACLE PIN — TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Randomised sweep of the oracle
against the reference's own code (see ref_live.py) over the layer's options: causal / bidirectional, padding mask,
look-back 0..2 and look-ahead 0..1 chunks, 1..3 hash rounds, bucket counts given as int, factor list or None, head sizes
and chunk lengths that are NOT the kernels' (the reference's tests use d_qk 7, d_v 17, chunk 5, seqlen 10, d_model 13 —
efficient_attention_test.py:138-151), `max_length_for_buckets`, and — every fourth case — attention + output dropout;
every fourth case instead checks `oracle/self_attention_oracle.py` against the reference's `SelfAttention` (EA:936).  For each case: bucket ids equal, float64 outputs equal
to 1e-11, and the analytic VJP equal to a central difference of the reference's forward along one random direction.

    python oracle/ref_live_sweep.py [n_cases] [seed]        # prints one line per case, exits non-zero on a mismatch
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lsh_oracle as O  # noqa: E402
from oracle import ref_live  # noqa: E402

EPS = 1e-6


def draw_case(rng, i):
  if i == 0:                                                         # the reference tests' own shape
    return dict(B=2, H=3, L=10, D=13, dq=7, dv=17, C=5, nb=1, na=0, nh=2, n_buckets=4, causal=True, masked=False, maxlen=None)
  C = int(rng.choice([4, 8, 16]))
  n_chunks = int(rng.choice([2, 3, 4, 8]))
  L = C * n_chunks
  kind = rng.choice(['int', 'list', 'none'])
  n_buckets = {'int': int(rng.choice([2, 4, 6, 8])), 'list': [int(rng.choice([2, 4])), int(rng.choice([2, 4, 6]))],
               'none': None}[kind]
  causal = bool(rng.random() < 0.6)
  return dict(B=int(rng.choice([1, 2])), H=int(rng.choice([1, 2, 3])), L=L, D=int(rng.choice([8, 13, 32])),
              dq=int(rng.choice([4, 7, 16])), dv=int(rng.choice([4, 9, 16])), C=C, nb=int(rng.choice([0, 1, 2])),
              na=0 if causal else int(rng.choice([0, 1])), nh=int(rng.choice([1, 2, 3])), n_buckets=n_buckets,
              causal=causal, masked=bool(rng.random() < 0.4), maxlen=None if rng.random() < 0.7 else 2 * L)


def run_case(R, c, rng):
  B, H, L, D = c['B'], c['H'], c['L'], c['D']
  x = rng.standard_normal((B, L, D))
  w = (rng.standard_normal((H, D, c['dq'])) / np.sqrt(D), rng.standard_normal((H, D, c['dv'])) / np.sqrt(D),
       rng.standard_normal((H, c['dv'], D)) / np.sqrt(c['dv']))
  mask = (rng.random((B, L)) > 0.3) if c['masked'] else None
  dout, direction = rng.standard_normal((B, L, D)), rng.standard_normal((B, L, D))
  if mask is not None:
    dout = dout * mask[:, :, None]
  kw = dict(n_heads=H, d_qk=c['dq'], d_v=c['dv'], causal=c['causal'], masked=c['masked'], chunk_len=c['C'],
            n_chunks_before=c['nb'], n_chunks_after=c['na'], n_hashes=c['nh'], n_buckets=c['n_buckets'],
            max_length_for_buckets=c['maxlen'])
  layer = R.EA.LSHSelfAttention(use_reference_code=True, **kw)
  sig = R.shapes.ShapeDtype((B, L, D), np.float64)
  layer.init((sig, R.shapes.ShapeDtype((B, L), np.bool_)) if c['masked'] else sig)
  layer.weights = w
  cfg = O.LSHConfig(**kw)
  seed = int(rng.integers(1 << 30))
  np.random.seed(seed)
  rot = np.stack([np.random.normal(size=O.rotations_shape(cfg, L)).astype(np.float64).astype(np.float32)
                  for _ in range(B * H)])
  np.random.seed(seed)
  y = np.asarray(layer((x, mask)) if c['masked'] else layer(x))
  ref_buckets = np.asarray(layer.state[0])
  out, buckets, _, _ = O.forward_and_or_backward(cfg, x, w, rotations=rot, mask=mask)
  assert ref_buckets.shape == buckets.shape, (ref_buckets.shape, buckets.shape)
  n_diff = int((ref_buckets != buckets).sum())
  err = float(np.abs(out - y).max())

  def fixed(xx):                                                     # the reference's forward with the buckets held
    res = np.zeros((B, L, D))
    for b in range(B):
      for h in range(H):
        args = (xx[b], mask[b]) if c['masked'] else (xx[b],)
        o, _ = layer.forward_unbatched(*args, weights=tuple(a[h] for a in w), state=(ref_buckets[b * H + h], None),
                                       rng=None, update_state=False)
        res[b] += o
    return res
  fd = float(((fixed(x + EPS * direction) - fixed(x - EPS * direction)) * dout).sum() / (2 * EPS))
  _, _, dx, _ = O.forward_and_or_backward(cfg, x, w, buckets=ref_buckets, mask=mask, output_grad=dout, update_state=False)
  an = float((dx * direction).sum())
  vjp_err = abs(an - fd) / max(abs(fd), 1e-3)
  return n_diff, err, vjp_err


def run_dropout_case(R, c, rng):
  """Attention dropout (EA:254-262: one (chunk_len, window) keep-mask broadcast over chunks, applied after the softmax,
  the log-sum-exp untouched) and output dropout (EA:271-280), one unit at a time with the buckets held.  The reference's
  NumPy backend draws the masks with numpy.random.binomial, so seeding reproduces them for the oracle."""
  H, L, D = c['H'], c['L'], c['D']
  a_rate, o_rate = 0.3, 0.25
  kw = dict(n_heads=H, d_qk=c['dq'], d_v=c['dv'], causal=c['causal'], masked=False, chunk_len=c['C'],
            n_chunks_before=c['nb'], n_chunks_after=c['na'], n_hashes=c['nh'], n_buckets=c['n_buckets'])
  layer = R.EA.LSHSelfAttention(use_reference_code=True, attention_dropout=a_rate, output_dropout=o_rate, mode='train', **kw)
  cfg = O.LSHConfig(**kw)
  x, direction, dout = (rng.standard_normal((L, D)) for _ in range(3))
  w = (rng.standard_normal((D, c['dq'])) / np.sqrt(D), rng.standard_normal((D, c['dv'])) / np.sqrt(D),
       rng.standard_normal((c['dv'], D)) / np.sqrt(c['dv']))
  factors = O.bucket_factors(c['n_buckets'], L, c['C'])
  nb_total = int(np.prod(factors))
  buckets = (rng.integers(0, nb_total, size=(c['nh'], L)) + np.arange(c['nh'])[:, None] * nb_total).astype(np.int32).reshape(-1)
  seed = int(rng.integers(1 << 30))
  key = np.zeros(2, np.uint32)                                       # the NumPy backend ignores it; it must not be None

  def live(xx):
    np.random.seed(seed)
    return np.asarray(layer.forward_unbatched(xx, weights=w, state=(buckets, None), rng=key, update_state=False)[0])
  y = live(x)
  window = c['C'] * (1 + c['nb'] + c['na'])
  np.random.seed(seed)
  attn_keep = np.random.binomial(1, 1 - a_rate, size=(c['C'], window)) / (1 - a_rate)
  out_keep = np.random.binomial(1, 1 - o_rate, size=(D,)) / (1 - o_rate)
  r = O.forward_unit(cfg, x, *w, buckets=buckets, attn_keep=attn_keep, out_keep=out_keep)
  err = float(np.abs(r.out - y).max())
  fd = float(((live(x + EPS * direction) - live(x - EPS * direction)) * dout).sum() / (2 * EPS))
  an = float((O.backward_unit(cfg, r, dout)[0] * direction).sum())
  return 0, err, abs(an - fd) / max(abs(fd), 1e-3)


def run_self_attention_case(R, rng, i):
  """`SelfAttention` (EA:936, the chunked local attention ReformerLM interleaves with the LSH layer) vs
  `oracle/self_attention_oracle.py`: both share_qk variants, chunked / unchunked, causal, masked."""
  from oracle import self_attention_oracle as S
  share_qk, causal, masked = bool(i % 2), bool(rng.random() < 0.6), bool(rng.random() < 0.4)
  chunked = rng.random() < 0.8
  C = int(rng.choice([4, 8, 16]))
  L = C * int(rng.choice([2, 3, 4]))
  B, H, D, dq, dv = int(rng.choice([1, 2])), int(rng.choice([1, 2, 3])), int(rng.choice([8, 13])), int(rng.choice([4, 7])), int(rng.choice([4, 9]))
  kw = dict(n_heads=H, d_qk=dq, d_v=dv, share_qk=share_qk, causal=causal, masked=masked, chunk_len=C if chunked else None,
            n_chunks_before=int(rng.choice([0, 1, 2])) if chunked else 0,
            n_chunks_after=int(rng.choice([0, 1])) if chunked and not causal else 0)
  layer = R.EA.SelfAttention(use_reference_code=True, **kw)
  x, dout, direction = (rng.standard_normal((B, L, D)) for _ in range(3))
  mask = (rng.random((B, L)) > 0.3) if masked else None
  if mask is not None:
    dout = dout * mask[:, :, None]
  n_w = 3 if share_qk else 4
  w = tuple(rng.standard_normal((H, D, dq)) / np.sqrt(D) for _ in range(n_w - 2)) + (
      rng.standard_normal((H, D, dv)) / np.sqrt(D), rng.standard_normal((H, dv, D)) / np.sqrt(dv))
  sig = R.shapes.ShapeDtype((B, L, D), np.float64)
  layer.init((sig, R.shapes.ShapeDtype((B, L), np.bool_)) if masked else sig)
  layer.weights = w

  def live(xx):
    return np.asarray(layer((xx, mask)) if masked else layer(xx))
  cfg = S.SelfAttentionConfig(**kw)
  out, dx, _ = S.forward_and_or_backward(cfg, x, w, mask=mask, output_grad=dout)
  err = float(np.abs(out - live(x)).max())
  fd = float(((live(x + EPS * direction) - live(x - EPS * direction)) * dout).sum() / (2 * EPS))
  an = float((dx * direction).sum())
  # weight gradients along one direction per weight
  werr = 0.0
  for j in range(n_w):
    dw = rng.standard_normal(w[j].shape)
    def shifted(e):
      layer.weights = tuple(a + (e * dw if jj == j else 0) for jj, a in enumerate(w))
      y = live(x)
      layer.weights = w
      return y
    fdw = float(((shifted(EPS) - shifted(-EPS)) * dout).sum() / (2 * EPS))
    gw = S.forward_and_or_backward(cfg, x, w, mask=mask, output_grad=dout)[2][j]
    werr = max(werr, abs(float((gw * dw).sum()) - fdw) / max(abs(fdw), 1e-3))
  return 0, err, max(abs(an - fd) / max(abs(fd), 1e-3), werr), dict(kw, B=B, L=L, D=D, kind='SelfAttention')


if __name__ == '__main__':
  n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
  rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
  R = ref_live.load()
  bad = 0
  for i in range(n_cases):
    c = draw_case(rng, i)
    if i % 4 == 2:
      n_diff, err, vjp_err, c = run_self_attention_case(R, rng, i // 4)
    else:
      dropout = i % 4 == 3 and not c['masked']
      n_diff, err, vjp_err = (run_dropout_case if dropout else run_case)(R, c, rng)
      c = dict(c, dropout=dropout)
    ok = n_diff == 0 and err <= 1e-11 and vjp_err <= 1e-5
    bad += not ok
    print('%s case %2d buckets_differ=%d out_err=%.1e vjp_rel_err=%.1e %s' % ('ok  ' if ok else 'FAIL', i, n_diff, err, vjp_err, c))
  sys.exit(1 if bad else 0)
