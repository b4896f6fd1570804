"""Writes tests/golden/reference_predict.npz BY RUNNING THE REFERENCE'S OWN fast-inference code (google/trax at
/root/reference, `mode='predict'`, `use_reference_code=True`) on CPU under its NumPy backend (oracle/ref_live.py,
oracle/ref_live_predict.py) at shapes the CUDA kernels take (d_qk = d_v = 64).

    python tests/golden/make_predict_golden.py [--out PATH]          # build container only (/root/reference must exist)

Cases (inputs are regenerated from the seeds in CASES; the fixture stores what the reference returned):
  lsh     LSHSelfAttention, memory 128 / drop 32, chunk 64, look-back 1, one round of 4 buckets: a prefix of 96 tokens, then
          single tokens until the memory has rolled twice.  n_hashes * chunk_len * 2 = 128 = the memory length, so every
          earlier slot is attended at every step and the OUTPUT does not depend on the bucket ids — which lets the GPU test
          compare its output with the reference's directly although the device hashes bf16 projections (near-ties of the
          argmax may fall the other way); the bucket memory itself is compared up to a small mismatch fraction.
  self    SelfAttention(share_qk=False), same memory: a prefix of 64 tokens (= chunk_len: the unchunked branch EA:1262-1267),
          then single tokens.
Stored per case: <case>/out (B, T, D) float32 — all calls concatenated; <case>/mem_end; <case>/mem float32;
lsh only: <case>/rot (B*H, 64, 1, 2) float32 — the rotations the reference drew; <case>/buckets, <case>/buckets_idx.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_live_predict  # noqa: E402

B, H, D, C, M, DROP = 2, 2, 128, 64, 128, 32
CASES = {
    'lsh': dict(kind='lsh', prefix=96, seed=101, kw=dict(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=C, n_chunks_before=1,
                                                        n_hashes=1, n_buckets=4)),
    'self': dict(kind='self', prefix=64, seed=102, kw=dict(n_heads=H, d_qk=64, d_v=64, share_qk=False, causal=True, chunk_len=C,
                                                          n_chunks_before=1)),
}


def calls(c):
  return [c['prefix']] + [1] * (M - c['prefix'] + 2 * DROP + 3)


def bf16_round(a):
  """Values representable in bf16 (round to nearest even), as float64: the device and the reference see the same numbers."""
  u = np.asarray(a, np.float32).view(np.uint32).astype(np.uint64)
  u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
  return u.astype(np.uint32).view(np.float32).astype(np.float64)


def inputs(name):
  c = CASES[name]
  rng = np.random.default_rng(c['seed'])
  n_w = 3 if c['kind'] == 'lsh' else 4
  shapes = [(H, D, 64)] * (n_w - 1) + [(H, 64, D)]
  w = tuple(bf16_round(rng.standard_normal(s) / np.sqrt(s[1])) for s in shapes)
  xs = bf16_round(rng.standard_normal((B, sum(calls(c)), D)))
  return w, xs


def run_reference(R, name):
  c = CASES[name]
  w, xs = inputs(name)
  cls = R.EA.LSHSelfAttention if c['kind'] == 'lsh' else R.EA.SelfAttention
  layer = cls(use_reference_code=True, mode='predict', predict_mem_len=M, predict_drop_len=DROP, **c['kw'])
  layer.init(R.shapes.ShapeDtype((B, 1, D), np.float64))
  layer.weights = w
  outs, t0 = [], 0
  res = {}
  if c['kind'] == 'lsh':
    np.random.seed(c['seed'])
    res[name + '/rot'] = np.stack([np.random.normal(size=(64, 1, 2)).astype(np.float64).astype(np.float32)
                                   for _ in range(B * H)])
  for n in calls(c):
    np.random.seed(c['seed'])                                        # predict mode re-uses the state's key: same draws each call
    outs.append(np.asarray(layer(xs[:, t0:t0 + n])))
    t0 += n
  st = layer.state
  res[name + '/out'] = np.concatenate(outs, axis=1).astype(np.float32)
  res[name + '/mem_end'] = np.asarray(int(st[0]), np.int32)
  res[name + '/mem'] = np.asarray(st[1][0], np.float32)
  if c['kind'] == 'lsh':
    res[name + '/buckets'] = np.asarray(st[2][0], np.int32)
    res[name + '/buckets_idx'] = np.asarray(st[2][1], np.int32)
  return res


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--out', default=os.path.join(HERE, 'reference_predict.npz'))
  args = ap.parse_args()
  R = ref_live_predict.load()
  res = {}
  for name in CASES:
    res.update(run_reference(R, name))
  np.savez_compressed(args.out, **res)
  print('wrote', args.out, {k: v.shape for k, v in res.items()})


if __name__ == '__main__':
  main()
