"""Seeded inputs of `reference_live.npz` — shared by the script that runs the reference (make_reference_golden.py) and by
the tests that compare the oracle / the CUDA path with what the reference returned.  Only NumPy; no oracle, no reference."""
import numpy as np

# kind 'lsh': LSHSelfAttention (EA:1729);  'pure': PureLSHSelfAttention (EA:2564);  'hash': LSHSelfAttention.hash_vectors alone;
# 'wrapper': PureLSHSelfAttentionWrapper (EA:3493);  'reversible': ReversibleHalfResidual around the LSH layer
CASES = {
    # the tcgen05 kernels' shape (chunk 128, look-back 1, causal, 2 rounds); also run on the GPU
    'lsh_c128': dict(kind='lsh', B=1, H=2, L=512, D=64, C=128, nb=1, na=0, nh=2, n_buckets=8, causal=True, masked=False, seed=101),
    # chunk 64, 4 rounds, n_buckets=None -> 2*L/C = 16 buckets (EA:1893-1902, int branch); two examples; also run on the GPU
    'lsh_c64_auto': dict(kind='lsh', B=2, H=2, L=512, D=32, C=64, nb=1, na=0, nh=4, n_buckets=None, causal=True, masked=False, seed=102),
    # padding mask, bidirectional, look-ahead chunk, factored buckets [4, 2] (EA:108-117, 1907-1909, 1968-1972)
    'lsh_masked_factored': dict(kind='lsh', B=1, H=2, L=256, D=48, C=64, nb=1, na=1, nh=2, n_buckets=[4, 2], causal=False, masked=True, seed=103),
    # the weight-less core at chunk 128; also run on the GPU
    'pure_c128': dict(kind='pure', B=1, H=2, L=256, D=64, C=128, nb=1, na=0, nh=2, n_buckets=4, causal=True, masked=False, seed=104),
    # n_buckets=None with 2*L/C = 260 > 128 -> factor list [32, 8] (EA:1896-1902)
    'hash_auto_factors': dict(kind='hash', B=1, H=1, L=4160, D=64, C=32, nb=1, na=0, nh=2, n_buckets=None, causal=True, masked=False, seed=105),
    # PureLSHSelfAttentionWrapper (EA:3493): Dense q, k, v with bias -> (q + k)/2 -> core -> Dense; also run on the GPU
    'wrapper_c128': dict(kind='wrapper', B=1, H=2, L=256, D=128, C=128, nb=1, na=0, nh=2, n_buckets=4, causal=True, masked=False,
                         num_weights=3, bias=True, rotary=False, seed=106),
    # two weights, no bias, rotary position embedding of qk (the hourglass config's settings)
    'wrapper_rotary': dict(kind='wrapper', B=2, H=2, L=128, D=128, C=64, nb=1, na=0, nh=2, n_buckets=4, causal=True, masked=False,
                           num_weights=2, bias=False, rotary=True, seed=107),
    # ReversibleHalfResidual(LayerNorm, attention_layer=LSHSelfAttention) (reversible.py:244-321); also run on the GPU
    'reversible_c128': dict(kind='reversible', B=1, H=2, L=256, D=256, C=128, nb=1, na=0, nh=2, n_buckets=4, causal=True,
                            masked=False, seed=108),
}
D_HEAD = 64


def bf16_representable(a):
  """Round-toward-zero to values a bf16 holds exactly, as float64 (so fp64 reference, fp32 oracle hash and bf16 kernels all
  start from identical numbers)."""
  return (np.asarray(a, np.float32).view(np.uint32) & 0xffff0000).view(np.float32).astype(np.float64)


def inputs(name):
  """dict of float64 inputs for a case: x / (qk, v), weights, mask, the output cotangent and the directions along which the
  derivative of <out, dout> is taken."""
  c = CASES[name]
  rng = np.random.default_rng(c['seed'])
  B, H, L, D = c['B'], c['H'], c['L'], c['D']
  n = lambda *s: bf16_representable(rng.standard_normal(s))
  d = dict(rot_seed=c['seed'] + 1000)
  if c['kind'] == 'pure':
    d.update(qk=n(B * H, L, D_HEAD), v=n(B * H, L, D_HEAD), dout=n(B * H, L, D_HEAD),
             dir_qk=n(B * H, L, D_HEAD), dir_v=n(B * H, L, D_HEAD))
  elif c['kind'] == 'wrapper':
    s = 1.0 / np.sqrt(D)
    nw = c['num_weights']
    dense = lambda: (n(D, D) * s, n(D) * 0.25) if c['bias'] else n(D, D) * s
    d.update(x=n(B, L, D), qkv=tuple(dense() for _ in range(nw)), dense=dense(), dout=n(B, L, D), dir_x=n(B, L, D),
             dir_qkv=tuple(dense() for _ in range(nw)), dir_dense=dense())
  elif c['kind'] == 'reversible':
    s = 1.0 / np.sqrt(D)
    d.update(x1=n(B, L, D), x2=n(B, L, D), ct_y1=n(B, L, D), scale=1.0 + 0.25 * n(D), bias=0.25 * n(D),
             w_q=n(H, D, D_HEAD) * s, w_v=n(H, D, D_HEAD) * s, w_o=n(H, D_HEAD, D) * 0.125,
             dir_x2=n(B, L, D), dir_scale=n(D), dir_bias=n(D), dir_w_q=n(H, D, D_HEAD) * s, dir_w_v=n(H, D, D_HEAD) * s,
             dir_w_o=n(H, D_HEAD, D) * 0.125)
  else:
    s = 1.0 / np.sqrt(D)
    d.update(x=n(B, L, D), w_q=n(H, D, D_HEAD) * s, w_v=n(H, D, D_HEAD) * s, w_o=n(H, D_HEAD, D) * 0.125,
             dout=n(B, L, D), dir_x=n(B, L, D), dir_w_q=n(H, D, D_HEAD) * s, dir_w_v=n(H, D, D_HEAD) * s,
             dir_w_o=n(H, D_HEAD, D) * 0.125)
  d['mask'] = (rng.random((B, L)) > 0.25) if c['masked'] else None
  if c['masked']:
    d['dout'] = d['dout'] * d['mask'][:, :, None]                  # masked positions' outputs are not meaningful
  return d
