"""Regenerates tests/golden/lsh_small.npz FROM THE ORACLE (oracle/lsh_oracle.py).

The reference layer (trax LSHSelfAttention) cannot be run in this image (JAX absent), and
the checkout ships no golden vectors for this path (F4), so this fixture pins the oracle against
regressions; it is not reference output.  Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import lsh_oracle as O  # noqa: E402

rng = np.random.default_rng(20261017)
B, L, D, H = 1, 128, 32, 2
cfg = O.LSHConfig(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=64, n_hashes=2, n_buckets=[4, 2])
x = rng.standard_normal((B, L, D)).astype(np.float32).astype(np.float64)
w_q, w_v, w_o = (w.astype(np.float64) for w in O.init_weights(H, D, 64, 64, seed=7))
rot = rng.standard_normal((B * H,) + O.rotations_shape(cfg, L)).astype(np.float32)
dout = rng.standard_normal((B, L, D)).astype(np.float32).astype(np.float64)
out, buckets, _, _ = O.forward_and_or_backward(cfg, x, (w_q, w_v, w_o), rotations=rot)
_, _, dx, dw = O.forward_and_or_backward(cfg, x, (w_q, w_v, w_o), buckets=buckets, output_grad=dout, update_state=False)
s, u = O.sort_buckets(buckets[0], L)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lsh_small.npz'),
                    x=x, w_q=w_q, w_v=w_v, w_o=w_o, rot=rot, dout=dout, out=out, buckets=buckets, sticker0=s, undo0=u,
                    dx=dx, dw_q=dw[0], dw_v=dw[1], dw_o=dw[2])
print('wrote lsh_small.npz')
