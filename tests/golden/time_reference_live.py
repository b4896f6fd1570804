"""Build-container only: times the reference's own `LSHSelfAttention(use_reference_code=True).forward` (NumPy backend, see
oracle/ref_live.py) on ONE (example, head) unit of the bench workload (c2: seq 65536, d_model 1024, chunk 128, 4 hashes)
next to the oracle's `forward_unit` on the same inputs, and counts bucket ids that differ.  Context for `cpu_baseline`
(kind "port") in bench.py: the port runs at the speed of the reference's eager NumPy path; the reference's jitted XLA-CPU
path cannot be timed here (no JAX).

    python tests/golden/time_reference_live.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import lsh_oracle as O  # noqa: E402
from oracle import ref_live  # noqa: E402

if __name__ == '__main__':
  R = ref_live.load()
  L, D, C, nh = 65536, 1024, 128, 4
  rng = np.random.default_rng(0)
  x = rng.standard_normal((1, L, D)).astype(np.float32)
  w = tuple(a[:1].astype(np.float32) for a in O.init_weights(8, D, 64, 64, seed=1))
  layer = R.EA.LSHSelfAttention(n_heads=1, d_qk=64, d_v=64, causal=True, chunk_len=C, n_hashes=nh, n_buckets=None,
                                use_reference_code=True)
  layer.init(R.shapes.ShapeDtype((1, L, D), np.float32))
  layer.weights = w
  for _ in range(2):
    np.random.seed(3)
    t0 = time.time()
    y = layer(x)
    t_live = time.time() - t0
  cfg = O.LSHConfig(n_heads=1, d_qk=64, d_v=64, causal=True, masked=False, chunk_len=C, n_chunks_before=1,
                    n_chunks_after=0, n_hashes=nh, n_buckets=None)
  np.random.seed(3)
  rot = np.random.normal(size=(64, nh, 32)).astype(np.float32)
  for _ in range(2):
    t0 = time.time()
    r = O.forward_unit(cfg, x[0], w[0][0], w[1][0], w[2][0], rotations=rot, dtype=np.float32)
    t_port = time.time() - t0
  r64 = O.forward_unit(cfg, x[0].astype(np.float64), w[0][0], w[1][0], w[2][0], buckets=r.buckets)
  print('reference forward (NumPy backend), 1 unit of c2: %.2f s = %.0f tokens/s' % (t_live, L / t_live))
  print('oracle forward_unit, same unit:                  %.2f s = %.0f tokens/s' % (t_port, L / t_port))
  print('bucket ids that differ: %d of %d' % ((r.buckets != np.asarray(layer.state[0][0])).sum(), r.buckets.size))
  print('max |reference - oracle(fp64)| = %.2e (max |out| %.2f)' % (np.abs(r64.out - y[0]).max(), np.abs(y).max()))
