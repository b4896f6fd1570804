"""ORACLE PIN — TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Checks `oracle/predict_oracle.py`
against the reference's own fast-inference code (`mode='predict'`, `use_reference_code=True`: the Python loop of
EA:2127-2170 around `_use_predict_mem` EA:2174-2244 and `_incremental_forward_unbatched` EA:1999-2109; for `SelfAttention`
EA:1200-1268, 1300-1336) run under the reference's NumPy backend (see ref_live.py), call by call: an optional prefix
(shorter than / equal to / longer than the memory, or a few tokens appended at the start) followed by single-token steps,
enough of them to roll the memory several times.  After EVERY call: outputs equal to 1e-11 (float64), `mem_end`, the input
memory, the bucket memory and `buckets_idx` equal exactly.

    python oracle/ref_live_predict.py [n_cases] [seed]       # one line per case, non-zero exit on a mismatch

The NumPy backend lacks three jax.lax primitives this path calls; each is stated here in NumPy with the jax semantics
(`trax/fastmath/jax.py:177, 183, 185`): cond (old five-argument form), dynamic_slice_in_dim and
dynamic_update_slice_in_dim (start index clamped so that the slice fits).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lsh_oracle as O  # noqa: E402
from oracle import predict_oracle as P  # noqa: E402
from oracle import ref_live  # noqa: E402
from oracle import self_attention_oracle as SA  # noqa: E402


def load():
  R = ref_live.load()
  from trax.fastmath.numpy import NUMPY_BACKEND

  def cond(pred, true_operand, true_fun, false_operand, false_fun):
    return true_fun(true_operand) if bool(pred) else false_fun(false_operand)

  def dynamic_slice_in_dim(operand, start_index, slice_size, axis=0):
    start = int(np.clip(int(start_index), 0, operand.shape[axis] - slice_size))
    return np.take(operand, np.arange(start, start + slice_size), axis=axis)

  def dynamic_update_slice_in_dim(operand, update, start_index, axis):
    start = int(np.clip(int(start_index), 0, operand.shape[axis] - update.shape[axis]))
    out = np.array(operand, copy=True)
    idx = [slice(None)] * out.ndim
    idx[axis] = slice(start, start + update.shape[axis])
    out[tuple(idx)] = update
    return out
  NUMPY_BACKEND.update(cond=cond, dynamic_slice_in_dim=dynamic_slice_in_dim,
                       dynamic_update_slice_in_dim=dynamic_update_slice_in_dim)
  return R


def draw_case(rng, i):
  C = int(rng.choice([2, 4, 8]))
  M = C * int(rng.choice([2, 4, 6]))
  drop = int(rng.choice([d for d in (1, 2, C, 2 * C) if d < M]))
  prefix = str(rng.choice(['none', 'append', 'short', 'full', 'long']))
  if i % 4 == 2:
    kind = rng.choice(['int', 'list', 'none'])
    d_head = int(rng.choice([4, 6]))                                 # the batched driver sizes its output by d_qk (EA:3128, 3239)
    return dict(kind='pure', B=int(rng.choice([1, 2])), H=int(rng.choice([1, 2])), D=0, dq=d_head, dv=d_head, C=C, nb=int(rng.choice([0, 1])), nh=int(rng.choice([1, 2])),
                n_buckets={'int': int(rng.choice([2, 4])), 'list': [2, 2], 'none': None}[str(kind)], M=M, drop=drop, prefix=prefix)
  if i % 4 == 3:
    return dict(kind='self', B=int(rng.choice([1, 2])), H=int(rng.choice([1, 2])), D=int(rng.choice([8, 12])),
                dq=int(rng.choice([4, 6])), dv=int(rng.choice([4, 5])), C=C, nb=int(rng.choice([0, 1])), M=M, drop=drop,
                share_qk=bool(rng.random() < 0.5), prefix=prefix)
  kind = rng.choice(['int', 'list', 'none'])
  n_buckets = {'int': int(rng.choice([2, 4, 6])), 'list': [2, int(rng.choice([2, 4]))], 'none': None}[str(kind)]
  return dict(kind='lsh', B=int(rng.choice([1, 2])), H=int(rng.choice([1, 2, 3])), D=int(rng.choice([8, 12])),
              dq=int(rng.choice([4, 6])), dv=int(rng.choice([4, 5])), C=C, nb=int(rng.choice([0, 1, 2])),
              nh=int(rng.choice([1, 2, 3])), n_buckets=n_buckets, M=M, drop=drop, prefix=prefix)


def schedule(c, rng):
  """List of call lengths: the prefix (if any), then single tokens until the memory has rolled at least twice."""
  M, drop, C = c['M'], c['drop'], c['C']
  calls = []
  if c['prefix'] == 'append' and drop >= 2:
    calls.append(int(rng.integers(2, min(drop, M - 1) + 1)))         # a few tokens at the start (EA:2179 branch, q_start 0)
  elif c['prefix'] == 'short' and drop + 1 < M:
    calls.append(int(rng.integers(drop + 1, M)))
  elif c['prefix'] == 'full':
    calls.append(M)
  elif c['prefix'] == 'long':
    calls.append(M + int(rng.integers(1, 2 * C + 1)))
  if c['kind'] == 'self' and calls and calls[0] > C and calls[0] % C:
    calls[0] -= calls[0] % C                                        # EA:1250-1252: longer than a chunk ⇒ whole chunks
    if not (calls[0] > drop or calls[0] == M):
      calls = []
  n_single = (M - (min(calls[0], M) if calls else 0)) + 2 * drop + 3
  return calls + [1] * n_single


def run_case(R, c, rng):
  B, H, D, M, drop = c['B'], c['H'], c['D'], c['M'], c['drop']
  pcfg = P.PredictConfig(predict_mem_len=M, predict_drop_len=drop)
  calls = schedule(c, rng)
  seed = int(rng.integers(1 << 30))
  if c['kind'] == 'pure':
    return run_pure_case(R, c, rng, calls, pcfg, seed)
  xs = rng.standard_normal((B, sum(calls), D))
  sig = R.shapes.ShapeDtype((B, 1, D), np.float64)
  if c['kind'] == 'lsh':
    kw = dict(n_heads=H, d_qk=c['dq'], d_v=c['dv'], causal=True, chunk_len=c['C'], n_chunks_before=c['nb'],
              n_hashes=c['nh'], n_buckets=c['n_buckets'])
    layer = R.EA.LSHSelfAttention(use_reference_code=True, mode='predict', predict_mem_len=M, predict_drop_len=drop, **kw)
    w = (rng.standard_normal((H, D, c['dq'])) / np.sqrt(D), rng.standard_normal((H, D, c['dv'])) / np.sqrt(D),
         rng.standard_normal((H, c['dv'], D)) / np.sqrt(c['dv']))
    cfg = O.LSHConfig(**kw)
    state = P.init_state(cfg, pcfg, B, D)
  else:
    kw = dict(n_heads=H, d_qk=c['dq'], d_v=c['dv'], share_qk=c['share_qk'], causal=True, chunk_len=c['C'],
              n_chunks_before=c['nb'])
    layer = R.EA.SelfAttention(use_reference_code=True, mode='predict', predict_mem_len=M, predict_drop_len=drop, **kw)
    w = [rng.standard_normal((H, D, c['dq'])) / np.sqrt(D)]
    if not c['share_qk']:
      w.append(rng.standard_normal((H, D, c['dq'])) / np.sqrt(D))
    w += [rng.standard_normal((H, D, c['dv'])) / np.sqrt(D), rng.standard_normal((H, c['dv'], D)) / np.sqrt(c['dv'])]
    w = tuple(w)
    cfg = SA.SelfAttentionConfig(**kw)
    state = (0, np.zeros((B, M, D)))
  layer.init(sig)
  layer.weights = w
  worst, n_state_diff, t0 = 0.0, 0, 0
  for n in calls:
    x = xs[:, t0:t0 + n]
    t0 += n
    np.random.seed(seed)                                             # the NumPy backend draws from the global generator
    y = np.asarray(layer(x))
    ref_state = layer.state
    if c['kind'] == 'lsh':
      def rotations_fn(unit, n_rows):
        np.random.seed(seed)                                         # hash_rng is not advanced in predict mode: same draws every call
        shape = O.rotations_shape(cfg, n_rows)
        draws = [np.random.normal(size=shape).astype(np.float64).astype(np.float32) for _ in range(unit + 1)]
        return draws[unit]
      out, state = P.predict_forward(cfg, pcfg, x, w, state, rotations_fn)
      ref_b, ref_i = np.asarray(ref_state[2][0]), np.asarray(ref_state[2][1])
      n_state_diff += int((ref_b != state[2][0]).sum()) + int((ref_i != state[2][1]).sum())
    else:
      out, state = P.self_attention_predict_forward(cfg, pcfg, x, w, state)
    n_state_diff += int(int(ref_state[0]) != int(state[0])) + int((np.asarray(ref_state[1][0]) != state[1]).sum())
    err = np.abs(out - y)
    worst = max(worst, float(np.nanmax(err)) if np.isfinite(y).any() else 0.0)
    n_state_diff += int((np.isnan(out) != np.isnan(y)).sum())
  return worst, n_state_diff, len(calls), calls[0]


def run_pure_case(R, c, rng, calls, pcfg, seed):
  """`PureLSHSelfAttention(mode='predict')` has no `use_reference_code` loop (EA:2946-2948): it runs through its batched
  driver in Python-loop mode (EA:3052-3265, `use_python_loop=True, n_parallel_heads=1`)."""
  BH, M = c['B'] * c['H'], c['M']
  kw = dict(n_heads=c['H'], d_qk=c['dq'], d_v=c['dv'], causal=True, chunk_len=c['C'], n_chunks_before=c['nb'],
            n_hashes=c['nh'], n_buckets=c['n_buckets'])
  layer = R.EA.PureLSHSelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=c['drop'], use_python_loop=True,
                                    n_parallel_heads=1, **kw)
  layer.init((R.shapes.ShapeDtype((BH, 1, c['dq']), np.float64), R.shapes.ShapeDtype((BH, 1, c['dv']), np.float64)))
  cfg = O.LSHConfig(**kw)
  qks, vs = rng.standard_normal((BH, sum(calls), c['dq'])), rng.standard_normal((BH, sum(calls), c['dv']))
  state = (0, (np.zeros((BH, M, c['dq'])), np.zeros((BH, M, c['dv']))),
           (np.zeros((BH, c['nh'] * M), np.int32), np.zeros((BH,), np.int32)))
  worst, n_state_diff, t0 = 0.0, 0, 0
  for n in calls:
    qk, v = qks[:, t0:t0 + n], vs[:, t0:t0 + n]
    t0 += n
    np.random.seed(seed)
    y = np.asarray(layer((qk, v)))
    ref_state = layer.state

    def rotations_fn(unit, n_rows):
      np.random.seed(seed)
      shape = O.rotations_shape(cfg, n_rows)
      return [np.random.normal(size=shape).astype(np.float64).astype(np.float32) for _ in range(unit + 1)][unit]
    out, state = P.pure_predict_forward(cfg, pcfg, qk, v, state, rotations_fn)
    worst = max(worst, float(np.abs(out - y).max()))
    n_state_diff += int(int(ref_state[0]) != int(state[0]))
    n_state_diff += int((np.asarray(ref_state[1][0]) != state[1][0]).sum()) + int((np.asarray(ref_state[1][1]) != state[1][1]).sum())
    n_state_diff += int((np.asarray(ref_state[2][0]) != state[2][0]).sum()) + int((np.asarray(ref_state[2][1]) != state[2][1]).sum())
  return worst, n_state_diff, len(calls), calls[0]


def main(n_cases=24, seed=0):
  R = load()
  rng = np.random.default_rng(seed)
  bad = 0
  for i in range(n_cases):
    c = draw_case(rng, i)
    worst, n_diff, n_calls, first = run_case(R, c, rng)
    ok = worst < 1e-11 and n_diff == 0
    bad += not ok
    print('%s case %2d  max|out-ref| %.2e  state mismatches %d  calls %3d (first %2d)  %s'
          % ('ok  ' if ok else 'FAIL', i, worst, n_diff, n_calls, first, c))
  print('%d / %d cases agree with the reference' % (n_cases - bad, n_cases))
  return 1 if bad else 0


if __name__ == '__main__':
  sys.exit(main(*(int(a) for a in sys.argv[1:3])))
