"""CPU tests of the fast-inference (`mode='predict'`) host logic and of the key selection the decode kernel runs.

* `trax_b200/csrc/predict_select.cuh` is compiled for the host (g++) and run phase by phase over all thread ids
  (tests/micro/predict_select_host.cpp): its "attended" flags must equal the reference's priority sort (EA:2073-2084,
  restated in oracle/predict_oracle.py and pinned against the live reference) restricted to the slots that can receive
  probability (i <= q_start).
* `trax_b200/predict.py`'s memory bookkeeping (EA:2174-2244, 2036-2053) against the oracle's, on CPU tensors.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import predict_oracle as P

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def select_lib(tmp_path_factory):
  so = str(tmp_path_factory.mktemp('predict_select') / 'libpredict_select_host.so')
  subprocess.check_call(['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-o', so,
                         os.path.join(REPO, 'tests', 'micro', 'predict_select_host.cpp')])
  lib = ctypes.CDLL(so)
  lib.predict_select_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
  lib.predict_select_host.restype = None
  return lib


def _reference_selection(buckets, qb, M, nh, q_start, k_sel):
  """EA:2073-2084 with q_len == 1: the attended slots, as a boolean vector over the slots that are not causally masked."""
  unflattened = buckets.reshape(nh, M)
  is_valid_target = np.any(unflattened == qb[:, None], axis=0)
  ar = np.arange(M, dtype=np.int32)
  pri = np.where(ar > (q_start + 1), -(M + ar), ar) + M * is_valid_target.astype(np.int32)
  kv_indices = np.argsort(pri, kind='stable')[-k_sel:]
  sel = np.zeros(M, bool)
  sel[kv_indices] = True
  sel[q_start + 1:] = False                                          # q_pos < kv_pos: exp(-1e9 - lse) == 0 (EA:150-152)
  return sel


@pytest.mark.parametrize('seed', range(6))
def test_kernel_key_selection_equals_the_reference_priority_sort(select_lib, seed):
  rng = np.random.default_rng(seed)
  for _ in range(60):
    nh = int(rng.choice([1, 2, 3, 4]))
    M = int(rng.choice([8, 64, 200, 256, 1000, 2048]))
    n_buckets = int(rng.choice([2, 4, 16, 64]))
    q_start = int(rng.choice([0, 1, M // 2, M - 2, M - 1, int(rng.integers(0, M))]))
    k_sel = int(rng.choice([1, 2, 8, 64, 256, 1024, 4096]))
    buckets = (rng.integers(0, n_buckets, (nh, M)) + n_buckets * np.arange(nh)[:, None]).astype(np.int32)
    if rng.random() < 0.3:
      buckets[:, q_start + 1:] = 0                                   # untouched tail of the bucket memory
    qb = (rng.integers(0, n_buckets, nh) + n_buckets * np.arange(nh)).astype(np.int32)
    buckets[:, q_start] = qb                                         # EA:2069-2071 happened before the selection
    flags = np.full(M, 7, np.uint8)
    select_lib.predict_select_host(buckets.ctypes.data, qb.ctypes.data, M, nh, q_start, k_sel, flags.ctypes.data)
    want = _reference_selection(buckets.reshape(-1), qb, M, nh, q_start, k_sel)
    got = flags[:q_start + 1] == 1
    np.testing.assert_array_equal(got, want[:q_start + 1], err_msg=str((nh, M, n_buckets, q_start, k_sel)))
    assert got[q_start] or k_sel < 2          # the query's own slot is always attended (k_sel >= chunk_len >= 32 in the layer)


# ---- trax_b200/predict.py control flow and memory bookkeeping, with the CUDA calls swapped for the oracle --------------------
from oracle import lsh_oracle as O                # noqa: E402
from oracle import self_attention_oracle as SA    # noqa: E402


def _schedule(kind, M, drop, C, rng):
  first = {'none': [], 'append': [int(rng.integers(2, drop + 1))] if drop >= 2 else [],
           'short': [int(rng.integers(drop + 1, M))], 'full': [M], 'long': [M + int(rng.integers(1, 2 * C))]}[kind]
  return first + [1] * ((M - min(first[0], M) if first else M) + 2 * drop + 3)


@pytest.mark.parametrize('prefix', ['none', 'append', 'short', 'full', 'long'])
def test_lsh_predict_host_logic_matches_the_oracle(prefix):
  """`predict._run` (roll of the memory and of the bucket memory, prefix handling, counters) with `lsh_predict_step` and the
  training-path forward replaced by the oracle's restatements of the same two pieces: outputs and every state leaf must
  follow `oracle.predict_oracle.predict_forward` (pinned against the live reference) call by call."""
  import trax_b200
  from trax_b200 import predict
  rng = np.random.default_rng(hash(prefix) % 1000)
  B, H, D, C, nh, M, drop, dq, dv = 2, 2, 12, 4, 2, 16, 4, 6, 5
  kw = dict(n_heads=H, d_qk=dq, d_v=dv, causal=True, chunk_len=C, n_chunks_before=1, n_hashes=nh, n_buckets=4)
  cfg, pcfg = O.LSHConfig(**kw), P.PredictConfig(predict_mem_len=M, predict_drop_len=drop)
  layer = trax_b200.LSHSelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=drop, **kw)
  w = (rng.standard_normal((H, D, dq)) / np.sqrt(D), rng.standard_normal((H, D, dv)) / np.sqrt(D),
       rng.standard_normal((H, dv, D)) / np.sqrt(dv))
  rot = rng.standard_normal((B * H,) + O.rotations_shape(cfg, 2)).astype(np.float32)
  layer._rotations_override = torch.from_numpy(rot)

  def fake_step(layer_, mem, weights, q_start, buckets, rotations, causal):
    assert causal and rotations is not None
    out = np.zeros((B, 1, D))
    for u in range(B * H):
      b, h = u // H, u % H
      o, nb, _ = P.incremental_forward_unit(cfg, pcfg, mem[b].numpy(), q_start, 1, w[0][h], w[1][h], w[2][h],
                                            buckets[u].numpy(), q_start, lambda n, _u=u: rotations[_u].numpy())
      out[b] += o
      buckets[u] = torch.from_numpy(nb)                              # in place, like the kernel
    return torch.from_numpy(out)

  def fake_train(x, weights, state):
    out, new_b, _, _ = O.forward_and_or_backward(cfg, x.numpy(), w, rotations=rot)
    return torch.from_numpy(out), (torch.from_numpy(new_b), state[1]), None, None

  calls = _schedule(prefix, M, drop, C, rng)
  xs = rng.standard_normal((B, sum(calls), D))
  ostate = P.init_state(cfg, pcfg, B, D)
  mem_end, mem = 0, torch.zeros((B, M, D), dtype=torch.float64)
  inner = (torch.zeros((B * H, nh * M), dtype=torch.int32), torch.zeros((B * H,), dtype=torch.int32), None)
  t0 = 0
  for n in calls:
    x = xs[:, t0:t0 + n]
    t0 += n
    want, ostate = P.predict_forward(cfg, pcfg, x, w, ostate, lambda u, n_rows: rot[u])
    mem_before, buckets_before = mem.clone(), inner[0].clone()
    out, (mem_end, mem, inner) = predict._run(layer, torch.from_numpy(x), w, mem_end, mem, inner, None, step=fake_step,
                                              train=fake_train)
    np.testing.assert_allclose(out.numpy(), want, rtol=1e-12, atol=1e-12)
    assert mem_end == ostate[0]
    np.testing.assert_array_equal(mem.numpy(), ostate[1])
    np.testing.assert_array_equal(inner[0].numpy(), ostate[2][0])
    np.testing.assert_array_equal(inner[1].numpy(), ostate[2][1])
    # states are values: the tensors handed in are untouched (EA returns new arrays)
    assert mem_before.data_ptr() != mem.data_ptr() and buckets_before.data_ptr() != inner[0].data_ptr()


@pytest.mark.parametrize('share_qk,prefix', [(False, 'none'), (True, 'append'), (False, 'short'), (True, 'full'), (False, 'long')])
def test_self_attention_predict_host_logic_matches_the_oracle(share_qk, prefix):
  import trax_b200
  from trax_b200 import predict
  rng = np.random.default_rng(11)
  B, H, D, C, M, drop, dq, dv = 2, 2, 12, 4, 16, 4, 6, 5
  kw = dict(n_heads=H, d_qk=dq, d_v=dv, share_qk=share_qk, causal=True, chunk_len=C, n_chunks_before=1)
  cfg, pcfg = SA.SelfAttentionConfig(**kw), P.PredictConfig(predict_mem_len=M, predict_drop_len=drop)
  layer = trax_b200.SelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=drop, **kw)
  w = [rng.standard_normal((H, D, dq)) / np.sqrt(D)] + ([] if share_qk else [rng.standard_normal((H, D, dq)) / np.sqrt(D)])
  w = tuple(w + [rng.standard_normal((H, D, dv)) / np.sqrt(D), rng.standard_normal((H, dv, D)) / np.sqrt(dv)])

  def fake_step(layer_, mem, weights, q_start, buckets, rotations, causal):
    assert buckets is None and rotations is None and causal
    out = np.zeros((B, 1, D))
    for u in range(B * H):
      out[u // H] += P.self_attention_incremental_unit(cfg, mem[u // H].numpy(), q_start, 1, tuple(a[u % H] for a in w))
    return torch.from_numpy(out)

  def fake_train(x, weights, state):
    out = SA.forward_and_or_backward(cfg, x.numpy(), w)[0]
    return torch.from_numpy(out), None, None, None

  calls = _schedule(prefix, M, drop, C, rng)
  if len(calls) and calls[0] > C and calls[0] % C:
    calls[0] -= calls[0] % C                                         # EA:1250-1252
    if not (calls[0] > drop or calls[0] == M):
      calls = calls[1:]
  xs = rng.standard_normal((B, sum(calls), D))
  ostate = (0, np.zeros((B, M, D)))
  mem_end, mem = 0, torch.zeros((B, M, D), dtype=torch.float64)
  t0 = 0
  for n in calls:
    x = xs[:, t0:t0 + n]
    t0 += n
    want, ostate = P.self_attention_predict_forward(cfg, pcfg, x, w, ostate)
    out, (mem_end, mem, inner) = predict._run(layer, torch.from_numpy(x), w, mem_end, mem, (), None, step=fake_step,
                                              train=fake_train)
    np.testing.assert_allclose(out.numpy(), want, rtol=1e-12, atol=1e-12)
    assert mem_end == ostate[0] and inner == ()
    np.testing.assert_array_equal(mem.numpy(), ostate[1])


def test_predict_mode_rejections():
  import trax_b200
  from trax_b200 import predict
  with pytest.raises(NotImplementedError):
    trax_b200.LSHSelfAttention(mode='predict', masked=True)
  with pytest.raises(ValueError):
    trax_b200.LSHSelfAttention(mode='predict', predict_mem_len=64, predict_drop_len=64)
  layer = trax_b200.LSHSelfAttention(mode='predict', predict_mem_len=16, predict_drop_len=4, chunk_len=4, n_buckets=4)
  mem = torch.zeros((1, 16, 8))
  with pytest.raises(ValueError, match='start of a sequence'):       # EA:2233-2243: the reference returns NaNs here
    predict.use_predict_mem(torch.zeros((1, 8, 8)), 3, mem, 16, 4)
  with pytest.raises(ValueError, match='start of a sequence'):       # EA:2005-2006
    predict._run(layer, torch.zeros((1, 2, 8)), (), 5, mem, (torch.zeros((2, 16), dtype=torch.int32),
                                                            torch.zeros((2,), dtype=torch.int32), None), None)
  with pytest.raises(NotImplementedError):                           # EA:2002-2003
    predict.forward_and_or_backward(layer, torch.zeros((1, 1, 8)), (), (), None, output_grad=torch.zeros((1, 1, 8)))


def _selector_weights(dq, dv):
  """x = [qk | v]: w_q picks qk, w_v picks v, w_o is the identity — the oracle's unit then is the weight-less core."""
  return (np.concatenate([np.eye(dq), np.zeros((dv, dq))]), np.concatenate([np.zeros((dq, dv)), np.eye(dv)]), np.eye(dv))


@pytest.mark.parametrize('prefix', ['none', 'append', 'short', 'full', 'long'])
def test_pure_core_predict_host_logic_matches_the_oracle(prefix):
  """`predict._run_pure` (PureLSHSelfAttention, EA:2823-3033) with `lsh_predict_attend` and the training-path core replaced
  by the oracle: outputs and every state leaf vs `oracle.predict_oracle.pure_predict_forward` (pinned against the live
  reference's batched driver)."""
  import trax_b200
  from trax_b200 import predict
  rng = np.random.default_rng(3 + len(prefix))
  B, H, C, nh, M, drop, d = 2, 2, 4, 2, 16, 4, 6
  BH = B * H
  kw = dict(n_heads=H, d_qk=d, d_v=d, causal=True, chunk_len=C, n_chunks_before=1, n_hashes=nh, n_buckets=4)
  cfg, pcfg = O.LSHConfig(**kw), P.PredictConfig(predict_mem_len=M, predict_drop_len=drop)
  cfg1 = cfg
  layer = trax_b200.PureLSHSelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=drop, **kw)
  rot = rng.standard_normal((BH,) + O.rotations_shape(cfg, 2)).astype(np.float32)
  layer._rotations_override = torch.from_numpy(rot)
  w_q, w_v, w_o = _selector_weights(d, d)

  def fake_step(layer_, qk_mem, v_mem, q_start, buckets, rotations):
    out = np.zeros((BH, 1, d))
    x = np.concatenate([qk_mem.numpy(), v_mem.numpy()], axis=-1)
    for u in range(BH):
      out[u], nb, _ = P.incremental_forward_unit(cfg, pcfg, x[u], q_start, 1, w_q, w_v, w_o, buckets[u].numpy(), q_start,
                                                 lambda n, _u=u: rotations[_u].numpy())
      buckets[u] = torch.from_numpy(nb)
    return torch.from_numpy(out)

  def fake_train(inputs, state):
    x = np.concatenate([inputs[0].numpy(), inputs[1].numpy()], axis=-1)
    res = [O.forward_unit(cfg1, x[u], w_q, w_v, w_o, rotations=rot[u]) for u in range(BH)]
    return (torch.from_numpy(np.stack([r.out for r in res])),
            (torch.from_numpy(np.stack([r.buckets for r in res])), state[1]), None)

  calls = _schedule(prefix, M, drop, C, rng)
  qks, vs = rng.standard_normal((BH, sum(calls), d)), rng.standard_normal((BH, sum(calls), d))
  ostate = (0, (np.zeros((BH, M, d)), np.zeros((BH, M, d))), (np.zeros((BH, nh * M), np.int32), np.zeros((BH,), np.int32)))
  mem_end, mems = 0, (torch.zeros((BH, M, d), dtype=torch.float64), torch.zeros((BH, M, d), dtype=torch.float64))
  inner = (torch.zeros((BH, nh * M), dtype=torch.int32), torch.zeros((BH,), dtype=torch.int32), None)
  t0 = 0
  for n in calls:
    qk, v = qks[:, t0:t0 + n], vs[:, t0:t0 + n]
    t0 += n
    want, ostate = P.pure_predict_forward(cfg, pcfg, qk, v, ostate, lambda u, n_rows: rot[u])
    out, (mem_end, mems, inner) = predict._run_pure(layer, torch.from_numpy(qk), torch.from_numpy(v), mem_end, mems, inner, None,
                                                    step=fake_step, train=fake_train)
    np.testing.assert_allclose(out.numpy(), want, rtol=1e-12, atol=1e-12)
    assert mem_end == ostate[0]
    np.testing.assert_array_equal(mems[0].numpy(), ostate[1][0])
    np.testing.assert_array_equal(mems[1].numpy(), ostate[1][1])
    np.testing.assert_array_equal(inner[0].numpy(), ostate[2][0])
    np.testing.assert_array_equal(inner[1].numpy(), ostate[2][1])


def test_kernel_key_selection_edge_cases(select_lib):
  """Largest memory the kernel takes (32 768 slots), many hash rounds, every slot valid / no slot valid, K below and above
  the number of candidates."""
  rng = np.random.default_rng(99)
  cases = []
  for M, nh in ((32768, 8), (4096, 64), (33, 1)):
    for fill in ('random', 'all_valid', 'none_valid'):
      for q_start in (0, M // 3, M - 1):
        for k_sel in (2, 300, 2 * M):
          cases.append((M, nh, fill, q_start, k_sel))
  for M, nh, fill, q_start, k_sel in cases:
    qb = (np.arange(nh) * 4 + rng.integers(0, 4, nh)).astype(np.int32)
    if fill == 'random':
      buckets = (rng.integers(0, 4, (nh, M)) + 4 * np.arange(nh)[:, None]).astype(np.int32)
    elif fill == 'all_valid':
      buckets = np.repeat(qb[:, None], M, axis=1).astype(np.int32)
    else:
      buckets = np.repeat(((qb + 1) % 4 + 4 * np.arange(nh))[:, None], M, axis=1).astype(np.int32)
    buckets[:, q_start] = qb
    flags = np.zeros(M, np.uint8)
    select_lib.predict_select_host(buckets.ctypes.data, qb.ctypes.data, M, nh, q_start, k_sel, flags.ctypes.data)
    want = _reference_selection(buckets.reshape(-1), qb, M, nh, q_start, k_sel)
    np.testing.assert_array_equal(flags[:q_start + 1] == 1, want[:q_start + 1], err_msg=str((M, nh, fill, q_start, k_sel)))


def test_reversible_block_does_not_fuse_the_residual_into_a_predict_mode_layer():
  """`ReversibleHalfResidual` folds `x1 + Attn(...)` into the training path's output-projection epilogue; the decode step has
  no such epilogue, so a predict-mode attention layer must take the separate-add path (the fused call would drop the
  residual silently)."""
  import trax_b200
  train = trax_b200.ReversibleHalfResidual(trax_b200.LSHSelfAttention(causal=True))
  pred = trax_b200.ReversibleHalfResidual(trax_b200.LSHSelfAttention(causal=True, mode='predict'))

  class _Dev(torch.Tensor):                                          # a CPU tensor that claims to live on the device
    is_cuda = True
  a = torch.zeros((1, 4, 8)).as_subclass(_Dev)
  assert train._fused(a, a) and not pred._fused(a, a)


def test_self_attention_without_chunks_decodes_in_predict_mode_only():
  """`SelfAttention(chunk_len=None)` — the reference's default — is dense attention; the training kernels are chunked, but a
  decode step attends over the whole memory whatever the chunk length (EA:1262-1267), and with `chunk_len=None` a prefix is
  never chunked either (EA:1250): token-by-token steps.  Control flow vs the oracle, CUDA step swapped for the oracle's."""
  import trax_b200
  from trax_b200 import predict
  with pytest.raises(NotImplementedError):
    trax_b200.SelfAttention(causal=True)
  rng = np.random.default_rng(17)
  B, H, D, M, drop, dq, dv = 1, 2, 12, 16, 4, 6, 5
  kw = dict(n_heads=H, d_qk=dq, d_v=dv, share_qk=False, causal=True, chunk_len=None)
  cfg, pcfg = SA.SelfAttentionConfig(**kw), P.PredictConfig(predict_mem_len=M, predict_drop_len=drop)
  layer = trax_b200.SelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=drop, **kw)
  w = tuple(rng.standard_normal(s) / 3 for s in ((H, D, dq), (H, D, dq), (H, D, dv), (H, dv, D)))

  def fake_step(layer_, mem, weights, q_start, buckets, rotations, causal):
    out = np.zeros((B, 1, D))
    for u in range(B * H):
      out[u // H] += P.self_attention_incremental_unit(cfg, mem[u // H].numpy(), q_start, 1, tuple(a[u % H] for a in w))
    return torch.from_numpy(out)

  def no_train(*a):
    raise AssertionError('a dense prefix must not reach the chunked training path')
  calls = [9] + [1] * 14                                             # a prefix longer than the stand-in chunk length would be
  xs = rng.standard_normal((B, sum(calls), D))
  ostate, mem_end, mem, t0 = (0, np.zeros((B, M, D))), 0, torch.zeros((B, M, D), dtype=torch.float64), 0
  layer._chunk_len = 4                                               # (the stand-in only sizes the C ABI's dims)
  for n in calls:
    x = xs[:, t0:t0 + n]
    t0 += n
    want, ostate = P.self_attention_predict_forward(cfg, pcfg, x, w, ostate)
    out, (mem_end, mem, _) = predict._run(layer, torch.from_numpy(x), w, mem_end, mem, (), None, step=fake_step, train=no_train)
    np.testing.assert_allclose(out.numpy(), want, rtol=1e-12, atol=1e-12)
    assert mem_end == ostate[0]
