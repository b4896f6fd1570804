"""Fast inference (`mode='predict'`) on the GPU vs `oracle/predict_oracle.py`, which is pinned call by call against the
reference's own predict-mode code (tests/test_reference_pin.py::test_predict_mode_sweep_against_the_live_reference).

  * one `lsh_predict_step` (EA:2032-2109) at chosen memory fill levels: the new token's bucket ids bit-exact against the
    oracle's hash of the device's own q row, the bucket memory after the call, and the output;
  * the same step without hashing = `SelfAttention` (EA:1200-1268), both `share_qk` settings;
  * the reference test's own invariant (`efficient_attention_test.py:158-206`): decoding token by token in predict mode
    reproduces the full forward pass of the layer;
  * the layer API end to end: a prefix, then enough single tokens to roll the memory twice, states compared leaf by leaf.

(The file name sorts after every other GPU test on purpose: the kernels under test here are the newest.)
"""
import numpy as np
import pytest
import torch

from oracle import lsh_oracle as O
from oracle import predict_oracle as P
from oracle import self_attention_oracle as SA
from tests import util

pytestmark = pytest.mark.gpu


def _weights(rng, H, D, separate_k=False):
  w = [rng.standard_normal((H, D, 64)) / np.sqrt(D)]
  if separate_k:
    w.append(rng.standard_normal((H, D, 64)) / np.sqrt(D))
  w += [rng.standard_normal((H, D, 64)) / np.sqrt(D), rng.standard_normal((H, 64, D)) / 8.0]
  return tuple(a.astype(np.float32) for a in w)


@pytest.mark.parametrize('M,C,nb,nh,n_buckets,q_start,dtype', [
    (512, 64, 0, 2, 8, 0, torch.float32),            # first token: attends to itself only (EA:153-155: -1e5, still the max)
    (512, 64, 0, 2, 8, 5, torch.float32),
    (512, 64, 0, 2, 8, 300, torch.float32),          # 128 of 301 slots attended: same-bucket ones, then the most recent
    (512, 64, 0, 2, 8, 511, torch.bfloat16),         # last slot of the memory
    (256, 128, 1, 1, [4, 2], 200, torch.bfloat16),   # factored bucket count; every earlier slot attended (K = 256)
    (1024, 32, 1, 4, 16, 777, torch.float32),
])
def test_predict_step_lsh_matches_oracle(M, C, nb, nh, n_buckets, q_start, dtype):
  import trax_b200
  from trax_b200 import _lib, ops, predict
  rng = np.random.default_rng(7 + q_start)
  B, H, D = 2, 2, 128
  kw = dict(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=C, n_chunks_before=nb, n_hashes=nh, n_buckets=n_buckets)
  cfg, pcfg = O.LSHConfig(**kw), P.PredictConfig(predict_mem_len=M, predict_drop_len=C)
  layer = trax_b200.LSHSelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=C, **kw)
  w = _weights(rng, H, D)
  mem = util.bf16_round(rng.standard_normal((B, M, D)))
  mem[:, q_start + 1:] = 0                                           # slots that have not been written yet
  factors = O.bucket_factors(n_buckets, 2, C)
  nbk = int(np.prod(factors))
  buckets = util.random_valid_buckets(rng, B * H, nh, M, nbk).reshape(B * H, nh, M)
  buckets[:, :, q_start:] = 0
  buckets = buckets.reshape(B * H, nh * M)
  rot = rng.standard_normal((B * H,) + O.rotations_shape(cfg, 2)).astype(np.float32)

  mem_d = torch.from_numpy(mem).cuda().to(dtype)
  w_d = tuple(torch.from_numpy(a).cuda() for a in w)
  buckets_d = torch.from_numpy(buckets.copy()).cuda()
  out_d = predict._step(layer, mem_d, w_d, q_start, buckets_d, torch.from_numpy(rot).cuda(), True)
  torch.cuda.synchronize()
  got_b = buckets_d.cpu().numpy()

  # (1) the new token's ids: the oracle's hash (EA:2064-2066: two copies of the row) of the DEVICE's bf16 q row
  dims = _lib.make_dims(B, H, M, D, 64, 64, C, nb, 0, nh, factors, True, False, _lib.LSH_DTYPE_BF16)
  wqv, _ = ops.pack_weights(dims, *w_d)
  qv = ops.project_qv(dims, mem_d.to(torch.bfloat16).contiguous(), wqv).float().cpu().numpy()       # (B, M, H, 128)
  new_ids = got_b.reshape(B * H, nh, M)[:, :, q_start]
  for u in range(B * H):
    q_row = qv[u // H, q_start, u % H, :64]
    want_ids = O.hash_vectors(cfg, np.stack([q_row, q_row]), rot[u]).reshape(nh, 2)[:, 0]
    np.testing.assert_array_equal(new_ids[u], want_ids, err_msg='unit %d' % u)
  # (2) bucket memory: only column q_start changed (EA:2069-2071)
  want_b = buckets.reshape(B * H, nh, M).copy()
  want_b[:, :, q_start] = new_ids
  np.testing.assert_array_equal(got_b, want_b.reshape(B * H, nh * M))
  # (3) output vs the oracle's step on the same ids
  want = np.zeros((B, 1, D))
  for u in range(B * H):
    b, h = u // H, u % H
    o, nb_u, idx_u = P.incremental_forward_unit(cfg, pcfg, mem[b], q_start, 1, w[0][h], w[1][h], w[2][h], buckets[u], q_start,
                                                None, new_ids=new_ids[u])
    want[b] += o
    np.testing.assert_array_equal(nb_u, got_b[u])
    assert idx_u == q_start + 1
  assert out_d.dtype == dtype and tuple(out_d.shape) == (B, 1, D)
  util.assert_close_layer(out_d.float().cpu().numpy(), want, 'out')


@pytest.mark.parametrize('share_qk,M,q_start,dtype', [
    (False, 256, 0, torch.float32), (False, 256, 100, torch.float32), (False, 256, 255, torch.bfloat16),
    (True, 256, 0, torch.float32), (True, 512, 301, torch.bfloat16),
])
def test_predict_step_self_attention_matches_oracle(share_qk, M, q_start, dtype):
  import trax_b200
  from trax_b200 import predict
  rng = np.random.default_rng(19 + q_start)
  B, H, D, C = 2, 2, 128, 64
  kw = dict(n_heads=H, d_qk=64, d_v=64, share_qk=share_qk, causal=True, chunk_len=C, n_chunks_before=1)
  cfg = SA.SelfAttentionConfig(**kw)
  layer = trax_b200.SelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=C, **kw)
  w = _weights(rng, H, D, separate_k=not share_qk)
  mem = util.bf16_round(rng.standard_normal((B, M, D)))
  mem[:, q_start + 1:] = 0
  out_d = predict._step(layer, torch.from_numpy(mem).cuda().to(dtype), tuple(torch.from_numpy(a).cuda() for a in w), q_start,
                        None, None, True)
  want = np.zeros((B, 1, D))
  for u in range(B * H):
    want[u // H] += P.self_attention_incremental_unit(cfg, mem[u // H], q_start, 1, tuple(a[u % H] for a in w))
  util.assert_close_layer(out_d.float().cpu().numpy(), want, 'out')


@pytest.mark.parametrize('share_qk', [False, True])
def test_self_attention_token_by_token_equals_the_full_forward(share_qk):
  """`_test_fast_inference` of the reference (efficient_attention_test.py:158-206; there: share_qk=False, chunk_len 5,
  n_chunks_before 1, seqlen 10): with seqlen == 2 * chunk_len every token's window is all of its past, so decoding one
  token at a time through `pure_fn` in predict mode must reproduce the full forward of the same layer in eval mode."""
  import trax_b200
  rng = np.random.default_rng(5)
  B, H, D, C = 2, 2, 128, 64
  L = 2 * C
  kw = dict(n_heads=H, d_qk=64, d_v=64, share_qk=share_qk, causal=True, chunk_len=C, n_chunks_before=1)
  ref_layer = trax_b200.SelfAttention(mode='eval', **kw)
  weights, state = ref_layer.init(trax_b200.ShapeDtype((B, L, D)))
  x = torch.from_numpy(util.bf16_round(rng.uniform(size=(B, L, D)))).cuda()
  ref_out, _ = ref_layer.pure_fn(x, weights, state, rng=np.array([0, 0], np.uint32))
  cfg = SA.SelfAttentionConfig(**kw)
  want = SA.forward_and_or_backward(cfg, x.cpu().numpy(), tuple(a.cpu().numpy().astype(np.float64) for a in weights))[0]
  util.assert_close_layer(ref_out.float().cpu().numpy(), want, 'full forward')
  test_layer = trax_b200.SelfAttention(mode='predict', predict_mem_len=L, predict_drop_len=C, **kw)
  cur_state = test_layer.init(trax_b200.ShapeDtype((B, 1, D)))[1]
  outs = []
  for i in range(L):
    cur_out, cur_state = test_layer.pure_fn(x[:, i:i + 1].contiguous(), weights, cur_state, np.array([0, 0], np.uint32))
    outs.append(cur_out)
  out = torch.cat(outs, dim=1)
  assert int(cur_state[0]) == L
  util.assert_close_layer(out.float().cpu().numpy(), want, 'token by token vs oracle')
  util.assert_close_layer(out.float().cpu().numpy(), ref_out.float().cpu().numpy(), 'token by token vs full forward')


@pytest.mark.parametrize('prefix_len,dtype', [(128, torch.float32), (100, torch.bfloat16), (0, torch.float32)])
def test_lsh_predict_layer_prefix_then_tokens(prefix_len, dtype):
  """Through `LSHSelfAttention(mode='predict').forward`: a prefix (a multiple of chunk_len, or one that needs the zero
  padding of EA:2007-2011), then single tokens until the memory has rolled twice.  After every call the output and every
  state leaf are compared with the oracle's; the oracle takes the bucket ids the device computed (their bit-exactness is
  the business of the step test above and of the training path's hash tests)."""
  import trax_b200
  rng = np.random.default_rng(23)
  B, H, D, C, nh, M, drop = 2, 2, 128, 64, 2, 256, 64
  kw = dict(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=C, n_chunks_before=0, n_hashes=nh, n_buckets=8)
  cfg, pcfg = O.LSHConfig(**kw), P.PredictConfig(predict_mem_len=M, predict_drop_len=drop)
  layer = trax_b200.LSHSelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=drop, **kw)
  layer.init(trax_b200.ShapeDtype((B, 1, D), dtype))
  assert [tuple(t.shape) for t in (layer.state[1][0],) + tuple(layer.state[2])] == [(B, M, D), (B * H, nh * M), (B * H,), (B * H, 2)]
  w = tuple(a.cpu().numpy().astype(np.float64) for a in layer.weights)
  rot = rng.standard_normal((B * H,) + O.rotations_shape(cfg, 2)).astype(np.float32)
  layer._rotations_override = torch.from_numpy(rot)
  calls = ([prefix_len] if prefix_len else []) + [1] * (M - prefix_len + 2 * drop + 3)
  xs = util.bf16_round(rng.standard_normal((B, sum(calls), D)))
  ostate = P.init_state(cfg, pcfg, B, D)
  t0 = 0
  for n in calls:
    x = xs[:, t0:t0 + n]
    t0 += n
    out = layer.forward(torch.from_numpy(x).cuda().to(dtype))
    mem_end, (mem,), (buckets, buckets_idx, _) = layer.state
    got_b = buckets.cpu().numpy().reshape(B * H, nh, M)
    if n > 1:
      padded = -(-n // C) * C

      def new_ids_fn(u):                                           # ids of the padded tail never reach the state: the pad rows are
        ids = np.zeros((nh, padded), np.int32)                       # zero vectors, whose bucket is 0 + the round's offset (EA:105)
        ids[:, :n] = got_b[u, :, :n]
        ids[:, n:] = (np.arange(nh) * 8)[:, None]
        return ids.reshape(-1)
    else:
      q_start = int(mem_end) - 1

      def new_ids_fn(u, _q=q_start):
        return got_b[u, :, _q]
    want, ostate = P.predict_forward(cfg, pcfg, x, w, ostate, None, new_ids_fn=new_ids_fn)
    assert int(mem_end) == ostate[0]
    np.testing.assert_array_equal(mem.float().cpu().numpy(), ostate[1].astype(np.float32))
    np.testing.assert_array_equal(got_b.reshape(B * H, nh * M), ostate[2][0])
    np.testing.assert_array_equal(buckets_idx.cpu().numpy(), ostate[2][1])
    assert out.dtype == dtype and tuple(out.shape) == (B, n, D)
    util.assert_close_layer(out.float().cpu().numpy(), want, 'out after %d tokens' % t0)
  assert int(layer.state[0]) < t0                                    # the memory did roll


@pytest.mark.parametrize('name', ['lsh', 'self'])
def test_predict_layers_against_the_reference_own_outputs(name):
  """tests/golden/reference_predict.npz: outputs of the REFERENCE's own predict-mode code (make_predict_golden.py) for a
  prefix followed by single tokens through two rolls of the memory.  The LSH case is sized so that every earlier slot is
  attended (n_hashes * chunk_len * 2 == predict_mem_len): its output does not depend on the bucket ids, so the device's
  output is compared with the reference's directly; the bucket memory (the device hashes bf16 projections, the reference
  fp64 ones) may differ at near-ties of the argmax, bounded here at 5 % like tests/test_gpu_golden.py does for the layer."""
  import os
  import trax_b200
  from tests.golden import make_predict_golden as G
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_predict.npz'))
  c = G.CASES[name]
  w, xs = G.inputs(name)
  cls = trax_b200.LSHSelfAttention if name == 'lsh' else trax_b200.SelfAttention
  layer = cls(mode='predict', predict_mem_len=G.M, predict_drop_len=G.DROP, **c['kw'])
  layer.init(trax_b200.ShapeDtype((G.B, 1, G.D)))
  layer.weights = tuple(torch.from_numpy(a.astype(np.float32)).cuda() for a in w)
  if name == 'lsh':
    layer._rotations_override = torch.from_numpy(g['lsh/rot'])
  outs, t0 = [], 0
  for n in G.calls(c):
    outs.append(layer.forward(torch.from_numpy(xs[:, t0:t0 + n].astype(np.float32)).cuda()))
    t0 += n
  out = torch.cat(outs, dim=1).cpu().numpy()
  util.assert_close_layer(out, g[name + '/out'], 'out')
  assert int(layer.state[0]) == int(g[name + '/mem_end'])
  np.testing.assert_array_equal(layer.state[1][0].cpu().numpy(), g[name + '/mem'])
  if name == 'lsh':
    got_b = layer.state[2][0].cpu().numpy()
    assert (got_b != g['lsh/buckets']).mean() <= 0.05
    np.testing.assert_array_equal(layer.state[2][1].cpu().numpy(), g['lsh/buckets_idx'])


# ---- the weight-less core and its wrapper (EA:2823-3033, 3493-3620) ----------------------------------------------------
def _selector_weights():
  return util.core_identity_weights()                                # x = [qk | v]: w_q / w_v pick the halves, w_o = I


@pytest.mark.parametrize('M,C,nb,nh,n_buckets,q_start,dtype', [
    (512, 64, 0, 2, 8, 0, torch.float32),
    (512, 64, 0, 2, 8, 300, torch.float32),
    (256, 128, 1, 1, [4, 2], 255, torch.bfloat16),
])
def test_pure_core_predict_step_matches_oracle(M, C, nb, nh, n_buckets, q_start, dtype):
  """`lsh_predict_attend` (the step of `PureLSHSelfAttention`, EA:2858-2932): bucket ids bit-exact vs the oracle's hash of
  the (bf16-rounded) qk row, bucket memory, output."""
  import trax_b200
  from trax_b200 import predict
  rng = np.random.default_rng(31 + q_start)
  B, H = 2, 2
  BH = B * H
  kw = dict(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=C, n_chunks_before=nb, n_hashes=nh, n_buckets=n_buckets)
  cfg, pcfg = O.LSHConfig(**kw), P.PredictConfig(predict_mem_len=M, predict_drop_len=C)
  layer = trax_b200.PureLSHSelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=C, **kw)
  qk_mem, v_mem = util.bf16_round(rng.standard_normal((BH, M, 64))), util.bf16_round(rng.standard_normal((BH, M, 64)))
  qk_mem[:, q_start + 1:] = 0
  v_mem[:, q_start + 1:] = 0
  nbk = int(np.prod(O.bucket_factors(n_buckets, 2, C)))
  buckets = util.random_valid_buckets(rng, BH, nh, M, nbk).reshape(BH, nh, M)
  buckets[:, :, q_start:] = 0
  buckets = buckets.reshape(BH, nh * M)
  rot = rng.standard_normal((BH,) + O.rotations_shape(cfg, 2)).astype(np.float32)
  buckets_d = torch.from_numpy(buckets.copy()).cuda()
  out_d = predict._pure_step(layer, torch.from_numpy(qk_mem).cuda().to(dtype), torch.from_numpy(v_mem).cuda().to(dtype), q_start,
                             buckets_d, torch.from_numpy(rot).cuda())
  got_b = buckets_d.cpu().numpy()
  new_ids = got_b.reshape(BH, nh, M)[:, :, q_start]
  w_q, w_v, w_o = _selector_weights()
  want = np.zeros((BH, 1, 64))
  for u in range(BH):
    q_row = qk_mem[u, q_start].astype(np.float32)
    want_ids = O.hash_vectors(cfg, np.stack([q_row, q_row]), rot[u]).reshape(nh, 2)[:, 0]
    np.testing.assert_array_equal(new_ids[u], want_ids, err_msg='unit %d' % u)
    x = np.concatenate([qk_mem[u], v_mem[u]], axis=-1)
    want[u], nb_u, _ = P.incremental_forward_unit(cfg, pcfg, x, q_start, 1, w_q, w_v, w_o, buckets[u], q_start, None,
                                                  new_ids=new_ids[u])
    np.testing.assert_array_equal(nb_u, got_b[u])
  assert out_d.dtype == dtype and tuple(out_d.shape) == (BH, 1, 64)
  util.assert_close(out_d.float().cpu().numpy(), want, 'out')


def test_pure_core_predict_layer_prefix_then_tokens():
  """`PureLSHSelfAttention(mode='predict')` through `pure_fn`: a prefix that needs padding, then single tokens through two
  rolls; output and state leaves vs the oracle on the device's bucket ids."""
  import trax_b200
  rng = np.random.default_rng(37)
  B, H, C, nh, M, drop = 2, 2, 64, 2, 256, 64
  BH = B * H
  kw = dict(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=C, n_chunks_before=0, n_hashes=nh, n_buckets=8)
  cfg, pcfg = O.LSHConfig(**kw), P.PredictConfig(predict_mem_len=M, predict_drop_len=drop)
  layer = trax_b200.PureLSHSelfAttention(mode='predict', predict_mem_len=M, predict_drop_len=drop, **kw)
  _, state = layer.init((trax_b200.ShapeDtype((BH, 1, 64)), trax_b200.ShapeDtype((BH, 1, 64))))
  rot = rng.standard_normal((BH,) + O.rotations_shape(cfg, 2)).astype(np.float32)
  layer._rotations_override = torch.from_numpy(rot)
  prefix_len = 100
  calls = [prefix_len] + [1] * (M - prefix_len + 2 * drop + 3)
  qks, vs = util.bf16_round(rng.standard_normal((BH, sum(calls), 64))), util.bf16_round(rng.standard_normal((BH, sum(calls), 64)))
  ostate = (0, (np.zeros((BH, M, 64)), np.zeros((BH, M, 64))), (np.zeros((BH, nh * M), np.int32), np.zeros((BH,), np.int32)))
  t0 = 0
  for n in calls:
    qk, v = qks[:, t0:t0 + n], vs[:, t0:t0 + n]
    t0 += n
    out, state = layer.pure_fn((torch.from_numpy(qk).cuda(), torch.from_numpy(v).cuda()), (), state, None)
    mem_end, (qk_mem, v_mem), (buckets, buckets_idx, _) = state
    got_b = buckets.cpu().numpy().reshape(BH, nh, M)
    if n > 1:
      padded = -(-n // C) * C

      def new_ids_fn(u):
        ids = np.zeros((nh, padded), np.int32)
        ids[:, :n] = got_b[u, :, :n]
        ids[:, n:] = (np.arange(nh) * 8)[:, None]                    # zero pad rows hash to bucket 0 of their round
        return ids.reshape(-1)
    else:
      def new_ids_fn(u, _q=int(mem_end) - 1):
        return got_b[u, :, _q]
    want, ostate = P.pure_predict_forward(cfg, pcfg, qk, v, ostate, None, new_ids_fn=new_ids_fn)
    assert int(mem_end) == ostate[0]
    np.testing.assert_array_equal(qk_mem.cpu().numpy(), ostate[1][0].astype(np.float32))
    np.testing.assert_array_equal(v_mem.cpu().numpy(), ostate[1][1].astype(np.float32))
    np.testing.assert_array_equal(got_b.reshape(BH, nh * M), ostate[2][0])
    np.testing.assert_array_equal(buckets_idx.cpu().numpy(), ostate[2][1])
    assert tuple(out.shape) == (BH, n, 64)
    util.assert_close(out.float().cpu().numpy(), want, 'out after %d tokens' % t0)
  assert int(state[0]) < t0


def test_pure_lsh_wrapper_predict_token_by_token():
  """`PureLSHSelfAttentionWrapper(mode='predict')` (EA:3493-3540: Dense projections of the NEW tokens, the core's predict
  mode, head merge, output Dense): a prefix, then single tokens; vs the same composition around the oracle's core."""
  import trax_b200
  rng = np.random.default_rng(41)
  B, H, C, nh, M, drop = 2, 2, 64, 1, 128, 32
  D = 64 * H
  kw = dict(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=C, n_chunks_before=1, n_hashes=nh, n_buckets=4)
  cfg, pcfg = O.LSHConfig(**kw), P.PredictConfig(predict_mem_len=M, predict_drop_len=drop)
  layer = trax_b200.PureLSHSelfAttentionWrapper(mode='predict', predict_mem_len=M, predict_drop_len=drop, bias=False,
                                                num_weights=2, **kw)
  weights, state = layer.init(trax_b200.ShapeDtype((B, 1, D)))
  rot = rng.standard_normal((B * H,) + O.rotations_shape(cfg, 2)).astype(np.float32)
  layer.sublayers[1]._rotations_override = torch.from_numpy(rot)
  w_qk, w_v = (w.cpu().numpy().astype(np.float64) for w in weights[0])
  w_out = weights[3].cpu().numpy().astype(np.float64)
  calls = [M] + [1] * (2 * drop + 3)                                # a prefix that fills the memory (two chunks), then two rolls
  xs = util.bf16_round(rng.standard_normal((B, sum(calls), D)))
  ostate = (0, (np.zeros((B * H, M, 64)), np.zeros((B * H, M, 64))), (np.zeros((B * H, nh * M), np.int32), np.zeros((B * H,), np.int32)))

  def split(t):                                                      # attention.py:347-364
    return t.reshape(B, -1, H, 64).transpose(0, 2, 1, 3).reshape(B * H, -1, 64)
  t0 = 0
  for n in calls:
    x = xs[:, t0:t0 + n]
    t0 += n
    out, state = layer.pure_fn(torch.from_numpy(x).cuda(), weights, state, None)
    want_core, ostate = P.pure_predict_forward(cfg, pcfg, split(x @ w_qk), split(x @ w_v), ostate, lambda u, n_rows: rot[u])
    want = want_core.reshape(B, H, n, 64).transpose(0, 2, 1, 3).reshape(B, n, D) @ w_out
    # every earlier slot is attended (n_hashes * chunk_len * 2 == predict_mem_len), so the output does not hinge on bucket ids
    util.assert_close_layer(out.float().cpu().numpy(), want, 'out after %d tokens' % t0)
    assert int(state[1][0]) == ostate[0]
