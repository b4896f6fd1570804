"""GPU parity tests, stage by stage, through the C ABI (trax_b200.ops → liblsh_attn_b200.so) against
the CPU oracle.  Integer results (bucket ids, permutations) must be bit-exact; floating-point
results must be within 2e-2 relative / 1e-3 absolute of the oracle's float64 result on identical
(bf16-representable) inputs — see tests/util.py for the exact form.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import lsh_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def _dims(B, H, L, D, C, nb, na, nh, factors, causal=True, masked=False, act=1):
  from trax_b200 import _lib
  return _lib.make_dims(B, H, L, D, 64, 64, C, nb, na, nh, factors, causal, masked, act)


def _cuda(a, dtype=None):
  t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
  return t if dtype is None else t.to(dtype)


# ---- T1 hash ----------------------------------------------------------------------------------------
HASH_CASES = [
    # (L, nh, n_buckets) — single int (C1: 32), non-power-of-2 factors, large int (R=96), [32,32], [128,128]
    (1024, 1, 32), (512, 2, [16, 12]), (768, 2, 192), (2048, 4, [32, 32]), (512, 8, [128, 128]),
    (320, 3, 6), (64, 1, 2),
]


@pytest.mark.parametrize('L,nh,n_buckets', HASH_CASES)
@pytest.mark.parametrize('masked', [False, True])
def test_hash_f32_bit_exact(L, nh, n_buckets, masked):
  from trax_b200 import ops
  rng = np.random.default_rng(L + nh)
  B, H = 2, 2
  factors = O.bucket_factors(n_buckets, L, 64)
  cfg = util.make_cfg(H=H, C=64, nh=nh, n_buckets=n_buckets, masked=masked)
  vecs = rng.standard_normal((B * H, L, 64)).astype(np.float32)
  vecs[0, 5] = 0.0                       # all-zero row -> bucket 0 of each round
  vecs[1, 7, :] = 0.0; vecs[1, 7, 0] = 1.0
  rot = rng.standard_normal((B * H, 64, nh, sum(factors) // 2)).astype(np.float32)
  if rot.shape[3] > 1:
    rot[1, 0, :, 1] = rot[1, 0, :, 0]    # exact tie between two rotation columns -> lowest index wins
  mask = (rng.random((B, L)) > 0.25) if masked else None
  dims = _dims(B, H, L, 64, 64 if L % 64 == 0 else 32, 1, 0, nh, factors, masked=masked)
  if (nh * L) % dims.C != 0:
    dims.C = 32
  if (nh * L) % 32 != 0:
    pytest.skip('shape not chunkable')
  got = ops.hash_f32(dims, _cuda(vecs), _cuda(rot), None if mask is None else _cuda(mask.astype(np.uint8)))
  got = got.cpu().numpy()
  for u in range(B * H):
    want = O.hash_vectors(cfg, vecs[u], rot[u], None if mask is None else mask[u // H])
    want_c = O.hash_vectors_c(cfg, vecs[u], rot[u], None if mask is None else mask[u // H])
    np.testing.assert_array_equal(want, want_c)
    np.testing.assert_array_equal(got[u], want)
  if not masked:
    nb_tot = int(np.prod(factors))
    assert (got[0].reshape(nh, L)[:, 5] == np.arange(nh) * nb_tot).all()


def test_hash_bf16_qv_bit_exact():
  from trax_b200 import ops
  rng = np.random.default_rng(3)
  B, H, L, nh = 2, 3, 512, 4
  factors = [16, 8]
  cfg = util.make_cfg(H=H, C=64, nh=nh, n_buckets=factors)
  qv = util.bf16_round(rng.standard_normal((B, L, H, 128)))
  rot = rng.standard_normal((B * H, 64, nh, 12)).astype(np.float32)
  dims = _dims(B, H, L, 64, 64, 1, 0, nh, factors)
  got = ops.hash_qv(dims, _cuda(qv, torch.bfloat16), _cuda(rot)).cpu().numpy()
  for u in range(B * H):
    b, h = divmod(u, H)
    want = O.hash_vectors(cfg, np.ascontiguousarray(qv[b, :, h, :64]), rot[u])
    np.testing.assert_array_equal(got[u], want)


# ---- T2 sort ----------------------------------------------------------------------------------------
SORT_CASES = [
    # (L, nh, n_buckets_total_per_round)
    (1024, 1, 32), (16384, 1, 256), (16384, 2, 256), (16384, 4, 256), (16384, 8, 256),
    (4096, 2, 16384), (4096, 3, 1025), (96, 2, 7), (2048 + 64, 1, 2049),
]


@pytest.mark.parametrize('L,nh,nbk', SORT_CASES)
def test_sort_bit_exact(L, nh, nbk):
  from trax_b200 import ops
  rng = np.random.default_rng(L * 7 + nh)
  BH = 3
  buckets = util.random_valid_buckets(rng, BH, nh, L, nbk)
  buckets[0] = (np.arange(nh * L) // L) * nbk               # everything in one bucket per round
  # factors only need to multiply to nbk for the sort: use [nbk] when even, else masked (+1)
  masked = nbk % 2 == 1
  factors = [nbk - 1] if masked else [nbk]
  dims = _dims(1, BH, L, 64, 32, 1, 0, nh, factors, masked=masked)
  sticker, undo = ops.sort(dims, _cuda(buckets))
  sticker, undo = sticker.cpu().numpy(), undo.cpu().numpy()
  for u in range(BH):
    ws, wu = O.sort_buckets(buckets[u], L)
    np.testing.assert_array_equal(sticker[u], ws)
    np.testing.assert_array_equal(undo[u], wu)
    np.testing.assert_array_equal(undo[u][sticker[u]], np.arange(nh * L))


@pytest.mark.parametrize('L,nh,nbk', [(1024, 3, 8), (192, 2, 4), (65536, 1, 1024)])
def test_chunk_possort_bit_exact(L, nh, nbk):
  """Internal order of the tcgen05 kernels: every 128-slot chunk of sticker re-ordered by position, ties (a chunk that
  straddles two hash rounds, L % 128 != 0) keeping slot order — against a stable NumPy argsort."""
  from trax_b200 import ops
  rng = np.random.default_rng(L + nh)
  BH = 2
  buckets = util.random_valid_buckets(rng, BH, nh, L, nbk)
  dims = _dims(1, BH, L, 64, 128, 1, 0, nh, [nbk])
  sticker, _ = ops.sort(dims, _cuda(buckets))
  s2, bounds = ops.chunk_possort(dims, sticker, with_bounds=True)
  s2, bounds = s2.cpu().numpy(), bounds.cpu().numpy()
  st = sticker.cpu().numpy()
  for u in range(BH):
    ch = st[u].reshape(-1, 128)
    order = np.argsort(ch % L, axis=1, kind='stable')
    np.testing.assert_array_equal(s2[u].reshape(-1, 128), np.take_along_axis(ch, order, axis=1))
    if L % 128:
      continue      # the interval bounds are only defined (and used) when positions inside a chunk are unique
    # neighbour-chunk bounds: cnt_prev | eq_prev << 8 | cnt_next << 16 | eq_next << 24 (cyclic inside the unit)
    pos = s2[u].reshape(-1, 128) % L
    bd = bounds[u].reshape(-1, 128)
    nc = pos.shape[0]
    for c in range(nc):
      for name, other, shift in (('prev', pos[(c - 1) % nc], 0), ('next', pos[(c + 1) % nc], 16)):
        cnt = (other[None, :] < pos[c][:, None]).sum(1)
        eq = (other[None, :] == pos[c][:, None]).any(1)
        np.testing.assert_array_equal((bd[c] >> shift) & 0xff, cnt, err_msg='%s cnt chunk %d' % (name, c))
        np.testing.assert_array_equal((bd[c] >> (shift + 8)) & 1, eq.astype(np.int64), err_msg='%s eq chunk %d' % (name, c))


def test_sort_rejects_int32_key_overflow():
  from trax_b200 import _lib
  lib = _lib.load()
  dims = _dims(1, 1, 1 << 20, 64, 128, 1, 0, 1, [128, 128])   # L*nb = 2^34
  assert lib.lsh_attn_check_dims(ctypes.byref(dims)) != 0
  assert b'wrap' in lib.lsh_attn_last_error()
  dims = _dims(1, 1, 1 << 20, 64, 128, 1, 0, 1, [32, 32])
  assert lib.lsh_attn_check_dims(ctypes.byref(dims)) == 0


# ---- T3/T4 attention core ---------------------------------------------------------------------------
def _core_case(seed, B, H, L, C, nb, na, nh, nbk, causal, masked):
  rng = np.random.default_rng(seed)
  qv = util.bf16_round(rng.standard_normal((B, L, H, 128)))
  # dims.masked adds the padding bucket (EA:1908): per-round offsets are then multiples of nbk + 1
  buckets = util.random_valid_buckets(rng, B * H, nh, L, nbk + (1 if masked else 0))
  mask = (rng.random((B, L)) > 0.2) if masked else None
  cfg = util.make_cfg(H=H, C=C, nb=nb, na=na, nh=nh, n_buckets=nbk, causal=causal, masked=masked)
  return cfg, qv, buckets, mask


def _oracle_core(cfg, qv, buckets, mask, B, H, dout=None, attn_keep=None):
  """Runs the oracle per unit with identity projections; returns dict of stacked results."""
  w_q, w_v, w_o = util.core_identity_weights()
  res, grads = [], []
  for u in range(B * H):
    b, h = divmod(u, H)
    r = O.forward_unit(cfg, qv[b, :, h, :].astype(np.float64), w_q, w_v, w_o, buckets=buckets[u],
                       mask=None if mask is None else mask[b], attn_keep=attn_keep)
    res.append(r)
    if dout is not None:
      grads.append(O.backward_unit(cfg, r, dout[b, :, h, :].astype(np.float64))[0])
  return res, grads


CORE_CASES = [
    # (B, H, L, C, nb, na, nh, nbk, causal, masked)
    (1, 2, 1024, 64, 1, 0, 1, 32, True, False),      # BASELINE config 1 core shape
    (2, 2, 512, 128, 1, 0, 4, 8, True, False),       # config 2 chunking, small
    (1, 2, 512, 64, 1, 1, 2, 16, False, False),      # encoder-style: non-causal, look-ahead
    (1, 2, 512, 64, 1, 0, 2, 16, True, True),        # padding mask
    (1, 1, 256, 64, 0, 0, 1, 8, True, False),        # no look-back (window = own chunk)
    (1, 1, 128, 128, 1, 0, 1, 2, True, False),       # one chunk: window = [itself, itself]
    (1, 2, 512, 256, 1, 0, 2, 4, True, False),
    (1, 1, 256, 32, 1, 0, 2, 8, True, False),        # forward only for C=32
    (1, 2, 192, 128, 1, 0, 2, 4, True, False),       # L % chunk_len != 0: a chunk straddles two hash rounds (same position twice)
    (1, 1, 320, 128, 1, 0, 2, 4, True, False),
]


@pytest.mark.parametrize('case', CORE_CASES)
def test_attend_fwd_and_combine(case):
  from trax_b200 import ops
  B, H, L, C, nb, na, nh, nbk, causal, masked = case
  cfg, qv, buckets, mask = _core_case(11, B, H, L, C, nb, na, nh, nbk, causal, masked)
  dims = _dims(B, H, L, 128, C, nb, na, nh, [nbk], causal, masked)
  mask_d = None if mask is None else _cuda(mask.astype(np.uint8))
  sticker, _ = ops.sort(dims, _cuda(buckets))
  o_r, logits = ops.attend_fwd(dims, _cuda(qv, torch.bfloat16), sticker, mask_d)
  o_c, lse_tot = ops.combine_fwd(dims, o_r, logits)
  res, _ = _oracle_core(cfg, qv, buckets, mask, B, H)
  o_r, logits = o_r.float().cpu().numpy(), logits.cpu().numpy()
  o_c, lse_tot = o_c.float().cpu().numpy(), lse_tot.cpu().numpy()
  for u in range(B * H):
    b, h = divmod(u, H)
    np.testing.assert_array_equal(sticker[u].cpu().numpy(), res[u].sticker)
    util.assert_close(o_r[u], res[u].o_rounds, 'o_rounds[%d]' % u)
    util.assert_close(logits[u], res[u].logits, 'logits[%d]' % u)
    util.assert_close(o_c[b, :, h, :], res[u].o, 'o_comb[%d]' % u)
    want_lse = O.logsumexp(res[u].logits.reshape(nh, L), axis=0)
    util.assert_close(lse_tot[u], want_lse, 'lse_tot[%d]' % u)


@pytest.mark.parametrize('case', [c for c in CORE_CASES if c[3] != 32])
def test_attend_bwd(case):
  from trax_b200 import ops
  B, H, L, C, nb, na, nh, nbk, causal, masked = case
  cfg, qv, buckets, mask = _core_case(12, B, H, L, C, nb, na, nh, nbk, causal, masked)
  rng = np.random.default_rng(5)
  do = util.bf16_round(rng.standard_normal((B, L, H, 64)))
  if mask is not None:
    # Padding queries whose whole window is padding have lse ~ -1e9, where fp32 cannot hold log(sum) (ulp 64): the
    # reference's own fp32 result is ill-defined there.  A real loss never back-propagates into padding outputs.
    do = do * mask[:, :, None, None]
  dims = _dims(B, H, L, 128, C, nb, na, nh, [nbk], causal, masked)
  mask_d = None if mask is None else _cuda(mask.astype(np.uint8))
  qv_d = _cuda(qv, torch.bfloat16)
  sticker, _ = ops.sort(dims, _cuda(buckets))
  o_r, logits = ops.attend_fwd(dims, qv_d, sticker, mask_d)
  o_c, lse_tot = ops.combine_fwd(dims, o_r, logits)
  dqv = ops.attend_bwd(dims, qv_d, sticker, o_c, lse_tot, _cuda(do, torch.bfloat16), mask_d)
  dqv = dqv.float().cpu().numpy()
  _, grads = _oracle_core(cfg, qv, buckets, mask, B, H, dout=do)
  for u in range(B * H):
    b, h = divmod(u, H)
    util.assert_close(dqv[b, :, h, :64], grads[u][:, :64], 'dq[%d]' % u)
    util.assert_close(dqv[b, :, h, 64:], grads[u][:, 64:], 'dv[%d]' % u)


DROPOUT_CASES = [
    # (B, H, L, C, nb, na, nh, nbk, causal, masked)
    (2, 2, 512, 128, 1, 0, 4, 8, True, False),       # the tcgen05 shape: dropout takes its slot-ordered (generic) instantiation
    (1, 3, 640, 128, 1, 0, 2, 4, True, False),       # odd chunk count per round: look-back of chunk 0 has the same row flip
    (1, 2, 1024, 64, 1, 0, 1, 32, True, False),      # mma.sync shape (BASELINE config 1)
    (1, 2, 512, 64, 1, 1, 2, 16, False, True),       # look-ahead window, padding mask
    (1, 2, 512, 128, 0, 1, 2, 8, False, False),      # tcgen05 generic path, window = [own, next]
    (1, 1, 512, 256, 1, 0, 2, 4, True, False),       # chunk_len 256 (reformer_enwik8.gin:40)
]


@pytest.mark.parametrize('case', DROPOUT_CASES)
@pytest.mark.parametrize('rate', [0.2, 0.5])
def test_attention_dropout_fwd_bwd(case, rate):
  """EA:254-262: ONE (chunk_len, window) keep multiplier, indexed by the ORIGINAL slot of query and key inside their
  chunks, applied to exp(dots - lse) before the product with v (the log-sum-exp does not see it) — and its VJP
  (dV += (P∘m)^T do, dS = P∘(m∘dP - D)).  reformer_enwik8.gin:94 / reformer_imagenet64.gin:69 train with 0.2."""
  from trax_b200 import ops
  B, H, L, C, nb, na, nh, nbk, causal, masked = case
  cfg, qv, buckets, mask = _core_case(41, B, H, L, C, nb, na, nh, nbk, causal, masked)
  rng = np.random.default_rng(int(rate * 100) + C)
  W = C * (1 + nb + na)
  keep = ((rng.random((C, W)) >= rate) / (1.0 - rate)).astype(np.float32)
  do = util.bf16_round(rng.standard_normal((B, L, H, 64)))
  if mask is not None:
    do = do * mask[:, :, None, None]
  dims = _dims(B, H, L, 128, C, nb, na, nh, [nbk], causal, masked)
  mask_d = None if mask is None else _cuda(mask.astype(np.uint8))
  qv_d, keep_d = _cuda(qv, torch.bfloat16), _cuda(keep)
  sticker, _ = ops.sort(dims, _cuda(buckets))
  o_r, logits = ops.attend_fwd(dims, qv_d, sticker, mask_d, attn_keep=keep_d)
  o_c, lse_tot = ops.combine_fwd(dims, o_r, logits)
  dqv = ops.attend_bwd(dims, qv_d, sticker, o_c, lse_tot, _cuda(do, torch.bfloat16), mask_d, attn_keep=keep_d)
  res, grads = _oracle_core(cfg, qv, buckets, mask, B, H, dout=do, attn_keep=keep.astype(np.float64))
  o_r, logits, dqv = o_r.float().cpu().numpy(), logits.cpu().numpy(), dqv.float().cpu().numpy()
  no_drop, _ = ops.attend_fwd(dims, qv_d, sticker, mask_d)
  assert not torch.equal(no_drop, torch.from_numpy(o_r).cuda().to(torch.bfloat16)), 'the keep matrix had no effect'
  for u in range(B * H):
    b, h = divmod(u, H)
    util.assert_close(o_r[u], res[u].o_rounds, 'o_rounds[%d]' % u)
    util.assert_close(logits[u], res[u].logits, 'logits[%d]' % u)
    util.assert_close(dqv[b, :, h, :64], grads[u][:, :64], 'dq[%d]' % u, frac_bad_max=0.02)
    util.assert_close(dqv[b, :, h, 64:], grads[u][:, 64:], 'dv[%d]' % u, frac_bad_max=0.02)


def test_forward_rows_whose_softmax_underflows_are_redone_exactly():
  """the tcgen05 forward shifts every row by its analytic self score; with large-norm queries whose visible keys
  all score far below it the exponentials flush to zero.  Such rows are queued and redone with the true maximum."""
  from trax_b200 import ops
  B, H, L, C, nh, nbk = 1, 2, 512, 128, 2, 4
  cfg, qv, buckets, _ = _core_case(43, B, H, L, C, 1, 0, nh, nbk, True, False)
  qv[..., :64] = util.bf16_round(qv[..., :64] * 24.0)          # |q| ~ 190: self score 8 r = 190, typical neighbour far below
  dims = _dims(B, H, L, 128, C, 1, 0, nh, [nbk], True, False)
  sticker, _ = ops.sort(dims, _cuda(buckets))
  o_r, logits = ops.attend_fwd(dims, _cuda(qv, torch.bfloat16), sticker)
  res, _ = _oracle_core(cfg, qv, buckets, None, B, H)
  o_r, logits = o_r.float().cpu().numpy(), logits.cpu().numpy()
  assert np.isfinite(logits).all() and (logits > -2e5).all()
  for u in range(B * H):
    util.assert_close(logits[u], res[u].logits, 'logits[%d]' % u)
    util.assert_close(o_r[u], res[u].o_rounds, 'o_rounds[%d]' % u, frac_bad_max=0.03)


def test_degenerate_rows():
  """position 0 under a causal mask attends only to itself -> o = v_0 exactly-ish."""
  from trax_b200 import ops
  B, H, L, C, nh, nbk = 1, 1, 256, 64, 2, 4
  cfg, qv, buckets, _ = _core_case(13, B, H, L, C, 1, 0, nh, nbk, True, False)
  dims = _dims(B, H, L, 128, C, 1, 0, nh, [nbk], True, False)
  sticker, _ = ops.sort(dims, _cuda(buckets))
  o_r, logits = ops.attend_fwd(dims, _cuda(qv, torch.bfloat16), sticker)
  o_c, _ = ops.combine_fwd(dims, o_r, logits)
  np.testing.assert_allclose(o_c[0, 0, 0].float().cpu().numpy(), qv[0, 0, 0, 64:], rtol=1e-2, atol=1e-3)
  assert (logits.cpu().numpy()[0].reshape(nh, L)[:, 0] < -9e4).all()   # only the -1e5 self entry is visible


def test_make_rotations_statistics_and_determinism():
  from trax_b200 import ops
  dims = _dims(2, 4, 4096, 64, 128, 1, 0, 4, [32, 32])
  keys = torch.arange(16, dtype=torch.int32, device='cuda').reshape(8, 2) * 77 + 1
  r1, k1 = ops.make_rotations(dims, keys)
  r2, k2 = ops.make_rotations(dims, keys)
  assert torch.equal(r1, r2) and torch.equal(k1, k2)
  assert not torch.equal(k1, keys)
  r3, _ = ops.make_rotations(dims, k1)
  assert not torch.equal(r1, r3)
  x = r1.flatten().double()
  assert abs(x.mean().item()) < 0.02 and abs(x.std().item() - 1.0) < 0.02
  assert abs((x ** 4).mean().item() - 3.0) < 0.2
  assert not torch.equal(r1[0], r1[1])


@pytest.mark.parametrize('max_ctas', ['1', '3', '7'])
@pytest.mark.parametrize('nh,L', [(4, 1024), (1, 2048), (3, 1280), (1, 384), (3, 384)])   # incl. odd chunk counts (look-back of chunk 0 has the same row order)
def test_persistent_walk_many_chunks_per_cta(max_ctas, nh, L, monkeypatch):
  """The tcgen05 kernels walk contiguous chunk ranges per CTA (tile ring, carried dQ, unit-boundary replays).  The
  default grid gives small problems one chunk per CTA, so force 1 / 3 / 7 CTAs: every CTA then crosses ring
  wrap-arounds and (with B*H = 3 units) unit boundaries.  Forward and backward vs the oracle."""
  from trax_b200 import ops
  monkeypatch.setenv('LSH_ATTN_MAX_CTAS', max_ctas)
  B, H, C, nbk = 1, 3, 128, 8
  cfg, qv, buckets, mask = _core_case(31 + nh, B, H, L, C, 1, 0, nh, nbk, True, False)
  rng = np.random.default_rng(6)
  do = util.bf16_round(rng.standard_normal((B, L, H, 64)))
  dims = _dims(B, H, L, 128, C, 1, 0, nh, [nbk], True, False)
  qv_d = _cuda(qv, torch.bfloat16)
  sticker, _ = ops.sort(dims, _cuda(buckets))
  o_r, logits = ops.attend_fwd(dims, qv_d, sticker)
  o_c, lse_tot = ops.combine_fwd(dims, o_r, logits)
  dqv = ops.attend_bwd(dims, qv_d, sticker, o_c, lse_tot, _cuda(do, torch.bfloat16)).float().cpu().numpy()
  res, grads = _oracle_core(cfg, qv, buckets, mask, B, H, dout=do)
  o_r, logits = o_r.float().cpu().numpy(), logits.cpu().numpy()
  for u in range(B * H):
    b, h = divmod(u, H)
    util.assert_close(o_r[u], res[u].o_rounds, 'o_rounds[%d]' % u)
    util.assert_close(logits[u], res[u].logits, 'logits[%d]' % u)
    util.assert_close(dqv[b, :, h, :64], grads[u][:, :64], 'dq[%d]' % u)
    util.assert_close(dqv[b, :, h, 64:], grads[u][:, 64:], 'dv[%d]' % u)
