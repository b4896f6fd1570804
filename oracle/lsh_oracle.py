"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package `trax_b200`).

CPU restatement (NumPy + a small C helper) of the reference's Reformer LSH-attention hot path,
`trax/layers/research/efficient_attention.py` ("EA"), class `LSHSelfAttention`.  Every function cites
the reference lines it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module.

PARITY PINNED AGAINST THE LIVE REFERENCE (forward directly, backward through derivatives of the
reference's forward).  JAX is not installable in this image and the checkout holds no golden vectors
for this path, but the reference is pure Python with a NumPy backend of its
own: `oracle/ref_live.py` executes the reference's files from /root/reference under that backend
(stubs only for the absent third-party packages; `use_reference_code=True` is the reference's own
Python loop over units around `forward_unbatched`).  `tests/golden/make_reference_golden.py` stores
what it returns in `tests/golden/reference_live.npz`; `tests/test_reference_pin.py` checks that this
restatement reproduces the reference's bucket ids bit for bit and its float64 outputs to 1e-12
(causal / bidirectional, masked, look-ahead chunks, single and factored bucket counts, the auto
factor-list rule, the weight-less PureLSH core, PureLSHSelfAttentionWrapper with and without rotary
embedding, the ReversibleHalfResidual block with LayerNorm), and that the analytic VJP below equals central
differences of the reference's forward to 1e-6 relative (observed 1e-9..1e-11) — the reference's
backward is `jax.vjp` of exactly that function (EA:2399-2421).  What is NOT pinned: XLA's own fp32
summation order inside the hash einsum on an accelerator (no XLA here; bucket ids can differ from a
TPU/GPU run of the reference at exact near-ties of the argmax, as they do between XLA backends), and
`jax.random` bit streams (rotations and dropout masks are inputs).  Further cross-checks: an
independent torch-autograd restatement (`lsh_oracle_torch.py`) and the reference tests' own
invariants re-run on it (tests/test_oracle.py).

Random rotations are an explicit input (EA:91-93 draws them with jax.random.normal, which is not
reproducible offline; under the reference's NumPy backend they come from numpy.random, which the pin
script seeds and records); so are dropout keep-masks.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional, Sequence, Union

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_helper() -> str:
  """Compiles oracle/hash_oracle.c → oracle/_build/liboracle.so (gcc).  Returns the .so path."""
  so = os.path.join(_HERE, '_build', 'liboracle.so')
  src = os.path.join(_HERE, 'hash_oracle.c')
  if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.check_call(['make', '-C', _HERE, '_build/liboracle.so'],
                          stdout=subprocess.DEVNULL)
  return so


def _lib():
  global _LIB
  if _LIB is None:
    _LIB = ctypes.CDLL(build_c_helper())
    _LIB.oracle_rotate_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                       ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
    _LIB.oracle_rotate_f32.restype = None
    _LIB.oracle_hash_vecs.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                      ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                      ctypes.c_void_p]
    _LIB.oracle_hash_vecs.restype = None
  return _LIB


# --------------------------------------------------------------------------------------------
# Configuration
# --------------------------------------------------------------------------------------------
@dataclass
class LSHConfig:
  """Hyper-parameters of the layer that the per-unit algorithm needs (EA:1732-1799)."""
  n_heads: int = 2
  d_qk: int = 64
  d_v: int = 64
  causal: bool = False
  masked: bool = False
  chunk_len: int = 128
  n_chunks_before: int = 1
  n_chunks_after: int = 0
  n_hashes: int = 1
  n_buckets: Union[None, int, Sequence[int]] = None
  max_length_for_buckets: Optional[int] = None


def bucket_factors(n_buckets, seqlen: int, chunk_len: int):
  """EA:1890-1902 — the bucket list `hash_vecs` receives.  Returns a python list of even ints."""
  if n_buckets is None:
    n = 2 * max(1, seqlen // chunk_len)
    if n <= 128:
      return [n]
    div = 2 ** math.ceil(math.log2(math.sqrt(n)))
    rest = 2 * (n // (2 * div))
    return [div, rest]
  if isinstance(n_buckets, (int, np.integer)):
    return [int(n_buckets)]
  return [int(f) for f in n_buckets]


def rotations_shape(cfg: LSHConfig, seqlen: int):
  """EA:79-91 — (d_qk, n_hashes, rot_size // 2)."""
  factors = bucket_factors(cfg.n_buckets, seqlen, cfg.chunk_len)
  for f in factors:
    assert f % 2 == 0  # EA:80, 87
  return (cfg.d_qk, cfg.n_hashes, sum(factors) // 2)


# --------------------------------------------------------------------------------------------
# Free functions (EA:54-317)
# --------------------------------------------------------------------------------------------
def length_normalized(x, epsilon=1e-6):
  """EA:54-57."""
  variance = np.mean(x ** 2, axis=-1, keepdims=True)
  return x / np.sqrt(variance + x.dtype.type(epsilon))


def rotate_f32(vecs: np.ndarray, rotations: np.ndarray) -> np.ndarray:
  """EA:95 `einsum('tf,fhb->htb')` in fp32 under the sequential-fmaf convention (hash_oracle.c).

  vecs (T, dq) fp32, rotations (dq, nh, R) fp32 → (nh, T, R) fp32.
  """
  vecs = np.ascontiguousarray(vecs, dtype=np.float32)
  rot = np.ascontiguousarray(rotations, dtype=np.float32)
  t, dq = vecs.shape
  assert rot.shape[0] == dq
  nh, r = rot.shape[1], rot.shape[2]
  out = np.empty((t, nh * r), dtype=np.float32)
  _lib().oracle_rotate_f32(vecs.ctypes.data, t, dq, rot.ctypes.data, nh * r, out.ctypes.data)
  return np.transpose(out.reshape(t, nh, r), (1, 0, 2))


def hash_vecs(vecs, n_buckets_in, n_hashes, rotations):
  """EA:60-119.  `rotations` replaces `rng` (shape (depth, n_hashes, rot_size//2), fp32).

  Returns (buckets int32 (n_hashes, T) WITHOUT per-round offsets, n_buckets).
  np.argmax returns the first maximal index, the same tie rule as jnp.argmax.
  """
  if isinstance(n_buckets_in, (int, np.integer)):
    n_buckets_in = [int(n_buckets_in)]
  rot_size, n_buckets = 0, 1
  for factor in n_buckets_in:
    assert factor % 2 == 0
    rot_size += factor
    n_buckets *= factor
  assert tuple(rotations.shape) == (vecs.shape[-1], n_hashes, rot_size // 2), rotations.shape
  rotated_vecs = rotate_f32(vecs, rotations)                      # EA:95
  buckets, cur_sum, cur_product = None, 0, 1                       # EA:108
  for factor in n_buckets_in:                                      # EA:109 (len==1 ≡ EA:103-105)
    rv = rotated_vecs[..., cur_sum:cur_sum + (factor // 2)]
    cur_sum += factor // 2
    rv = np.concatenate([rv, -rv], axis=-1)
    am = np.argmax(rv, axis=-1).astype(np.int32)
    buckets = am if buckets is None else buckets + cur_product * am
    cur_product *= factor
  return buckets, n_buckets


def hash_vectors(cfg: LSHConfig, vecs, rotations, mask=None):
  """EA:1889-1916.  Returns flat int32 (n_hashes*T,) bucket ids INCLUDING per-round offsets."""
  factors = bucket_factors(cfg.n_buckets, vecs.shape[0], cfg.chunk_len)
  buckets, n_buckets = hash_vecs(vecs, factors, cfg.n_hashes, rotations)
  if mask is not None:
    n_buckets += 1                                                 # EA:1908
    buckets = np.where(mask[None, :], buckets, n_buckets - 1)      # EA:1909
  offsets = np.arange(cfg.n_hashes, dtype=np.int32)
  offsets = np.reshape(offsets * n_buckets, (-1, 1))               # EA:1913-1914
  return np.reshape(buckets + offsets, (-1,)).astype(np.int32)     # EA:1915


def hash_vectors_c(cfg: LSHConfig, vecs, rotations, mask=None):
  """Same as `hash_vectors` but through the loop-level C restatement (second opinion)."""
  factors = np.asarray(bucket_factors(cfg.n_buckets, vecs.shape[0], cfg.chunk_len), np.int32)
  vecs = np.ascontiguousarray(vecs, np.float32)
  rot = np.ascontiguousarray(rotations, np.float32)
  t, dq = vecs.shape
  out = np.empty((cfg.n_hashes, t), np.int32)
  m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
  _lib().oracle_hash_vecs(vecs.ctypes.data, t, dq, rot.ctypes.data, cfg.n_hashes, rot.shape[2],
                          factors.ctypes.data, len(factors),
                          None if m is None else m.ctypes.data, out.ctypes.data)
  return out.reshape(-1)


def sort_buckets(buckets: np.ndarray, seqlen: int):
  """EA:1946-1958.  int32 key = seqlen*buckets + ticker%seqlen (wraps like XLA int32), stable sort.

  Returns (sticker, undo_sort), both int32 (n_hashes*seqlen,).
  """
  n = buckets.shape[0]
  ticker = np.arange(n, dtype=np.int32)
  key64 = np.int64(seqlen) * buckets.astype(np.int64) + (ticker % seqlen).astype(np.int64)
  key = key64.astype(np.int32)                                     # silent wrap-around (F5)
  sticker = np.argsort(key, kind='stable').astype(np.int32)        # sort_key_val(key, ticker)
  undo_sort = np.argsort(sticker, kind='stable').astype(np.int32)  # sort_key_val(sticker, ticker)
  return sticker, undo_sort


def key_overflows(seqlen: int, n_hashes: int, n_buckets_total: int) -> bool:
  """True when the int32 sort key of EA:1947 would wrap (max key = L*(nh*nb-1) + L-1)."""
  return seqlen * n_hashes * n_buckets_total - 1 >= 2 ** 31


def look_adjacent(x, n_chunks_before, n_chunks_after):
  """EA:122-142."""
  if n_chunks_before == 0 and n_chunks_after == 0:
    return x
  slices = []
  for i in range(-n_chunks_before, n_chunks_after + 1):
    if i == 0:
      slices.append(x)
    else:
      slices.append(np.concatenate([x[i:, ...], x[:i, ...]], axis=0))
  return np.concatenate(slices, axis=1)


def mask_self_attention(dots, q_info, kv_info, causal=True, exclude_self=True, masked=False):
  """EA:145-160 — subtractive masks, positions compared as fp32, arithmetic in dots.dtype."""
  q_info = q_info.astype(np.float32)
  kv_info = kv_info.astype(np.float32)
  dt = dots.dtype.type
  if causal:
    dots = dots - dt(1e9) * (q_info < kv_info).astype(dots.dtype)
  if exclude_self:
    dots = dots - dt(1e5) * (q_info == kv_info).astype(dots.dtype)
  if masked:
    dots = dots - dt(1e9) * (kv_info < 0).astype(dots.dtype)
  return dots


def logsumexp(x, axis, keepdims=False):
  """jax.scipy.special.logsumexp: max-shifted."""
  m = np.max(x, axis=axis, keepdims=True)
  s = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m
  return s if keepdims else np.squeeze(s, axis=axis)


def attend(q, v, q_chunk_len, n_chunks_before, n_chunks_after, mask_fn, q_info, kv_info=None,
           keep_multiplier=None):
  """EA:163-268 for the shared-QK case (k=None).  Returns (out, lse, cache-for-backward).

  `keep_multiplier` (chunk_len, W) = keep/keep_prob replaces the rng draw of EA:254-262.
  """
  q_info = q_info + 1                                              # EA:201
  if kv_info is not None:
    kv_info = kv_info + 1                                          # EA:206
  q = np.reshape(q, (-1, q_chunk_len, q.shape[-1]))                # EA:210
  q_info = np.reshape(q_info, (-1, q_chunk_len))                   # EA:211
  k = q                                                            # EA:215
  if kv_info is None:
    kv_info = q_info                                               # EA:218
  else:
    kv_info = np.reshape(kv_info, (-1, q_chunk_len))               # EA:221
  v = np.reshape(v, (-1, q_chunk_len, v.shape[-1]))                # EA:227
  k = length_normalized(k)                                         # EA:230
  k = k / np.sqrt(k.dtype.type(k.shape[-1]))                       # EA:231
  k = look_adjacent(k, n_chunks_before, n_chunks_after)            # EA:239
  v = look_adjacent(v, n_chunks_before, n_chunks_after)            # EA:240
  kv_info = look_adjacent(kv_info, n_chunks_before, n_chunks_after)  # EA:241
  dots = np.matmul(q, np.swapaxes(k, -1, -2))                      # EA:244
  dots = mask_fn(dots, q_info[..., :, None], kv_info[..., None, :])  # EA:248
  dots_logsumexp = logsumexp(dots, axis=-1, keepdims=True)         # EA:251
  p = np.exp(dots - dots_logsumexp)                                # EA:252
  pd = p if keep_multiplier is None else p * keep_multiplier.astype(p.dtype)  # EA:262
  out = np.matmul(pd, v)                                           # EA:265
  out = np.reshape(out, (-1, out.shape[-1]))
  lse = np.reshape(dots_logsumexp, (-1,))
  return out, lse, dict(p=p, k=k, v=v, q=q, keep=keep_multiplier)


# --------------------------------------------------------------------------------------------
# One unit = one (example, head) pair: EA:1918-1997 forward, SURVEY Appendix B backward
# --------------------------------------------------------------------------------------------
@dataclass
class UnitResult:
  out: np.ndarray
  buckets: np.ndarray
  sticker: np.ndarray
  undo_sort: np.ndarray
  q: np.ndarray
  v: np.ndarray
  o: np.ndarray            # combined per-head output before w_o, (L, dv)
  o_rounds: np.ndarray     # un-sorted per-round outputs (nh*L, dv)
  logits: np.ndarray       # un-sorted per-round log-sum-exp (nh*L,)
  cache: dict = field(default_factory=dict)


def forward_unit(cfg: LSHConfig, x, w_q, w_v, w_o, *, buckets=None, rotations=None, mask=None,
                 attn_keep=None, out_keep=None, dtype=np.float64, hash_q=None) -> UnitResult:
  """EA:1918-1997 `forward_unbatched` for one (example, head).

  Either `buckets` (update_state=False, EA:1939-1941) or `rotations` (update_state=True,
  EA:1926-1937) must be given.  `hash_q` optionally overrides the fp32 vectors that get hashed
  (used to feed a device-computed q to the bit-exact bucket check).
  """
  x = np.asarray(x, dtype)
  w_q, w_v, w_o = (np.asarray(w, dtype) for w in (w_q, w_v, w_o))
  seqlen = x.shape[0]
  q = np.matmul(x, w_q)                                            # EA:1923
  v = np.matmul(x, w_v)                                            # EA:1924
  if buckets is None:
    hq = q.astype(np.float32) if hash_q is None else hash_q
    buckets = hash_vectors(cfg, hq, rotations, mask)               # EA:1929
  else:
    buckets = np.asarray(buckets, np.int32)[:cfg.n_hashes * seqlen]  # EA:1941
  assert int(buckets.shape[0]) == cfg.n_hashes * seqlen            # EA:1944
  sticker, undo_sort = sort_buckets(buckets, seqlen)               # EA:1946-1956
  st = sticker % seqlen                                            # EA:1958
  sq = np.take(q, st, axis=0)                                      # EA:1959
  sv = np.take(v, st, axis=0)                                      # EA:1960
  mask_fn = lambda d, qi, ki: mask_self_attention(                 # EA:1962-1963
      d, qi, ki, causal=cfg.causal, exclude_self=True, masked=cfg.masked)
  q_info = st
  assert (mask is not None) == cfg.masked                          # EA:1966
  kv_info = None
  if cfg.masked:
    smask = np.take(np.asarray(mask, bool), st, axis=0)            # EA:1970
    kv_info = q_info * np.where(smask, 1, -1).astype(np.int32)     # EA:1972
  so, slogits, acache = attend(sq, sv, cfg.chunk_len, cfg.n_chunks_before, cfg.n_chunks_after,
                               mask_fn, q_info, kv_info, keep_multiplier=attn_keep)  # EA:1974
  o_rounds = np.take(so, undo_sort, axis=0)                        # EA:1985
  logits = np.take(slogits, undo_sort, axis=0)                     # EA:1986 (≡ sort by sticker)
  probs = None
  if cfg.n_hashes > 1:                                             # EA:1988-1992
    o3 = np.reshape(o_rounds, (cfg.n_hashes, seqlen, o_rounds.shape[-1]))
    l3 = np.reshape(logits, (cfg.n_hashes, seqlen, 1))
    probs = np.exp(l3 - logsumexp(l3, axis=0, keepdims=True))
    o = np.sum(o3 * probs, axis=0)
  else:
    o = o_rounds
  out = np.matmul(o, w_o)                                          # EA:1995
  if out_keep is not None:                                         # EA:1996, 271-280
    out = out * out_keep.astype(out.dtype)
  cache = dict(x=x, w_q=w_q, w_v=w_v, w_o=w_o, sq=sq, sv=sv, so=so, probs=probs, attend=acache,
               out_keep=out_keep, seqlen=seqlen)
  return UnitResult(out=out, buckets=buckets, sticker=sticker, undo_sort=undo_sort, q=q, v=v, o=o,
                    o_rounds=o_rounds, logits=logits, cache=cache)


def backward_unit(cfg: LSHConfig, res: UnitResult, dout):
  """VJP of `forward_unit` with buckets fixed (EA:2413-2421 = jax.vjp; formulas: SURVEY App. B).

  Returns (dx, dw_q, dw_v, dw_o).
  """
  c = res.cache
  dtype = c['x'].dtype
  dout = np.asarray(dout, dtype)
  if c['out_keep'] is not None:
    dout = dout * c['out_keep'].astype(dtype)
  seqlen, nh, cl = c['seqlen'], cfg.n_hashes, cfg.chunk_len
  dq_, dv_ = cfg.d_qk, cfg.d_v
  # B1
  do = dout @ c['w_o'].T
  dw_o = res.o.T @ dout
  # B2 combine
  if nh > 1:
    probs = c['probs']                                             # (nh, L, 1)
    o3 = res.o_rounds.reshape(nh, seqlen, dv_)
    do_h = probs * do[None]                                        # (nh, L, dv)
    dlogit = probs[..., 0] * (np.sum(do[None] * o3, -1) - np.sum(do * res.o, -1)[None])
    do_flat = do_h.reshape(nh * seqlen, dv_)
    dlogit_flat = dlogit.reshape(nh * seqlen)
  else:
    do_flat, dlogit_flat = do, np.zeros(seqlen, dtype)
  # B3 re-sort = gather by the inverse permutation of undo_sort, i.e. by sticker (EA:293-297)
  dso = np.take(do_flat, res.sticker, axis=0)
  dlse = np.take(dlogit_flat, res.sticker, axis=0)
  # B4 per chunk
  a = c['attend']
  p, k, v, q, keep = a['p'], a['k'], a['v'], a['q'], a['keep']     # p (nc, C, W); k,v (nc, W, d)
  nc = p.shape[0]
  dso_c = dso.reshape(nc, cl, dv_)
  dlse_c = dlse.reshape(nc, cl, 1)
  dpd = dso_c @ np.swapaxes(v, -1, -2)                             # cotangent of (p*keep)
  pd = p if keep is None else p * keep.astype(dtype)
  dp = dpd if keep is None else dpd * keep.astype(dtype)
  delta = np.sum(p * dp, -1, keepdims=True)
  ds = p * (dp - delta + dlse_c)
  dq_query = ds @ k                                                # (nc, C, dq)
  dk_win = np.swapaxes(ds, -1, -2) @ q                             # (nc, W, dq)
  dv_win = np.swapaxes(pd, -1, -2) @ dso_c                         # (nc, W, dv)

  def un_look_adjacent(dwin, d):
    """Transpose of look_adjacent (EA:136-142): window slot j of chunk c came from chunk c+i."""
    acc = np.zeros((nc, cl, d), dtype)
    j = 0
    for i in range(-cfg.n_chunks_before, cfg.n_chunks_after + 1):
      part = dwin[:, j * cl:(j + 1) * cl]                          # contributions to chunk (c+i)%nc
      acc += np.roll(part, i, axis=0)
      j += 1
    return acc
  dk = un_look_adjacent(dk_win, dq_)
  dsv = un_look_adjacent(dv_win, dv_).reshape(nh * seqlen, dv_)
  # B5 key-side normalisation VJP (k = sq / sqrt(mean(sq^2)+eps) / sqrt(dq))
  sq_c = q                                                         # (nc, C, dq) un-normalised
  r = np.sqrt(np.mean(sq_c ** 2, -1, keepdims=True) + dtype.type(1e-6))
  sdq = np.sqrt(dtype.type(dq_))
  dq_key = dk / (r * sdq) - sq_c * np.sum(dk * sq_c, -1, keepdims=True) / (dq_ * r ** 3 * sdq)
  dsq = (dq_query + dq_key).reshape(nh * seqlen, dq_)
  # B6 un-sort and sum the nh copies of each token
  dq_tok = np.take(dsq, res.undo_sort, axis=0).reshape(nh, seqlen, dq_).sum(0)
  dv_tok = np.take(dsv, res.undo_sort, axis=0).reshape(nh, seqlen, dv_).sum(0)
  # B7
  x = c['x']
  dw_q = x.T @ dq_tok
  dw_v = x.T @ dv_tok
  dx = dq_tok @ c['w_q'].T + dv_tok @ c['w_v'].T
  return dx, dw_q, dw_v, dw_o


# --------------------------------------------------------------------------------------------
# Batched driver: EA:2261-2561 (`forward_and_or_backward`, n_parallel_heads == 1 loop)
# --------------------------------------------------------------------------------------------
def forward_and_or_backward(cfg: LSHConfig, x, weights, buckets=None, rotations=None, mask=None,
                            output_grad=None, compute_output=True, update_state=True,
                            attn_keep=None, out_keep=None, dtype=np.float64, hash_q=None):
  """Returns (output, new_buckets, inputs_grad, weights_grad) like EA:2283-2288.

  x (B, L, D); weights = (w_q (H,D,dq), w_v (H,D,dv), w_o (H,dv,D)); buckets (B*H, nh*Lb) when
  update_state is False; rotations (B*H, dq, nh, R) when update_state is True; mask (B, L) bool;
  hash_q optional (B*H, L, dq) fp32 vectors to hash instead of the oracle's own q.
  """
  w_q, w_v, w_o = weights
  bsz, seqlen, d_model = x.shape
  nheads = cfg.n_heads
  out = np.zeros((bsz, seqlen, d_model), dtype) if compute_output else None
  new_b = None
  if update_state:
    length = cfg.n_hashes * (cfg.max_length_for_buckets or seqlen)
    new_b = np.zeros((bsz * nheads, length), np.int32)
  dx = dws = None
  if output_grad is not None:
    dx = np.zeros((bsz, seqlen, d_model), dtype)
    dws = [np.zeros(w.shape, dtype) for w in weights]
  for idx in range(bsz * nheads):                                  # EA:2403-2432
    b, h = idx // nheads, idx % nheads
    res = forward_unit(
        cfg, x[b], w_q[h], w_v[h], w_o[h],
        buckets=None if update_state else buckets[idx],
        rotations=rotations[idx] if update_state else None,
        mask=None if mask is None else mask[b], attn_keep=attn_keep, out_keep=out_keep,
        dtype=dtype, hash_q=None if hash_q is None else hash_q[idx])
    if compute_output:
      out[b] += res.out                                            # EA:2426
    if update_state:
      new_b[idx, :res.buckets.shape[0]] = res.buckets              # EA:1930-1937, 2428
    if output_grad is not None:
      g = backward_unit(cfg, res, output_grad[b])                  # EA:2418-2421
      dx[b] += g[0]                                                # EA:2430
      for acc, gi in zip(dws, g[1:]):
        acc[h] += gi                                               # EA:2431
  return out, new_b, dx, (None if dws is None else tuple(dws))


def init_weights(n_heads, d_model, d_qk, d_v, seed=1):
  """EA:1801-1808, 1845-1868 shapes and the Glorot-uniform limit (values from NumPy, not threefry)."""
  rng = np.random.default_rng(seed)
  def ki(shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1] * n_heads))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)
  w_q = np.stack([ki((d_model, d_qk)) for _ in range(n_heads)])
  w_v = np.stack([ki((d_model, d_v)) for _ in range(n_heads)])
  w_o = np.stack([ki((d_model, d_v)).T for _ in range(n_heads)])
  return w_q, w_v, np.ascontiguousarray(w_o)


# ---- the reversible block around the layer (trax/layers/reversible.py:244-412, layers/normalization.py:121-142) -------
def layernorm(x, scale, bias, epsilon=1e-6):
  """normalization.py:129-136 (center=True), fp64."""
  x = np.asarray(x, np.float64)
  mean = x.mean(axis=-1, keepdims=True)
  centered = x - mean
  variance = (centered * centered).mean(axis=-1, keepdims=True)
  return centered / np.sqrt(variance + epsilon) * scale + bias


def layernorm_vjp(x, scale, dz, epsilon=1e-6):
  """VJP of `layernorm` at x for cotangent dz: (dx, d_scale, d_bias) — what fastmath.vjp(call_compute_residual) returns
  at reversible.py:352-353 / 384-385."""
  x, dz = np.asarray(x, np.float64), np.asarray(dz, np.float64)
  mean = x.mean(axis=-1, keepdims=True)
  centered = x - mean
  rstd = 1.0 / np.sqrt((centered * centered).mean(axis=-1, keepdims=True) + epsilon)
  norm = centered * rstd
  g = dz * scale
  dx = rstd * (g - g.mean(axis=-1, keepdims=True) - norm * (g * norm).mean(axis=-1, keepdims=True))
  red = tuple(range(x.ndim - 1))
  return dx, (dz * norm).sum(axis=red), dz.sum(axis=red)


def reversible_half_forward(cfg: LSHConfig, x1, x2, ln_weights, attn_weights, rotations, epsilon=1e-6):
  """ReversibleHalfResidual(LayerNorm(), attention_layer=LSHSelfAttention).forward (reversible.py:296-321):
  returns ((y1, x2), buckets)."""
  z = layernorm(x2, ln_weights[0], ln_weights[1], epsilon)
  res, buckets, _, _ = forward_and_or_backward(cfg, z, attn_weights, rotations=rotations, update_state=True)
  return (np.asarray(x1, np.float64) + res, x2), buckets


def reversible_half_reverse_and_grad(cfg: LSHConfig, y1, x2, ct_y1, ct_x2, ln_weights, attn_weights, buckets, epsilon=1e-6):
  """reverse_and_grad (reversible.py:326-412): returns ((x1, x2), ((ct_y1, ct_x2'), ((d_scale, d_bias), attn dW)))."""
  z = layernorm(x2, ln_weights[0], ln_weights[1], epsilon)
  res, _, dz, dw = forward_and_or_backward(cfg, z, attn_weights, buckets=buckets, output_grad=ct_y1, update_state=False)
  dx2, d_scale, d_bias = layernorm_vjp(x2, ln_weights[0], dz, epsilon)
  return (np.asarray(y1, np.float64) - res, x2), ((ct_y1, np.asarray(ct_x2, np.float64) + dx2), ((d_scale, d_bias), dw))


# ---- PureLSHSelfAttentionWrapper (EA:3493-3620) with _ProjectAndSplitHeads weights_format='model' (EA:3294-3312, 3360-3372) ---
def rotary(x):
  """research/rotary_positional_embedding.py:27-44 on (B, L, d), fp64."""
  x = np.asarray(x, np.float64)
  _, l, d = x.shape
  inv_freq = np.exp(np.arange(0, d, 2) * -(np.log(10000.0) / d))
  freqs = np.arange(l)[:, None] * inv_freq[None, :]
  emb = np.concatenate((freqs, freqs), axis=-1)
  half = d // 2
  rot_half = np.concatenate((-x[..., half:], x[..., :half]), axis=-1)
  return x * np.cos(emb) + rot_half * np.sin(emb)


def rotary_vjp(g):
  """VJP of `rotary` (linear in x): g cos + rotate_half^T(g sin), rotate_half^T(u) = (u2, -u1)."""
  g = np.asarray(g, np.float64)
  _, l, d = g.shape
  inv_freq = np.exp(np.arange(0, d, 2) * -(np.log(10000.0) / d))
  freqs = np.arange(l)[:, None] * inv_freq[None, :]
  emb = np.concatenate((freqs, freqs), axis=-1)
  u = g * np.sin(emb)
  half = d // 2
  return g * np.cos(emb) + np.concatenate((u[..., half:], -u[..., :half]), axis=-1)


def _dense(x, w):
  """core.py:79-98: w is `(kernel, bias)` or a bare kernel."""
  if isinstance(w, (tuple, list)):
    return np.matmul(x, np.asarray(w[0], np.float64)) + np.asarray(w[1], np.float64)
  return np.matmul(x, np.asarray(w, np.float64))


def _dense_vjp(x, w, dy):
  """(dx, dw) with dw shaped like w."""
  kernel = np.asarray(w[0] if isinstance(w, (tuple, list)) else w, np.float64)
  x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, dy.shape[-1])
  dk = np.matmul(x2.T, dy2)
  dx = np.matmul(dy, kernel.T)
  return dx, ((dk, dy2.sum(axis=0)) if isinstance(w, (tuple, list)) else dk)


def split_heads(x, n_heads):
  """attention.py:347-364: (B, L, H d) -> (B H, L, d)."""
  b, l, f = x.shape
  return x.reshape(b, l, n_heads, f // n_heads).transpose(0, 2, 1, 3).reshape(b * n_heads, l, f // n_heads)


def merge_heads(x, n_heads):
  """attention.py:369-388: (B H, L, d) -> (B, L, H d)."""
  bh, l, d = x.shape
  return x.reshape(bh // n_heads, n_heads, l, d).transpose(0, 2, 1, 3).reshape(bh // n_heads, l, n_heads * d)


def pure_lsh_wrapper(cfg: LSHConfig, x, qkv_weights, dense_weights, *, buckets=None, rotations=None, mask=None,
                     output_grad=None, rotary_position_emb=False):
  """`PureLSHSelfAttentionWrapper` = Serial(_ProjectAndSplitHeads('model'), PureLSHSelfAttention, MergeHeads, Dense)
  (EA:3512-3540) and its `forward_and_or_backward` (EA:3542-3620): x (B, L, d_model); `qkv_weights` holds two
  (qk, v: EA:3360-3372) or three (q, k, v with qk = (q + k)/2: EA:3294-3312) Dense weights; either `rotations`
  (B H, dq, nh, R) or stored `buckets` (B H, nh L).  Returns (out, buckets, dx, (d_qkv_weights, d_dense_weights))."""
  x = np.asarray(x, np.float64)
  H = cfg.n_heads
  proj = [_dense(x, w) for w in qkv_weights]
  n_rot = len(proj) - 1                                            # q (and k) are rotated, v never (EA:3303-3304, 3363-3365)
  rot = [rotary(p) if (rotary_position_emb and i < n_rot) else p for i, p in enumerate(proj)]
  qk = (rot[0] + rot[1]) / 2.0 if len(proj) == 3 else rot[0]       # EA:3306
  v = rot[-1]
  qk_h, v_h = split_heads(qk, H), split_heads(v, H)                # EA:3309-3311
  eye_q = np.concatenate([np.eye(cfg.d_qk), np.zeros((cfg.d_v, cfg.d_qk))], axis=0)
  eye_v = np.concatenate([np.zeros((cfg.d_qk, cfg.d_v)), np.eye(cfg.d_v)], axis=0)
  units, new_buckets = [], []
  for u in range(qk_h.shape[0]):                                   # PureLSH forward_unbatched (EA:2739-2826) per unit
    xu = np.concatenate([qk_h[u], v_h[u]], axis=1)
    r = forward_unit(cfg, xu, eye_q, eye_v, np.eye(cfg.d_v),
                     buckets=None if buckets is None else buckets[u],
                     rotations=None if rotations is None else rotations[u],
                     mask=None if mask is None else mask[u // H])
    units.append(r)
    new_buckets.append(r.buckets)
  merged = merge_heads(np.stack([r.out for r in units]), H)        # EA:3531
  out = _dense(merged, dense_weights)                              # EA:3533
  if output_grad is None:
    return out, np.stack(new_buckets), None, None
  d_merged, d_dense = _dense_vjp(merged, dense_weights, np.asarray(output_grad, np.float64))   # EA:3590-3592
  d_heads = split_heads(d_merged, H)                               # EA:3596-3597 (vjp of MergeHeads)
  g = np.stack([backward_unit(cfg, units[u], d_heads[u])[0] for u in range(len(units))])      # EA:3600-3602
  d_qk, d_v = merge_heads(g[..., :cfg.d_qk], H), merge_heads(g[..., cfg.d_qk:], H)
  d_rot = [d_qk / 2.0, d_qk / 2.0, d_v] if len(proj) == 3 else [d_qk, d_v]
  d_proj = [rotary_vjp(d) if (rotary_position_emb and i < n_rot) else d for i, d in enumerate(d_rot)]
  dx = np.zeros_like(x)
  d_qkv = []
  for w, d in zip(qkv_weights, d_proj):                            # EA:3605
    dxi, dw = _dense_vjp(x, w, d)
    dx += dxi
    d_qkv.append(dw)
  return out, np.stack(new_buckets), dx, (tuple(d_qkv), d_dense)
