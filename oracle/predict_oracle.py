"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's fast-inference path (`mode='predict'`) of
`LSHSelfAttention` (EA:1999-2109 `_incremental_forward_unbatched`, EA:2174-2244 `_use_predict_mem`, driver EA:2127-2170 /
2333-2350, state EA:1833-1841, 1879-1887) and of the chunked local `SelfAttention` (EA:1200-1268, same memory logic) —
SURVEY.md §8 row a15 / §8(f) rank 4.  No product code imports this file.

PINNED against the reference's own code run under its NumPy backend (`oracle/ref_live.py`;
`oracle/ref_live_predict.py` is the sweep, `tests/test_reference_pin.py` runs it): memory contents, `mem_end`, bucket
memory, `buckets_idx` bit for bit and float64 outputs to 1e-11, token by token over sequences long enough to roll the memory
(`predict_drop_len`) several times, after prefixes shorter / equal / longer than the memory.

State layout (stacked over units like the reference's, EA:1829-1841):
  (mem_end int, mem (B, M, D), (buckets int32 (B*H, nh*M), buckets_idx int32 (B*H,)))   LSH (the rng leaf is carried by the
  caller; rotations are an explicit input here as everywhere in the oracle), and (mem_end, mem) for SelfAttention.
Quirks kept on purpose (they are the reference's behaviour):
  * predict mode hashes with `hash_rng` itself, un-split (EA:2014, 2066) — the SAME rotations at every call;
  * `n_buckets=None` resolves from the number of rows hashed: the padded prefix length, but 2 rows (the duplicated query,
    EA:2064) at a single-token step ⇒ 2 buckets;
  * a single-token step attends to the `n_hashes·chunk_len·(1+n_chunks_before)` highest-priority memory slots — same-bucket
    tokens first, then the most recent others, whatever their bucket (the TODO at EA:2098-2099) — with the causal / self
    masks hard-wired (EA:2091-2092);
  * `dynamic_slice_in_dim` / `dynamic_update_slice_in_dim` clamp their start index (jax.lax semantics).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from oracle import lsh_oracle as O
from oracle import self_attention_oracle as SA


@dataclass
class PredictConfig:
  predict_mem_len: int = 2048
  predict_drop_len: int = 256


def init_state(cfg: O.LSHConfig, pcfg: PredictConfig, batch: int, d_model: int, dtype=np.float64):
  """EA:1833-1841, 1883-1887."""
  m = pcfg.predict_mem_len
  return (0, np.zeros((batch, m, d_model), dtype),
          (np.zeros((batch * cfg.n_heads, cfg.n_hashes * m), np.int32), np.zeros((batch * cfg.n_heads,), np.int32)))


def _dynamic_slice(operand, start, size, axis):
  """jax.lax.dynamic_slice_in_dim: the start index is clamped so that the slice fits."""
  start = int(np.clip(int(start), 0, operand.shape[axis] - size))
  return np.take(operand, np.arange(start, start + size), axis=axis)


def _dynamic_update_slice(operand, update, start, axis):
  """jax.lax.dynamic_update_slice_in_dim (clamped start)."""
  start = int(np.clip(int(start), 0, operand.shape[axis] - update.shape[axis]))
  out = np.array(operand, copy=True)
  idx = [slice(None)] * out.ndim
  idx[axis] = slice(start, start + update.shape[axis])
  out[tuple(idx)] = update
  return out


def use_predict_mem(pcfg: PredictConfig, x, mem_end: int, mem):
  """EA:2174-2244.  Returns (inputs, q_start, new_mem, new_mem_end); `inputs` is what the units see."""
  seqlen = x.shape[1]
  m, drop = pcfg.predict_mem_len, pcfg.predict_drop_len
  if seqlen <= drop and seqlen < m:                                  # EA:2179: a few tokens appended
    if mem_end + seqlen > m:                                         # EA:2189-2197 roll_mem
      mem = np.concatenate([mem[:, drop:], np.zeros_like(mem[:, :drop])], axis=1)
      mem_end = mem_end - drop
    if seqlen == 1:
      mem = np.array(mem, copy=True)
      mem[:, mem_end] = x[:, 0]                                      # EA:2201-2202 (index_update wraps like NumPy)
    else:
      mem = _dynamic_update_slice(mem, x, mem_end, axis=1)           # EA:2204-2205
    return mem, mem_end, mem, mem_end + seqlen                       # EA:2207
  assert seqlen > drop or seqlen == m                                # EA:2209
  if seqlen == m:                                                    # EA:2218-2231
    new_mem = x
  elif seqlen > m:
    new_mem = x[:, -m:]
  else:
    new_mem = np.concatenate([x, np.zeros(x.shape[:1] + (m - seqlen,) + x.shape[2:], x.dtype)], axis=1)
  if mem_end != 0:                                                   # EA:2236-2243: only valid at the start of a sequence
    x = x * np.nan
  return x, 0, new_mem, min(seqlen, m)


def incremental_forward_unit(cfg: O.LSHConfig, pcfg: PredictConfig, x, q_start: int, q_len: int, w_q, w_v, w_o, buckets,
                             buckets_idx: int, rotations_fn, new_ids=None):
  """EA:1999-2109 for one (example, head).  `rotations_fn(n_rows)` returns the (dq, nh, R) rotations `hash_vectors` would
  draw for `n_rows` hashed rows (the shape depends on n_rows only through `n_buckets=None`, EA:1893-1902).
  `new_ids` (GPU parity tests): bucket ids to use instead of hashing here — (nh * padded_len,) for a prefix, (nh,) for one
  token — so that a comparison downstream of the hash does not hinge on fp32-vs-bf16 argmax near-ties.
  Returns (out (q_len, D), new_buckets, new_buckets_idx)."""
  nh, m = cfg.n_hashes, pcfg.predict_mem_len
  x = np.asarray(x, np.float64)
  w_q, w_v, w_o = (np.asarray(w, np.float64) for w in (w_q, w_v, w_o))
  if q_len > 1:                                                      # EA:2004-2030: a prefix, at the start only
    if x.shape[0] % cfg.chunk_len:
      x_padded = np.pad(x, ((0, cfg.chunk_len - x.shape[0] % cfg.chunk_len), (0, 0)), mode='constant')
    else:
      x_padded = x
    q = x_padded @ w_q
    if new_ids is None:
      buckets_update = O.hash_vectors(cfg, q.astype(np.float32), rotations_fn(x_padded.shape[0]))    # EA:2014
    else:
      buckets_update = np.asarray(new_ids, np.int32).reshape(-1)
    res = O.forward_unit(cfg, x_padded, w_q, w_v, w_o, buckets=buckets_update)                        # EA:2016-2018
    out = res.out[:q_len]
    buckets = np.reshape(buckets, (nh, -1))
    buckets_update = np.reshape(buckets_update, (nh, -1))[:, :q_len]
    if q_len > m:
      buckets_update = buckets_update[:, -m:]
    buckets = _dynamic_update_slice(buckets, buckets_update, q_start, axis=1)                         # EA:2026-2027
    return out, np.reshape(buckets, (-1,)), buckets_idx + q_len
  assert q_len == 1
  if buckets_idx > q_start:                                          # EA:2036-2053 roll_buckets
    b2 = np.reshape(buckets, (nh, -1))
    b2 = np.concatenate([b2, np.zeros((nh, pcfg.predict_drop_len), b2.dtype)], axis=1)
    buckets = np.reshape(_dynamic_slice(b2, buckets_idx - q_start, m, axis=1), (-1,))
  q = np.concatenate([x[q_start:q_start + 1]] * 2, 0) @ w_q          # EA:2064 (the duplicated row)
  if new_ids is None:
    q_buckets = O.hash_vectors(cfg, q.astype(np.float32), rotations_fn(2))                            # EA:2066
    q_buckets = np.reshape(q_buckets, (nh, 2))[:, :1]
  else:
    q_buckets = np.asarray(new_ids, np.int32).reshape(nh, 1)
  unflattened = _dynamic_update_slice(np.reshape(buckets, (nh, -1)), q_buckets, q_start, axis=1)      # EA:2069-2071
  buckets = np.reshape(unflattened, (-1,))
  is_valid_target = np.any(unflattened == q_buckets, axis=0)         # EA:2073
  seqlen = x.shape[0]
  ar = np.arange(seqlen, dtype=np.int32)
  kv_priorities = np.where(ar > (q_start + q_len), -(seqlen + ar), ar)                                # EA:2078-2080
  kv_priorities = kv_priorities + seqlen * is_valid_target.astype(np.int32)
  kv_indices = np.argsort(kv_priorities, kind='stable').astype(np.int32)                              # EA:2082
  kv_indices = kv_indices[-nh * cfg.chunk_len * (1 + cfg.n_chunks_before):]                           # EA:2083-2084
  assert cfg.n_chunks_after == 0
  x_attend_to = x[kv_indices]
  k = O.length_normalized(x_attend_to @ w_q)                         # EA:2088
  v = x_attend_to @ w_v
  k = k / np.sqrt(k.shape[-1])                                       # attend, EA:232
  dots = q @ k.T                                                     # (2, K)
  q_info = (q_start + np.arange(q_len, dtype=np.int32)) + 1          # EA:2093, 201
  kv_info = kv_indices + 1
  dots = O.mask_self_attention(dots, q_info[:, None], kv_info[None, :], causal=True, exclude_self=True, masked=True)
  p = np.exp(dots - O.logsumexp(dots, axis=-1, keepdims=True))
  out = (p @ v) @ w_o
  return out[:1], buckets, q_start + q_len                           # EA:2104-2109


def predict_forward(cfg: O.LSHConfig, pcfg: PredictConfig, x, weights, state, rotations_fn, new_ids_fn=None):
  """One call of the layer in predict mode (EA:2127-2170): x (B, seqlen, D); `rotations_fn(unit, n_rows)` as above;
  `new_ids_fn(unit)` optionally supplies the call's bucket ids (see incremental_forward_unit).
  Returns (output (B, seqlen, D), new_state)."""
  mem_end, mem, (buckets, buckets_idx) = state
  w_q, w_v, w_o = weights
  bsz, seqlen, d_model = x.shape
  inputs, q_start, new_mem, new_mem_end = use_predict_mem(pcfg, np.asarray(x, np.float64), int(mem_end), mem)
  out = np.zeros((bsz, seqlen, d_model))
  nb, ni = np.array(buckets, copy=True), np.array(buckets_idx, copy=True)
  for idx in range(bsz * cfg.n_heads):
    b, h = idx // cfg.n_heads, idx % cfg.n_heads
    o, nb[idx], ni[idx] = incremental_forward_unit(cfg, pcfg, inputs[b], q_start, seqlen, w_q[h], w_v[h], w_o[h], buckets[idx],
                                                   int(buckets_idx[idx]), lambda n, _u=idx: rotations_fn(_u, n),
                                                   None if new_ids_fn is None else new_ids_fn(idx))
    out[b] += o
  return out, (new_mem_end, new_mem, (nb, ni))


# ---- PureLSHSelfAttention (EA:2823-2932, 2955-3033, 3122-3146) ---------------------------------------------------------------
def pure_predict_forward(cfg: O.LSHConfig, pcfg: PredictConfig, qk, v, state, rotations_fn, new_ids_fn=None):
  """One call of the weight-less core in predict mode: inputs qk (B*H, seqlen, d_qk), v (B*H, seqlen, d_v); state
  (mem_end, (qk_mem, v_mem), (buckets, buckets_idx)).  EA:2823-2932 is EA:1999-2109 with the projections taken out — the
  memory holds qk and v themselves — so this is `incremental_forward_unit` on x = [qk | v] with selector weights
  (w_q = [I; 0], w_v = [0; I], w_o = I), one "head" per leading row.  Returns (output (B*H, seqlen, d_v), new_state)."""
  mem_end, (qk_mem, v_mem), (buckets, buckets_idx) = state
  dq, dv = qk.shape[-1], v.shape[-1]
  w_q = np.concatenate([np.eye(dq), np.zeros((dv, dq))], axis=0)
  w_v = np.concatenate([np.zeros((dq, dv)), np.eye(dv)], axis=0)
  w_o = np.eye(dv)
  x = np.concatenate([np.asarray(qk, np.float64), np.asarray(v, np.float64)], axis=-1)
  mem = np.concatenate([qk_mem, v_mem], axis=-1)
  seqlen = x.shape[1]
  inputs, q_start, new_mem, new_mem_end = use_predict_mem(pcfg, x, int(mem_end), mem)
  out = np.zeros((x.shape[0], seqlen, dv))
  nb, ni = np.array(buckets, copy=True), np.array(buckets_idx, copy=True)
  for u in range(x.shape[0]):
    out[u], nb[u], ni[u] = incremental_forward_unit(cfg, pcfg, inputs[u], q_start, seqlen, w_q, w_v, w_o, buckets[u],
                                                    int(buckets_idx[u]), lambda n, _u=u: rotations_fn(_u, n),
                                                    None if new_ids_fn is None else new_ids_fn(u))
  return out, (new_mem_end, (new_mem[..., :dq], new_mem[..., dq:]), (nb, ni))


# ---- SelfAttention (EA:1200-1268) -------------------------------------------------------------------------------------
def self_attention_incremental_unit(cfg: SA.SelfAttentionConfig, x, q_start: int, q_len: int, weights):
  """EA:1200-1268 for one (example, head), no input mask.  Returns out (q_len, D)."""
  x = np.asarray(x, np.float64)
  if cfg.share_qk:
    w_q, w_v, w_o = (np.asarray(w, np.float64) for w in weights)
  else:
    w_q, w_k, w_v, w_o = (np.asarray(w, np.float64) for w in weights)
  q_range = q_start + np.arange(q_len, dtype=np.int32)
  q = x[q_range] @ w_q                                               # EA:1229-1238 (the duplicated row changes nothing here)
  k = O.length_normalized(x @ w_q) if cfg.share_qk else x @ w_k      # EA:1239-1242
  v = x @ w_v
  kv_info = np.arange(k.shape[0], dtype=np.int32) + 1
  q_info = q_range + 1
  k = k / np.sqrt(k.shape[-1])
  if cfg.chunk_len is not None and q_len > cfg.chunk_len:            # EA:1250-1261
    assert q_start == 0 and q_len % cfg.chunk_len == 0
    cl = cfg.chunk_len
    qc, q_info_c = q.reshape(-1, cl, q.shape[-1]), q_info.reshape(-1, cl)
    kw = O.look_adjacent(k.reshape(-1, cl, k.shape[-1]), cfg.n_chunks_before, cfg.n_chunks_after)
    vw = O.look_adjacent(v.reshape(-1, cl, v.shape[-1]), cfg.n_chunks_before, cfg.n_chunks_after)
    kv_info_w = O.look_adjacent(kv_info.reshape(-1, cl), cfg.n_chunks_before, cfg.n_chunks_after)
    dots = qc @ np.swapaxes(kw, -1, -2)
    dots = O.mask_self_attention(dots, q_info_c[..., :, None], kv_info_w[..., None, :], causal=cfg.causal,
                                 exclude_self=cfg.share_qk, masked=cfg.masked)
    p = np.exp(dots - O.logsumexp(dots, axis=-1, keepdims=True))
    o = (p @ vw).reshape(q_len, -1)
  else:                                                              # EA:1262-1267: every memory slot is a key
    dots = q @ k.T
    dots = O.mask_self_attention(dots, q_info[:, None], kv_info[None, :], causal=cfg.causal, exclude_self=cfg.share_qk,
                                 masked=cfg.masked)
    p = np.exp(dots - O.logsumexp(dots, axis=-1, keepdims=True))
    o = p @ v
  return o @ w_o


def self_attention_predict_forward(cfg: SA.SelfAttentionConfig, pcfg: PredictConfig, x, weights, state):
  """EA:1300-1336 in predict mode; state = (mem_end, mem).  Returns (output, new_state)."""
  mem_end, mem = state
  bsz, seqlen, d_model = x.shape
  inputs, q_start, new_mem, new_mem_end = use_predict_mem(pcfg, np.asarray(x, np.float64), int(mem_end), mem)
  out = np.zeros((bsz, seqlen, d_model))
  for idx in range(bsz * cfg.n_heads):
    b, h = idx // cfg.n_heads, idx % cfg.n_heads
    out[b] += self_attention_incremental_unit(cfg, inputs[b], q_start, seqlen, tuple(w[h] for w in weights))
  return out, (new_mem_end, new_mem)
