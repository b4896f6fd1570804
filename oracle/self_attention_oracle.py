"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's chunked local attention,
`trax.layers.research.efficient_attention.SelfAttention` (EA:936-1726, train path), the layer ReformerLM interleaves with
the LSH layer (`reformer_enwik8.gin:23-28`: 3 of 4 layers, chunk_len 128, n_chunks_before 1) — SURVEY.md §8f rank 4.  It is
the same `attend` (EA:163-268) without hashing and sorting; by default (`share_qk=False`) keys have their own projection,
are NOT length-normalised and a token may attend to itself (EA:1175-1178 `exclude_self=self._share_qk`).

Written before any kernel for it exists (oracle first): PINNED against the reference's own code run under its NumPy
backend (`oracle/ref_live_sweep.py`, see `ref_live.py`): float64 outputs to 1e-11 and the analytic VJP below against
central differences of the reference's forward.  No product code imports this file.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from oracle.lsh_oracle import length_normalized, logsumexp, look_adjacent, mask_self_attention


@dataclass
class SelfAttentionConfig:
  n_heads: int = 2
  d_qk: int = 64
  d_v: int = 64
  share_qk: bool = False
  causal: bool = False
  masked: bool = False
  chunk_len: Optional[int] = None
  n_chunks_before: int = 0
  n_chunks_after: int = 0


def forward_unit(cfg: SelfAttentionConfig, x, weights, mask=None):
  """EA:1136-1197 `forward_unbatched` for one (example, head): weights = (w_q, w_v, w_o) if share_qk else
  (w_q, w_k, w_v, w_o) (EA:1112-1128, no bias).  Returns (out, cache)."""
  x = np.asarray(x, np.float64)
  if cfg.share_qk:
    w_q, w_v, w_o = (np.asarray(w, np.float64) for w in weights)
    w_k = None
  else:
    w_q, w_k, w_v, w_o = (np.asarray(w, np.float64) for w in weights)
  seqlen = x.shape[0]
  q = x @ w_q                                                       # EA:1159
  k = None if cfg.share_qk else x @ w_k                             # EA:1160-1162
  v = x @ w_v                                                       # EA:1163
  q_info = np.arange(seqlen, dtype=np.int32)                        # EA:1180
  kv_info = q_info
  assert (mask is not None) == cfg.masked                           # EA:1182
  if cfg.masked:
    kv_info = kv_info * np.where(np.asarray(mask, bool), 1, -1).astype(np.int32)   # EA:1185-1186
  cl = cfg.chunk_len or seqlen                                      # chunk_len None: one chunk (EA:208-226 skip the reshape)
  if cfg.chunk_len is None:
    assert cfg.n_chunks_before == 0 and cfg.n_chunks_after == 0     # EA:236
  q_info = q_info + 1                                               # EA:201
  kv_info = kv_info + 1                                             # EA:206
  qc = q.reshape(-1, cl, q.shape[-1])                               # EA:210
  q_info_c = q_info.reshape(-1, cl)
  kc = qc if cfg.share_qk else k.reshape(-1, cl, k.shape[-1])      # EA:215 / 224
  kv_info_c = kv_info.reshape(-1, cl)
  vc = v.reshape(-1, cl, v.shape[-1])                               # EA:228
  kn = length_normalized(kc) if cfg.share_qk else kc                # EA:230-231
  kn = kn / np.sqrt(kn.shape[-1])                                   # EA:232
  kw = look_adjacent(kn, cfg.n_chunks_before, cfg.n_chunks_after)   # EA:239-241
  vw = look_adjacent(vc, cfg.n_chunks_before, cfg.n_chunks_after)
  kv_info_w = look_adjacent(kv_info_c, cfg.n_chunks_before, cfg.n_chunks_after)
  dots = qc @ np.swapaxes(kw, -1, -2)                               # EA:244
  dots = mask_self_attention(dots, q_info_c[..., :, None], kv_info_w[..., None, :], causal=cfg.causal,
                             exclude_self=cfg.share_qk, masked=cfg.masked)        # EA:248, 1175-1178
  p = np.exp(dots - logsumexp(dots, axis=-1, keepdims=True))        # EA:251-252
  o = (p @ vw).reshape(seqlen, -1)                                  # EA:265
  out = o @ w_o                                                     # EA:1195
  return out, dict(x=x, w_q=w_q, w_k=w_k, w_v=w_v, w_o=w_o, qc=qc, kc=kc, kw=kw, vw=vw, p=p, o=o, cl=cl)


def backward_unit(cfg: SelfAttentionConfig, cache, dout):
  """VJP of `forward_unit` (what `jax.vjp` returns at EA:1564-1590): (dx, weight grads in the order of `weights`)."""
  c = cache
  dout = np.asarray(dout, np.float64)
  seqlen, cl = c['x'].shape[0], c['cl']
  nc = seqlen // cl
  do = dout @ c['w_o'].T
  dw_o = c['o'].T @ dout
  do_c = do.reshape(nc, cl, -1)
  p = c['p']
  dp = do_c @ np.swapaxes(c['vw'], -1, -2)
  ds = p * (dp - np.sum(p * dp, -1, keepdims=True))
  dq_c = ds @ c['kw']                                               # query side
  dk_win = np.swapaxes(ds, -1, -2) @ c['qc']                        # (nc, W, dq) w.r.t. the scaled (normalised) keys
  dv_win = np.swapaxes(p, -1, -2) @ do_c

  def un_look_adjacent(dwin):
    acc = np.zeros((nc, cl, dwin.shape[-1]))
    for j, i in enumerate(range(-cfg.n_chunks_before, cfg.n_chunks_after + 1)):
      acc += np.roll(dwin[:, j * cl:(j + 1) * cl], i, axis=0)       # window slot j of chunk c came from chunk (c+i) % nc
    return acc
  dkn = un_look_adjacent(dk_win)
  dv_tok = un_look_adjacent(dv_win).reshape(seqlen, -1)
  sdq = np.sqrt(float(c['kc'].shape[-1]))
  x = c['x']
  if cfg.share_qk:                                                  # k = q / sqrt(mean(q^2) + 1e-6) / sqrt(dq)
    qc = c['qc']
    r = np.sqrt(np.mean(qc ** 2, -1, keepdims=True) + 1e-6)
    dq_key = dkn / (r * sdq) - qc * np.sum(dkn * qc, -1, keepdims=True) / (qc.shape[-1] * r ** 3 * sdq)
    dq_tok = (dq_c + dq_key).reshape(seqlen, -1)
    dx = dq_tok @ c['w_q'].T + dv_tok @ c['w_v'].T
    return dx, (x.T @ dq_tok, x.T @ dv_tok, dw_o)
  dq_tok = dq_c.reshape(seqlen, -1)
  dk_tok = (dkn / sdq).reshape(seqlen, -1)
  dx = dq_tok @ c['w_q'].T + dk_tok @ c['w_k'].T + dv_tok @ c['w_v'].T
  return dx, (x.T @ dq_tok, x.T @ dk_tok, x.T @ dv_tok, dw_o)


def forward_and_or_backward(cfg: SelfAttentionConfig, x, weights, mask=None, output_grad=None):
  """Batched driver (EA:1426-1726): heads summed into the output and dx, examples summed into the weight gradients.
  weights are stacked over heads.  Returns (output, inputs_grad, weights_grad)."""
  x = np.asarray(x, np.float64)
  B, H = x.shape[0], cfg.n_heads
  out = np.zeros_like(x)
  dx = None if output_grad is None else np.zeros_like(x)
  dws = None if output_grad is None else [np.zeros(np.shape(w), np.float64) for w in weights]
  for b in range(B):
    for h in range(H):
      o, cache = forward_unit(cfg, x[b], tuple(w[h] for w in weights), None if mask is None else mask[b])
      out[b] += o
      if output_grad is not None:
        g, gw = backward_unit(cfg, cache, output_grad[b])
        dx[b] += g
        for acc, gi in zip(dws, gw):
          acc[h] += gi
  return out, dx, (None if dws is None else tuple(dws))
