"""ORACLE — TEST INFRASTRUCTURE ONLY.  Second, independent CPU restatement of
`LSHSelfAttention.forward_unbatched` (EA:1918-1997) written with differentiable torch ops (float64
on CPU) so that `torch.autograd` plays the role `jax.vjp` plays in the reference (EA:2399, 2418).
Used by tests/test_oracle.py to check the analytic backward of `lsh_oracle.backward_unit`.
The permutation (sticker) is an input: the reference stops gradients through it (EA:1948, 1954-1956).
"""
import torch


def forward_unit_torch(x, w_q, w_v, w_o, sticker, *, seqlen, chunk_len, n_hashes, n_chunks_before,
                       n_chunks_after, causal, masked=False, mask=None, attn_keep=None,
                       out_keep=None):
  sticker = torch.as_tensor(sticker, dtype=torch.long)
  undo_sort = torch.argsort(sticker)
  q = x @ w_q                                                      # EA:1923
  v = x @ w_v                                                      # EA:1924
  st = sticker % seqlen                                            # EA:1958
  sq, sv = q[st], v[st]                                            # EA:1959-1960
  q_info = st + 1                                                  # EA:201
  if masked:
    sm = torch.as_tensor(mask, dtype=torch.bool)[st]
    kv_info = st * torch.where(sm, 1, -1) + 1                      # EA:1972, 206
  else:
    kv_info = q_info
  d = sq.shape[-1]
  cq = sq.reshape(-1, chunk_len, d)                                # EA:210
  qi = q_info.reshape(-1, chunk_len)
  ki = kv_info.reshape(-1, chunk_len)
  k = cq / torch.sqrt((cq ** 2).mean(-1, keepdim=True) + 1e-6)     # EA:54-57, 230
  k = k / (d ** 0.5)                                               # EA:231
  cv = sv.reshape(-1, chunk_len, sv.shape[-1])

  def look(t):                                                     # EA:122-142
    if n_chunks_before == 0 and n_chunks_after == 0:
      return t
    return torch.cat([t if i == 0 else torch.roll(t, -i, 0)
                      for i in range(-n_chunks_before, n_chunks_after + 1)], dim=1)
  k, cv, ki = look(k), look(cv), look(ki)
  dots = cq @ k.transpose(-1, -2)                                  # EA:244
  qf, kf = qi[:, :, None].to(dots.dtype), ki[:, None, :].to(dots.dtype)   # (float32 inputs: the timing path of bench.py)
  if causal:
    dots = dots - 1e9 * (qf < kf).to(dots.dtype)                   # EA:150-152
  dots = dots - 1e5 * (qf == kf).to(dots.dtype)                    # EA:153-155
  if masked:
    dots = dots - 1e9 * (kf < 0).to(dots.dtype)                    # EA:156-159
  lse = torch.logsumexp(dots, -1, keepdim=True)                    # EA:251
  p = torch.exp(dots - lse)                                        # EA:252
  if attn_keep is not None:
    p = p * torch.as_tensor(attn_keep, dtype=p.dtype)              # EA:262
  so = (p @ cv).reshape(-1, cv.shape[-1])                          # EA:265-266
  slogits = lse.reshape(-1)
  o = so[undo_sort]                                                # EA:1985
  logits = slogits[undo_sort]                                      # EA:1986
  if n_hashes > 1:                                                 # EA:1988-1992
    o = o.reshape(n_hashes, seqlen, -1)
    logits = logits.reshape(n_hashes, seqlen, 1)
    probs = torch.exp(logits - torch.logsumexp(logits, 0, keepdim=True))
    o = (o * probs).sum(0)
  out = o @ w_o                                                    # EA:1995
  if out_keep is not None:
    out = out * torch.as_tensor(out_keep, dtype=out.dtype)         # EA:1996
  return out
