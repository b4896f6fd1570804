/*
 * lsh_attn.h — C ABI of the B200-native Reformer LSH-attention path (liblsh_attn_b200.so).
 *
 * The reference (google/trax) has NO native code and no FFI for this path: the hot path is
 * `trax/layers/research/efficient_attention.py` ("EA") class LSHSelfAttention, pure Python over
 * jax.numpy.  These entry points are what a `jax.ffi` custom call (or the ctypes binding shipped
 * in trax_b200/_lib.py) binds in place of the XLA ops that EA lowers to; each declaration cites the
 * reference lines it replaces.  See INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *   - plain C, device pointers only (unless a name ends in _host), no allocation inside, no
 *     synchronisation inside: every call is asynchronous on the caller-supplied CUDA stream
 *     (`stream` is a cudaStream_t passed as void*).
 *   - scratch memory comes from the caller: query `*_workspace_bytes`, pass `ws`/`ws_bytes`.
 *   - return value 0 = OK; non-zero = error, message in lsh_attn_last_error() (thread-local).
 *   - a "unit" is one (example b, head h) pair, unit index u = b*H + h (EA:2406-2407).
 *   - N = nh*L is the sorted length of one unit; "ticker" index = round*L + position (EA:1946).
 *
 * Device layouts (all row-major, innermost last)
 *   x, out, dout, dx : (B, L, D)            act_dtype (f32 or bf16)
 *   w_q              : (H, D, dq) f32       w_v : (H, D, dv) f32       w_o : (H, dv, D) f32   (EA:1845-1868)
 *   wqv (packed)     : (D, H, dq+dv) bf16   — columns [h][0:dq]=w_q[h], [h][dq:dq+dv]=w_v[h]
 *                      (dims.separate_k: (D, H, dq+dv+dq), columns [h][dq+dv:] = w_k[h]; likewise qv, dqv below)
 *   wo  (packed)     : (H*dv, D) bf16
 *   qv               : (B, L, H, dq+dv) bf16 — q then v of head h for token (b,t)
 *   rotations        : (B*H, dq, nh, R) f32 (EA:91, one draw per unit because the hash rng lives in
 *                      per-unit state, EA:1927-1929); R = sum(factors)/2
 *   mask             : (B, L) uint8, 1 = valid token (only when dims.masked)
 *   buckets          : (B*H, buckets_stride) int32, first nh*L entries of each row used (EA:1913-1916, 1930-1941)
 *   sticker, undo    : (B*H, nh*L) int32 (EA:1951-1953)
 *   o_rounds         : (B*H, nh*L, dv) bf16 in TICKER order (= EA:1985 `o` before the combine)
 *   logits           : (B*H, nh*L) f32 in ticker order (EA:1986)
 *   o_comb           : (B, L, H, dv) bf16 (EA:1992 `o`, heads side by side so that w_o sums heads, EA:2426)
 */
#ifndef LSH_ATTN_H_
#define LSH_ATTN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 4: + lsh_layer_fwd_res / lsh_layer_bwd_res (residual epilogue), lsh_pack_heads / lsh_unpack_heads, lsh_layernorm_fwd_bf16,
 *    LshAttnDims.x_bf16 (was reserved[0]; 0 keeps the v3 behaviour).  v3 entry points are unchanged. */
/* 5: + lsh_predict_step / lsh_predict_attend (+ their *_workspace_bytes): fast inference, mode='predict'.  v4 entry points
 *    are unchanged. */
#define LSH_ATTN_ABI_VERSION 5

enum { LSH_DTYPE_F32 = 0, LSH_DTYPE_BF16 = 1 };

/* Hyper-parameters of one call: the constructor arguments of EA:1732-1748 that the train path uses,
 * plus the input shape.  `factors` is the hash factor list after EA:1893-1902 resolved n_buckets. */
typedef struct LshAttnDims {
  int32_t B, H, L, D;      /* batch, n_heads, seqlen, d_model */
  int32_t dq, dv;          /* d_qk, d_v */
  int32_t C, nb, na, nh;   /* chunk_len, n_chunks_before, n_chunks_after, n_hashes */
  int32_t n_factors;       /* 1..4 */
  int32_t factors[4];      /* even ints; n_buckets = prod(factors) (+1 when masked, EA:1908) */
  int32_t causal, masked;  /* bools */
  int32_t act_dtype;       /* LSH_DTYPE_* of x/out/dout/dx */
  int32_t separate_k;      /* 0: shared-QK (LSHSelfAttention; SelfAttention(share_qk=True)): keys are the length-normalised
                            * queries, a token does not attend to itself (EA:229-231, 153-155).
                            * 1: SelfAttention(share_qk=False) (EA:1133-1197): keys have their own projection w_k, are NOT
                            * normalised (only divided by sqrt(d_qk), EA:232) and self-attention is allowed (EA:1175-1178);
                            * every "qv" buffer then has dq+dv+dq columns per head: q | v | k.  n_hashes must be 1. */
  int32_t x_bf16;          /* layer calls with act_dtype = F32 only: 1 = the INPUT x is already bf16 (the caller's LayerNorm wrote
                            * it that way, lsh_layernorm_fwd_bf16): no conversion pass; out / dout / dx stay f32. */
  int32_t reserved[1];
} LshAttnDims;

int lsh_attn_abi_version(void);
/* Content hash of the sources this library was compiled from (trax_b200/build.py); the loader refuses a stale build. */
const char *lsh_attn_source_hash(void);
const char *lsh_attn_last_error(void);

/* 0 when the shape is supported by the sm_100a kernels; otherwise non-zero + message.  There is no
 * CPU or generic fallback: unsupported shapes are rejected. */
int lsh_attn_check_dims(const LshAttnDims *dims);

/* ---- stage-level entry points ---------------------------------------------------------------- */

/* Weight layout change + f32→bf16 (replaces nothing in EA; prepares operands for EA:1923-1924, 1995). */
/* w_k (H, D, dq) f32: the key projection of SelfAttention(share_qk=False) (EA:1112-1128); NULL unless dims.separate_k. */
int lsh_pack_weights(const LshAttnDims *dims, const float *w_q, const float *w_v, const float *w_o, const float *w_k,
                     void *wqv_bf16, void *wo_bf16, void *stream);

/* EA:1923-1924  q = x·w_q ; v = x·w_v for every unit at once.  x_bf16 (B,L,D) bf16. */
int lsh_project_qv(const LshAttnDims *dims, const void *x_bf16, const void *wqv_bf16, void *qv_bf16,
                   void *ws, size_t ws_bytes, void *stream);

/* EA:1889-1916 hash_vectors + EA:60-119 hash_vecs: fp32 sequential-fmaf rotation, argmax over
 * [x,-x] per factor (first max wins), factor combine, mask bucket, per-round offsets. */
int lsh_hash(const LshAttnDims *dims, const void *qv_bf16, const float *rotations,
             const uint8_t *mask, int32_t *buckets, int64_t buckets_stride, void *stream);

/* Same hash on caller-supplied fp32 vectors (BH, L, dq) — the `hash_vecs`-granularity entry. */
int lsh_hash_f32(const LshAttnDims *dims, const float *vecs, const float *rotations,
                 const uint8_t *mask, int32_t *buckets, int64_t buckets_stride, void *stream);

/* EA:1946-1956 the two sort_key_val calls: stable sort of key = L*bucket + position per unit.
 * Emits sticker and (if non-null) undo_sort. */
size_t lsh_sort_workspace_bytes(const LshAttnDims *dims);
int lsh_sort(const LshAttnDims *dims, const int32_t *buckets, int64_t buckets_stride,
             int32_t *sticker, int32_t *undo_sort, void *ws, size_t ws_bytes, void *stream);

/* EA:1958-1986: gather by sticker, `attend` (EA:163-268) with look-back window, masks
 * (EA:145-160), per-row log-sum-exp, and the un-sort of EA:1985-1986 (rows are written straight to
 * their ticker slot).  o_rounds may alias o_comb's layout when nh==1 via lsh_attend_fwd_strided. */
/* The workspace holds the per-call auxiliaries of the tcgen05 path: per-token key scale, normalised keys
 * q / (8 r) (EA:229-231), {query scale, softmax shift} pairs and the position-sorted chunks (see lsh_chunk_possort). */
/* attn_keep (may be NULL): the attention-dropout multiplier of EA:254-262, a (chunk_len, window) f32 matrix with values in
 * {0, 1 / (1 - rate)} drawn by the caller (`bernoulli(rng, keep_prob, (C, W)) / keep_prob`, EA:258-262; a JAX host draws it
 * with jax.random from `attend_rng`, EA:1920) — ONE matrix per call, shared by every chunk, unit and hash round.  It
 * multiplies exp(dots - lse) before the product with v; the returned log-sum-exp does not see it. */
size_t lsh_attend_fwd_workspace_bytes(const LshAttnDims *dims);
int lsh_attend_fwd(const LshAttnDims *dims, const void *qv_bf16, const int32_t *sticker,
                   const uint8_t *mask, const float *attn_keep, void *o_rounds_bf16, float *logits, void *ws,
                   size_t ws_bytes, void *stream);

/* Internal order used by the tcgen05 attention kernels (chunk_len 128): sticker2 = sticker with every 128-slot
 * chunk re-ordered by token position (ticker % L, ascending; ties keep slot order).  Attention inside a chunk
 * window (EA:209-268) is a sum over keys and its rows are un-sorted by ticker (EA:1985-1986), so results do not
 * depend on the order inside a chunk; the reference's sticker / undo_sort (EA:1951-1956) are untouched.
 * bounds (B*H, nh*L) int32, may be NULL: per row of sticker2, where the row's position falls inside the two neighbour
 * chunks of its unit (cyclic, EA:137-141): cnt_prev | eq_prev << 8 | cnt_next << 16 | eq_next << 24 with cnt = number of
 * the neighbour's tokens at a lower position (0..128), eq = the neighbour holds the same position (the token's copy from
 * the adjacent hash round).  These are the column intervals the causal / self masks of EA:150-155 reduce to. */
int lsh_chunk_possort(const LshAttnDims *dims, const int32_t *sticker, int32_t *sticker2, int32_t *bounds,
                      void *stream);

/* EA:1988-1992 multi-round combine; also emits lse_tot = logsumexp_h(logits) (BH, L) when non-null. */
int lsh_combine_fwd(const LshAttnDims *dims, const void *o_rounds_bf16, const float *logits,
                    void *o_comb_bf16, float *lse_tot, void *stream);

/* EA:1995 out = o·w_o summed over heads (EA:2426).  out dtype = dims->act_dtype. */
int lsh_project_out(const LshAttnDims *dims, const void *o_comb_bf16, const void *wo_bf16, void *out,
                    void *ws, size_t ws_bytes, void *stream);

/* VJP of EA:1958-1992 with the permutation fixed (jax.vjp at EA:2418-2421; formulas in SURVEY.md
 * App. B): recomputes S and P per chunk from qv + sticker; do_comb is the cotangent of o_comb
 * (B,L,H,dv) bf16.  Writes dqv (B,L,H,dq+dv) bf16, the cotangent of qv. */
size_t lsh_attend_bwd_workspace_bytes(const LshAttnDims *dims);
int lsh_attend_bwd(const LshAttnDims *dims, const void *qv_bf16, const int32_t *sticker,
                   const uint8_t *mask, const float *attn_keep, const void *o_comb_bf16, const float *lse_tot,
                   const void *do_comb_bf16, void *dqv_bf16, void *ws, size_t ws_bytes,
                   void *stream);

/* ---- neighbours of the layer inside the reversible block -------------------
 * trax/layers/reversible.py:244-412 `ReversibleHalfResidual(LayerNorm(), attention_layer=LSHSelfAttention)`:
 * forward  y1 = x1 + Attn(LN(x2));  reverse_and_grad  x1 = y1 - Attn(LN(x2)), ct_x2 += LN_vjp(dz).
 * Activations (rows, d_model) in dims-style act_dtype (LSH_DTYPE_F32 / LSH_DTYPE_BF16); d_model in {256,512,1024,2048}. */

/* trax/layers/normalization.py:129-136 (center=True): z = (x - mean) / sqrt(var + epsilon) * scale + bias.
 * stats (rows, 2) f32 receives {mean, 1/sqrt(var + epsilon)} for the backward (may be NULL). */
int lsh_layernorm_fwd(int64_t rows, int d_model, int act_dtype, const void *x, const float *scale,
                      const float *bias, void *z, float *stats, float epsilon, void *stream);

/* The same normalisation for f32 activations with z written as bf16 — what the attention layer makes of its input anyway
 * (same values: one rounding of the same fp32 number) — for layer calls with dims.x_bf16 = 1. */
int lsh_layernorm_fwd_bf16(int64_t rows, int d_model, const float *x, const float *scale, const float *bias,
                           void *z_bf16, float *stats, float epsilon, void *stream);

/* VJP of the above: ct_out = (ct_in ? ct_in : 0) + dx (reversible.py:397-398 adds it to the context cotangent);
 * d_scale, d_bias (d_model) f32 are overwritten. */
int lsh_layernorm_bwd(int64_t rows, int d_model, int act_dtype, const void *x, const void *dz,
                      const void *ct_in, const float *stats, const float *scale, void *ct_out,
                      float *d_scale, float *d_bias, void *stream);

/* reversible.py:400 reconstructed_x = accumulator_output - residual (n elements, n % 8 == 0; out may alias a). */
int lsh_residual_sub(int64_t n, int act_dtype, const void *a, const void *b, void *out, void *stream);
/* reversible.py:318 output = accumulator + residual. */
int lsh_residual_add(int64_t n, int act_dtype, const void *a, const void *b, void *out, void *stream);

/* Head layout plumbing of the weight-less core, PureLSHSelfAttention (EA:2564; inputs (batch*heads, seqlen, d_head),
 * EA:3052-3070):  a (BH, L, da) [, b (BH, L, db), or NULL with db = 0] in act_dtype  ->  dst (B, L, H, da + db) bf16, the
 * (token, head) row layout `lsh_hash` / `lsh_attend_fwd` / `lsh_attend_bwd` read ([qk | v] side by side).  Widths % 8 == 0. */
int lsh_pack_heads(int B, int H, int L, int act_dtype, const void *a, int da, const void *b, int db, void *dst, void *stream);
/* src (B, L, H, d_total) bf16, columns [col0, col0 + d)  ->  dst (BH, L, d) in act_dtype: the core's output (from o_comb) and
 * its input cotangents (dqk, dv from dqv), EA:3245-3265. */
int lsh_unpack_heads(int B, int H, int L, int act_dtype, const void *src, int d_total, int col0, int d, void *dst, void *stream);

/* ---- layer-level entry points (EA:2261-2561 forward_and_or_backward) -------------------------- */

size_t lsh_layer_workspace_bytes(const LshAttnDims *dims, int with_grad);

/* compute_output=True.  update_state=True when `rotations` != NULL: buckets are computed and
 * written (EA:1926-1937); otherwise buckets are read (EA:1939-1941). */
int lsh_layer_fwd(const LshAttnDims *dims, const void *x, const float *w_q, const float *w_v,
                  const float *w_o, const float *w_k, const float *rotations, const uint8_t *mask, const float *attn_keep,
                  int32_t *buckets, int64_t buckets_stride, void *out, void *ws, size_t ws_bytes, void *stream);

/* output_grad given, update_state=False: recomputes the forward from the stored buckets, then the
 * backward.  `out` may be NULL (compute_output=False, EA:2256-2258) or a buffer
 * (compute_output=True, the call ReversibleHalfResidual makes, reversible.py:374-378).
 * dw_q/dw_v/dw_o are fully overwritten with the sum over examples (EA:2431); dx (B,L,D) with the
 * sum over heads (EA:2430).  Work is ordered on `stream`; one GEMM (do = dout·w_o^T) runs on an internal
 * helper stream that is forked from and joined back into `stream` by events (no host wait, capture-safe).
 * ev_dwo_ready / ev_dwqv_ready (cudaEvent_t passed as void*, either may be NULL) are recorded on `stream` the moment
 * dw_o, resp. dw_q and dw_v, are final — dw_o before the attention-gradient kernels start, dw_q|dw_v before the last
 * GEMM (dx) — so a data-parallel caller can start the gradient all-reduce (`psum`, trax/optimizers/trainer.py:197-199)
 * on its own communication stream underneath the rest of the call. */
int lsh_layer_bwd(const LshAttnDims *dims, const void *x, const float *w_q, const float *w_v,
                  const float *w_o, const float *w_k, const uint8_t *mask, const float *attn_keep, const int32_t *buckets,
                  int64_t buckets_stride, const void *dout, void *out, void *dx, float *dw_q,
                  float *dw_v, float *dw_o, float *dw_k, void *ws, size_t ws_bytes, void *ev_dwo_ready,
                  void *ev_dwqv_ready, void *stream);

/* The same two calls with the residual of the enclosing reversible block fused into the output projection's epilogue
 * (layers/reversible.py:318  output = accumulator + residual;  :400  reconstructed_x = accumulator_output - residual):
 * out = residual + acc_sign * attention_output, `residual` (B, L, D) in act_dtype (may alias `out`), acc_sign = +1 / -1.
 * residual == NULL gives the plain calls above. */
int lsh_layer_fwd_res(const LshAttnDims *dims, const void *x, const float *w_q, const float *w_v,
                      const float *w_o, const float *w_k, const float *rotations, const uint8_t *mask, const float *attn_keep,
                      int32_t *buckets, int64_t buckets_stride, void *out, const void *residual, float acc_sign,
                      void *ws, size_t ws_bytes, void *stream);
int lsh_layer_bwd_res(const LshAttnDims *dims, const void *x, const float *w_q, const float *w_v,
                      const float *w_o, const float *w_k, const uint8_t *mask, const float *attn_keep,
                      const int32_t *buckets, int64_t buckets_stride, const void *dout, void *out, void *dx,
                      float *dw_q, float *dw_v, float *dw_o, float *dw_k, void *ws, size_t ws_bytes,
                      void *ev_dwo_ready, void *ev_dwqv_ready, const void *residual, float acc_sign, void *stream);

/* ---- fast inference, mode='predict' (EA:1999-2109, 1200-1268) --------------------------------- */

/* ONE new token per example against the layer's input memory — the single-token branch of
 * `LSHSelfAttention._incremental_forward_unbatched` (EA:2032-2109) for every unit at once, and with rotations == buckets ==
 * NULL the q_len == 1 case of `SelfAttention._incremental_forward_unbatched` (EA:1200-1268; dims.separate_k / dims.causal as
 * in the layer calls).  dims.L is the MEMORY length (`predict_mem_len`), dims.factors the bucket list of THIS step (EA:2066
 * hashes two rows, so `n_buckets=None` resolves to [2], EA:1893-1902).
 *   mem      (B, M, D) act_dtype: the input memory AFTER `_use_predict_mem` (EA:2174-2207) stored the new token at q_start
 *   buckets  (B*H, buckets_stride) int32, rows (nh, M): the bucket memory AFTER the caller's roll (EA:2036-2053); the call
 *            writes the new token's ids into column q_start (EA:2069-2071) — in place
 *   rotations (B*H, dq, nh, R) f32: drawn from the state's hash key itself, not from a split of it (EA:2066)
 *   out      (B, 1, D) act_dtype
 * Attended slots: the n_hashes * chunk_len * (1 + n_chunks_before) memory slots of highest priority (same-bucket slots first,
 * then the most recent ones, EA:2073-2084) under the causal and self masks of EA:2091-2092.  The memory roll, `mem_end` and
 * `buckets_idx` are host bookkeeping (trax_b200/predict.py).  Prefixes (q_len > 1, EA:2004-2030) are lsh_layer_fwd calls. */
size_t lsh_predict_workspace_bytes(const LshAttnDims *dims);
int lsh_predict_step(const LshAttnDims *dims, const void *mem, const float *w_q, const float *w_v, const float *w_o,
                     const float *w_k, const float *rotations, int32_t *buckets, int64_t buckets_stride, int32_t q_start,
                     void *out, void *ws, size_t ws_bytes, void *stream);

/* The same step on caller-supplied projections — `PureLSHSelfAttention._incremental_forward_unbatched` (EA:2858-2932), whose
 * memory holds qk and v themselves: qv (B, M, H, dq+dv) bf16 is that memory in the kernels' row layout (lsh_pack_heads) with
 * the new token stored at q_start; o (B*H, dv) f32 receives the attention output of the new token (there is no output
 * projection, EA:2928).  buckets / rotations / q_start as above. */
size_t lsh_predict_attend_workspace_bytes(const LshAttnDims *dims);
int lsh_predict_attend(const LshAttnDims *dims, const void *qv_bf16, const float *rotations, int32_t *buckets,
                       int64_t buckets_stride, int32_t q_start, float *o, void *ws, size_t ws_bytes, void *stream);

/* ---- helpers ---------------------------------------------------------------------------------- */

/* Counter-based N(0,1) draws for the rotations (stands in for fastmath.random.normal at EA:92;
 * NOT bit-compatible with jax.random — a JAX host passes its own rotations instead).
 * keys: (B*H, 2) uint32 per-unit hash keys (EA:1870-1882 state rng); new_keys (B*H,2) receives the
 * advanced key (EA:1928 split). */
int lsh_make_rotations(const LshAttnDims *dims, const uint32_t *keys, uint32_t *new_keys,
                       float *rotations, void *stream);

/* Number of kernels this library has launched on this thread since the last reset (bench.py's
 * gpu_launches claim). */
int64_t lsh_attn_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif  /* LSH_ATTN_H_ */
