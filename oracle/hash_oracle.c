/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Not shipped, not on the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object.
 *
 * CPU restatement of the floating-point contraction inside
 *   trax/layers/research/efficient_attention.py:91-101  (hash_vecs: rotated_vecs = einsum('tf,fhb->htb'))
 * under ONE fixed fp32 accumulation convention: a single fp32 accumulator per output,
 * fused multiply-add (fmaf, one rounding per step), contraction index f ascending 0..dq-1,
 * accumulator starting at +0.0f.  The CUDA kernel (trax_b200/csrc/hash.cu) follows the same
 * convention with __fmaf_rn, which is what makes bucket ids bit-exact.
 *
 * XLA's own summation order is not observable here (JAX is not installed), so the
 * convention is OURS.  It is checked against the live reference run under the reference's NumPy
 * backend (float64 np.dot; tests/test_reference_pin.py: every bucket id of the fixture cases equal,
 * 20 608 ids); at an exact near-tie of the argmax two summation orders can legitimately differ.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o _build/liboracle.so hash_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>

/* vecs: [n_vecs][dq] row-major fp32; rot: [dq][ncols] row-major fp32 (ncols = n_hashes * rot_size/2,
 * i.e. the reference's (dq, n_hashes, rot_size//2) array flattened over its last two axes);
 * out: [n_vecs][ncols]. */
void oracle_rotate_f32(const float *vecs, int64_t n_vecs, int32_t dq,
                       const float *rot, int32_t ncols, float *out) {
  for (int64_t t = 0; t < n_vecs; ++t) {
    const float *v = vecs + t * (int64_t)dq;
    float *o = out + t * (int64_t)ncols;
    for (int32_t c = 0; c < ncols; ++c) {
      float acc = 0.0f;
      for (int32_t f = 0; f < dq; ++f) acc = fmaf(v[f], rot[(int64_t)f * ncols + c], acc);
      o[c] = acc;
    }
  }
}

/* argmax over concat([rv, -rv]) with first-max-wins ties (np.argmax / jnp.argmax semantics),
 * EA:103-105 and EA:110-116.  rv: [half] fp32.  Returns index in [0, 2*half). */
int32_t oracle_argmax_pm(const float *rv, int32_t half) {
  int32_t best = 0;
  float bestv = rv[0];
  for (int32_t i = 1; i < 2 * half; ++i) {
    float x = (i < half) ? rv[i] : -rv[i - half];
    if (x > bestv) { bestv = x; best = i; }
  }
  return best;
}

/* Full hash_vecs for a factor list (EA:60-119) + per-round offsets (EA:1913-1916), used as a second,
 * loop-level restatement next to the NumPy one.  buckets_out: [n_hashes][n_vecs] int32 WITH offsets. */
void oracle_hash_vecs(const float *vecs, int64_t n_vecs, int32_t dq, const float *rot,
                      int32_t n_hashes, int32_t rot_half_total, const int32_t *factors,
                      int32_t n_factors, const uint8_t *mask /* may be NULL; 1 = valid */,
                      int32_t *buckets_out) {
  int32_t ncols = n_hashes * rot_half_total;
  int32_t n_buckets = 1;
  for (int32_t i = 0; i < n_factors; ++i) n_buckets *= factors[i];
  int32_t n_buckets_total = n_buckets + (mask ? 1 : 0);
  float rv[4096];
  for (int64_t t = 0; t < n_vecs; ++t) {
    const float *v = vecs + t * (int64_t)dq;
    for (int32_t h = 0; h < n_hashes; ++h) {
      for (int32_t c = 0; c < rot_half_total; ++c) {
        float acc = 0.0f;
        int32_t col = h * rot_half_total + c;
        for (int32_t f = 0; f < dq; ++f) acc = fmaf(v[f], rot[(int64_t)f * ncols + col], acc);
        rv[c] = acc;
      }
      int32_t bucket = 0, cur = 0, prod = 1;
      for (int32_t i = 0; i < n_factors; ++i) {
        int32_t half = factors[i] / 2;
        bucket += prod * oracle_argmax_pm(rv + cur, half);
        cur += half;
        prod *= factors[i];
      }
      if (mask && !mask[t]) bucket = n_buckets_total - 1;
      buckets_out[(int64_t)h * n_vecs + t] = bucket + h * n_buckets_total;
    }
  }
}
