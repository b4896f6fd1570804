// Parameter block of the tcgen05 backward attention kernel.
#pragma once
#include "common.cuh"

namespace lsh {

struct AttendBwdTcParams {
  const __nv_bfloat16 *qv;        // (B, L, H, 128)
  const int32_t *sticker;         // (BH, N)
  const int32_t *sticker2;        // (BH, N) sticker with every 128-slot chunk re-ordered by position (chunk_possort_kernel)
  const int32_t *bounds;          // (BH, N) neighbour-chunk interval bounds per row of sticker2 (chunk_possort_kernel)
  const __nv_bfloat16 *do_comb;   // (B, L, H, 64)
  const float *qscale;            // (BH, L)  log2e / (sqrt(mean(q^2)+eps) sqrt(dq))
  const float *lse2;              // (BH, L)  -(log2e * lse_tot (+ log2e*1e5 for self-only rows))   } stored negated
  const float *dvec;              // (BH, L)  -(do . o)                                              }
  const float *qcmp;              // (BH, L)  pos + 1 (+ 0.5 for self-only rows)
  __nv_bfloat16 *dq_out;          // (BH, N, 64) ticker order: dq_query + dq_key of every token copy
  __nv_bfloat16 *dv_out;          // (BH, N, 64)
  int L, H, N, n_chunks;
  // Attention dropout (EA:254-262): keep bits by window column (AttnKeep::bits_t, (2 C, C / 32)) and the 1 / (1 - rate)
  // multiplier; null = none.  Dropout calls pass `sticker` (slot order) as sticker2: the keep matrix is indexed by slot, so
  // tiles are NOT re-ordered by position and every block takes the per-element position compare (no block skipping).
  const uint32_t *keep_bits_t;
  const float *keep_scale;
  int slot_order;
  long long *trace;               // debug: per-phase clock64 stamps of CTA 0 (null = off)
};

int attend_bwd_tc_run(const AttendBwdTcParams &p, int BH, cudaStream_t stream);

}  // namespace lsh
