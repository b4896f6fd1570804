// Parameter block shared by the two forward attention kernels (mma.sync path and tcgen05 path).
#pragma once
#include "common.cuh"

namespace lsh {

struct AttendFwdParams {
  const __nv_bfloat16 *qv;      // (B, L, H, 128)
  const int32_t *sticker;       // (BH, N)
  const uint8_t *mask;          // (B, L) or null
  __nv_bfloat16 *o;             // rows addressed as b*o_sb + h*o_sh + round*o_sr + pos*o_sp
  int64_t o_sb, o_sh, o_sr, o_sp;
  float *lse;                   // (BH, N) ticker order
  const float *qscale;          // (BH, L) per-token key scale (tcgen05 path only)
  long long *trace;             // debug: per-phase clock64 stamps of CTA 0 (null = off)
  unsigned stagger_ns;          // start delay of the second softmax warpgroup (tcgen05 path)
  int L, H, N, n_chunks, nb, nwin, causal, masked;
};

int attend_fwd_tc_run(const AttendFwdParams &p, int BH, cudaStream_t stream);

}  // namespace lsh
