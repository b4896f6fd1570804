// Parameter block shared by the two forward attention kernels (mma.sync path and tcgen05 path).
#pragma once
#include <cuda.h>   // CUtensorMap

#include "common.cuh"

namespace lsh {

struct AttendFwdParams {
  const __nv_bfloat16 *qv;      // (B, L, H, 128)
  const int32_t *sticker;       // (BH, N)
  const uint8_t *mask;          // (B, L) or null
  __nv_bfloat16 *o;             // rows addressed as b*o_sb + h*o_sh + round*o_sr + pos*o_sp
  int64_t o_sb, o_sh, o_sr, o_sp;
  float *lse;                   // (BH, N) ticker order
  const float *qscale;          // (BH, L) per-token key scale (unused by the forward kernels; kept for the backward)
  const __nv_bfloat16 *qhat;    // (BH, L, 64) normalised keys q / (8 r)                 } tcgen05 path only
  const float2 *rowmeta;        // (BH, L) {a = 8 r log2e, m2 = a |qhat|^2}               }
  const int32_t *sticker2;      // (BH, N) sticker with every chunk re-ordered by position }
  const int32_t *bounds;        // (BH, N) per row of sticker2: neighbour-chunk interval bounds (chunk_possort_kernel)
  const uint32_t *keep_bits;    // attention dropout (EA:254-262): (C, W / 32) bit rows, bit j of word w = keep[i][32 w + j]; null = none
  const float *keep_scale;      // device scalar: the keep multiplier 1 / (1 - rate)
  int *redo;                    // tcgen05 path: [0] = number of queued rows, [2 + 2 i], [3 + 2 i] = {unit * n_chunks + chunk, ticker}
  long long *trace;             // debug: per-phase clock64 stamps of CTA 0 (null = off)
  int L, H, N, n_chunks, nb, nwin, causal, masked;
  CUtensorMap tm_k, tm_v;       // tcgen05 path: TMA row-gather descriptors of qhat (BH*L, 64) and of qv viewed as (B*L*H, row)
  int row;                      // elements per (token, head) row of qv: 128 (q | v) or 192 (q | v | k, separate keys)
  int ksep;                     // separate, un-normalised keys, self-attention allowed (SelfAttention(share_qk=False), EA:1133-1197)
};

int attend_fwd_tc_run(const AttendFwdParams &p, int BH, cudaStream_t stream);

// Attention dropout (EA:254-262): ONE (chunk_len, window) keep multiplier per layer call — shared by every chunk, unit and
// hash round — with values in {0, 1 / (1 - rate)}.  The kernels consume it as bit rows (by query slot; `bits_t`: by window
// column, for the key-centric backward) plus the scalar; attn_keep_prepare converts the caller's float matrix.
struct AttnKeep {
  const uint32_t *bits;     // (C, W / 32)
  const uint32_t *bits_t;   // (W, C / 32)
  const float *scale;       // device scalar
};
size_t attn_keep_bytes(const LshAttnDims &d);
int attn_keep_prepare(const LshAttnDims &d, const float *keep_f32, void *ws, AttnKeep *out, cudaStream_t stream);

// Auxiliary per-call buffers of the tcgen05 forward path (produced by qscale_run / chunk_possort_run).
struct FwdAux {
  float *qscale;
  float2 *rowmeta;
  void *qhat;
  int32_t *sticker2;
  int32_t *bounds;
  int *redo;
};
bool attend_fwd_uses_tc(const LshAttnDims &d);
size_t fwd_aux_bytes(const LshAttnDims &d);
FwdAux fwd_aux_carve(const LshAttnDims &d, void *ws);
int fwd_aux_prepare(const LshAttnDims &d, const void *qv, const int32_t *sticker, const FwdAux &aux, cudaStream_t stream,
                    bool scales_done = false);

}  // namespace lsh
