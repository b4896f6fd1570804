// Chunked look-back attention, backward — the VJP of EA:1958-1992 that `jax.vjp` synthesises at
// EA:2418-2421, restated per SURVEY.md App. B with the multi-round combine folded in:
//   P_tot[i][j] = exp(S[i][j] - lse_tot[i])          (= w_round * P_round)
//   dP[i][j]    = do[i] · v[j]                        (do = cotangent of the COMBINED o, EA:1992)
//   dS          = P_tot ∘ (dP - D[i]),  D[i] = do[i]·o[i]
//   dV[j] += P_tot^T do ; dK^[j] += dS^T q ; dQ[i] += dS k^ ; then the length-normalisation VJP.
//
// v1 compute path: bf16 mma.sync m16n8k16, fp32 accumulation.  KEY-centric: one CTA owns one key
// chunk (its dK^/dV accumulators stay in registers, no atomics) and visits the (1+nb+na) query
// chunks that look at it, 64 queries at a time, in the transposed orientation (S^T, dP^T) so that
// P^T / dS^T are already A-operand fragments for the dV / dK^ products.  dS^T goes through shared
// memory once for dQ = dS·k^.  dQ partials are written per window slot ("kind"); sum_rounds_kernel
// adds kinds and hash rounds (App. B6) — deterministic, no atomics.
#include <stdlib.h>
#include <string.h>

#include "attend_bwd_params.cuh"
#include "attend_params.cuh"

namespace lsh {

extern long long *g_fwd_trace;   // shared debug trace buffer (attend_fwd.cu)

struct AttendBwdParams {
  const __nv_bfloat16 *qv;       // (B, L, H, 128)
  const int32_t *sticker;        // (BH, N)
  const uint8_t *mask;           // (B, L) or null
  const __nv_bfloat16 *do_comb;  // (B, L, H, 64)
  const float *lse_tot;          // (BH, L)
  const float *dvec;             // (BH, L)
  __nv_bfloat16 *dq_part;        // (nwin+1, BH, N, 64): kinds 0..nwin-1 = query side, kind nwin = key side
  __nv_bfloat16 *dv_part;        // (BH, N, 64)
  int64_t kind_stride;
  const uint32_t *keep_bits_t;   // attention dropout: (W, C / 32) bit rows by window column (null = none), see AttnKeep
  const float *keep_scale;
  int L, H, N, n_chunks, nb, nwin, causal, masked;
  int row, ksep;                 // see AttendFwdParams
};

constexpr int QB = 64;   // queries per sub-block

template <int C>
__global__ void __launch_bounds__(2 * C, 1) attend_bwd_kernel(const AttendBwdParams p) {
  constexpr int NT = 2 * C;
  constexpr int NWARP = C / 16;
  constexpr int NSPLIT = NWARP / 4;            // how many warps share one 16-row dQ stripe
  constexpr int NTW = 8 / NSPLIT;              // n-tiles of dQ per warp
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *Ks = smem;                          // [C][64] bf16  raw q rows of the key chunk
  uint8_t *Vs = Ks + C * 128;                  // [C][64] bf16
  uint8_t *Qs = Vs + C * 128;                  // [QB][64] bf16 raw queries of the sub-block
  uint8_t *Ds = Qs + QB * 128;                 // [QB][64] bf16 do rows
  uint8_t *Ts = Ds + QB * 128;                 // [C][QB] bf16 (dS ∘ kscale)^T staged for dQ
  float *rnorm = reinterpret_cast<float *>(Ts + C * 128);   // [C]
  int *kpos = reinterpret_cast<int *>(rnorm + C);           // [C] 0-based key positions
  int *ktk = kpos + C;                                      // [C] key tickers
  int *kinfo = ktk + C;                                     // [C]
  int *qtk = kinfo + C;                                     // [QB]
  int *qpos = qtk + QB;                                     // [QB]
  float *qlse = reinterpret_cast<float *>(qpos + QB);       // [QB]
  float *qD = qlse + QB;                                    // [QB]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3, mi = lane >> 3;
  const int u = blockIdx.x / p.n_chunks, kc = blockIdx.x % p.n_chunks;
  const int b = u / p.H, h = u % p.H;
  const int32_t *stk = p.sticker + static_cast<int64_t>(u) * p.N;
  const uint32_t ks_base = smem_u32(Ks), vs_base = smem_u32(Vs), qs_base = smem_u32(Qs),
                 ds_base = smem_u32(Ds), ts_base = smem_u32(Ts);

  // ---- key chunk: metadata, gather, normalise ------------------------------------------------------
  for (int j = tid; j < C; j += NT) {
    const int tk = stk[kc * C + j];
    const int pos = tk % p.L;
    bool valid = true;
    if (p.masked) valid = p.mask[static_cast<int64_t>(b) * p.L + pos] != 0;
    ktk[j] = tk; kpos[j] = pos; kinfo[j] = (valid ? pos : -pos) + 1;
  }
  __syncthreads();
  for (int i = tid; i < C * 16; i += NT) {
    const int j = i >> 4, ch = i & 15;
    // key rows: the q columns (shared-QK) or the k columns (separate keys); value rows: the v columns
    const __nv_bfloat16 *src = p.qv + ((static_cast<int64_t>(b) * p.L + kpos[j]) * p.H + h) * p.row + ch * 8 + ((p.ksep && ch < 8) ? 128 : 0);
    cp_async16((ch < 8) ? ks_base + swz(j, ch) : vs_base + swz(j, ch - 8), src);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  for (int j = tid >> 3; j < C; j += NT / 8) {
    const int ch = tid & 7;
    const uint4 raw = *reinterpret_cast<const uint4 *>(Ks + swz(j, ch));
    float2 f0 = unpack_bf16(raw.x), f1 = unpack_bf16(raw.y), f2 = unpack_bf16(raw.z), f3 = unpack_bf16(raw.w);
    float ss = f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y + f2.x * f2.x + f2.y * f2.y +
               f3.x * f3.x + f3.y * f3.y;
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    if (ch == 0) rnorm[j] = p.ksep ? 1.f : sqrtf(ss * (1.0f / 64) + 1e-6f);   // separate keys are not normalised (EA:229-232)
  }
  __syncthreads();

  // A-operand fragments of this warp's 16 keys: k^ (for S^T) and v (for dP^T)
  const int krow0 = warp * 16;
  uint32_t ka[4][4], va[4][4];
  {
    const int row = krow0 + (lane & 7) + 8 * (mi & 1);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldmatrix_x4(ks_base + swz(row, ks * 2 + (mi >> 1)), ka[ks][0], ka[ks][1], ka[ks][2], ka[ks][3]);
      ldmatrix_x4(vs_base + swz(row, ks * 2 + (mi >> 1)), va[ks][0], va[ks][1], va[ks][2], va[ks][3]);
    }
  }
  const float ki0 = static_cast<float>(kinfo[krow0 + g]), ki1 = static_cast<float>(kinfo[krow0 + g + 8]);
  // k^ = q / (r * sqrt(dq)) is never materialised: the per-key factor scales the fp32 score row (and dS below)
  const float ksc0 = 0.125f / rnorm[krow0 + g], ksc1 = 0.125f / rnorm[krow0 + g + 8];

  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }

  // ---- visit the query chunks whose window contains key chunk kc ----------------------------------
  for (int wslot = 0; wslot < p.nwin; ++wslot) {
    // window slot `wslot` of query chunk qc holds chunk qc + (wslot - nb)  (EA:137-141)  =>  qc = kc - (wslot - nb)
    int qc = kc - (wslot - p.nb);
    qc = (qc % p.n_chunks + p.n_chunks) % p.n_chunks;
    for (int sub = 0; sub < C / QB; ++sub) {
      __syncthreads();   // previous sub-block fully consumed (Qs, Ds, Ts, q-meta)
      if (tid < QB) {
        const int tk = stk[qc * C + sub * QB + tid];
        const int pos = tk % p.L;
        qtk[tid] = tk; qpos[tid] = pos;
        qlse[tid] = p.lse_tot[static_cast<int64_t>(u) * p.L + pos];
        qD[tid] = p.dvec[static_cast<int64_t>(u) * p.L + pos];
      }
      __syncthreads();
      for (int i = tid; i < QB * 16; i += NT) {
        const int j = i >> 4, ch = i & 15;
        const int64_t tokrow = (static_cast<int64_t>(b) * p.L + qpos[j]) * p.H + h;
        if (ch < 8) cp_async16(qs_base + swz(j, ch), p.qv + tokrow * p.row + ch * 8);
        else cp_async16(ds_base + swz(j, ch - 8), p.do_comb + tokrow * 64 + (ch - 8) * 8);
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();

      // S^T (16 keys x 64 queries) and dP^T
      float s[8][4], dp[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int ntp = 0; ntp < 4; ++ntp) {
          const int qrow = ntp * 16 + (lane & 7) + 8 * (mi >> 1);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(qs_base + swz(qrow, ks * 2 + (mi & 1)), b0, b1, b2, b3);
          mma_bf16(s[2 * ntp], ka[ks], b0, b1);
          mma_bf16(s[2 * ntp + 1], ka[ks], b2, b3);
          ldmatrix_x4(ds_base + swz(qrow, ks * 2 + (mi & 1)), b0, b1, b2, b3);
          mma_bf16(dp[2 * ntp], va[ks], b0, b1);
          mma_bf16(dp[2 * ntp + 1], va[ks], b2, b3);
        }
      }
      // masks (EA:145-160), P_tot^T, dS^T -> fragments + shared memory
      // attention dropout m (EA:254-262): so = (P∘m) V  =>  dV += (P∘m)^T do,  dS = P ∘ (m∘dP - D)   
      const uint32_t *kt0 = p.keep_bits_t ? p.keep_bits_t + static_cast<size_t>(wslot * C + krow0 + g) * (C / 32) + sub * 2 : nullptr;
      const uint32_t *kt1 = p.keep_bits_t ? kt0 + 8 * (C / 32) : nullptr;
      const float kscale_m = p.keep_bits_t ? __ldg(p.keep_scale) : 1.f;
      uint32_t pa[4][4], dsa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float pv[4], dsv[4];
        uint32_t kw0 = 0xffffffffu, kw1 = 0xffffffffu;
        if (kt0) { kw0 = __ldg(kt0 + (nt >> 2)) >> ((nt & 3) * 8 + 2 * t); kw1 = __ldg(kt1 + (nt >> 2)) >> ((nt & 3) * 8 + 2 * t); }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = nt * 8 + 2 * t + (e & 1);
          const float qi = static_cast<float>(qpos[col] + 1);
          const float ki = (e < 2) ? ki0 : ki1;
          float v = s[nt][e] * ((e < 2) ? ksc0 : ksc1);
          if (p.causal && qi < ki) v = v - 1e9f;
          if (!p.ksep && qi == ki) v = v - 1e5f;           // exclude_self = share_qk (EA:1175-1178)
          if (p.masked && ki < 0.f) v = v - 1e9f;
          const float pt = exp2f((v - qlse[col]) * kLog2e);
          const float mk = (((e < 2 ? kw0 : kw1) >> (e & 1)) & 1u) ? kscale_m : 0.f;
          pv[e] = pt * mk;
          dsv[e] = pt * (mk * dp[nt][e] - qD[col]);
        }
        const int kk = nt >> 1, hi = (nt & 1) * 2;
        pa[kk][hi] = pack_bf16(pv[0], pv[1]);   pa[kk][hi + 1] = pack_bf16(pv[2], pv[3]);
        dsa[kk][hi] = pack_bf16(dsv[0], dsv[1]); dsa[kk][hi + 1] = pack_bf16(dsv[2], dsv[3]);
        // dQ = dS·k^ = (dS ∘ kscale_j)·q_raw: the staged copy carries the key factor
        *reinterpret_cast<uint32_t *>(Ts + swz(krow0 + g, nt) + 4 * t) = pack_bf16(dsv[0] * ksc0, dsv[1] * ksc0);
        *reinterpret_cast<uint32_t *>(Ts + swz(krow0 + g + 8, nt) + 4 * t) = pack_bf16(dsv[2] * ksc1, dsv[3] * ksc1);
      }
      // dV += P^T·do ; dK^ += dS^T·q   (contraction over the 64 queries)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int ntp = 0; ntp < 4; ++ntp) {
          const int row = kk * 16 + (lane & 7) + 8 * (mi & 1);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(ds_base + swz(row, ntp * 2 + (mi >> 1)), b0, b1, b2, b3);
          mma_bf16(dv[2 * ntp], pa[kk], b0, b1);
          mma_bf16(dv[2 * ntp + 1], pa[kk], b2, b3);
          ldmatrix_x4_trans(qs_base + swz(row, ntp * 2 + (mi >> 1)), b0, b1, b2, b3);
          mma_bf16(dk[2 * ntp], dsa[kk], b0, b1);
          mma_bf16(dk[2 * ntp + 1], dsa[kk], b2, b3);
        }
      }
      __syncthreads();   // Ts complete

      // dQ partial (64 queries x 64) = dS · k^ : contraction over the C keys
      {
        const int mt = warp & 3, nq = warp >> 2;
        float dq[NTW][4];
#pragma unroll
        for (int i = 0; i < NTW; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < C / 16; ++kk) {
          uint32_t a[4];
          ldmatrix_x4_trans(ts_base + swz(kk * 16 + (lane & 7) + 8 * (mi >> 1), 2 * mt + (mi & 1)),
                            a[0], a[1], a[2], a[3]);
#pragma unroll
          for (int ntp = 0; ntp < NTW / 2; ++ntp) {
            const int row = kk * 16 + (lane & 7) + 8 * (mi & 1);
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(ks_base + swz(row, (nq * NTW / 2 + ntp) * 2 + (mi >> 1)), b0, b1, b2, b3);
            mma_bf16(dq[2 * ntp], a, b0, b1);
            mma_bf16(dq[2 * ntp + 1], a, b2, b3);
          }
        }
        __nv_bfloat16 *out = p.dq_part + static_cast<int64_t>(wslot) * p.kind_stride +
                             static_cast<int64_t>(u) * p.N * 64;
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        __nv_bfloat16 *o0 = out + static_cast<int64_t>(qtk[r0]) * 64, *o1 = out + static_cast<int64_t>(qtk[r1]) * 64;
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt) {
          const int col = (nq * NTW + nt) * 8 + 2 * t;
          *reinterpret_cast<uint32_t *>(o0 + col) = pack_bf16(dq[nt][0], dq[nt][1]);
          *reinterpret_cast<uint32_t *>(o1 + col) = pack_bf16(dq[nt][2], dq[nt][3]);
        }
      }
    }
  }

  // ---- key side: length-normalisation VJP (App. B5) and row stores ---------------------------------
  {
    const int r0 = krow0 + g, r1 = r0 + 8;
    float qraw[8][4];
    float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t *>(Ks + swz(r0, nt) + 4 * t));
      const float2 c = unpack_bf16(*reinterpret_cast<const uint32_t *>(Ks + swz(r1, nt) + 4 * t));
      qraw[nt][0] = a.x; qraw[nt][1] = a.y; qraw[nt][2] = c.x; qraw[nt][3] = c.y;
      dot0 += dk[nt][0] * a.x + dk[nt][1] * a.y;
      dot1 += dk[nt][2] * c.x + dk[nt][3] * c.y;
    }
    dot0 = quad_sum(dot0); dot1 = quad_sum(dot1);
    const float rr0 = rnorm[r0], rr1 = rnorm[r1];
    // dq_key = dk^/(r*sqrt(dq)) - q * (dk^·q) / (dq * r^3 * sqrt(dq)),  dq = 64
    const float a0 = 0.125f / rr0, a1 = 0.125f / rr1;
    // (separate keys: k = x w_k / sqrt(dq) only, so the key-side cotangent is dk^ / sqrt(dq) — no normalisation term)
    const float c0 = p.ksep ? 0.f : dot0 * 0.125f / (64.f * rr0 * rr0 * rr0), c1 = p.ksep ? 0.f : dot1 * 0.125f / (64.f * rr1 * rr1 * rr1);
    __nv_bfloat16 *oq = p.dq_part + static_cast<int64_t>(p.nwin) * p.kind_stride + static_cast<int64_t>(u) * p.N * 64;
    __nv_bfloat16 *ov = p.dv_part + static_cast<int64_t>(u) * p.N * 64;
    const int64_t t0 = ktk[r0], t1 = ktk[r1];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      *reinterpret_cast<uint32_t *>(oq + t0 * 64 + col) =
          pack_bf16(dk[nt][0] * a0 - qraw[nt][0] * c0, dk[nt][1] * a0 - qraw[nt][1] * c0);
      *reinterpret_cast<uint32_t *>(oq + t1 * 64 + col) =
          pack_bf16(dk[nt][2] * a1 - qraw[nt][2] * c1, dk[nt][3] * a1 - qraw[nt][3] * c1);
      *reinterpret_cast<uint32_t *>(ov + t0 * 64 + col) = pack_bf16(dv[nt][0], dv[nt][1]);
      *reinterpret_cast<uint32_t *>(ov + t1 * 64 + col) = pack_bf16(dv[nt][2], dv[nt][3]);
    }
  }
}

template <int C>
static int launch_attend_bwd(const AttendBwdParams &p, int BH, cudaStream_t stream) {
  size_t smem = static_cast<size_t>(C) * 128 * 3 + QB * 128 * 2 + C * 16 + QB * 16;
  LSH_OPT_IN_SMEM(attend_bwd_kernel<C>);
  attend_bwd_kernel<C><<<BH * p.n_chunks, 2 * C, smem, stream>>>(p);
  LSH_CHECK_LAUNCH("attend_bwd_kernel");
  return 0;
}

size_t attend_bwd_workspace_bytes(const LshAttnDims &d) {
  Derived dr = derive(d);
  size_t rows = static_cast<size_t>(dr.BH) * dr.N;
  size_t b = rows * 64 * 2 * (dr.nwin + 2);       // dq kinds (nwin+1) + dv
  b += static_cast<size_t>(dr.BH) * d.L * 4 * 4 + 1024;   // dvec, lse2, qcmp, qscale
  return (b + 255) / 256 * 256 + 512;
}

int bwd_prep_run(const LshAttnDims &d, const void *do_comb, const void *o_comb, float *dvec,
                 cudaStream_t stream);
int sum_rounds_run(const LshAttnDims &d, const void *dq_part, const void *dv_part, void *dqv,
                   int n_kinds, cudaStream_t stream);
int bwd_prep_tc_run(const LshAttnDims &d, const void *do_comb, const void *o_comb, const float *lse_tot, float *dvec,
                    float *lse2, float *qcmp, cudaStream_t stream);
int qscale_run(const LshAttnDims &d, const void *qv, float *qscale, float2 *rowmeta, void *qhat, cudaStream_t stream);
int chunk_possort_run(const LshAttnDims &d, const int32_t *sticker, int32_t *sticker2, int32_t *bounds, cudaStream_t stream);
bool attend_tc_uses_bounds();

// Measurement hook (lsh_debug_set_bwd_parts, bench.py): which parts of the stage run — bit 0 the per-token preparation
// kernels, bit 1 the attention-gradient kernel, bit 2 the sum over hash rounds.  7 = everything (the only product setting);
// a part that is skipped leaves the workspace of an earlier full call in place, so one part can be timed alone.
int g_bwd_parts = 7;

static bool force_mma_bwd() {
  static const bool f = [] { const char *e = getenv("LSH_ATTN_BWD"); return e && strcmp(e, "mma") == 0; }();
  return f;
}

// Whether attend_bwd_run takes the tcgen05 path for this call (long-sequence shape, no dropout / separate keys).
bool attend_bwd_uses_tc(const LshAttnDims &d, const AttnKeep *keep) {
  (void)keep;   // dropout runs on this path too: slot-ordered tiles, keep bits by window column (AttendBwdTcParams::slot_order)
  return d.C == 128 && d.nb == 1 && d.na == 0 && d.causal && !d.masked && d.L % 128 == 0 && !force_mma_bwd() && !d.separate_k;
}

// The per-token scalar buffers inside the stage workspace (same carve-up as attend_bwd_run below).
static void bwd_token_buffers(const LshAttnDims &d, void *ws, float **dvec, float **lse2, float **qcmp, float **qscale_ws) {
  Derived dr = derive(d);
  const size_t rows = static_cast<size_t>(dr.BH) * dr.N;
  char *w = static_cast<char *>(ws);
  w += rows * 64 * 2 * (dr.nwin + 1);
  w += rows * 64 * 2;
  w = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(w) + 255) / 256 * 256);
  const size_t tok = (static_cast<size_t>(dr.BH) * d.L * 4 + 255) / 256 * 256;
  *dvec = reinterpret_cast<float *>(w);
  *lse2 = reinterpret_cast<float *>(w + tok);
  *qcmp = reinterpret_cast<float *>(w + 2 * tok);
  *qscale_ws = reinterpret_cast<float *>(w + 3 * tok);
}

// For the layer's backward call: where the combine of the recompute should leave dvec / lse2 / qcmp (see BwdPrepOut).
int attend_bwd_prep_target(const LshAttnDims &d, void *ws, size_t ws_bytes, const void *do_comb, BwdPrepOut *out) {
  if (ws_bytes < attend_bwd_workspace_bytes(d)) return set_error("lsh_attend_bwd: workspace too small");
  float *qs;
  out->do_comb = do_comb;
  bwd_token_buffers(d, ws, &out->dvec, &out->lse2, &out->qcmp, &qs);
  return 0;
}

int attend_bwd_run(const LshAttnDims &d, const void *qv, const int32_t *sticker, const uint8_t *mask,
                   const void *o_comb, const float *lse_tot, const void *do_comb, const float *qscale_in,
                   const int32_t *sticker2_in, const int32_t *bounds_in, const AttnKeep *keep, void *dqv, void *ws, size_t ws_bytes,
                   cudaStream_t stream, bool prep_done) {
  Derived dr = derive(d);
  if (ws_bytes < attend_bwd_workspace_bytes(d))
    return set_error("lsh_attend_bwd: workspace too small (%zu < %zu)", ws_bytes, attend_bwd_workspace_bytes(d));
  if (d.masked && !mask) return set_error("attend_bwd: dims.masked set but mask == NULL");
  const size_t rows = static_cast<size_t>(dr.BH) * dr.N;
  char *w = static_cast<char *>(ws);
  __nv_bfloat16 *dq_part = reinterpret_cast<__nv_bfloat16 *>(w);
  w += rows * 64 * 2 * (dr.nwin + 1);
  __nv_bfloat16 *dv_part = reinterpret_cast<__nv_bfloat16 *>(w);
  w += rows * 64 * 2;
  float *dvec, *lse2, *qcmp, *qscale_ws;
  bwd_token_buffers(d, ws, &dvec, &lse2, &qcmp, &qscale_ws);
  int rc;
  // tcgen05 path for the long-sequence shape; LSH_ATTN_BWD=mma forces the mma.sync path
  const bool dropout = keep && keep->bits_t;
  if (attend_bwd_uses_tc(d, keep)) {
    const float *qscale = qscale_in;
    const bool prep = g_bwd_parts & 1;
    if (!qscale) {
      if (prep && (rc = qscale_run(d, qv, qscale_ws, nullptr, nullptr, stream))) return rc;
      qscale = qscale_ws;
    }
    if (prep && !prep_done && (rc = bwd_prep_tc_run(d, do_comb, o_comb, lse_tot, dvec, lse2, qcmp, stream))) return rc;
    // position-sorted chunks: from the forward pass of the same layer call if it made them, else into the (unused on this
    // path) second partial-dq area of the workspace
    const int32_t *sticker2 = sticker2_in, *bounds = bounds_in;
    if (dropout) {                       // the keep matrix is indexed by slot: tiles stay in the reference's order
      sticker2 = sticker; bounds = nullptr;
    } else if (!sticker2 || (!bounds && attend_tc_uses_bounds())) {
      int32_t *s2 = reinterpret_cast<int32_t *>(dq_part + rows * 64), *bd = attend_tc_uses_bounds() ? s2 + rows : nullptr;
      if (prep && (rc = chunk_possort_run(d, sticker, s2, bd, stream))) return rc;
      sticker2 = s2; bounds = bd;
    }
    AttendBwdTcParams t;
    t.qv = static_cast<const __nv_bfloat16 *>(qv); t.sticker = sticker; t.sticker2 = sticker2; t.bounds = bounds;
    t.do_comb = static_cast<const __nv_bfloat16 *>(do_comb); t.qscale = qscale; t.lse2 = lse2; t.dvec = dvec; t.qcmp = qcmp;
    t.keep_bits_t = dropout ? keep->bits_t : nullptr; t.keep_scale = dropout ? keep->scale : nullptr; t.slot_order = dropout ? 1 : 0;
    t.trace = g_fwd_trace; t.dq_out = dq_part; t.dv_out = dv_part; t.L = d.L; t.H = d.H; t.N = dr.N; t.n_chunks = dr.n_chunks;
    if ((g_bwd_parts & 2) && (rc = attend_bwd_tc_run(t, dr.BH, stream))) return rc;
    return (g_bwd_parts & 4) ? sum_rounds_run(d, dq_part, dv_part, dqv, 1, stream) : 0;
  }
  rc = (g_bwd_parts & 1) ? bwd_prep_run(d, do_comb, o_comb, dvec, stream) : 0;
  if (rc) return rc;
  AttendBwdParams p;
  p.qv = static_cast<const __nv_bfloat16 *>(qv); p.sticker = sticker; p.mask = d.masked ? mask : nullptr;
  p.do_comb = static_cast<const __nv_bfloat16 *>(do_comb); p.lse_tot = lse_tot; p.dvec = dvec;
  p.dq_part = dq_part; p.dv_part = dv_part; p.kind_stride = static_cast<int64_t>(rows) * 64;
  p.L = d.L; p.H = d.H; p.N = dr.N; p.n_chunks = dr.n_chunks; p.nb = d.nb; p.nwin = dr.nwin;
  p.causal = d.causal; p.masked = d.masked;
  p.row = dr.QV; p.ksep = d.separate_k ? 1 : 0;
  p.keep_bits_t = dropout ? keep->bits_t : nullptr; p.keep_scale = dropout ? keep->scale : nullptr;
  if (g_bwd_parts & 2) switch (d.C) {
    case 64: rc = launch_attend_bwd<64>(p, dr.BH, stream); break;
    case 128: rc = launch_attend_bwd<128>(p, dr.BH, stream); break;
    case 256: rc = launch_attend_bwd<256>(p, dr.BH, stream); break;
    default: return set_error("attend_bwd: chunk_len %d unsupported (64, 128, 256)", d.C);
  }
  if (rc) return rc;
  return (g_bwd_parts & 4) ? sum_rounds_run(d, dq_part, dv_part, dqv, dr.nwin + 1, stream) : 0;
}

}  // namespace lsh
