// Chunked look-back attention, forward — replaces EA:1958-1986:
//   gather by sticker (EA:1959-1960), `attend` (EA:163-268: key length-normalisation EA:229-231,
//   look_adjacent EA:239-241, dots EA:244, masks EA:145-160, logsumexp/exp EA:251-252, P·V EA:265),
//   and the un-sort of EA:1985-1986 (each query row is written straight to its ticker slot).
//
// v1 compute path: bf16 mma.sync m16n8k16 with fp32 accumulation, one CTA per query chunk
// (chunk_len/16 warps, one 16-row stripe each), K/V window staged once in swizzled shared memory
// by cp.async row gathers, flash-style online softmax over 64-key blocks.
#include <stdlib.h>
#include <string.h>

#include "attend_params.cuh"

namespace lsh {

long long *g_fwd_trace = nullptr;   // set through lsh_debug_set_trace (profiling aid)

template <int C>
__global__ void __launch_bounds__(2 * C) attend_fwd_kernel(const AttendFwdParams p) {
  constexpr int NT = 2 * C;            // threads
  constexpr int D = 64;                // dq == dv == 64
  const int QVROW = p.row;             // elements per (token, head) row of qv: q | v (| k)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int W = C * p.nwin;
  uint8_t *Ks = smem;                                  // [W][64] bf16 swizzled (raw q rows: queries AND keys)
  uint8_t *Vs = smem + static_cast<size_t>(W) * 128;   // [W][64] bf16 swizzled
  int *kinfo = reinterpret_cast<int *>(Vs + static_cast<size_t>(W) * 128);   // [W] kv_info (+1 applied)
  int *spos = kinfo + W;                               // [W] 0-based positions
  int *tkq = spos + W;                                 // [C] ticker of the query rows
  float *kscale = reinterpret_cast<float *>(tkq + C);  // [W] 1 / (sqrt(mean(q^2)+eps) * sqrt(dq)) per key row
  uint8_t *Qs = reinterpret_cast<uint8_t *>(kscale + W);   // [C][64] bf16 swizzled: the chunk's queries (separate keys only)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u = blockIdx.x / p.n_chunks, c = blockIdx.x % p.n_chunks;
  const int b = u / p.H, h = u % p.H;
  const int32_t *stk = p.sticker + static_cast<int64_t>(u) * p.N;

  // ---- window metadata (EA:1958, 1964-1972, 198-206, 239-241) -------------------------------------
  for (int j = tid; j < W; j += NT) {
    const int blk = j / C;
    int src_chunk = c + blk - p.nb;
    src_chunk = (src_chunk % p.n_chunks + p.n_chunks) % p.n_chunks;   // cyclic (EA:141)
    const int tk = stk[src_chunk * C + (j - blk * C)];
    const int pos = tk % p.L;
    bool valid = true;
    if (p.masked) valid = p.mask[static_cast<int64_t>(b) * p.L + pos] != 0;
    kinfo[j] = (valid ? pos : -pos) + 1;
    spos[j] = pos;
    if (blk == p.nb) tkq[j - blk * C] = tk;
  }
  __syncthreads();

  // ---- gather rows: 16 lanes move one 256-byte (q|v) row ------------------------------------------
  {
    const uint32_t ks_base = smem_u32(Ks), vs_base = smem_u32(Vs);
    for (int i = tid; i < W * 16; i += NT) {
      const int j = i >> 4, ch = i & 15;
      // shared-QK: the q half is query AND key; separate keys (EA:1160-1162): the key tile comes from the k columns
      const __nv_bfloat16 *src =
          p.qv + ((static_cast<int64_t>(b) * p.L + spos[j]) * p.H + h) * QVROW + ch * 8 + ((p.ksep && ch < 8) ? 128 : 0);
      const uint32_t dst = (ch < 8) ? ks_base + swz(j, ch) : vs_base + swz(j, ch - 8);
      cp_async16(dst, src);
    }
    if (p.ksep) {
      const uint32_t qs_base = smem_u32(Qs);
      for (int i = tid; i < C * 8; i += NT) {
        const int j = i >> 3, ch = i & 7;
        cp_async16(qs_base + swz(j, ch), p.qv + ((static_cast<int64_t>(b) * p.L + spos[p.nb * C + j]) * p.H + h) * QVROW + ch * 8);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();

  // ---- Q fragments (un-normalised queries) of this warp's 16 rows ---------------------------------
  const int qrow0 = p.nb * C + warp * 16;
  uint32_t qa[4][4];
  {
    const uint32_t q_base = p.ksep ? smem_u32(Qs) : smem_u32(Ks);
    const int mi = lane >> 3;
    const int row = (p.ksep ? warp * 16 : qrow0) + (lane & 7) + 8 * (mi & 1);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldmatrix_x4(q_base + swz(row, ks * 2 + (mi >> 1)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
  }

  // ---- keys: k = q / sqrt(mean(q^2) + 1e-6) / sqrt(dq)  (EA:54-57, 229-231) -------------------------
  // The rows stay un-normalised in shared memory (exact bf16 operands); the per-key factor is applied
  // to the fp32 score column instead, which saves one bf16 rounding of every key.
  for (int j = tid >> 3; j < W; j += NT / 8) {
    const int ch = tid & 7;
    const uint4 raw = *reinterpret_cast<const uint4 *>(Ks + swz(j, ch));
    float2 f0 = unpack_bf16(raw.x), f1 = unpack_bf16(raw.y), f2 = unpack_bf16(raw.z), f3 = unpack_bf16(raw.w);
    float ss = f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y + f2.x * f2.x + f2.y * f2.y +
               f3.x * f3.x + f3.y * f3.y;
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    // separate keys are not length-normalised (EA:229-231 apply to shared-QK only), just divided by sqrt(dq) (EA:232)
    if (ch == 0) kscale[j] = p.ksep ? 0.125f : 0.125f / sqrtf(ss * (1.0f / D) + 1e-6f);   // 1/sqrt(64) = 0.125
  }
  __syncthreads();

  // ---- main loop over 64-key blocks ---------------------------------------------------------------
  const int g = lane >> 2, t = lane & 3;
  const float qi0 = static_cast<float>(spos[qrow0 + g] + 1);       // q_info = pos + 1 (EA:201)
  const float qi1 = static_cast<float>(spos[qrow0 + g + 8] + 1);
  float oacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const uint32_t ks_base = smem_u32(Ks), vs_base = smem_u32(Vs);
  const int n_kb = W / 64;
  // attention dropout (EA:254-262): rows of the (C, W) keep matrix of my two query slots; tiles are in slot order here
  const uint32_t *kb0 = p.keep_bits ? p.keep_bits + static_cast<size_t>(warp * 16 + g) * (W / 32) : nullptr;
  const uint32_t *kb1 = p.keep_bits ? kb0 + 8 * (W / 32) : nullptr;
  for (int kb = 0; kb < n_kb; ++kb) {
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int ntp = 0; ntp < 4; ++ntp) {
        const int mi = lane >> 3;
        const int krow = kb * 64 + ntp * 16 + (lane & 7) + 8 * (mi >> 1);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(ks_base + swz(krow, ks * 2 + (mi & 1)), b0, b1, b2, b3);
        mma_bf16(s[2 * ntp], qa[ks], b0, b1);
        mma_bf16(s[2 * ntp + 1], qa[ks], b2, b3);
      }
    }
    // masks (EA:145-160), fp32 arithmetic, same order: causal, self, padding
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = kb * 64 + nt * 8 + 2 * t + (e & 1);
        const float ki = static_cast<float>(kinfo[col]);
        const float qi = (e < 2) ? qi0 : qi1;
        float v = s[nt][e] * kscale[col];
        if (p.causal && qi < ki) v = v - 1e9f;
        if (!p.ksep && qi == ki) v = v - 1e5f;             // exclude_self = share_qk (EA:1175-1178)
        if (p.masked && ki < 0.f) v = v - 1e9f;
        s[nt][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = quad_max(mx0); mx1 = quad_max(mx1);
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float al0 = exp2f((m0 - mn0) * kLog2e), al1 = exp2f((m1 - mn1) * kLog2e);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f((s[nt][0] - mn0) * kLog2e), p1 = exp2f((s[nt][1] - mn0) * kLog2e);
      const float p2 = exp2f((s[nt][2] - mn1) * kLog2e), p3 = exp2f((s[nt][3] - mn1) * kLog2e);
      rs0 += p0 + p1; rs1 += p2 + p3;                       // the log-sum-exp does not see the dropout (EA:251-262)
      float d0 = p0, d1 = p1, d2 = p2, d3 = p3;
      if (kb0) {
        const int sh = (nt & 3) * 8 + 2 * t;
        const uint32_t w0 = __ldg(kb0 + kb * 2 + (nt >> 2)) >> sh, w1 = __ldg(kb1 + kb * 2 + (nt >> 2)) >> sh;
        d0 = (w0 & 1u) ? p0 : 0.f; d1 = (w0 & 2u) ? p1 : 0.f;
        d2 = (w1 & 1u) ? p2 : 0.f; d3 = (w1 & 2u) ? p3 : 0.f;
      }
      const int kk = nt >> 1;
      if ((nt & 1) == 0) { pa[kk][0] = pack_bf16(d0, d1); pa[kk][1] = pack_bf16(d2, d3); }
      else               { pa[kk][2] = pack_bf16(d0, d1); pa[kk][3] = pack_bf16(d2, d3); }
    }
    l0 = l0 * al0 + rs0; l1 = l1 * al1 + rs1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { oacc[i][0] *= al0; oacc[i][1] *= al0; oacc[i][2] *= al1; oacc[i][3] *= al1; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int ntp = 0; ntp < 4; ++ntp) {
        const int mi = lane >> 3;
        const int vrow = kb * 64 + kk * 16 + (lane & 7) + 8 * (mi & 1);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(vs_base + swz(vrow, ntp * 2 + (mi >> 1)), b0, b1, b2, b3);
        mma_bf16(oacc[2 * ntp], pa[kk], b0, b1);
        mma_bf16(oacc[2 * ntp + 1], pa[kk], b2, b3);
      }
    }
  }
  l0 = quad_sum(l0); l1 = quad_sum(l1);
  const float ksc = p.keep_bits ? __ldg(p.keep_scale) : 1.f;        // keep / (1 - rate): folded into the normalisation
  const float il0 = ksc / l0, il1 = ksc / l1;

  // ---- epilogue: stage the 16x64 stripe in shared memory, then 128-byte row stores -----------------
  __syncthreads();   // every warp is done reading Ks/Vs
  {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t *>(Ks + swz(r0, nt) + 4 * t) = pack_bf16(oacc[nt][0] * il0, oacc[nt][1] * il0);
      *reinterpret_cast<uint32_t *>(Ks + swz(r1, nt) + 4 * t) = pack_bf16(oacc[nt][2] * il1, oacc[nt][3] * il1);
    }
    if (t == 0) {
      float *lse_u = p.lse + static_cast<int64_t>(u) * p.N;
      lse_u[tkq[r0]] = m0 + logf(l0);
      lse_u[tkq[r1]] = m1 + logf(l1);
    }
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int piece = it * 32 + lane;
    const int row = warp * 16 + (piece >> 3), ch = piece & 7;
    const int tk = tkq[row];
    const int round = tk / p.L, pos = tk - round * p.L;
    __nv_bfloat16 *dst = p.o + b * p.o_sb + h * p.o_sh + round * p.o_sr + pos * p.o_sp + ch * 8;
    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(Ks + swz(row, ch));
  }
}

template <int C>
static int launch_attend_fwd(const AttendFwdParams &p, int BH, cudaStream_t stream) {
  const int W = C * p.nwin;
  size_t smem = static_cast<size_t>(W) * 256 + static_cast<size_t>(W) * 12 + C * 4 + (p.ksep ? C * 128 : 0);
  if (smem > 227 * 1024) return set_error("attend_fwd: window of %d keys needs %zu B shared memory", W, smem);
  LSH_OPT_IN_SMEM(attend_fwd_kernel<C>);
  attend_fwd_kernel<C><<<BH * p.n_chunks, 2 * C, smem, stream>>>(p);
  LSH_CHECK_LAUNCH("attend_fwd_kernel");
  return 0;
}

static bool force_mma_fwd() {
  static const bool f = [] { const char *e = getenv("LSH_ATTN_FWD"); return e && strcmp(e, "mma") == 0; }();
  return f;
}
// tcgen05 path for the long-sequence shape (chunk 128, 2-chunk window); LSH_ATTN_FWD=mma forces the mma.sync path
// (L % 128 == 0: every chunk lies inside one hash round, so positions inside a tile are unique — the position-sorted
// interval masks rely on that; other lengths take the mma.sync path)
bool attend_fwd_uses_tc(const LshAttnDims &d) {
  return d.C == 128 && 1 + d.nb + d.na == 2 && d.L % 128 == 0 && !d.separate_k && !force_mma_fwd();
}

bool attend_fwd_tc_uses_bounds();
bool attend_bwd_tc_uses_bounds();
bool attend_tc_uses_bounds() { return attend_fwd_tc_uses_bounds() || attend_bwd_tc_uses_bounds(); }
int qscale_run(const LshAttnDims &d, const void *qv, float *qscale, float2 *rowmeta, void *qhat, cudaStream_t stream);
int chunk_possort_run(const LshAttnDims &d, const int32_t *sticker, int32_t *sticker2, int32_t *bounds, cudaStream_t stream);

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }
size_t fwd_aux_bytes(const LshAttnDims &d) {
  Derived dr = derive(d);
  const size_t rows = static_cast<size_t>(dr.BH) * d.L;
  return align256(rows * 4) + align256(rows * 8) + align256(rows * 128) + 2 * align256(static_cast<size_t>(dr.BH) * dr.N * 4) +
         align256(static_cast<size_t>(dr.BH) * dr.N * 8 + 16) + 256;
}
FwdAux fwd_aux_carve(const LshAttnDims &d, void *ws) {
  Derived dr = derive(d);
  const size_t rows = static_cast<size_t>(dr.BH) * d.L;
  char *b = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  FwdAux a;
  a.qscale = reinterpret_cast<float *>(b); b += align256(rows * 4);
  a.rowmeta = reinterpret_cast<float2 *>(b); b += align256(rows * 8);
  a.qhat = b; b += align256(rows * 128);
  a.sticker2 = reinterpret_cast<int32_t *>(b); b += align256(static_cast<size_t>(dr.BH) * dr.N * 4);
  a.bounds = reinterpret_cast<int32_t *>(b); b += align256(static_cast<size_t>(dr.BH) * dr.N * 4);
  a.redo = reinterpret_cast<int *>(b);
  return a;
}
// Everything the attention kernels need besides qv and sticker: per-token scales (always: the backward kernel reads
// qscale), and for the tcgen05 forward path the normalised keys and the position-sorted chunks.
int fwd_aux_prepare(const LshAttnDims &d, const void *qv, const int32_t *sticker, const FwdAux &aux, cudaStream_t stream,
                    bool scales_done) {
  const bool tc = attend_fwd_uses_tc(d);
  if (!scales_done && !d.separate_k) {     // (a forward call that hashes gets them from the hash kernel, which already holds q;
                                           //  separate keys are not normalised: no per-token scales at all)
    if (int rc = qscale_run(d, qv, aux.qscale, tc ? aux.rowmeta : nullptr, tc ? aux.qhat : nullptr, stream)) return rc;
  }
  if (tc && sticker) return chunk_possort_run(d, sticker, aux.sticker2, attend_tc_uses_bounds() ? aux.bounds : nullptr, stream);
  return 0;
}

int attend_fwd_run(const LshAttnDims &d, const void *qv, const int32_t *sticker, const uint8_t *mask,
                   void *o, int64_t o_sb, int64_t o_sh, int64_t o_sr, int64_t o_sp, float *lse,
                   const FwdAux *aux, const AttnKeep *keep, cudaStream_t stream) {
  Derived dr = derive(d);
  AttendFwdParams p;
  p.qv = static_cast<const __nv_bfloat16 *>(qv); p.sticker = sticker;
  p.mask = d.masked ? mask : nullptr; p.o = static_cast<__nv_bfloat16 *>(o);
  p.o_sb = o_sb; p.o_sh = o_sh; p.o_sr = o_sr; p.o_sp = o_sp; p.lse = lse; p.trace = g_fwd_trace;
  p.qscale = aux ? aux->qscale : nullptr;
  p.qhat = aux ? static_cast<const __nv_bfloat16 *>(aux->qhat) : nullptr;
  p.rowmeta = aux ? aux->rowmeta : nullptr;
  p.sticker2 = aux ? aux->sticker2 : nullptr;
  p.bounds = aux ? aux->bounds : nullptr;
  p.redo = aux ? aux->redo : nullptr;
  p.keep_bits = keep ? keep->bits : nullptr; p.keep_scale = keep ? keep->scale : nullptr;
  p.L = d.L; p.H = d.H; p.N = dr.N; p.n_chunks = dr.n_chunks; p.nb = d.nb; p.nwin = dr.nwin;
  p.causal = d.causal; p.masked = d.masked;
  p.row = dr.QV; p.ksep = d.separate_k ? 1 : 0;
  if (d.masked && !mask) return set_error("attend_fwd: dims.masked set but mask == NULL");
  if (attend_fwd_uses_tc(d)) {
    if (!aux) return set_error("attend_fwd: the tcgen05 path needs the auxiliary workspace");
    return attend_fwd_tc_run(p, dr.BH, stream);
  }
  switch (d.C) {
    case 32: return launch_attend_fwd<32>(p, dr.BH, stream);
    case 64: return launch_attend_fwd<64>(p, dr.BH, stream);
    case 128: return launch_attend_fwd<128>(p, dr.BH, stream);
    case 256: return launch_attend_fwd<256>(p, dr.BH, stream);
    default: return set_error("attend_fwd: chunk_len %d unsupported (32, 64, 128, 256)", d.C);
  }
}

}  // namespace lsh
