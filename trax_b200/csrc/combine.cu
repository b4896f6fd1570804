// HBM-bound row kernels around the chunk attention:
//   combine_fwd : EA:1988-1992  o = sum_r o_r * exp(logit_r - logsumexp_r(logit))      (+ lse_tot)
//   bwd_prep    : D[u][t] = do[t]·o[t]   (-Delta + dlse collapses to -w_r * (do·o))
//   sum_rounds  : App. B6: dq[t] = sum over the nh copies (and partial kinds) of each token
// One 64-wide bf16 row = 128 bytes = 8 lanes x 16 bytes; a warp moves 4 rows per step.
#include "common.cuh"

namespace lsh {

constexpr int ROW_THREADS = 256;

__device__ __forceinline__ void bf16x8_to_f32(const uint4 &v, float (&f)[8]) {
  float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 f32_to_bf16x8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]);
  v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
  return v;
}

// o_rounds (BH, nh, L, 64) bf16, logits (BH, nh, L) f32 -> o_comb (B, L, H, 64) bf16, lse_tot (BH, L)
// NH > 0: the round count is a compile-time constant (1, 2, 4, 8: every BASELINE config) — all rows of the token are in
// flight at once (one 16-byte load per round per lane) and no predicated-off round is issued; NH = 0: any round count.
template <int NH>
__global__ void __launch_bounds__(ROW_THREADS) combine_fwd_kernel(
    const __nv_bfloat16 *__restrict__ o_rounds, const float *__restrict__ logits,
    __nv_bfloat16 *__restrict__ o_comb, float *__restrict__ lse_tot, int L, int H, int nh_rt,
    int64_t total_rows, BwdPrepOut prep) {
  const int nh = NH > 0 ? NH : nh_rt;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * ROW_THREADS + threadIdx.x) >> 3;   // (u, t)
  const int ch = threadIdx.x & 7;
  if (row >= total_rows) return;                    // (whole warps: total_rows * 8 is a multiple of 32 whenever prep is set)
  const int64_t u = static_cast<uint32_t>(row) / static_cast<uint32_t>(L);   // rows < 2^31 (checked on the host): 32-bit divide
  const int t = static_cast<int>(row - u * L);
  const int64_t b = static_cast<uint32_t>(u) / static_cast<uint32_t>(H), h = u - b * H;
  const float *lg = logits + u * nh * L + t;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float lse;
  if constexpr (NH > 0) {
    float lgv[NH];
    uint4 ov[NH];
    const uint4 *src = reinterpret_cast<const uint4 *>(o_rounds + (u * NH * L + t) * 64) + ch;
#pragma unroll
    for (int r = 0; r < NH; ++r) {
      lgv[r] = __ldg(lg + static_cast<int64_t>(r) * L);
      ov[r] = __ldg(src + static_cast<int64_t>(r) * L * 8);
    }
    float mx = lgv[0];
#pragma unroll
    for (int r = 1; r < NH; ++r) mx = fmaxf(mx, lgv[r]);
    float den = 0.f;
#pragma unroll
    for (int r = 0; r < NH; ++r) den += __expf(lgv[r] - mx);
    lse = mx + __logf(den);                                           // logsumexp over rounds (EA:1991); MUFU-based exp / log (2 ulp)
#pragma unroll
    for (int r = 0; r < NH; ++r) {
      const float w = __expf(lgv[r] - lse);
      float f[8];
      bf16x8_to_f32(ov[r], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, f[i], acc[i]);
    }
  } else {
    float mx = -INFINITY;
    for (int r = 0; r < nh; ++r) mx = fmaxf(mx, __ldg(lg + static_cast<int64_t>(r) * L));
    float den = 0.f;
    for (int r = 0; r < nh; ++r) den += expf(__ldg(lg + static_cast<int64_t>(r) * L) - mx);
    lse = mx + logf(den);
    for (int r = 0; r < nh; ++r) {
      const float w = expf(__ldg(lg + static_cast<int64_t>(r) * L) - lse);
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(o_rounds + ((u * nh + r) * L + t) * 64) + ch);
      float f[8];
      bf16x8_to_f32(v, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, f[i], acc[i]);
    }
  }
  const uint4 packed = f32_to_bf16x8(acc);
  *(reinterpret_cast<uint4 *>(o_comb + ((b * L + t) * H + h) * 64) + ch) = packed;
  if (lse_tot != nullptr && ch == 0) lse_tot[u * L + t] = lse;
  if (prep.do_comb != nullptr) {
    // backward call: D = do . o of the row just combined (the bf16 values that were stored, same operation order as
    // bwd_prep_tc_kernel) and the other two per-token inputs of the gradient kernel — no second pass over o_comb
    float a[8], c[8];
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(static_cast<const __nv_bfloat16 *>(prep.do_comb) + ((b * L + t) * H + h) * 64) + ch), a);
    bf16x8_to_f32(packed, c);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[i], c[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ch == 0) {
      const int64_t o = u * L + t;
      const bool self_only = lse < -5e4f;
      prep.dvec[o] = -s;
      prep.lse2[o] = -(lse * kLog2e + (self_only ? 1e5f * kLog2e : 0.f));
      prep.qcmp[o] = static_cast<float>(t + 1) + (self_only ? 0.5f : 0.f);
    }
  }
}

int combine_fwd_run(const LshAttnDims &d, const void *o_rounds, const float *logits, void *o_comb,
                    float *lse_tot, cudaStream_t stream, const BwdPrepOut *prep_in) {
  Derived dr = derive(d);
  const int64_t rows = static_cast<int64_t>(dr.BH) * d.L;
  BwdPrepOut prep = {nullptr, nullptr, nullptr, nullptr};
  if (prep_in) {
    if ((rows * 8) % 32 != 0) return set_error("combine_fwd: fused backward preparation needs whole warps");
    prep = *prep_in;
  }
  const unsigned blocks = static_cast<unsigned>((rows * 8 + ROW_THREADS - 1) / ROW_THREADS);
  const __nv_bfloat16 *o = static_cast<const __nv_bfloat16 *>(o_rounds);
  __nv_bfloat16 *oc = static_cast<__nv_bfloat16 *>(o_comb);
  switch (d.nh) {
    case 1: combine_fwd_kernel<1><<<blocks, ROW_THREADS, 0, stream>>>(o, logits, oc, lse_tot, d.L, d.H, d.nh, rows, prep); break;
    case 2: combine_fwd_kernel<2><<<blocks, ROW_THREADS, 0, stream>>>(o, logits, oc, lse_tot, d.L, d.H, d.nh, rows, prep); break;
    case 4: combine_fwd_kernel<4><<<blocks, ROW_THREADS, 0, stream>>>(o, logits, oc, lse_tot, d.L, d.H, d.nh, rows, prep); break;
    case 8: combine_fwd_kernel<8><<<blocks, ROW_THREADS, 0, stream>>>(o, logits, oc, lse_tot, d.L, d.H, d.nh, rows, prep); break;
    default: combine_fwd_kernel<0><<<blocks, ROW_THREADS, 0, stream>>>(o, logits, oc, lse_tot, d.L, d.H, d.nh, rows, prep); break;
  }
  LSH_CHECK_LAUNCH("combine_fwd_kernel");
  return 0;
}

// D[u][t] = sum_d do[b,t,h,d] * o[b,t,h,d]; both (B, L, H, 64) bf16
__global__ void __launch_bounds__(ROW_THREADS) bwd_prep_kernel(
    const __nv_bfloat16 *__restrict__ do_comb, const __nv_bfloat16 *__restrict__ o_comb,
    float *__restrict__ dvec, int L, int H, int64_t total_rows) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * ROW_THREADS + threadIdx.x) >> 3;   // (b,t,h)
  const int ch = threadIdx.x & 7;
  const bool ok = row < total_rows;
  float s = 0.f;
  if (ok) {
    float a[8], c[8];
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(do_comb + row * 64) + ch), a);
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(o_comb + row * 64) + ch), c);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[i], c[i], s);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && ch == 0) {
    const uint32_t bt32 = static_cast<uint32_t>(row) / static_cast<uint32_t>(H);
    const int64_t h = static_cast<uint32_t>(row) - bt32 * static_cast<uint32_t>(H), bt = bt32;
    const int64_t b = static_cast<uint32_t>(bt) / static_cast<uint32_t>(L), t = bt - b * L;
    dvec[(b * H + h) * L + t] = s;
  }
}

int bwd_prep_run(const LshAttnDims &d, const void *do_comb, const void *o_comb, float *dvec,
                 cudaStream_t stream) {
  Derived dr = derive(d);
  const int64_t rows = static_cast<int64_t>(dr.BH) * d.L;
  const int64_t blocks = (rows * 8 + ROW_THREADS - 1) / ROW_THREADS;
  bwd_prep_kernel<<<static_cast<unsigned>(blocks), ROW_THREADS, 0, stream>>>(
      static_cast<const __nv_bfloat16 *>(do_comb), static_cast<const __nv_bfloat16 *>(o_comb), dvec,
      d.L, d.H, rows);
  LSH_CHECK_LAUNCH("bwd_prep_kernel");
  return 0;
}

// Token-level inputs of the tcgen05 backward kernel: -D = -(do.o), -lse2 = -log2e*lse_tot and the causal compare value.
// Rows whose log-sum-exp sits at the "-1e5" level saw only their own key class (EA:153-155): they keep exactly
// that class (compare against pos + 1.5) and their shift absorbs the 1e5 again.
__global__ void __launch_bounds__(ROW_THREADS) bwd_prep_tc_kernel(
    const __nv_bfloat16 *__restrict__ do_comb, const __nv_bfloat16 *__restrict__ o_comb,
    const float *__restrict__ lse_tot, float *__restrict__ dvec, float *__restrict__ lse2,
    float *__restrict__ qcmp, int L, int H, int64_t total_rows) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * ROW_THREADS + threadIdx.x) >> 3;   // (b,t,h)
  const int ch = threadIdx.x & 7;
  const bool ok = row < total_rows;
  float s = 0.f;
  if (ok) {
    float a[8], c[8];
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(do_comb + row * 64) + ch), a);
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(o_comb + row * 64) + ch), c);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[i], c[i], s);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && ch == 0) {
    const uint32_t bt32 = static_cast<uint32_t>(row) / static_cast<uint32_t>(H);
    const int64_t h = static_cast<uint32_t>(row) - bt32 * static_cast<uint32_t>(H), bt = bt32;
    const int64_t b = static_cast<uint32_t>(bt) / static_cast<uint32_t>(L), t = bt - b * L;
    const int64_t o = (b * H + h) * L + t;
    const float lse = lse_tot[o];
    const bool self_only = lse < -5e4f;
    // both are stored NEGATED: the kernel folds them into packed FFMA2 / FADD2 without a sign flip
    dvec[o] = -s;
    lse2[o] = -(lse * kLog2e + (self_only ? 1e5f * kLog2e : 0.f));
    qcmp[o] = static_cast<float>(t + 1) + (self_only ? 0.5f : 0.f);
  }
}

int bwd_prep_tc_run(const LshAttnDims &d, const void *do_comb, const void *o_comb, const float *lse_tot, float *dvec,
                    float *lse2, float *qcmp, cudaStream_t stream) {
  Derived dr = derive(d);
  const int64_t rows = static_cast<int64_t>(dr.BH) * d.L;
  const int64_t blocks = (rows * 8 + ROW_THREADS - 1) / ROW_THREADS;
  bwd_prep_tc_kernel<<<static_cast<unsigned>(blocks), ROW_THREADS, 0, stream>>>(
      static_cast<const __nv_bfloat16 *>(do_comb), static_cast<const __nv_bfloat16 *>(o_comb), lse_tot, dvec, lse2, qcmp,
      d.L, d.H, rows);
  LSH_CHECK_LAUNCH("bwd_prep_tc_kernel");
  return 0;
}

// Per-token key normalisation of EA:54-57, 229-231, done once per layer call for the tcgen05 kernels:
//   qscale[u][t]  = log2(e) / (r * 8),  r = sqrt(mean(q^2) + 1e-6)        (per-key factor; backward kernel)
//   qhat[u][t][:] = bf16(q / (8 r))                                        (the reference's normalised key)
//   rowmeta[u][t] = {a, m2}:  a = 8 r log2(e) restores the un-normalised query, q_i·k_j = 8 r_i (qhat_i·qhat_j), so the
//                   forward kernel uses ONE gathered row as query and as key and its softmax needs no per-key data;
//                   m2 = a |qhat|^2 = the row's self score in the log2 domain (the analytic softmax shift).
__global__ void __launch_bounds__(ROW_THREADS) qscale_kernel(const __nv_bfloat16 *__restrict__ qv,
                                                            float *__restrict__ qscale, float2 *__restrict__ rowmeta,
                                                            __nv_bfloat16 *__restrict__ qhat, int L, int H,
                                                            int64_t total_rows) {
  // rows in (unit, position) order: the three outputs are written contiguously; the q reads are 128-byte pieces at a
  // stride of H * 256 bytes (whole sectors either way)
  const int64_t ut = (static_cast<int64_t>(blockIdx.x) * ROW_THREADS + threadIdx.x) >> 3;    // (u, t)
  const int ch = threadIdx.x & 7;
  const bool ok = ut < total_rows;
  const int64_t u = static_cast<uint32_t>(ut) / static_cast<uint32_t>(L), t = ut - u * L;
  const int64_t b = static_cast<uint32_t>(u) / static_cast<uint32_t>(H), h = u - b * H;
  const int64_t row = (b * L + t) * H + h;                                                      // row of qv
  float s = 0.f;
  float a[8];
  if (ok) {
    bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(qv + row * 128) + ch), a);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[i], a[i], s);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float r = sqrtf(s * (1.0f / 64) + 1e-6f);
  float s2 = 0.f;
  if (ok && qhat != nullptr) {
    const float c = 0.125f / r;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = a[i] * c;
    const uint4 packed = f32_to_bf16x8(f);
    bf16x8_to_f32(packed, f);                     // |qhat|^2 of the rounded values the tensor cores will see
#pragma unroll
    for (int i = 0; i < 8; ++i) s2 = fmaf(f[i], f[i], s2);
    *(reinterpret_cast<uint4 *>(qhat + ut * 64) + ch) = packed;
  }
  s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
  if (ok && ch == 0) {
    if (qscale != nullptr) qscale[ut] = 0.125f * kLog2e / r;
    if (rowmeta != nullptr) {
      const float am = 8.f * r * kLog2e;
      rowmeta[ut] = make_float2(am, am * s2);
    }
  }
}

int qscale_run(const LshAttnDims &d, const void *qv, float *qscale, float2 *rowmeta, void *qhat, cudaStream_t stream) {
  Derived dr = derive(d);
  const int64_t rows = static_cast<int64_t>(dr.BH) * d.L;
  const int64_t blocks = (rows * 8 + ROW_THREADS - 1) / ROW_THREADS;
  qscale_kernel<<<static_cast<unsigned>(blocks), ROW_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16 *>(qv), qscale, rowmeta,
                                                                          static_cast<__nv_bfloat16 *>(qhat), d.L, d.H, rows);
  LSH_CHECK_LAUNCH("qscale_kernel");
  return 0;
}

// dqv[b,t,h,0:64]   = sum_r sum_kind dq_part[kind][u][r*L+t][:]
// dqv[b,t,h,64:128] = sum_r dv_part[u][r*L+t][:]
// NH > 0: tcgen05 path with a compile-time round count (one dq and one dv row per round, all 2 NH loads in flight).
template <int NH>
__global__ void __launch_bounds__(ROW_THREADS) sum_rounds_kernel(
    const __nv_bfloat16 *__restrict__ dq_part, const __nv_bfloat16 *__restrict__ dv_part,
    __nv_bfloat16 *__restrict__ dqv, int L, int H, int nh, int n_kinds, int64_t kind_stride,
    int64_t total_rows, int row_elems, int ksep) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * ROW_THREADS + threadIdx.x) >> 3;   // (u, t)
  const int ch = threadIdx.x & 7;
  if (row >= total_rows) return;
  const int64_t u = static_cast<uint32_t>(row) / static_cast<uint32_t>(L);   // rows < 2^31 (checked on the host): 32-bit divide
  const int t = static_cast<int>(row - u * L);
  const int64_t b = static_cast<uint32_t>(u) / static_cast<uint32_t>(H), h = u - b * H;
  float aq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, av[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float ak[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // separate keys: the key-side kind is the cotangent of k, not of q
  if constexpr (NH > 0) {
    uint4 vq[NH], vv[NH];
    const int64_t off0 = (u * NH * L + t) * 64;
#pragma unroll
    for (int r = 0; r < NH; ++r) {
      vq[r] = __ldg(reinterpret_cast<const uint4 *>(dq_part + off0) + static_cast<int64_t>(r) * L * 8 + ch);
      vv[r] = __ldg(reinterpret_cast<const uint4 *>(dv_part + off0) + static_cast<int64_t>(r) * L * 8 + ch);
    }
#pragma unroll
    for (int r = 0; r < NH; ++r) {
      float f[8];
      bf16x8_to_f32(vq[r], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) aq[i] += f[i];
      bf16x8_to_f32(vv[r], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] += f[i];
    }
  } else {
    for (int r = 0; r < nh; ++r) {
      const int64_t off = ((u * nh + r) * L + t) * 64;
      float f[8];
      for (int k = 0; k < n_kinds; ++k) {
        bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(dq_part + k * kind_stride + off) + ch), f);
        if (ksep && k == n_kinds - 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) ak[i] += f[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) aq[i] += f[i];
        }
      }
      bf16x8_to_f32(__ldg(reinterpret_cast<const uint4 *>(dv_part + off) + ch), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] += f[i];
    }
  }
  __nv_bfloat16 *dst = dqv + ((b * L + t) * H + h) * row_elems;
  *(reinterpret_cast<uint4 *>(dst) + ch) = f32_to_bf16x8(aq);
  *(reinterpret_cast<uint4 *>(dst + 64) + ch) = f32_to_bf16x8(av);
  if (ksep) *(reinterpret_cast<uint4 *>(dst + 128) + ch) = f32_to_bf16x8(ak);
}

int sum_rounds_run(const LshAttnDims &d, const void *dq_part, const void *dv_part, void *dqv,
                   int n_kinds, cudaStream_t stream) {
  Derived dr = derive(d);
  const int64_t rows = static_cast<int64_t>(dr.BH) * d.L;
  const int64_t blocks = (rows * 8 + ROW_THREADS - 1) / ROW_THREADS;
  const unsigned nb = static_cast<unsigned>(blocks);
  const __nv_bfloat16 *dq = static_cast<const __nv_bfloat16 *>(dq_part), *dv = static_cast<const __nv_bfloat16 *>(dv_part);
  __nv_bfloat16 *out = static_cast<__nv_bfloat16 *>(dqv);
  const int64_t kstride = static_cast<int64_t>(dr.BH) * dr.N * 64;
  const int ksep = d.separate_k ? 1 : 0;
  const int fast = (n_kinds == 1) ? d.nh : 0;
#define LSH_SUM_ROUNDS(NH_) sum_rounds_kernel<NH_><<<nb, ROW_THREADS, 0, stream>>>(dq, dv, out, d.L, d.H, d.nh, n_kinds, kstride, rows, dr.QV, ksep)
  switch (fast) {
    case 1: LSH_SUM_ROUNDS(1); break;
    case 2: LSH_SUM_ROUNDS(2); break;
    case 4: LSH_SUM_ROUNDS(4); break;
    case 8: LSH_SUM_ROUNDS(8); break;
    default: LSH_SUM_ROUNDS(0); break;
  }
#undef LSH_SUM_ROUNDS
  LSH_CHECK_LAUNCH("sum_rounds_kernel");
  return 0;
}

}  // namespace lsh
