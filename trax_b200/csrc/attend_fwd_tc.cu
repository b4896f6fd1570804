// Chunked look-back attention, forward, on the 5th-generation tensor cores (tcgen05 + TMEM).
// Same contract as attend_fwd.cu (EA:1958-1986); specialised for chunk_len 128 with a 2-chunk window
// (n_chunks_before + n_chunks_after == 1), dq = dv = 64 — the shape of every long-sequence config.
//
// One CTA = one query chunk; two CTAs share an SM (256 TMEM columns each), so one CTA's softmax
// overlaps the other's MMAs and gathers.
//   smem : window rows [256][64] bf16 for q (queries AND keys, un-normalised) and v, SWIZZLE_128B
//          (cp.async row gathers with the chunk ^= row&7 pattern = the UMMA canonical K-major /
//          MN-major SW128 layouts)
//   TMEM : S = Q·K^T  128 lanes x 256 fp32 columns [0,256)           (tcgen05.mma SS, M128 N256 K16 x4)
//          P (bf16, 2 per column) written back in place to columns [0,128) by the softmax warps
//          O = P·V    128 x 64 fp32 at columns [128,192)              (tcgen05.mma TS, M128 N64 K16 x16)
//   warps 0-3: one query row per thread (TMEM lane = row): key scale + masks + 2-pass softmax, epilogue
//   warp  4  : TMEM allocation, MMA issue (one lane), commit -> mbarriers
//   warp  5  : helps with the gathers
#include "attend_params.cuh"
#include "tc_common.cuh"

namespace lsh {

constexpr int TC_C = 128, TC_W = 256, TC_THREADS = 192;
constexpr uint32_t TC_TMEM_COLS = 256;
constexpr uint32_t TC_IDESC_S = make_idesc_bf16(128, 256, 0, 0);    // Q (K-major) x K (K-major)
constexpr uint32_t TC_IDESC_O = make_idesc_bf16(128, 64, 0, 1);     // P (TMEM)    x V (MN-major)

struct __align__(16) TcShared {
  float kinfo[TC_W];     // kv_info (+1 applied, negative = padding) as fp32 (EA:148-149 compares in fp32)
  float kscl[TC_W];      // kscale * log2(e)
  int spos[TC_W];
  int tkq[TC_C];
  uint64_t bar_s, bar_p, bar_o;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(TC_THREADS, 2) attend_fwd_tc_kernel(const AttendFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16-byte aligned: round up to the 1024-byte swizzle atom
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *Ks = smem;                       // [256][128 B]
  uint8_t *Vs = smem + TC_W * 128;          // [256][128 B]
  TcShared &sh = *reinterpret_cast<TcShared *>(smem + 2 * TC_W * 128);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u = blockIdx.x / p.n_chunks, c = blockIdx.x % p.n_chunks;
  const int b = u / p.H, h = u % p.H;
  const int32_t *stk = p.sticker + static_cast<int64_t>(u) * p.N;

  if (warp == 4) tmem_alloc(&sh.tmem_base, TC_TMEM_COLS);
  if (tid == 0) {
    mbar_init(&sh.bar_s, 1);
    mbar_init(&sh.bar_p, 128);
    mbar_init(&sh.bar_o, 1);
    fence_mbar_init();
  }
  // ---- window metadata ------------------------------------------------------------------------------
  for (int j = tid; j < TC_W; j += TC_THREADS) {
    const int blk = j >> 7;
    int src_chunk = c + blk - p.nb;
    src_chunk = (src_chunk % p.n_chunks + p.n_chunks) % p.n_chunks;
    const int tk = stk[src_chunk * TC_C + (j & 127)];
    const int pos = tk % p.L;
    bool valid = true;
    if (p.masked) valid = p.mask[static_cast<int64_t>(b) * p.L + pos] != 0;
    sh.kinfo[j] = static_cast<float>((valid ? pos : -pos) + 1);
    sh.spos[j] = pos;
    if (blk == p.nb) sh.tkq[j & 127] = tk;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  // ---- gather q|v rows -------------------------------------------------------------------------------
  const uint32_t ks_base = smem_u32(Ks), vs_base = smem_u32(Vs);
  for (int i = tid; i < TC_W * 16; i += TC_THREADS) {
    const int j = i >> 4, ch = i & 15;
    const __nv_bfloat16 *src = p.qv + ((static_cast<int64_t>(b) * p.L + sh.spos[j]) * p.H + h) * 128 + ch * 8;
    cp_async16((ch < 8) ? ks_base + swz(j, ch) : vs_base + swz(j, ch - 8), src);
  }
  cp_async_commit();
  cp_async_wait<0>();
  fence_proxy_async();          // this thread's gathered bytes -> async proxy (UMMA operand reads)
  __syncthreads();

  if (warp == 4) {
    // ---- MMA issuer --------------------------------------------------------------------------------
    if (lane == 0) {
      // S[128 x 256] = Q (window rows 128 * nb .. +127) x K^T (all 256 rows); K = 64 = 4 x UMMA_K
      const uint32_t q_addr = ks_base + p.nb * TC_C * 128;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t a_desc = make_smem_desc(q_addr + ks * 32, 16, 1024);
        const uint64_t b_desc = make_smem_desc(ks_base + ks * 32, 16, 1024);
        umma_ss(tmem, a_desc, b_desc, TC_IDESC_S, ks > 0);
      }
      umma_commit(&sh.bar_s);
      // O[128 x 64] = P (TMEM cols [0,128), 8 columns per K=16 step) x V (MN-major, 16 key rows per step)
      mbar_wait(&sh.bar_p, 0);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const uint64_t b_desc = make_smem_desc(vs_base + kk * 2048, 1024, 1024);
        umma_ts(tmem + 128, tmem + kk * 8, b_desc, TC_IDESC_O, kk > 0);
      }
      umma_commit(&sh.bar_o);
    }
    __syncwarp();
  } else if (warp < 4) {
    // ---- key scale 1/(sqrt(mean(q^2)+eps)*sqrt(dq)) * log2(e), 2 rows per thread -------------------------
    for (int j = tid; j < TC_W; j += 128) {
      float ss = 0.f;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(Ks + swz(j, ch));
        const float2 f0 = unpack_bf16(raw.x), f1 = unpack_bf16(raw.y), f2 = unpack_bf16(raw.z), f3 = unpack_bf16(raw.w);
        ss += f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y + f2.x * f2.x + f2.y * f2.y + f3.x * f3.x + f3.y * f3.y;
      }
      sh.kscl[j] = 0.125f * kLog2e / sqrtf(ss * (1.0f / 64) + 1e-6f);
    }
    asm volatile("bar.sync 1, 128;\n" ::: "memory");   // kscl visible to the 4 softmax warps

    const int row = warp * 32 + lane;                      // query row == TMEM lane
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const float qi = static_cast<float>(sh.spos[p.nb * TC_C + row] + 1);
    constexpr float kBig = 1e9f * kLog2e, kSelf = 1e5f * kLog2e;   // masks in the log2 domain
    mbar_wait(&sh.bar_s, 0);
    tc_fence_after();
    // pass 1: row maximum of the scaled + masked scores
    float m = -INFINITY;
    for (int k = 0; k < 8; ++k) {
      uint32_t r[32];
      tmem_ld32(t_lane + k * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int c4 = 0; c4 < 32; c4 += 4) {
        const float4 ki4 = *reinterpret_cast<const float4 *>(&sh.kinfo[k * 32 + c4]);
        const float4 sc4 = *reinterpret_cast<const float4 *>(&sh.kscl[k * 32 + c4]);
        const float kis[4] = {ki4.x, ki4.y, ki4.z, ki4.w}, scs[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = __uint_as_float(r[c4 + e]) * scs[e];
          if (p.causal && qi < kis[e]) v -= kBig;
          if (qi == kis[e]) v -= kSelf;
          if (p.masked && kis[e] < 0.f) v -= kBig;
          m = fmaxf(m, v);
        }
      }
    }
    // pass 2: P = exp2(t - m) -> bf16, written in place (columns [16k, 16k+16) after reading [32k, 32k+32))
    float l = 0.f;
    for (int k = 0; k < 8; ++k) {
      uint32_t r[32];
      tmem_ld32(t_lane + k * 32, r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int c4 = 0; c4 < 32; c4 += 4) {
        const float4 ki4 = *reinterpret_cast<const float4 *>(&sh.kinfo[k * 32 + c4]);
        const float4 sc4 = *reinterpret_cast<const float4 *>(&sh.kscl[k * 32 + c4]);
        const float kis[4] = {ki4.x, ki4.y, ki4.z, ki4.w}, scs[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = __uint_as_float(r[c4 + e]) * scs[e];
          if (p.causal && qi < kis[e]) v -= kBig;
          if (qi == kis[e]) v -= kSelf;
          if (p.masked && kis[e] < 0.f) v -= kBig;
          pv[e] = exp2f(v - m);
        }
        l += (pv[0] + pv[1]) + (pv[2] + pv[3]);
        pk[c4 >> 1] = pack_bf16(pv[0], pv[1]);
        pk[(c4 >> 1) + 1] = pack_bf16(pv[2], pv[3]);
      }
      tmem_st16(t_lane + k * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&sh.bar_p);

    // ---- epilogue: O / l -> bf16 row, written straight to its ticker slot (EA:1985-1986) ---------------
    mbar_wait(&sh.bar_o, 0);
    tc_fence_after();
    const float il = 1.f / l;
    const int tk = sh.tkq[row];
    const int round = tk / p.L, pos = tk - round * p.L;
    __nv_bfloat16 *dst = p.o + b * p.o_sb + h * p.o_sh + round * p.o_sr + pos * p.o_sp;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
      tmem_ld32(t_lane + 128 + half * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 v;
        v.x = pack_bf16(__uint_as_float(r[8 * q4 + 0]) * il, __uint_as_float(r[8 * q4 + 1]) * il);
        v.y = pack_bf16(__uint_as_float(r[8 * q4 + 2]) * il, __uint_as_float(r[8 * q4 + 3]) * il);
        v.z = pack_bf16(__uint_as_float(r[8 * q4 + 4]) * il, __uint_as_float(r[8 * q4 + 5]) * il);
        v.w = pack_bf16(__uint_as_float(r[8 * q4 + 6]) * il, __uint_as_float(r[8 * q4 + 7]) * il);
        *reinterpret_cast<uint4 *>(dst + half * 32 + q4 * 8) = v;
      }
    }
    p.lse[static_cast<int64_t>(u) * p.N + tk] = (m + log2f(l)) * kLn2;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TC_TMEM_COLS);
}

int attend_fwd_tc_run(const AttendFwdParams &p, int BH, cudaStream_t stream) {
  // > 76 KB so that at most two CTAs (2 x 256 TMEM columns) are resident per SM
  const size_t smem = 2 * TC_W * 128 + sizeof(TcShared) + 1024 + 12 * 1024;
  LSH_OPT_IN_SMEM(attend_fwd_tc_kernel);
  attend_fwd_tc_kernel<<<BH * p.n_chunks, TC_THREADS, smem, stream>>>(p);
  LSH_CHECK_LAUNCH("attend_fwd_tc_kernel");
  return 0;
}

}  // namespace lsh
