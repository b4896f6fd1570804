// Dense projections of the layer on the 5th-generation tensor cores — C[M, N] = A[M, K] · B[N, K]^T, bf16 operands (both
// K-major: rows of K contiguous elements), fp32 accumulation in TMEM, C bf16 or fp32 — for the activation x weight products
// of the path:  q|v(|k) = x · wqv (EA:1923-1924),  out = o · w_o (EA:1995, heads summed by the contraction, EA:2426) and, in the
// backward pass,  do = dout · w_o^T  and  dx = dqv · wqv^T  (the weight operand is packed K-major for each of them by
// pack.cu, so one kernel serves all four).
//
// Persistent, warp-specialised, TMA-fed:
//   warp 0   producer: one elected lane issues `cp.async.bulk.tensor.2d` tile loads (SWIZZLE_128B boxes of 64 K-elements x
//            128 rows) into a 4-stage ring; completion = transaction bytes on the stage's `full` mbarrier.
//   warp 1   MMA issuer: per K block of 64, four `tcgen05.mma.kind::f16` M128 x N{256,128} x K16 on shared-memory descriptors,
//            accumulating into one of TWO TMEM accumulators (2 x N columns); `tcgen05.commit` frees the stage / hands the
//            finished accumulator to the epilogue.
//   warps 2-9 epilogue (two warpgroups, each owns half of the tile's columns; a warp reads the TMEM lane quarter warp % 4):
//            TMEM -> registers (32 columns at a time, next load in flight) -> per-warp staging tile -> bf16 / fp32 row
//            segments in global memory, while the issuer is already accumulating the next tile in the other accumulator.
// CLUSTER = 2: the two CTAs of a cluster form ONE `cta_group::2` tensor-core pair computing a 256 x BN tile: each CTA loads
// its own 128 rows of A and HALF of the weight tile (B rows [rank BN/2, +BN/2)) — `cp.async.bulk.tensor ... cta_group::2`
// signals the LEADER's (rank 0) barrier from both CTAs — and the leader's issuer alone issues `tcgen05.mma.cta_group::2`
// (M 256: the hardware reads A / B halves from both shared memories and accumulates 128 rows in each CTA's TMEM); stages and
// accumulators are handed back with multicast commits.  Per CTA and K block this moves 16 + 16 KB through shared memory
// instead of 16 + 32 KB (the 1-CTA form, and the round-2 multicast form, are bound by the 128 B/clk shared-memory port: 48 KB
// written by TMA + 48 KB read by the MMAs per 512 tensor cycles), and halves the L2 -> SM traffic of B.
// Shapes the kernel does not cover (N % 128, K % 64, unaligned pointers) are reported to the caller, which then uses cuBLAS.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace lsh {

constexpr int GM_BM = 128, GM_BK = 64;
constexpr int GM_EPI_WARPS = 8;                     // two warpgroups, each drains half of the tile's columns
constexpr int GM_THREADS = 64 + 32 * GM_EPI_WARPS;

struct GemmTcParams {
  CUtensorMap tm_a, tm_b;
  void *c;
  int64_t ldc;
  int M, N, K, c_f32, m_blocks, n_blocks;
  int splits;          // MN mode: the K range is cut into `splits` pieces, piece s writes its fp32 partial to c + s * M * ldc
  // q|v projection only (all null otherwise): the per-token key normalisation of EA:229-231 straight from the epilogue — for
  // every (token, head) whose 64 q columns this thread has just rounded to bf16: qscale = log2e / (8 r), the normalised key
  // qhat = bf16(q / (8 r)) and rowmeta = {8 r log2e, that times |qhat|^2}, r = sqrt(mean(q^2) + 1e-6) (see qscale_kernel)
  float *qscale;
  float2 *rowmeta;
  __nv_bfloat16 *qhat;
  int L, H;
  // residual epilogue: C = resid + acc_sign * acc (null = plain product); same pitch and element type as C
  const void *resid;
  float acc_sign;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// shared::cluster address of the same shared-memory offset in the LEADER CTA (cluster rank 0) of the pair
__device__ __forceinline__ uint32_t leader_addr(const void *p) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;\n" : "=r"(r) : "r"(smem_u32(p)));
  return r;
}
// cta_group::2 forms: the load's completion bytes go to the LEADER CTA's barrier, whichever CTA of the pair issues it
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_ss2_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives (once all MMAs issued so far have completed) on the barrier at this offset in EVERY CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t *bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
// ordinary arrival on the LEADER CTA's barrier at this offset (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
  // (default semantics, as CUTLASS's umma_arrive_2x1SM_sm0: a `.release.cluster` arrive compiles to ERRBAR + a cluster-scope
  // fence that waits for the warp's outstanding global stores of the C tile — measured: ~40 % of the epilogue's samples.  What
  // the issuer needs ordered before this arrival are the TMEM reads: tcgen05.wait::ld + tcgen05.fence::before_thread_sync.)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(leader_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

constexpr int GM_MAX_STAGES = 6;
struct __align__(16) GemmShared {
  uint64_t full[GM_MAX_STAGES], empty[GM_MAX_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// MN = false: A (M, K) and B (N, K), K contiguous (activation x weight).  MN = true: A (K, M) and B (K, N), M / N contiguous —
// the weight-gradient products dW = act^T · cotangent, contraction over the B L token rows; tiles are boxes of 64 columns x 64
// K-rows (MN-major SWIZZLE_128B operands, 8 KB per 64-wide group), K is split over the clusters (fp32 partials, summed by
// sum_partials_kernel in a fixed order).
__host__ __device__ constexpr int gemm_stages(int cl) { return cl == 2 ? 6 : 4; }

// RES: residual epilogue (its own instantiations: the plain products keep the code they were tuned with)
template <int BN, int CL, bool MN, bool RES = false>
__global__ void __launch_bounds__(GM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = GM_BM * 128, B_BYTES = (BN / CL) * 128, STAGE_BYTES = A_BYTES + B_BYTES;   // per CTA
  constexpr int GM_STAGES = gemm_stages(CL);
  __shared__ GemmShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const uint32_t smem_u = smem_u32(smem);
  const uint32_t stage_u = smem_u + GM_STAGES * STAGE_BYTES;          // epilogue staging: GM_EPI_WARPS x 4 KB

  if (warp == 1) { if (CL == 2) tmem_alloc_2sm(&sh.tmem_base, 512); else tmem_alloc(&sh.tmem_base, 512); }
  if (tid == 0) {
    tma_prefetch_desc(&p.tm_a);
    tma_prefetch_desc(&p.tm_b);
    // CL = 2: `full` and `acc_empty` are used in the leader CTA only (both CTAs' loads / epilogue warps arrive there);
    // `empty` and `acc_full` receive one multicast commit per phase in each CTA
    for (int i = 0; i < GM_STAGES; ++i) { mbar_init(&sh.full[i], 1); mbar_init(&sh.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh.acc_full[i], 1); mbar_init(&sh.acc_empty[i], CL * GM_EPI_WARPS); }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  const int n_clusters = gridDim.x / CL, cid = blockIdx.x / CL;
  const int m_groups = (p.m_blocks + CL - 1) / CL, tiles = m_groups * p.n_blocks, total = tiles * p.splits, kb_all = p.K / GM_BK;
  // work item g = (K piece, row-block group, column block), column block fastest; its K blocks are [kb0, kb1)
  auto k_range = [&](int g, int &kb0, int &kb1) {
    const int sp = g / tiles;
    kb0 = static_cast<int>(static_cast<int64_t>(kb_all) * sp / p.splits);
    kb1 = static_cast<int>(static_cast<int64_t>(kb_all) * (sp + 1) / p.splits);
  };

  if (warp == 0) {
    // ================================ TMA producer ======================================================
    int it = 0;                                                     // k-block counter across tiles (ring position)
    for (int g = cid; g < total; g += n_clusters) {
      const int n_blk = g % p.n_blocks, m_blk = ((g % tiles) / p.n_blocks) * CL + static_cast<int>(rank);
      int kb0, kb1;
      k_range(g, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % GM_STAGES;
        mbar_wait(&sh.empty[s], ((it / GM_STAGES) & 1) ^ 1);
        if (elect_one()) {
          const uint32_t a_dst = smem_u + s * STAGE_BYTES, b_dst = a_dst + A_BYTES;
          if (CL == 1) mbar_arrive_expect_tx(&sh.full[s], STAGE_BYTES);
          else if (rank == 0) mbar_arrive_expect_tx(&sh.full[s], 2 * STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
          if constexpr (!MN) {
            if (CL == 1) {
              tma_load_2d(a_dst, &p.tm_a, &sh.full[s], kb * GM_BK, m_blk * GM_BM);     // rows past M read as zeros
              tma_load_2d(b_dst, &p.tm_b, &sh.full[s], kb * GM_BK, n_blk * BN);
            } else {                                                   // my rows of A, my half of the weight tile
              tma_load_2d_2sm(a_dst, &p.tm_a, leader_addr(&sh.full[s]), kb * GM_BK, m_blk * GM_BM);
              tma_load_2d_2sm(b_dst, &p.tm_b, leader_addr(&sh.full[s]), kb * GM_BK, n_blk * BN + static_cast<int>(rank) * (BN / 2));
            }
          } else {
            // 64-column groups: 2 of A (mine), BN / 64 / CL of B (CL = 2: my half of the tile's column groups)
#pragma unroll
            for (int gq = 0; gq < GM_BM / 64; ++gq) {
              if (CL == 1) tma_load_2d(a_dst + gq * 8192, &p.tm_a, &sh.full[s], m_blk * GM_BM + gq * 64, kb * GM_BK);
              else tma_load_2d_2sm(a_dst + gq * 8192, &p.tm_a, leader_addr(&sh.full[s]), m_blk * GM_BM + gq * 64, kb * GM_BK);
            }
#pragma unroll
            for (int gq = 0; gq < BN / 64 / CL; ++gq) {
              const int grp = static_cast<int>(rank) * (BN / 64 / CL) + gq;
              if (CL == 1) tma_load_2d(b_dst + gq * 8192, &p.tm_b, &sh.full[s], n_blk * BN + grp * 64, kb * GM_BK);
              else tma_load_2d_2sm(b_dst + gq * 8192, &p.tm_b, leader_addr(&sh.full[s]), n_blk * BN + grp * 64, kb * GM_BK);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (CL = 2: the leader CTA's, for the pair) ==================
    constexpr uint32_t HI = desc_hi(1024);
    constexpr uint32_t IDESC = make_idesc_bf16(GM_BM * CL, BN, MN ? 1 : 0, MN ? 1 : 0);
    if (CL == 1 || rank == 0) {
    // K-major: 128-byte rows, next K16 step = +32 bytes inside the row.  MN-major: 64-wide groups 8 KB apart (LBO), next K16
    // step = 16 K-rows = +2048 bytes
    constexpr uint32_t LBO = MN ? 8192 : 16, KSTEP = MN ? 128 : 2;
    int it = 0, t = 0;
    for (int g = cid; g < total; g += n_clusters, ++t) {
      const int acc = t & 1;
      int kb0, kb1;
      k_range(g, kb0, kb1);
      mbar_wait(&sh.acc_empty[acc], ((t >> 1) & 1) ^ 1);            // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_t = tmem + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % GM_STAGES;
        mbar_wait(&sh.full[s], (it / GM_STAGES) & 1);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smem_u + s * STAGE_BYTES, LBO), b_lo = desc_lo(smem_u + s * STAGE_BYTES + A_BYTES, LBO);
        if (elect_one()) {
          if (CL == 1) {
#pragma unroll
            for (int ks = 0; ks < GM_BK / 16; ++ks)
              umma_ss2(d_t, a_lo + ks * KSTEP, HI, b_lo + ks * KSTEP, HI, IDESC, (kb > kb0 || ks > 0) ? 1u : 0u);
            umma_commit(&sh.empty[s]);
            if (kb == kb1 - 1) umma_commit(&sh.acc_full[acc]);
          } else {
#pragma unroll
            for (int ks = 0; ks < GM_BK / 16; ++ks)
              umma_ss2_2sm(d_t, a_lo + ks * KSTEP, HI, b_lo + ks * KSTEP, HI, IDESC, (kb > kb0 || ks > 0) ? 1u : 0u);
            umma_commit_2sm_mc(&sh.empty[s], 3);                       // the stage is free in both CTAs
            if (kb == kb1 - 1) umma_commit_2sm_mc(&sh.acc_full[acc], 3);   // both CTAs' epilogues drain their 128 rows
          }
        }
        __syncwarp();
      }
    }
    }
  } else {
    // ================================ epilogue warps ====================================================
    const int q = warp & 3;                                          // TMEM lane quarter of this warp
    constexpr int CPG = BN / 32 / (GM_EPI_WARPS / 4);                // 32-column chunks per warpgroup (even: heads stay whole or q | v)
    const int c_first = ((warp - 2) >> 2) * CPG;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(q * 32) << 16);
    int t = 0;
    for (int g = cid; g < total; g += n_clusters, ++t) {
      const int acc = t & 1;
      const int n_blk = g % p.n_blocks, m_blk = ((g % tiles) / p.n_blocks) * CL + static_cast<int>(rank);
      const int64_t row = static_cast<int64_t>(m_blk) * GM_BM + q * 32 + lane;
      const bool live = row < p.M;
      mbar_wait(&sh.acc_full[acc], (t >> 1) & 1);
      tc_fence_after();
      uint32_t ra[32], rb[32];
      tmem_ld32(t_lane + acc * BN + c_first * 32, ra);
      uint32_t qlo[16];                                              // first half of a head's q row, as stored (bf16 pairs)
      float qss = 0.f;
      // key normalisation of one head from the bf16 values just stored: `c` = chunk of 32 columns; a head owns 4 chunks,
      // q | q | v | v (columns [128 h', 128 h' + 64) are its q)
      auto q_stats = [&](const uint32_t (&r)[32], int c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pk[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
          const float2 f = unpack_bf16(pk[i]);
          qss = fmaf(f.x, f.x, qss); qss = fmaf(f.y, f.y, qss);
        }
        if ((c & 1) == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) qlo[i] = pk[i];
          return;
        }
        const float rr = sqrtf(qss * (1.0f / 64) + 1e-6f), sc = 0.125f / rr;
        qss = 0.f;
        const int h = n_blk * (BN / 128) + (c >> 2);
        const int64_t b = row / p.L, ut = (b * p.H + h) * p.L + (row - b * p.L);
        float s2 = 0.f;
        uint4 *dst = reinterpret_cast<uint4 *>(p.qhat + ut * 64);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16(half ? pk[4 * i + e] : qlo[4 * i + e]);
              o[e] = pack_bf16(f.x * sc, f.y * sc);
              const float2 g = unpack_bf16(o[e]);                     // |qhat|^2 of the rounded values the tensor cores see
              s2 = fmaf(g.x, g.x, s2); s2 = fmaf(g.y, g.y, s2);
            }
            uint4 v; v.x = o[0]; v.y = o[1]; v.z = o[2]; v.w = o[3];
            dst[half * 4 + i] = v;
          }
        }
        p.qscale[ut] = 0.125f * kLog2e / rr;
        const float am = 8.f * rr * kLog2e;
        p.rowmeta[ut] = make_float2(am, am * s2);
      };
      // Rows leave through a per-warp staging tile (32 rows x 128 bytes, 16-byte pieces XOR-swizzled by the row): a thread
      // owns a ROW of the accumulator, but a store instruction should cover contiguous bytes — after the transpose 8 lanes
      // write one 128-byte row segment (4 rows per instruction) instead of 32 lanes writing 16 bytes of 32 different rows.
      // One round = 32 fp32 columns, or 64 bf16 columns (two accumulator chunks).
      const uint32_t stg = stage_u + (warp - 2) * 4096;
      auto put = [&](int piece, uint32_t a, uint32_t b, uint32_t c2, uint32_t d) {
        sts128(stg + lane * 128 + ((static_cast<uint32_t>(piece) ^ (lane & 7)) << 4), a, b, c2, d);
      };
      auto flush = [&](int64_t col_bytes) {                          // col_bytes: byte offset of the round inside the C row
        __syncwarp();
        const int piece = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + (lane >> 3);
          const int64_t grow = static_cast<int64_t>(m_blk) * GM_BM + q * 32 + r;
          uint4 v = lds128u(stg + r * 128 + ((static_cast<uint32_t>(piece) ^ (r & 7)) << 4));
          if (grow < p.M) {
            const int64_t off = (grow + static_cast<int64_t>(g / tiles) * p.M) * p.ldc * (p.c_f32 ? 4 : 2) + col_bytes + piece * 16;
            if constexpr (RES) {                                       // y = x + attn / x = y - attn of the reversible block
              const uint4 rv = *reinterpret_cast<const uint4 *>(static_cast<const char *>(p.resid) + off);
              const float sg = p.acc_sign;
              if (p.c_f32) {
                v.x = __float_as_uint(fmaf(sg, __uint_as_float(v.x), __uint_as_float(rv.x)));
                v.y = __float_as_uint(fmaf(sg, __uint_as_float(v.y), __uint_as_float(rv.y)));
                v.z = __float_as_uint(fmaf(sg, __uint_as_float(v.z), __uint_as_float(rv.z)));
                v.w = __float_as_uint(fmaf(sg, __uint_as_float(v.w), __uint_as_float(rv.w)));
              } else {
                auto mix = [&](uint32_t a, uint32_t b) {
                  const float2 fa = unpack_bf16(a), fb = unpack_bf16(b);
                  return pack_bf16(fmaf(sg, fa.x, fb.x), fmaf(sg, fa.y, fb.y));
                };
                v.x = mix(v.x, rv.x); v.y = mix(v.y, rv.y); v.z = mix(v.z, rv.z); v.w = mix(v.w, rv.w);
              }
            }
            *reinterpret_cast<uint4 *>(static_cast<char *>(p.c) + off) = v;
          }
        }
        __syncwarp();
      };
      auto store = [&](const uint32_t (&r)[32], int c) {
        if (live && p.qhat != nullptr && (c & 2) == 0) q_stats(r, c);
        const int64_t col = static_cast<int64_t>(n_blk) * BN + c * 32;
        if (p.c_f32) {
#pragma unroll
          for (int i = 0; i < 8; ++i) put(i, r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          flush(col * 4);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            put((c & 1) * 4 + i, pack_bf16(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1])),
                pack_bf16(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3])),
                pack_bf16(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5])),
                pack_bf16(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])));
          if (c & 1) flush((col - 32) * 2);
        }
      };
#pragma unroll
      for (int cc = 0; cc < CPG; cc += 2) {
        const int c = c_first + cc;
        tmem_ld_wait_dep(ra);
        tmem_ld32(t_lane + acc * BN + (c + 1) * 32, rb);
        store(ra, c);
        tmem_ld_wait_dep(rb);
        if (cc + 2 < CPG) tmem_ld32(t_lane + acc * BN + (c + 2) * 32, ra);
        store(rb, c + 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CL == 1) mbar_arrive(&sh.acc_empty[acc]); else mbar_arrive_leader(&sh.acc_empty[acc]); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it
  if (warp == 1) { if (CL == 2) tmem_dealloc_2sm(tmem, 512); else tmem_dealloc(tmem, 512); }
}

int make_tile_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_cols,
                  uint32_t box_rows);

template <int BN, int CL, bool MN, bool RES = false>
static int gemm_tc_launch(const GemmTcParams &p, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(gemm_stages(CL)) * (GM_BM * 128 + (BN / CL) * 128) + GM_EPI_WARPS * 4096 + 1024;
  auto kernel = gemm_tc_kernel<BN, CL, MN, RES>;
  LSH_OPT_IN_SMEM(kernel);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int groups = ((p.m_blocks + CL - 1) / CL) * p.n_blocks * p.splits;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(GM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  int n_clusters = sms / CL;
  if (CL > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(sms / CL * CL);
    static thread_local int max_clusters[64] = {};
    if (dev < 64 && max_clusters[dev] == 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = sms / CL / 2; }
      max_clusters[dev] = n;
    }
    if (dev < 64 && max_clusters[dev] < n_clusters) n_clusters = max_clusters[dev];
  }
  if (groups < n_clusters) n_clusters = groups;
  cfg.gridDim = dim3(n_clusters * CL);
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
  if (e != cudaSuccess) return set_error("gemm_tc_kernel launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

// A[M, K] (row pitch lda elements) · B[N, K]^T (row pitch ldb) -> C[M, N] (row pitch ldc), bf16 in, bf16 or f32 out.
// Returns 0 on launch, > 0 on error, -1 when the shape is outside what the kernel covers (the caller falls back to cuBLAS).
int gemm_tc_run(int64_t M, int64_t N, int64_t K, const void *A, int64_t lda, const void *B, int64_t ldb, void *C, int64_t ldc,
                bool c_f32, cudaStream_t stream, const GemmQStats *qs, const GemmResidual *res) {
  static const int mode = [] {        // LSH_GEMM=cublas: library GEMMs everywhere; LSH_GEMM=cluster1: no weight-tile multicast
    const char *e = getenv("LSH_GEMM");
    if (!e) return 2;
    if (strcmp(e, "cublas") == 0) return 0;
    if (strcmp(e, "cluster1") == 0) return 1;
    return 2;
  }();
  if (mode == 0) return -1;
  if (N % 128 != 0 || K % GM_BK != 0 || M < 1 || M >= (1ll << 31) || lda % 8 != 0 || ldb % 8 != 0 || ldc % 8 != 0) return -1;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) return -1;
  const int bn = (N % 256 == 0) ? 256 : 128;
  GemmTcParams p;
  p.c = C; p.ldc = ldc; p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K); p.c_f32 = c_f32 ? 1 : 0;
  p.m_blocks = static_cast<int>((M + GM_BM - 1) / GM_BM); p.n_blocks = static_cast<int>(N / bn);
  p.splits = 1;
  p.qscale = nullptr; p.rowmeta = nullptr; p.qhat = nullptr; p.L = 1; p.H = 1;
  p.resid = res ? res->resid : nullptr; p.acc_sign = res ? res->acc_sign : 1.f;
  if (res && res->resid && (reinterpret_cast<uintptr_t>(res->resid) & 15)) return -1;
  if (qs) {           // fused key normalisation: rows are (b, t), columns (h, q | v) — needs bf16 output of 128-column heads
    if (c_f32 || N != static_cast<int64_t>(qs->H) * 128) return set_error("gemm_tc: q statistics need the (B L, H 128) bf16 projection");
    p.qscale = qs->qscale; p.rowmeta = qs->rowmeta; p.qhat = static_cast<__nv_bfloat16 *>(qs->qhat); p.L = qs->L; p.H = qs->H;
  }
  const int cl = (mode == 2 && p.m_blocks >= 2) ? 2 : 1;
  if (int rc = make_tile_map(&p.tm_a, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda) * 2, GM_BK, GM_BM)) return rc;
  if (int rc = make_tile_map(&p.tm_b, B, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldb) * 2, GM_BK, bn / cl)) return rc;
  if (p.resid != nullptr) {
    if (bn == 256) return cl == 2 ? gemm_tc_launch<256, 2, false, true>(p, stream) : gemm_tc_launch<256, 1, false, true>(p, stream);
    return cl == 2 ? gemm_tc_launch<128, 2, false, true>(p, stream) : gemm_tc_launch<128, 1, false, true>(p, stream);
  }
  if (bn == 256) return cl == 2 ? gemm_tc_launch<256, 2, false>(p, stream) : gemm_tc_launch<256, 1, false>(p, stream);
  return cl == 2 ? gemm_tc_launch<128, 2, false>(p, stream) : gemm_tc_launch<128, 1, false>(p, stream);
}

// fp32 partial sums of the K pieces, added in piece order (deterministic): dst[i] = sum_s part[s * n + i]
__global__ void __launch_bounds__(256) sum_partials_kernel(const float4 *__restrict__ part, float4 *__restrict__ dst, int64_t n4, int splits) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 a = part[i];
    for (int s = 1; s < splits; ++s) {
      const float4 b = part[s * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dst[i] = a;
  }
}

// Same sum, but of the (D, H, dq | dv (| dq)) weight gradient of the q|v(|k) projection, written straight into the reference's
// per-head layouts dw_q (H, D, dq), dw_v (H, D, dv) (, dw_k (H, D, dq)): the separate unpack kernel's launch and its 2 x 4 MB
// of traffic are folded into the pass that reads the partials anyway.
__global__ void __launch_bounds__(256) sum_partials_unpack_kernel(const float4 *__restrict__ part, int64_t n4, int splits, WgradUnpack up) {
  const int QV = up.dq + up.dv + (up.dw_k ? up.dq : 0), NQV = up.H * QV;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 a = part[i];
    for (int s = 1; s < splits; ++s) {
      const float4 b = part[s * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    const int64_t e = i * 4;                                        // element (dm, h, c), four consecutive c (dq, dv multiples of 4)
    const int dm = static_cast<int>(e / NQV), col = static_cast<int>(e - static_cast<int64_t>(dm) * NQV);
    const int h = col / QV, c = col - h * QV;
    float *dst = c < up.dq ? up.dw_q + (static_cast<int64_t>(h) * up.D + dm) * up.dq + c
                 : c < up.dq + up.dv ? up.dw_v + (static_cast<int64_t>(h) * up.D + dm) * up.dv + (c - up.dq)
                                     : up.dw_k + (static_cast<int64_t>(h) * up.D + dm) * up.dq + (c - up.dq - up.dv);
    *reinterpret_cast<float4 *>(dst) = a;
  }
}

// Weight gradient C[M, N] (f32, row pitch N) = A^T · B with A (K, M) (row pitch lda) and B (K, N) (row pitch ldb), bf16, the
// contraction running over the K = B L token rows (EA:2431 sums examples: they are rows of the same product).  `scratch`
// holds the split-K partials.  Returns like gemm_tc_run (-1: shape not covered).
int gemm_tc_wgrad_run(int64_t M, int64_t N, int64_t K, const void *A, int64_t lda, const void *B, int64_t ldb, float *C,
                      void *scratch, size_t scratch_bytes, cudaStream_t stream, const WgradUnpack *up, bool *unpacked) {
  if (unpacked) *unpacked = false;
  static const int mode = [] {
    const char *e = getenv("LSH_GEMM");
    if (!e) return 2;
    if (strcmp(e, "cublas") == 0 || strcmp(e, "cublas_wgrad") == 0) return 0;
    if (strcmp(e, "cluster1") == 0) return 1;
    return 2;
  }();
  if (mode == 0) return -1;
  if (M % 128 != 0 || N % 128 != 0 || K % GM_BK != 0 || K >= (1ll << 31) || lda % 8 != 0 || ldb % 8 != 0) return -1;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(scratch)) & 15) return -1;
  const int bn = (N % 256 == 0) ? 256 : 128;
  GemmTcParams p;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K); p.c_f32 = 1; p.ldc = N;
  p.m_blocks = static_cast<int>(M / GM_BM); p.n_blocks = static_cast<int>(N / bn);
  p.qscale = nullptr; p.rowmeta = nullptr; p.qhat = nullptr; p.L = 1; p.H = 1; p.resid = nullptr; p.acc_sign = 1.f;
  const int cl = (mode == 2 && p.m_blocks % 2 == 0) ? 2 : 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = (p.m_blocks / cl) * p.n_blocks;
  int64_t splits = (sms / cl) / tiles;                       // one wave of clusters
  const int64_t kb_all = K / GM_BK, fit = static_cast<int64_t>(scratch_bytes / (static_cast<size_t>(M) * N * 4));
  if (splits > kb_all) splits = kb_all;
  if (splits > fit) splits = fit;
  if (splits < 1) splits = 1;
  p.splits = static_cast<int>(splits);
  p.c = splits > 1 ? scratch : static_cast<void *>(C);
  if (splits > 1 && !scratch) return -1;
  if (int rc = make_tile_map(&p.tm_a, A, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(lda) * 2, 64, GM_BK)) return rc;
  if (int rc = make_tile_map(&p.tm_b, B, static_cast<uint64_t>(K), static_cast<uint64_t>(N), static_cast<uint64_t>(ldb) * 2, 64, GM_BK)) return rc;
  int rc;
  if (bn == 256) rc = cl == 2 ? gemm_tc_launch<256, 2, true>(p, stream) : gemm_tc_launch<256, 1, true>(p, stream);
  else rc = cl == 2 ? gemm_tc_launch<128, 2, true>(p, stream) : gemm_tc_launch<128, 1, true>(p, stream);
  if (rc || splits == 1) return rc;
  const int64_t n4 = M * N / 4;
  const unsigned nb = static_cast<unsigned>((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
  if (up && unpacked && up->dq % 4 == 0 && up->dv % 4 == 0 && M == up->D &&
      N == static_cast<int64_t>(up->H) * (up->dq + up->dv + (up->dw_k ? up->dq : 0))) {
    sum_partials_unpack_kernel<<<nb, 256, 0, stream>>>(static_cast<const float4 *>(scratch), n4, static_cast<int>(splits), *up);
    LSH_CHECK_LAUNCH("sum_partials_unpack_kernel");
    *unpacked = true;
    return 0;
  }
  sum_partials_kernel<<<nb, 256, 0, stream>>>(static_cast<const float4 *>(scratch), reinterpret_cast<float4 *>(C), n4, static_cast<int>(splits));
  LSH_CHECK_LAUNCH("sum_partials_kernel");
  return 0;
}

}  // namespace lsh
