// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX; encodings follow the PTX ISA as
// exercised by CUTLASS's cute/arch/mma_sm100_desc.hpp, mma_sm100_umma.hpp, copy_sm100.hpp).
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched from the driver at run time, no libcuda link)

#include "common.cuh"

namespace lsh {

// ---- TMA descriptors (host) ---------------------------------------------------------------------------
// 2-D bf16 tensor of `rows` rows x `cols` columns (row pitch `pitch_bytes`), box = `box_cols` columns x 1 row,
// SWIZZLE_128B: the descriptor form `cp.async.bulk.tensor.2d ... tile::gather4` takes (four independent row
// coordinates per instruction, rows land in consecutive 128-byte rows of the swizzle atom; CUTLASS builds the same map for
// SM100_TMA_LOAD_2D_GATHER4, cute/atom/copy_traits_sm90_tma.hpp).
int make_row_gather_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_cols);

// Optional epilogue of the q|v projection GEMM (gemm_tc.cu): per-token key scales / normalised keys, see GemmTcParams.
struct GemmQStats {
  float *qscale;
  float2 *rowmeta;
  void *qhat;
  int L, H;
};
// Optional residual epilogue (layers/reversible.py:318, 400): C = resid + acc_sign * (A B^T), `resid` laid out and typed like C
// (it may BE C: every element is read and written by the same thread).
struct GemmResidual {
  const void *resid;
  float acc_sign;
};
int gemm_tc_run(int64_t M, int64_t N, int64_t K, const void *A, int64_t lda, const void *B, int64_t ldb, void *C, int64_t ldc,
                bool c_f32, cudaStream_t stream, const GemmQStats *qs = nullptr, const GemmResidual *res = nullptr);
// Optional destination of the q|v(|k) weight gradient in the reference's per-head layouts (see sum_partials_unpack_kernel).
struct WgradUnpack {
  float *dw_q, *dw_v, *dw_k;
  int H, D, dq, dv;
};
int gemm_tc_wgrad_run(int64_t M, int64_t N, int64_t K, const void *A, int64_t lda, const void *B, int64_t ldb, float *C,
                      void *scratch, size_t scratch_bytes, cudaStream_t stream, const WgradUnpack *up = nullptr,
                      bool *unpacked = nullptr);

#ifdef __CUDACC__

// ---- TMA (device) -------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// Four rows (r0..r3, any order) of a 2-D tensor, columns [c0, c0 + box_cols), into four consecutive 128-byte rows at
// `dst` (swizzled by the shared-memory address like every SWIZZLE_128B tile); completion = bytes on the mbarrier.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, uint64_t *bar, int c0, int r0, int r1, int r2,
                                            int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
// Potentially-blocking probe: the hardware suspends the thread until the phase completes or `hint_ns` expires
// (wake-up on completion is immediate), so a waiting warp neither burns issue slots nor oversleeps.
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity, uint32_t hint_ns = 20000) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// non-blocking probe
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#ifdef LSH_DEBUG_SPIN
  uint32_t spins = 0;
#endif
  while (!mbar_try_wait(bar, parity)) {
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
#ifdef LSH_DEBUG_SPIN
    if (++spins > (1u << 16)) __trap();   // a protocol bug becomes a trap instead of a hung GPU
#endif
  }
}
// The mbarrier receives one (pre-counted) arrival once every cp.async issued so far by this thread has landed: tiles are
// signalled by the copy engine itself, so a producer never blocks on its own loads and many tiles stay in flight.
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// generic-proxy writes (st.shared / cp.async) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- TMEM allocation --------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), base_offset [49,52), layout_type [61,64) (2 = SWIZZLE_128B).
// Tiles here are rows of 128 bytes (64 bf16) in 1024-byte swizzle atoms of 8 rows:
//   K-major  (rows = M/N index, 64 K-elements per row): SBO = 1024 (next 8 rows), LBO unused (1)
//   MN-major (rows = K index, 64 MN-elements per row) : SBO = 1024 (next 8 K-rows), LBO = stride between
//            64-wide MN groups (unused for MN extent 64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, bf16 x bf16 -> f32.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                       // c_format = F32
         | (1u << 7)                     // a_format = BF16
         | (1u << 10)                    // b_format = BF16
         | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// Split form: hi word is a per-operand-kind constant, lo word = start address field (+ LBO); advancing the start
// address by `bytes` is `lo + (bytes >> 4)` (the 14-bit field cannot overflow for shared-memory addresses).
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B
}
// One lane of a converged warp (the MMA / commit instructions are issued by it alone).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}

// ---- MMA issue (one thread) -------------------------------------------------------------------------
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// Arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ---- packed fp32 pairs ----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  return static_cast<uint64_t>(__float_as_uint(hi)) << 32 | __float_as_uint(lo);
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) { return static_cast<uint64_t>(hi) << 32 | lo; }
__device__ __forceinline__ float lo32(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v)); }
__device__ __forceinline__ float hi32(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }
// Packed fp32 pairs (FFMA2 / FADD2): same FMA throughput per lane, half the issue slots.
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---- TMEM <-> registers (32 lanes x 32-bit, N consecutive columns; lane = thread) ----------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// wait::ld tied to the destination registers of two earlier (prefetched) x16 loads
__device__ __forceinline__ void tmem_ld_wait_dep16(uint32_t (&a)[16], uint32_t (&b)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]),
                 "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(b[0]),
                 "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]),
                 "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// Same wait, but data-dependent on the destination registers of an earlier (prefetched) tcgen05.ld so that
// the compiler cannot hoist their uses above the wait.
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

#endif  // __CUDACC__
}  // namespace lsh
