// Shared device/host helpers for the LSH-attention kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lsh_attn.h"

namespace lsh {

// ---- host side -----------------------------------------------------------------------------------
int set_error(const char *fmt, ...);   // stores thread-local message, returns non-zero
void count_launch(int n = 1);

struct Derived {
  int BH, N, n_chunks, W, QV, R, n_buckets, nwin;
};

// Token-level inputs of the tcgen05 backward kernel, written by the multi-round combine of the backward call's recompute
// when it already has the token's output row in registers (see combine_fwd_kernel): do_comb (B, L, H, 64) bf16 in,
// dvec / lse2 / qcmp (BH, L) f32 out (same values as bwd_prep_tc_kernel).
struct BwdPrepOut {
  const void *do_comb;
  float *dvec, *lse2, *qcmp;
};

inline Derived derive(const LshAttnDims &d) {
  Derived r;
  r.BH = d.B * d.H;
  r.N = d.nh * d.L;
  r.n_chunks = d.C > 0 ? r.N / d.C : 0;
  r.nwin = 1 + d.nb + d.na;
  r.W = d.C * r.nwin;
  r.QV = d.dq + d.dv + (d.separate_k ? d.dq : 0);   // q | v (| k): columns of one (token, head) row
  r.R = 0;
  r.n_buckets = 1;
  for (int i = 0; i < d.n_factors && i < 4; ++i) {
    r.R += d.factors[i] / 2;
    r.n_buckets *= d.factors[i];
  }
  if (d.masked) r.n_buckets += 1;
  return r;
}

// Opt a kernel in to the full 227 KB of dynamic shared memory once per (thread, kernel).
#define LSH_OPT_IN_SMEM(kernel)                                                                      \
  do {                                                                                               \
    static thread_local int done_dev_ = -1;                                                          \
    int dev_ = 0;                                                                                    \
    cudaGetDevice(&dev_);                                                                            \
    if (done_dev_ != dev_) {                                                                         \
      cudaFuncAttributes fa_;                                                                        \
      cudaError_t e_ = cudaFuncGetAttributes(&fa_, kernel);                                          \
      if (e_ == cudaSuccess)                                                                         \
        e_ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                                  227 * 1024 - static_cast<int>(fa_.sharedSizeBytes));               \
      if (e_ != cudaSuccess) return lsh::set_error("cudaFuncSetAttribute(%s): %s", #kernel, cudaGetErrorString(e_)); \
      done_dev_ = dev_;                                                                              \
    }                                                                                                \
  } while (0)

#define LSH_CHECK_LAUNCH(name)                                                      \
  do {                                                                              \
    cudaError_t e_ = cudaGetLastError();                                            \
    if (e_ != cudaSuccess) return lsh::set_error("%s: %s", name, cudaGetErrorString(e_)); \
    lsh::count_launch();                                                            \
  } while (0)

// ---- device side ---------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// 128-byte rows (64 bf16) with the 16-byte-chunk XOR swizzle (chunk ^= row & 7) — the same pattern as
// the TMA/UMMA SWIZZLE_128B atom, conflict-free for ldmatrix.
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2,
                                            uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1,
                                                  uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// D += A(16x16, row) * B(16x8, col), bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162 *>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// 2^x on the SFU (MUFU.EX2), flush-to-zero: one instruction, ~2 ulp; exp2(-huge) == 0 exactly.
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

#endif  // __CUDACC__
}  // namespace lsh
