// Layout / dtype plumbing around the projection GEMMs (EA:1923-1924, 1995 and their VJPs).
#include "attend_params.cuh"

namespace lsh {

// w_q (H, D, dq) f32, w_v (H, D, dv) f32 [, w_k (H, D, dq)] -> wqv (D, H, dq+dv[+dq]) bf16 ; w_o (H, dv, D) f32 -> bf16 same layout
__global__ void pack_wqv_kernel(const float *__restrict__ w_q, const float *__restrict__ w_v, const float *__restrict__ w_k,
                                __nv_bfloat16 *__restrict__ wqv, int H, int D, int dq, int dv) {
  const int QV = dq + dv + (w_k ? dq : 0);
  const int64_t n = static_cast<int64_t>(D) * H * QV;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % QV);
    const int h = static_cast<int>((i / QV) % H);
    const int dm = static_cast<int>(i / (static_cast<int64_t>(QV) * H));
    const float v = (c < dq) ? w_q[(static_cast<int64_t>(h) * D + dm) * dq + c]
                    : (c < dq + dv) ? w_v[(static_cast<int64_t>(h) * D + dm) * dv + (c - dq)]
                                    : w_k[(static_cast<int64_t>(h) * D + dm) * dq + (c - dq - dv)];
    wqv[i] = __float2bfloat16_rn(v);
  }
}

__global__ void f32_to_bf16_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + i);
    uint2 o;
    o.x = pack_bf16(v.x, v.y); o.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2 *>(dst)[i] = o;
  }
  for (int64_t i = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// dwqv (D, H, dq+dv[+dq]) f32 -> dw_q (H, D, dq), dw_v (H, D, dv) [, dw_k (H, D, dq)] f32
__global__ void unpack_dwqv_kernel(const float *__restrict__ dwqv, float *__restrict__ dw_q,
                                   float *__restrict__ dw_v, float *__restrict__ dw_k, int H, int D, int dq, int dv) {
  const int QV = dq + dv + (dw_k ? dq : 0);
  const int64_t n = static_cast<int64_t>(D) * H * QV;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % QV);
    const int h = static_cast<int>((i / QV) % H);
    const int dm = static_cast<int>(i / (static_cast<int64_t>(QV) * H));
    if (c < dq) dw_q[(static_cast<int64_t>(h) * D + dm) * dq + c] = dwqv[i];
    else if (c < dq + dv) dw_v[(static_cast<int64_t>(h) * D + dm) * dv + (c - dq)] = dwqv[i];
    else dw_k[(static_cast<int64_t>(h) * D + dm) * dq + (c - dq - dv)] = dwqv[i];
  }
}

// (C, W) float keep multiplier -> bit rows in both orientations + the multiplier value (see AttnKeep)
__global__ void __launch_bounds__(256) attn_keep_kernel(const float *__restrict__ keep, uint32_t *__restrict__ bits,
                                                        uint32_t *__restrict__ bits_t, float *__restrict__ scale, int C, int W) {
  __shared__ float smax[8];
  float mx = 0.f;
  const int wpr = W / 32, cpr = C / 32;
  for (int i = threadIdx.x; i < C * wpr; i += blockDim.x) {              // bits[r][w]
    const int r = i / wpr, w = i - r * wpr;
    uint32_t v = 0u;
    for (int j = 0; j < 32; ++j) {
      const float k = keep[static_cast<int64_t>(r) * W + 32 * w + j];
      mx = fmaxf(mx, k);
      v |= (k != 0.f ? 1u : 0u) << j;
    }
    bits[i] = v;
  }
  for (int i = threadIdx.x; i < W * cpr; i += blockDim.x) {              // bits_t[col][w]: bit j = keep[32 w + j][col]
    const int col = i / cpr, w = i - col * cpr;
    uint32_t v = 0u;
    for (int j = 0; j < 32; ++j) v |= (keep[static_cast<int64_t>(32 * w + j) * W + col] != 0.f ? 1u : 0u) << j;
    bits_t[i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, smax[i]);
    *scale = mx;
  }
}

size_t attn_keep_bytes(const LshAttnDims &d) {
  Derived dr = derive(d);
  return static_cast<size_t>(d.C) * dr.W / 8 * 2 + 512;
}

int attn_keep_prepare(const LshAttnDims &d, const float *keep_f32, void *ws, AttnKeep *out, cudaStream_t stream) {
  Derived dr = derive(d);
  out->bits = out->bits_t = nullptr; out->scale = nullptr;
  if (!keep_f32) return 0;
  if (!ws) return set_error("attention dropout: workspace missing");
  char *b = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  uint32_t *bits = reinterpret_cast<uint32_t *>(b), *bits_t = bits + static_cast<size_t>(d.C) * dr.W / 32;
  float *scale = reinterpret_cast<float *>(bits_t + static_cast<size_t>(d.C) * dr.W / 32);
  attn_keep_kernel<<<1, 256, 0, stream>>>(keep_f32, bits, bits_t, scale, d.C, dr.W);
  LSH_CHECK_LAUNCH("attn_keep_kernel");
  out->bits = bits; out->bits_t = bits_t; out->scale = scale;
  return 0;
}

static unsigned grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

int pack_weights_run(const LshAttnDims &d, const float *w_q, const float *w_v, const float *w_o, const float *w_k,
                     void *wqv, void *wo, cudaStream_t stream) {
  if ((d.separate_k != 0) != (w_k != nullptr)) return set_error("pack_weights: w_k must be given exactly when dims.separate_k is set");
  const int64_t n = static_cast<int64_t>(d.D) * d.H * derive(d).QV;
  pack_wqv_kernel<<<grid_for(n, 256), 256, 0, stream>>>(w_q, w_v, w_k, static_cast<__nv_bfloat16 *>(wqv), d.H,
                                                        d.D, d.dq, d.dv);
  LSH_CHECK_LAUNCH("pack_wqv_kernel");
  const int64_t no = static_cast<int64_t>(d.H) * d.dv * d.D;
  f32_to_bf16_kernel<<<grid_for(no / 4 + 1, 256), 256, 0, stream>>>(w_o, static_cast<__nv_bfloat16 *>(wo), no);
  LSH_CHECK_LAUNCH("f32_to_bf16_kernel");
  return 0;
}

// All four packed bf16 weight layouts of a layer call in ONE launch (the separate pack / convert / two transposes cost four
// latency-bound launches, ~28 us per call at config 2, for 6 MB of weights):
//   wqv   (D, H*QV)   = [w_q | w_v (| w_k)] per head, row = model dimension       (weight-gradient layout; B operand of dx)
//   wqv_t (H*QV, D)   = its transpose                                              (B operand of the q|v projection)
//   wo    (H*dv, D)   = w_o                                                        (B operand of do = dout w_o^T)
//   wo_t  (D, H*dv)   = its transpose                                              (B operand of the output projection)
// One CTA = one 32 x 32 tile of one head's matrix; blockIdx.x enumerates the w_q | w_v | w_k tiles, then the w_o tiles.
__global__ void __launch_bounds__(256) pack_all_kernel(const float *__restrict__ w_q, const float *__restrict__ w_v,
                                                       const float *__restrict__ w_k, const float *__restrict__ w_o,
                                                       __nv_bfloat16 *__restrict__ wqv, __nv_bfloat16 *__restrict__ wqv_t,
                                                       __nv_bfloat16 *__restrict__ wo, __nv_bfloat16 *__restrict__ wo_t,
                                                       int H, int D, int dq, int dv, int n_qv_tiles) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int QV = dq + dv + (w_k ? dq : 0), NQV = H * QV, KO = H * dv;
  int t = blockIdx.x;
  if (t < n_qv_tiles) {
    // tile (h, dm0, c0) of the (D, QV) matrix of head h
    const int ct = QV / 32, dt = D / 32;
    const int c0 = (t % ct) * 32, dm0 = ((t / ct) % dt) * 32, h = t / (ct * dt);
    const float *src; int cs, width;
    if (c0 < dq) { src = w_q; cs = c0; width = dq; }
    else if (c0 < dq + dv) { src = w_v; cs = c0 - dq; width = dv; }
    else { src = w_k; cs = c0 - dq - dv; width = dq; }
    for (int i = ty; i < 32; i += 8) {
      const __nv_bfloat16 v = __float2bfloat16_rn(src[(static_cast<int64_t>(h) * D + dm0 + i) * width + cs + tx]);
      tile[i][tx] = v;
      wqv[static_cast<int64_t>(dm0 + i) * NQV + h * QV + c0 + tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) wqv_t[static_cast<int64_t>(h * QV + c0 + i) * D + dm0 + tx] = tile[tx][i];
  } else {
    t -= n_qv_tiles;
    // tile (h, e0, dm0) of the (dv, D) matrix of head h
    const int dt = D / 32, et = dv / 32;
    const int dm0 = (t % dt) * 32, e0 = ((t / dt) % et) * 32, h = t / (dt * et);
    for (int i = ty; i < 32; i += 8) {
      const __nv_bfloat16 v = __float2bfloat16_rn(w_o[(static_cast<int64_t>(h) * dv + e0 + i) * D + dm0 + tx]);
      tile[i][tx] = v;
      wo[static_cast<int64_t>(h * dv + e0 + i) * D + dm0 + tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) wo_t[static_cast<int64_t>(dm0 + i) * KO + h * dv + e0 + tx] = tile[tx][i];
  }
}

// Returns -1 when the shape is outside the fused kernel's tiling (the caller then runs the separate kernels).
int pack_all_run(const LshAttnDims &d, const float *w_q, const float *w_v, const float *w_o, const float *w_k, void *wqv,
                 void *wqv_t, void *wo, void *wo_t, cudaStream_t stream) {
  if ((d.separate_k != 0) != (w_k != nullptr)) return set_error("pack_weights: w_k must be given exactly when dims.separate_k is set");
  if (d.D % 32 != 0 || d.dq % 32 != 0 || d.dv % 32 != 0) return -1;
  const int QV = derive(d).QV;
  const int n_qv = d.H * (d.D / 32) * (QV / 32), n_o = d.H * (d.dv / 32) * (d.D / 32);
  pack_all_kernel<<<n_qv + n_o, 256, 0, stream>>>(w_q, w_v, w_k, w_o, static_cast<__nv_bfloat16 *>(wqv),
                                                  static_cast<__nv_bfloat16 *>(wqv_t), static_cast<__nv_bfloat16 *>(wo),
                                                  static_cast<__nv_bfloat16 *>(wo_t), d.H, d.D, d.dq, d.dv, n_qv);
  LSH_CHECK_LAUNCH("pack_all_kernel");
  return 0;
}

// src (R, C) bf16 -> dst (C, R) bf16: the K-major copy of a packed weight for the tensor-core GEMM (weights only: a few MB)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16 *__restrict__ src, __nv_bfloat16 *__restrict__ dst,
                                                             int R, int C) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < R && c0 + tx < C) tile[i][tx] = src[static_cast<int64_t>(r0 + i) * C + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && r0 + tx < R) dst[static_cast<int64_t>(c0 + i) * R + r0 + tx] = tile[tx][i];
}

int transpose_bf16_run(const void *src, void *dst, int R, int C, cudaStream_t stream) {
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  transpose_bf16_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(src), static_cast<__nv_bfloat16 *>(dst), R, C);
  LSH_CHECK_LAUNCH("transpose_bf16_kernel");
  return 0;
}

int f32_to_bf16_run(const float *src, void *dst, int64_t n, cudaStream_t stream) {
  f32_to_bf16_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, stream>>>(src, static_cast<__nv_bfloat16 *>(dst), n);
  LSH_CHECK_LAUNCH("f32_to_bf16_kernel");
  return 0;
}

int unpack_dwqv_run(const LshAttnDims &d, const float *dwqv, float *dw_q, float *dw_v, float *dw_k, cudaStream_t stream) {
  if ((d.separate_k != 0) != (dw_k != nullptr)) return set_error("unpack_dwqv: dw_k must be given exactly when dims.separate_k is set");
  const int64_t n = static_cast<int64_t>(d.D) * d.H * derive(d).QV;
  unpack_dwqv_kernel<<<grid_for(n, 256), 256, 0, stream>>>(dwqv, dw_q, dw_v, dw_k, d.H, d.D, d.dq, d.dv);
  LSH_CHECK_LAUNCH("unpack_dwqv_kernel");
  return 0;
}

// ---- head layout plumbing of the weight-less core (EA:2564 PureLSHSelfAttention: inputs (batch*heads, seqlen, d_head)) ------
// a (BH, L, da) [, b (BH, L, db)] in f32 or bf16  ->  dst (B, L, H, da + db) bf16: the (token, head) row layout every kernel
// gathers from ([qk | v] side by side).  One thread = 8 consecutive elements (16 bytes of bf16).
template <typename T>
__global__ void __launch_bounds__(256) pack_heads_kernel(const T *__restrict__ a, int da, const T *__restrict__ b, int db,
                                                         __nv_bfloat16 *__restrict__ dst, int H, int L, int64_t n_chunks) {
  const int cpr = (da + db) >> 3;                                    // 16-byte chunks per destination row
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_chunks; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cpr);
    const int64_t row = i / cpr;                                     // (b, t, h)
    const int h = static_cast<int>(row % H);
    const int64_t bt = row / H, bb = bt / L, t = bt - bb * L;
    const int64_t srow = (bb * H + h) * L + t;                       // (b h, t)
    const int col = c * 8;
    const T *src = col < da ? a + srow * da + col : b + srow * db + (col - da);
    uint4 o;
    if constexpr (sizeof(T) == 4) {
      const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src)), v1 = __ldg(reinterpret_cast<const float4 *>(src) + 1);
      o.x = pack_bf16(v0.x, v0.y); o.y = pack_bf16(v0.z, v0.w); o.z = pack_bf16(v1.x, v1.y); o.w = pack_bf16(v1.z, v1.w);
    } else {
      o = __ldg(reinterpret_cast<const uint4 *>(src));
    }
    *reinterpret_cast<uint4 *>(dst + row * (da + db) + col) = o;
  }
}

// src (B, L, H, d_total) bf16, columns [col0, col0 + d)  ->  dst (BH, L, d) in f32 or bf16 (the core's outputs / cotangents)
template <typename T>
__global__ void __launch_bounds__(256) unpack_heads_kernel(const __nv_bfloat16 *__restrict__ src, int d_total, int col0, int d,
                                                           T *__restrict__ dst, int H, int L, int64_t n_chunks) {
  const int cpr = d >> 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_chunks; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cpr);
    const int64_t drow = i / cpr;                                    // (b h, t)
    const int64_t t = drow % L, bh = drow / L, bb = bh / H, h = bh - bb * H;
    const int64_t srow = (bb * L + t) * H + h;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + srow * d_total + col0 + c * 8));
    if constexpr (sizeof(T) == 4) {
      const float2 f0 = unpack_bf16(v.x), f1 = unpack_bf16(v.y), f2 = unpack_bf16(v.z), f3 = unpack_bf16(v.w);
      float4 *o = reinterpret_cast<float4 *>(dst + drow * d + c * 8);
      o[0] = make_float4(f0.x, f0.y, f1.x, f1.y);
      o[1] = make_float4(f2.x, f2.y, f3.x, f3.y);
    } else {
      *reinterpret_cast<uint4 *>(dst + drow * d + c * 8) = v;
    }
  }
}

int pack_heads_run(int B, int H, int L, int act_dtype, const void *a, int da, const void *b, int db, void *dst, cudaStream_t stream) {
  if (B < 1 || H < 1 || L < 1 || da < 8 || da % 8 != 0 || db < 0 || db % 8 != 0 || (db > 0) != (b != nullptr))
    return set_error("lsh_pack_heads: widths must be multiples of 8 (da=%d, db=%d)", da, db);
  const int64_t n = static_cast<int64_t>(B) * L * H * ((da + db) >> 3);
  const unsigned grid = grid_for(n, 256);
  if (act_dtype == LSH_DTYPE_F32)
    pack_heads_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(a), da, static_cast<const float *>(b), db,
                                                       static_cast<__nv_bfloat16 *>(dst), H, L, n);
  else
    pack_heads_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(a), da, static_cast<const __nv_bfloat16 *>(b), db,
                                                               static_cast<__nv_bfloat16 *>(dst), H, L, n);
  LSH_CHECK_LAUNCH("pack_heads_kernel");
  return 0;
}

int unpack_heads_run(int B, int H, int L, int act_dtype, const void *src, int d_total, int col0, int d, void *dst, cudaStream_t stream) {
  if (B < 1 || H < 1 || L < 1 || d < 8 || d % 8 != 0 || col0 < 0 || col0 % 8 != 0 || d_total % 8 != 0 || col0 + d > d_total)
    return set_error("lsh_unpack_heads: column range [%d, %d) of %d must be multiples of 8", col0, col0 + d, d_total);
  const int64_t n = static_cast<int64_t>(B) * L * H * (d >> 3);
  const unsigned grid = grid_for(n, 256);
  if (act_dtype == LSH_DTYPE_F32)
    unpack_heads_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(src), d_total, col0, d, static_cast<float *>(dst), H, L, n);
  else
    unpack_heads_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(src), d_total, col0, d,
                                                                 static_cast<__nv_bfloat16 *>(dst), H, L, n);
  LSH_CHECK_LAUNCH("unpack_heads_kernel");
  return 0;
}

// ---- rotations: counter-based N(0,1) (stands in for jax.random.normal at EA:92) ---------------------
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// Threefry-2x32, 20 rounds (Salmon et al. 2011), the generator family JAX uses; bit-compatibility with
// jax.random is NOT claimed (not verifiable offline).
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t &x0, uint32_t &x1) {
  const uint32_t ks[3] = {k0, k1, 0x1BD11BDAu ^ k0 ^ k1};
  const int rot[8] = {13, 15, 26, 6, 17, 29, 16, 24};
  x0 += ks[0]; x1 += ks[1];
#pragma unroll
  for (int r = 0; r < 5; ++r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x0 += x1; x1 = rotl32(x1, rot[(r & 1) * 4 + i]); x1 ^= x0;
    }
    x0 += ks[(r + 1) % 3]; x1 += ks[(r + 2) % 3] + static_cast<uint32_t>(r + 1);
  }
}

__global__ void make_rotations_kernel(const uint32_t *__restrict__ keys, uint32_t *__restrict__ new_keys,
                                      float *__restrict__ rot, int per_unit) {
  const int u = blockIdx.y;
  const uint32_t k0 = keys[2 * u], k1 = keys[2 * u + 1];
  // split (EA:1928): child 0 = next state key, child 1 = key for this draw
  uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 1;
  threefry2x32(k0, k1, a0, a1);
  threefry2x32(k0, k1, b0, b1);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // pair index
  if (2 * i < per_unit) {
    uint32_t x0 = static_cast<uint32_t>(i), x1 = 0x9E3779B9u;
    threefry2x32(b0, b1, x0, x1);
    // Box-Muller on two uniforms in (0,1]
    const float u1 = (static_cast<float>(x0 >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float u2 = static_cast<float>(x1 >> 8) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    float *dst = rot + static_cast<int64_t>(u) * per_unit;
    dst[2 * i] = rad * cs;
    if (2 * i + 1 < per_unit) dst[2 * i + 1] = rad * sn;
  }
  if (new_keys != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    new_keys[2 * u] = a0; new_keys[2 * u + 1] = a1;
  }
}

int make_rotations_run(const LshAttnDims &d, const uint32_t *keys, uint32_t *new_keys, float *rot,
                       cudaStream_t stream) {
  Derived dr = derive(d);
  const int per_unit = d.dq * d.nh * dr.R;
  dim3 grid((per_unit / 2 + 1 + 127) / 128, dr.BH);
  make_rotations_kernel<<<grid, 128, 0, stream>>>(keys, new_keys, rot, per_unit);
  LSH_CHECK_LAUNCH("make_rotations_kernel");
  return 0;
}

}  // namespace lsh
