// The two memory-bound neighbours of the attention layer inside Trax's reversible block
// (`ReversibleHalfResidual(LayerNorm(), attention_layer=LSHSelfAttention)`, trax/layers/reversible.py:244-412):
//   layernorm_fwd  : trax/layers/normalization.py:129-136  z = (x - mean) / sqrt(var + eps) * scale + bias
//   layernorm_bwd  : its VJP (what fastmath.vjp(call_compute_residual) gives at reversible.py:352-353, 384-385), added
//                    into the context cotangent (reversible.py:397-398), + d_scale / d_bias
//   residual_sub   : reconstructed_x = accumulator_output - residual (reversible.py:400)
// One warp per row; rows are streamed once with 16-byte accesses (HBM-bound: bytes per row = the reads and writes
// listed at each kernel).  Statistics are fp32; activations are f32 or bf16 (template).
#include "common.cuh"

namespace lsh {

constexpr int LN_THREADS = 256;       // 8 rows per CTA
constexpr int LN_MAX_V = 8;           // 8-element vectors per lane: D = 256 * NV, NV in {1, 2, 4, 8}

template <typename T>
__device__ __forceinline__ void load8(const T *p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float *p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16 *p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
  const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <typename T>
__device__ __forceinline__ void store8(T *p, const float (&f)[8]);
template <>
__device__ __forceinline__ void store8<float>(float *p, const float (&f)[8]) {
  reinterpret_cast<float4 *>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4 *>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16 *p, const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]); v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
  *reinterpret_cast<uint4 *>(p) = v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reads x (row), writes z (row) and {mean, rstd} (8 bytes per row).
// TZ: type of z — T, or bf16 for f32 activations when the consumer is the attention layer (which rounds its input to bf16
// anyway: the same values, half the bytes, and no separate conversion pass in the layer).
template <typename T, int NV, typename TZ = T>
__global__ void __launch_bounds__(LN_THREADS) layernorm_fwd_kernel(const T *__restrict__ x, const float *__restrict__ scale,
                                                                 const float *__restrict__ bias, TZ *__restrict__ z,
                                                                 float2 *__restrict__ stats, int64_t rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (LN_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    {
      load8<T>(x + row * D + (i * 32 + lane) * 8, v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    {
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[i][e] -= mean; q = fmaf(v[i][e], v[i][e], q); }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    {
      float sc[8], bi[8], o[8];
      load8<float>(scale + (i * 32 + lane) * 8, sc);
      load8<float>(bias + (i * 32 + lane) * 8, bi);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(v[i][e] * rstd, sc[e], bi[e]);
      store8<TZ>(z + row * D + (i * 32 + lane) * 8, o);
    }
  }
  if (lane == 0 && stats != nullptr) stats[row] = make_float2(mean, rstd);
}

// Reads x, dz, ct_in (rows) and stats; writes ct_out = ct_in + dx (row); accumulates d_scale, d_bias (D floats each,
// zeroed by the caller) through per-CTA shared-memory partials and one atomicAdd per feature per CTA.
template <typename T, int NV>
__global__ void __launch_bounds__(LN_THREADS) layernorm_bwd_kernel(const T *__restrict__ x, const T *__restrict__ dz,
                                                                 const T *ct_in, const float2 *__restrict__ stats,
                                                                 const float *__restrict__ scale, T *ct_out,
                                                                 float *__restrict__ d_scale, float *__restrict__ d_bias,
                                                                 int64_t rows, int D, int rows_per_cta) {
  extern __shared__ float part[];          // [2][D]
  for (int i = threadIdx.x; i < 2 * D; i += LN_THREADS) part[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ds_acc[NV][8], db_acc[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) { ds_acc[i][e] = 0.f; db_acc[i][e] = 0.f; }
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
  for (int64_t row = r0 + warp; row < r0 + rows_per_cta && row < rows; row += LN_THREADS / 32) {
    const float2 st = stats[row];
    float g[NV][8], nrm[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      {
        float xv[8], sc[8];
        load8<T>(x + row * D + (i * 32 + lane) * 8, xv);
        load8<T>(dz + row * D + (i * 32 + lane) * 8, g[i]);
        load8<float>(scale + (i * 32 + lane) * 8, sc);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          nrm[i][e] = (xv[e] - st.x) * st.y;
          ds_acc[i][e] = fmaf(g[i][e], nrm[i][e], ds_acc[i][e]);
          db_acc[i][e] += g[i][e];
          g[i][e] *= sc[e];                                   // d(norm)
          s1 += g[i][e];
          s2 = fmaf(g[i][e], nrm[i][e], s2);
        }
      }
    }
    const float m1 = warp_sum(s1) / D, m2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      {
        float c[8], o[8];
        if (ct_in != nullptr) load8<T>(ct_in + row * D + (i * 32 + lane) * 8, c);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (ct_in != nullptr ? c[e] : 0.f) + st.y * (g[i][e] - m1 - nrm[i][e] * m2);
        store8<T>(ct_out + row * D + (i * 32 + lane) * 8, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&part[(i * 32 + lane) * 8 + e], ds_acc[i][e]);
        atomicAdd(&part[D + (i * 32 + lane) * 8 + e], db_acc[i][e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += LN_THREADS) {
    atomicAdd(d_scale + i, part[i]);
    atomicAdd(d_bias + i, part[D + i]);
  }
}

// out = a + sign * b (elementwise; n % 8 == 0; out may alias a)
template <typename T>
__global__ void __launch_bounds__(256) residual_sub_kernel(const T *a, const T *__restrict__ b, T *out, int64_t n8, float sign) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * 256) {
    float x[8], y[8];
    load8<T>(a + i * 8, x);
    load8<T>(b + i * 8, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = fmaf(sign, y[e], x[e]);
    store8<T>(out + i * 8, x);
  }
}

static int check_ln(int64_t rows, int D) {
  if (rows < 1 || (D != 256 && D != 512 && D != 1024 && D != 2048))
    return set_error("layernorm: d_model=%d unsupported (256, 512, 1024, 2048)", D);
  return 0;
}

#define LN_DISPATCH_NV(D, CALL)                \
  switch ((D) / 256) {                         \
    case 1: { constexpr int NV = 1; CALL; } break; \
    case 2: { constexpr int NV = 2; CALL; } break; \
    case 4: { constexpr int NV = 4; CALL; } break; \
    default: { constexpr int NV = 8; CALL; } break; \
  }

int layernorm_fwd_run(int64_t rows, int D, int dtype, const void *x, const float *scale, const float *bias, void *z,
                      float2 *stats, float eps, cudaStream_t stream, bool z_bf16) {
  if (int rc = check_ln(rows, D)) return rc;
  const unsigned grid = static_cast<unsigned>((rows + LN_THREADS / 32 - 1) / (LN_THREADS / 32));
  if (dtype == LSH_DTYPE_F32 && z_bf16) {
    LN_DISPATCH_NV(D, (layernorm_fwd_kernel<float, NV, __nv_bfloat16><<<grid, LN_THREADS, 0, stream>>>(
        static_cast<const float *>(x), scale, bias, static_cast<__nv_bfloat16 *>(z), stats, rows, D, eps)));
  } else if (dtype == LSH_DTYPE_F32) {
    LN_DISPATCH_NV(D, (layernorm_fwd_kernel<float, NV><<<grid, LN_THREADS, 0, stream>>>(
        static_cast<const float *>(x), scale, bias, static_cast<float *>(z), stats, rows, D, eps)));
  } else {
    LN_DISPATCH_NV(D, (layernorm_fwd_kernel<__nv_bfloat16, NV><<<grid, LN_THREADS, 0, stream>>>(
        static_cast<const __nv_bfloat16 *>(x), scale, bias, static_cast<__nv_bfloat16 *>(z), stats, rows, D, eps)));
  }
  LSH_CHECK_LAUNCH("layernorm_fwd_kernel");
  return 0;
}

int layernorm_bwd_run(int64_t rows, int D, int dtype, const void *x, const void *dz, const void *ct_in, const float2 *stats,
                      const float *scale, void *ct_out, float *d_scale, float *d_bias, cudaStream_t stream) {
  if (int rc = check_ln(rows, D)) return rc;
  cudaMemsetAsync(d_scale, 0, sizeof(float) * D, stream);
  cudaMemsetAsync(d_bias, 0, sizeof(float) * D, stream);
  // ~4 CTAs per SM; each CTA sweeps a contiguous block of rows and publishes one partial per feature
  int64_t ctas = 148 * 4;
  if (ctas > (rows + 7) / 8) ctas = (rows + 7) / 8;
  const int rows_per_cta = static_cast<int>((rows + ctas - 1) / ctas);
  const unsigned grid = static_cast<unsigned>((rows + rows_per_cta - 1) / rows_per_cta);
  const size_t smem = 2 * static_cast<size_t>(D) * sizeof(float);
  if (dtype == LSH_DTYPE_F32) {
    LN_DISPATCH_NV(D, (layernorm_bwd_kernel<float, NV><<<grid, LN_THREADS, smem, stream>>>(
        static_cast<const float *>(x), static_cast<const float *>(dz), static_cast<const float *>(ct_in), stats, scale,
        static_cast<float *>(ct_out), d_scale, d_bias, rows, D, rows_per_cta)));
  } else {
    LN_DISPATCH_NV(D, (layernorm_bwd_kernel<__nv_bfloat16, NV><<<grid, LN_THREADS, smem, stream>>>(
        static_cast<const __nv_bfloat16 *>(x), static_cast<const __nv_bfloat16 *>(dz), static_cast<const __nv_bfloat16 *>(ct_in),
        stats, scale, static_cast<__nv_bfloat16 *>(ct_out), d_scale, d_bias, rows, D, rows_per_cta)));
  }
  LSH_CHECK_LAUNCH("layernorm_bwd_kernel");
  return 0;
}

int residual_sub_run(int64_t n, int dtype, const void *a, const void *b, void *out, float sign, cudaStream_t stream) {
  if (n % 8 != 0) return set_error("residual_sub: element count must be a multiple of 8");
  const int64_t n8 = n / 8;
  int64_t blocks = (n8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == LSH_DTYPE_F32)
    residual_sub_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const float *>(a), static_cast<const float *>(b),
                                                                                 static_cast<float *>(out), n8, sign);
  else
    residual_sub_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16 *>(a), static_cast<const __nv_bfloat16 *>(b), static_cast<__nv_bfloat16 *>(out), n8, sign);
  LSH_CHECK_LAUNCH("residual_sub_kernel");
  return 0;
}

}  // namespace lsh
