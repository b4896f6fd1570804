// LSH bucket hashing — replaces EA:1889-1916 (hash_vectors) + EA:60-119 (hash_vecs).
//
// Contract (bit-exact with oracle/hash_oracle.c): every rotated coordinate is ONE fp32 accumulator,
// updated with __fmaf_rn over the contraction index f = 0..dq-1 in ascending order, starting from
// +0.0f.  argmax over concat([rv, -rv]) takes the first maximum.  CUDA-core fp32 work by design:
// tensor-core (bf16/tf32) products would change bucket ids.
//
// One thread owns one token: its dq-vector lives in registers, the rotation columns of one hash
// round are staged in shared memory and read as warp-wide broadcasts (LDS.128 feeds 4 FFMA/lane).
#include "common.cuh"

namespace lsh {

constexpr int HASH_THREADS = 128;
constexpr int HASH_COLS = 8;   // rotation columns processed together per thread

// Optional per-token by-products of the hash kernel (it already holds q in registers): exactly what qscale_kernel
// (combine.cu) writes, in the same operation order, so a forward call with hashing skips that kernel.
struct HashAux {
  float *qscale;            // (BH, L)      or null
  float2 *rowmeta;          // (BH, L)      or null
  __nv_bfloat16 *qhat;      // (BH, L, 64)  or null
};

struct HashParams {
  const void *vecs;        // bf16 or f32
  int64_t stride_b, stride_h, stride_t;   // element strides of vecs for (example, head, token)
  const float *rot;        // (BH, dq, nh, R)
  const uint8_t *mask;     // (B, L) or null
  int32_t *buckets;
  int64_t buckets_stride;
  int L, H, nh, R, Rpad, n_factors, n_buckets;
  int factors[4];
};

template <typename T>
__device__ __forceinline__ void load_vec64(const T *p, float (&q)[64]);

template <>
__device__ __forceinline__ void load_vec64<float>(const float *p, float (&q)[64]) {
  const float4 *p4 = reinterpret_cast<const float4 *>(p);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float4 v = __ldg(p4 + i);
    q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
  }
}
template <>
__device__ __forceinline__ void load_vec64<__nv_bfloat16>(const __nv_bfloat16 *p, float (&q)[64]) {
  const uint4 *p4 = reinterpret_cast<const uint4 *>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 v = __ldg(p4 + i);
    float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
    q[8 * i] = a.x; q[8 * i + 1] = a.y; q[8 * i + 2] = b.x; q[8 * i + 3] = b.y;
    q[8 * i + 4] = c.x; q[8 * i + 5] = c.y; q[8 * i + 6] = d.x; q[8 * i + 7] = d.y;
  }
}

template <typename T>
__global__ void __launch_bounds__(HASH_THREADS) hash_kernel(const HashParams p) {
  constexpr int DQ = 64;
  extern __shared__ __align__(16) float s_rot[];   // [DQ][Rpad]
  const int u = blockIdx.y;
  const int b = u / p.H, h = u % p.H;
  const int t = blockIdx.x * HASH_THREADS + threadIdx.x;
  const bool active = t < p.L;

  float q[DQ];
  if (active) {
    const T *src = reinterpret_cast<const T *>(p.vecs) + b * p.stride_b + h * p.stride_h +
                   static_cast<int64_t>(t) * p.stride_t;
    load_vec64<T>(src, q);
  } else {
#pragma unroll
    for (int i = 0; i < DQ; ++i) q[i] = 0.f;
  }
  bool valid_tok = true;
  if (p.mask != nullptr && active) valid_tok = p.mask[static_cast<int64_t>(b) * p.L + t] != 0;

  const float *rot_u = p.rot + static_cast<int64_t>(u) * DQ * p.nh * p.R;
  for (int round = 0; round < p.nh; ++round) {
    __syncthreads();
    for (int i = threadIdx.x; i < DQ * p.Rpad; i += HASH_THREADS) {
      int f = i / p.Rpad, c = i % p.Rpad;
      s_rot[i] = (c < p.R) ? __ldg(rot_u + (static_cast<int64_t>(f) * p.nh + round) * p.R + c) : 0.f;
    }
    __syncthreads();

    // argmax state machine over the factor list
    int fi = 0, pos = 0, half = p.factors[0] >> 1;
    float best_pos = -INFINITY, best_neg = INFINITY;   // max of rv, min of rv
    int idx_pos = 0, idx_neg = 0;
    int bucket = 0, prod = 1;
    for (int c0 = 0; c0 < p.Rpad; c0 += HASH_COLS) {
      float acc[HASH_COLS];
#pragma unroll
      for (int j = 0; j < HASH_COLS; ++j) acc[j] = 0.f;
      const float *sr = s_rot + c0;
#pragma unroll
      for (int f = 0; f < DQ; ++f) {
        const float4 r0 = *reinterpret_cast<const float4 *>(sr + f * p.Rpad);
        const float4 r1 = *reinterpret_cast<const float4 *>(sr + f * p.Rpad + 4);
        acc[0] = __fmaf_rn(q[f], r0.x, acc[0]);
        acc[1] = __fmaf_rn(q[f], r0.y, acc[1]);
        acc[2] = __fmaf_rn(q[f], r0.z, acc[2]);
        acc[3] = __fmaf_rn(q[f], r0.w, acc[3]);
        acc[4] = __fmaf_rn(q[f], r1.x, acc[4]);
        acc[5] = __fmaf_rn(q[f], r1.y, acc[5]);
        acc[6] = __fmaf_rn(q[f], r1.z, acc[6]);
        acc[7] = __fmaf_rn(q[f], r1.w, acc[7]);
      }
#pragma unroll
      for (int j = 0; j < HASH_COLS; ++j) {
        if (c0 + j < p.R) {
          const float x = acc[j];
          if (x > best_pos) { best_pos = x; idx_pos = pos; }
          if (x < best_neg) { best_neg = x; idx_neg = pos; }
          ++pos;
          if (pos == half) {
            // concat([rv, -rv]): the +half wins ties (lower index), EA:104-105 / 112-116
            const int am = (best_pos >= -best_neg) ? idx_pos : half + idx_neg;
            bucket += prod * am;
            prod *= p.factors[fi];
            ++fi;
            half = (fi < p.n_factors) ? (p.factors[fi] >> 1) : 0x7fffffff;
            pos = 0; best_pos = -INFINITY; best_neg = INFINITY; idx_pos = 0; idx_neg = 0;
          }
        }
      }
    }
    if (active) {
      if (!valid_tok) bucket = p.n_buckets - 1;                      // EA:1908-1909
      p.buckets[static_cast<int64_t>(u) * p.buckets_stride + static_cast<int64_t>(round) * p.L + t] =
          bucket + round * p.n_buckets;                              // EA:1913-1915
    }
  }
}

// Fast variant: the rotation columns of ALL hash rounds stay resident in shared memory (one load per CTA, no per-round
// barriers), the row stride TC = n_hashes * Rpad is a compile-time constant (every LDS has an immediate offset), and each
// thread owns NT tokens so that one broadcast LDS.128 feeds 4 * NT FFMA per lane.  Measured on B200
// (tests/micro/ffma2_bench.cu): the register-file write port is the shared resource — an LDS.128 costs as many issue
// cycles as 4 FFMA — so with one token per thread the FMA pipe cannot exceed 50 %; two tokens lift the bound to 67 %.
// Same arithmetic per token, same order => same bits.  Optionally emits the per-token key scale (qscale).
template <typename T, int TC, int NT>
#ifndef LSH_HASH_MINB
#define LSH_HASH_MINB 2
#endif
__global__ void __launch_bounds__(HASH_THREADS, NT == 1 ? 4 : LSH_HASH_MINB) hash_all_rounds_kernel(const HashParams p, const HashAux aux) {
  constexpr int DQ = 64;
  extern __shared__ __align__(16) float s_rot[];   // [DQ][TC], column = round * Rpad + c
  const int u = blockIdx.y;
  const int b = u / p.H, h = u % p.H;
  const float *rot_u = p.rot + static_cast<int64_t>(u) * DQ * p.nh * p.R;
  for (int i = threadIdx.x; i < DQ * TC; i += HASH_THREADS) {
    const int f = i / TC, col = i % TC;
    const int round = col / p.Rpad, c = col % p.Rpad;
    s_rot[i] = (c < p.R) ? __ldg(rot_u + (static_cast<int64_t>(f) * p.nh + round) * p.R + c) : 0.f;
  }
  __syncthreads();
  // grid-stride over groups of NT * HASH_THREADS tokens of this unit: the rotations are loaded once per CTA
  for (int grp = blockIdx.x; grp * NT * HASH_THREADS < p.L; grp += gridDim.x) {
  float q[NT][DQ];
  int t[NT];
  bool active[NT], valid_tok[NT];
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    t[k] = (grp * NT + k) * HASH_THREADS + threadIdx.x;
    active[k] = t[k] < p.L;
    valid_tok[k] = true;
    if (active[k]) {
      const T *src = reinterpret_cast<const T *>(p.vecs) + b * p.stride_b + h * p.stride_h + static_cast<int64_t>(t[k]) * p.stride_t;
      load_vec64<T>(src, q[k]);
      if (p.mask != nullptr) valid_tok[k] = p.mask[static_cast<int64_t>(b) * p.L + t[k]] != 0;
    } else {
#pragma unroll
      for (int i = 0; i < DQ; ++i) q[k][i] = 0.f;
    }
    if (aux.qscale != nullptr && active[k]) {
      // sum of squares in qscale_kernel's order: eight 8-element fmaf chains, then the xor-shuffle tree of lane 0
      auto tree_sumsq = [](const float (&v)[DQ]) {
        float part[8];
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc = fmaf(v[8 * c8 + i], v[8 * c8 + i], acc);
          part[c8] = acc;
        }
        return ((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]));
      };
      const float r = sqrtf(tree_sumsq(q[k]) * (1.0f / 64) + 1e-6f);
      const int64_t ut = static_cast<int64_t>(u) * p.L + t[k];
      aux.qscale[ut] = 0.125f * kLog2e / r;
      if (aux.qhat != nullptr) {
        const float c = 0.125f / r;
        float qh[DQ];
        uint4 *dst = reinterpret_cast<uint4 *>(aux.qhat + ut * 64);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint4 pk;
          pk.x = pack_bf16(q[k][8 * c8 + 0] * c, q[k][8 * c8 + 1] * c); pk.y = pack_bf16(q[k][8 * c8 + 2] * c, q[k][8 * c8 + 3] * c);
          pk.z = pack_bf16(q[k][8 * c8 + 4] * c, q[k][8 * c8 + 5] * c); pk.w = pack_bf16(q[k][8 * c8 + 6] * c, q[k][8 * c8 + 7] * c);
          dst[c8] = pk;
          const float2 a2 = unpack_bf16(pk.x), b2 = unpack_bf16(pk.y), c2 = unpack_bf16(pk.z), d2 = unpack_bf16(pk.w);
          qh[8 * c8 + 0] = a2.x; qh[8 * c8 + 1] = a2.y; qh[8 * c8 + 2] = b2.x; qh[8 * c8 + 3] = b2.y;
          qh[8 * c8 + 4] = c2.x; qh[8 * c8 + 5] = c2.y; qh[8 * c8 + 6] = d2.x; qh[8 * c8 + 7] = d2.y;
        }
        const float am = 8.f * r * kLog2e;
        aux.rowmeta[ut] = make_float2(am, am * tree_sumsq(qh));
      }
    }
  }
  for (int round = 0; round < p.nh; ++round) {
    int fi = 0, pos = 0, half = p.factors[0] >> 1, prod = 1;
    float best_pos[NT], best_neg[NT];
    int idx_pos[NT], idx_neg[NT], bucket[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) { best_pos[k] = -INFINITY; best_neg[k] = INFINITY; idx_pos[k] = 0; idx_neg[k] = 0; bucket[k] = 0; }
    for (int c0 = 0; c0 < p.Rpad; c0 += HASH_COLS) {
      float acc[NT][HASH_COLS];
#pragma unroll
      for (int k = 0; k < NT; ++k)
#pragma unroll
        for (int j = 0; j < HASH_COLS; ++j) acc[k][j] = 0.f;
      const float *sr = s_rot + round * p.Rpad + c0;
#pragma unroll
      for (int f = 0; f < DQ; ++f) {
        const float4 r0 = *reinterpret_cast<const float4 *>(sr + f * TC);
        const float4 r1 = *reinterpret_cast<const float4 *>(sr + f * TC + 4);
#pragma unroll
        for (int k = 0; k < NT; ++k) {
          acc[k][0] = __fmaf_rn(q[k][f], r0.x, acc[k][0]); acc[k][1] = __fmaf_rn(q[k][f], r0.y, acc[k][1]);
          acc[k][2] = __fmaf_rn(q[k][f], r0.z, acc[k][2]); acc[k][3] = __fmaf_rn(q[k][f], r0.w, acc[k][3]);
          acc[k][4] = __fmaf_rn(q[k][f], r1.x, acc[k][4]); acc[k][5] = __fmaf_rn(q[k][f], r1.y, acc[k][5]);
          acc[k][6] = __fmaf_rn(q[k][f], r1.z, acc[k][6]); acc[k][7] = __fmaf_rn(q[k][f], r1.w, acc[k][7]);
        }
      }
#pragma unroll
      for (int j = 0; j < HASH_COLS; ++j) {
        if (c0 + j < p.R) {                              // (token-independent bookkeeping: pos, half, fi, prod)
#pragma unroll
          for (int k = 0; k < NT; ++k) {
            const float x = acc[k][j];
            if (x > best_pos[k]) { best_pos[k] = x; idx_pos[k] = pos; }
            if (x < best_neg[k]) { best_neg[k] = x; idx_neg[k] = pos; }
          }
          ++pos;
          if (pos == half) {
#pragma unroll
            for (int k = 0; k < NT; ++k) {
              const int am = (best_pos[k] >= -best_neg[k]) ? idx_pos[k] : half + idx_neg[k];
              bucket[k] += prod * am;
              best_pos[k] = -INFINITY; best_neg[k] = INFINITY; idx_pos[k] = 0; idx_neg[k] = 0;
            }
            prod *= p.factors[fi];
            ++fi;
            half = (fi < p.n_factors) ? (p.factors[fi] >> 1) : 0x7fffffff;
            pos = 0;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      if (active[k]) {
        const int bk = valid_tok[k] ? bucket[k] : p.n_buckets - 1;
        p.buckets[static_cast<int64_t>(u) * p.buckets_stride + static_cast<int64_t>(round) * p.L + t[k]] = bk + round * p.n_buckets;
      }
    }
  }
  }   // token groups
}

template <typename T, int TC>
static int launch_hash_fast(const HashParams &p, int BH, const HashAux &aux, cudaStream_t stream) {
  // two tokens per thread when there are enough CTAs to fill the machine twice over
  const int64_t ctas2 = static_cast<int64_t>((p.L + 2 * HASH_THREADS - 1) / (2 * HASH_THREADS)) * BH;
  if (ctas2 >= 2 * 148) {
    LSH_OPT_IN_SMEM((hash_all_rounds_kernel<T, TC, 2>));
    const int groups = (p.L + 2 * HASH_THREADS - 1) / (2 * HASH_THREADS), per_unit = (LSH_HASH_MINB * 148 + BH - 1) / BH;
    dim3 grid(groups < per_unit ? groups : per_unit, BH);      // LSH_HASH_MINB resident CTAs per SM, one wave
    hash_all_rounds_kernel<T, TC, 2><<<grid, HASH_THREADS, 64 * TC * sizeof(float), stream>>>(p, aux);
  } else {
    LSH_OPT_IN_SMEM((hash_all_rounds_kernel<T, TC, 1>));
    dim3 grid((p.L + HASH_THREADS - 1) / HASH_THREADS, BH);
    hash_all_rounds_kernel<T, TC, 1><<<grid, HASH_THREADS, 64 * TC * sizeof(float), stream>>>(p, aux);
  }
  LSH_CHECK_LAUNCH("hash_all_rounds_kernel");
  return 0;
}

template <typename T>
static int launch_hash(const LshAttnDims &d, const void *vecs, int64_t sb, int64_t sh, int64_t st,
                       const float *rot, const uint8_t *mask, int32_t *buckets, int64_t bstride,
                       const HashAux &aux, cudaStream_t stream) {
  Derived dr = derive(d);
  HashParams p;
  p.vecs = vecs; p.stride_b = sb; p.stride_h = sh; p.stride_t = st;
  p.rot = rot; p.mask = d.masked ? mask : nullptr; p.buckets = buckets; p.buckets_stride = bstride;
  p.L = d.L; p.H = d.H; p.nh = d.nh; p.R = dr.R; p.Rpad = (dr.R + HASH_COLS - 1) / HASH_COLS * HASH_COLS;
  p.n_factors = d.n_factors; p.n_buckets = dr.n_buckets;
  for (int i = 0; i < 4; ++i) p.factors[i] = i < d.n_factors ? d.factors[i] : 2;
  size_t smem = static_cast<size_t>(64) * p.Rpad * sizeof(float);
  if (smem > 200 * 1024) return set_error("lsh_hash: sum(factors)/2 = %d too large for shared memory", dr.R);
  if (d.masked && mask == nullptr) return set_error("lsh_hash: dims.masked set but mask == NULL");
  switch (d.nh * p.Rpad) {          // all rounds resident, compile-time stride
    case 16: return launch_hash_fast<T, 16>(p, dr.BH, aux, stream);
    case 32: return launch_hash_fast<T, 32>(p, dr.BH, aux, stream);
    case 64: return launch_hash_fast<T, 64>(p, dr.BH, aux, stream);
    case 128: return launch_hash_fast<T, 128>(p, dr.BH, aux, stream);
    case 192: return launch_hash_fast<T, 192>(p, dr.BH, aux, stream);
    case 256: return launch_hash_fast<T, 256>(p, dr.BH, aux, stream);
    default: break;
  }
  if (aux.qscale != nullptr) return set_error("lsh_hash: fused qscale needs a specialised column count (got %d)", d.nh * p.Rpad);
  LSH_OPT_IN_SMEM(hash_kernel<T>);
  dim3 grid((d.L + HASH_THREADS - 1) / HASH_THREADS, dr.BH);
  hash_kernel<T><<<grid, HASH_THREADS, smem, stream>>>(p);
  LSH_CHECK_LAUNCH("hash_kernel");
  return 0;
}

int hash_bf16_qv(const LshAttnDims &d, const void *qv, const float *rot, const uint8_t *mask,
                 int32_t *buckets, int64_t bstride, cudaStream_t stream) {
  Derived dr = derive(d);
  return launch_hash<__nv_bfloat16>(d, qv, static_cast<int64_t>(d.L) * d.H * dr.QV, dr.QV,
                                    static_cast<int64_t>(d.H) * dr.QV, rot, mask, buckets, bstride,
                                    HashAux{nullptr, nullptr, nullptr}, stream);
}

// Hash + the per-token by-products of qscale_kernel in one pass over q (specialised column counts only).
bool hash_can_fuse_aux(const LshAttnDims &d) {
  Derived dr = derive(d);
  const int tc = d.nh * ((dr.R + HASH_COLS - 1) / HASH_COLS * HASH_COLS);
  return tc == 16 || tc == 32 || tc == 64 || tc == 128 || tc == 192 || tc == 256;
}
int hash_bf16_qv_aux(const LshAttnDims &d, const void *qv, const float *rot, const uint8_t *mask, int32_t *buckets,
                     int64_t bstride, float *qscale, float2 *rowmeta, void *qhat, cudaStream_t stream) {
  Derived dr = derive(d);
  return launch_hash<__nv_bfloat16>(d, qv, static_cast<int64_t>(d.L) * d.H * dr.QV, dr.QV,
                                    static_cast<int64_t>(d.H) * dr.QV, rot, mask, buckets, bstride,
                                    HashAux{qscale, rowmeta, static_cast<__nv_bfloat16 *>(qhat)}, stream);
}

int hash_f32_vecs(const LshAttnDims &d, const float *vecs, const float *rot, const uint8_t *mask,
                  int32_t *buckets, int64_t bstride, cudaStream_t stream) {
  return launch_hash<float>(d, vecs, static_cast<int64_t>(d.H) * d.L * d.dq,
                            static_cast<int64_t>(d.L) * d.dq, d.dq, rot, mask, buckets, bstride,
                            HashAux{nullptr, nullptr, nullptr}, stream);
}

}  // namespace lsh
