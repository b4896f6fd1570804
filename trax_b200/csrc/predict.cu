// Fast inference (`mode='predict'`): one new token per call against the layer's input memory.
//   LSHSelfAttention._incremental_forward_unbatched, single-token branch   EA:2032-2109
//   SelfAttention._incremental_forward_unbatched, q_len == 1                EA:1200-1268
// for every (example, head) unit at once.  The step is launch- and latency-bound (one query per unit, at most a few thousand
// keys of 128 bytes), so these are plain CUDA-core kernels: one CTA per unit walks the memory's projected rows once.
//
//   predict_attend_kernel  per unit: takes the query's bucket ids from this step's hash, writes them into the bucket memory
//                          (EA:2069-2071), ranks the memory slots (predict_select.cuh, EA:2073-2084), then one pass of online
//                          softmax over the selected slots: 8 warps x strided slots, a slot's key / value row (2 x 128 B,
//                          coalesced) per warp, fp32 statistics; the key normalisation of EA:2088 / EA:232 is computed from
//                          the row in flight.  Bytes per unit: nh * M * 4 (bucket memory) + 256 per attended slot.
//   predict_out_kernel     out[b] = sum_h o[b, h] . w_o[h]  (EA:2104, heads summed as at EA:2161), fp32 weights as stored.
#include "common.cuh"
#include "predict_select.cuh"

namespace lsh {

struct PredictAttendParams {
  const __nv_bfloat16 *qv;   // (B, M, H, QV) bf16: q | v (| k) of every memory slot, this step's projection
  int32_t *buckets;          // (BH, bstride) bucket memory, rows (nh, M); NULL without hashing (SelfAttention)
  const int32_t *hashed;     // (BH, nh * M) this step's bucket ids of every slot (only column q_start is read); NULL likewise
  float *o;                  // (BH, 64) f32: the step's attention output per unit
  int64_t bstride;
  int H, M, QV, nh, q_start, k_sel;
  int kcol;                  // first key column inside a (slot, head) row: 0 (shared-QK) or 128 (separate k)
  int normalize;             // keys are length-normalised (shared-QK, EA:2088 / EA:230)
  int exclude_self;          // the query's own slot gets -1e5 (EA:2091-2092 / EA:1246)
  int causal;                // without hashing: slots after q_start are masked (EA:1245)
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(PREDICT_THREADS) predict_attend_kernel(const PredictAttendParams p) {
  extern __shared__ uint8_t flags[];                 // M bytes: validity, then "attended" flags
  __shared__ float qf[64];
  __shared__ int qb[64];
  __shared__ int seg_valid[PREDICT_THREADS], seg_invalid[PREDICT_THREADS];
  __shared__ float red_m[PREDICT_THREADS / 32], red_l[PREDICT_THREADS / 32], red_o[PREDICT_THREADS / 32][64];
  const int u = blockIdx.x, b = u / p.H, h = u % p.H, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __nv_bfloat16 *unit_rows = p.qv + (static_cast<int64_t>(b) * p.M * p.H + h) * p.QV;   // slot i at + i * H * QV
  const int64_t slot_stride = static_cast<int64_t>(p.H) * p.QV;
  if (tid < 64) qf[tid] = __bfloat162float(unit_rows[p.q_start * slot_stride + tid]);
  int last = p.q_start;                              // highest slot that can receive probability
  if (p.buckets != nullptr) {
    int32_t *bk = p.buckets + static_cast<int64_t>(u) * p.bstride;
    if (tid < p.nh) {
      const int32_t v = p.hashed[(static_cast<int64_t>(u) * p.nh + tid) * p.M + p.q_start];
      qb[tid] = v;
      bk[static_cast<int64_t>(tid) * p.M + p.q_start] = v;           // EA:2069-2071
    }
    __syncthreads();
    const PredictSelect sel = {bk, qb, p.M, p.nh, p.q_start, p.k_sel};
    predict_select_count(sel, tid, flags, seg_valid, seg_invalid);
    __syncthreads();
    predict_select_rank(sel, tid, flags, seg_valid, seg_invalid);
  } else {
    if (!p.causal) last = p.M - 1;
    for (int i = tid; i <= last; i += PREDICT_THREADS) flags[i] = 1;
  }
  __syncthreads();

  // one pass of online softmax; warp w takes slots w, w + 8, ...; lane l holds columns 2l, 2l + 1 of the key and the value
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  const float q0 = qf[2 * lane], q1 = qf[2 * lane + 1];
  for (int i = warp; i <= last; i += PREDICT_THREADS / 32) {
    if (!flags[i]) continue;                         // (warp-uniform)
    const __nv_bfloat16 *row = unit_rows + i * slot_stride;
    const float2 kk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(row + p.kcol + 2 * lane));
    const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(row + 64 + 2 * lane));
    float s = warp_sum_f(fmaf(q0, kk.x, q1 * kk.y));
    if (p.normalize) {
      const float ss = warp_sum_f(fmaf(kk.x, kk.x, kk.y * kk.y));
      s *= rsqrtf(ss * (1.f / 64.f) + 1e-6f);        // length_normalized, EA:54-57
    }
    s *= 0.125f;                                     // k / sqrt(d_qk), EA:232
    if (p.exclude_self && i == p.q_start) s -= 1e5f; // EA:153-155
    const float m_new = fmaxf(m, s);
    const float scale = m == -INFINITY ? 0.f : expf(m - m_new);
    const float pr = expf(s - m_new);
    l = fmaf(l, scale, pr);
    o0 = fmaf(o0, scale, pr * vv.x);
    o1 = fmaf(o1, scale, pr * vv.y);
    m = m_new;
  }
  if (lane == 0) { red_m[warp] = m; red_l[warp] = l; }
  red_o[warp][2 * lane] = o0;
  red_o[warp][2 * lane + 1] = o1;
  __syncthreads();
  if (tid < 64) {
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < PREDICT_THREADS / 32; ++w) mm = fmaxf(mm, red_m[w]);
    float ll = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < PREDICT_THREADS / 32; ++w) {
      if (red_m[w] == -INFINITY) continue;
      const float f = expf(red_m[w] - mm);
      ll = fmaf(red_l[w], f, ll);
      oo = fmaf(red_o[w][tid], f, oo);
    }
    p.o[static_cast<int64_t>(u) * 64 + tid] = ll > 0.f ? oo / ll : 0.f;
  }
}

// out (B, D) = sum over heads and value columns of o (B, H, 64) . w_o (H, 64, D); one thread per output column.
template <typename T>
__global__ void __launch_bounds__(128) predict_out_kernel(const float *__restrict__ o, const float *__restrict__ w_o,
                                                        T *__restrict__ out, int H, int D) {
  extern __shared__ float os[];                      // H * 64
  const int b = blockIdx.y, d = blockIdx.x * 128 + threadIdx.x;
  for (int i = threadIdx.x; i < H * 64; i += 128) os[i] = o[static_cast<int64_t>(b) * H * 64 + i];
  __syncthreads();
  if (d >= D) return;
  float acc = 0.f;
  for (int k = 0; k < H * 64; ++k) acc = fmaf(os[k], __ldg(w_o + static_cast<int64_t>(k) * D + d), acc);
  if constexpr (sizeof(T) == 4) out[static_cast<int64_t>(b) * D + d] = acc;
  else out[static_cast<int64_t>(b) * D + d] = __float2bfloat16(acc);
}

int predict_attend_run(const LshAttnDims &d, const void *qv, int32_t *buckets, int64_t bstride, const int32_t *hashed, int q_start,
                       float *o, cudaStream_t stream) {
  Derived dr = derive(d);
  if (d.L > 32768) return set_error("predict: memory of %d slots exceeds the kernel's 32768", d.L);
  if (q_start < 0 || q_start >= d.L) return set_error("predict: q_start=%d outside the memory of %d slots", q_start, d.L);
  if (buckets != nullptr && d.nh > 64) return set_error("predict: n_hashes=%d > 64", d.nh);
  PredictAttendParams p;
  p.qv = static_cast<const __nv_bfloat16 *>(qv);
  p.buckets = buckets;
  p.hashed = hashed;
  p.o = o;
  p.bstride = bstride;
  p.H = d.H; p.M = d.L; p.QV = dr.QV; p.nh = d.nh; p.q_start = q_start;
  p.k_sel = d.nh * d.C * (1 + d.nb);                 // EA:2083-2084
  p.kcol = d.separate_k ? d.dq + d.dv : 0;
  p.normalize = d.separate_k ? 0 : 1;
  p.exclude_self = d.separate_k ? 0 : 1;
  p.causal = buckets != nullptr ? 1 : d.causal;
  predict_attend_kernel<<<dr.BH, PREDICT_THREADS, d.L, stream>>>(p);
  LSH_CHECK_LAUNCH("predict_attend_kernel");
  return 0;
}

int predict_out_run(const LshAttnDims &d, const float *o, const float *w_o, void *out, cudaStream_t stream) {
  const dim3 grid((d.D + 127) / 128, d.B);
  const size_t smem = static_cast<size_t>(d.H) * 64 * sizeof(float);
  if (smem > 48 * 1024) return set_error("predict: n_heads=%d too large for the output kernel", d.H);
  if (d.act_dtype == LSH_DTYPE_F32)
    predict_out_kernel<float><<<grid, 128, smem, stream>>>(o, w_o, static_cast<float *>(out), d.H, d.D);
  else
    predict_out_kernel<__nv_bfloat16><<<grid, 128, smem, stream>>>(o, w_o, static_cast<__nv_bfloat16 *>(out), d.H, d.D);
  LSH_CHECK_LAUNCH("predict_out_kernel");
  return 0;
}

}  // namespace lsh
