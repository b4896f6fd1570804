// extern "C" surface of liblsh_attn_b200.so (see include/lsh_attn.h) + the layer-level orchestration
// that mirrors LSHSelfAttention.forward_and_or_backward (EA:2261-2561) for every unit at once.
//
// The D-contractions (x·w_q|w_v EA:1923-1924, o·w_o EA:1995 and their VJPs) run on the hand-written tcgen05 + TMA GEMM of
// gemm_tc.cu (bf16 operands, fp32 accumulation); cuBLAS is only the fallback for shapes outside that kernel's tiling
// (N % 128 != 0 or K % 64 != 0 — none of the BASELINE configs) and the A/B reference (LSH_GEMM=cublas).
#include <cublas_v2.h>
#include <stdarg.h>
#include <stdio.h>

#include <mutex>

#include "attend_params.cuh"
#include "tc_common.cuh"

namespace lsh {

// ---- error + launch bookkeeping --------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

int set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
void count_launch(int n) { g_launches += n; }

extern long long *g_fwd_trace;   // attend_fwd.cu (debug trace buffer)
extern int g_bwd_parts;          // attend_bwd.cu (measurement hook)

// ---- kernels implemented in the other translation units --------------------------------------------
int hash_bf16_qv(const LshAttnDims &, const void *, const float *, const uint8_t *, int32_t *, int64_t, cudaStream_t);
int hash_f32_vecs(const LshAttnDims &, const float *, const float *, const uint8_t *, int32_t *, int64_t, cudaStream_t);
bool hash_can_fuse_aux(const LshAttnDims &);
int hash_bf16_qv_aux(const LshAttnDims &, const void *, const float *, const uint8_t *, int32_t *, int64_t, float *, float2 *, void *,
                     cudaStream_t);
size_t sort_workspace_bytes(const LshAttnDims &);
int sort_run(const LshAttnDims &, const int32_t *, int64_t, int32_t *, int32_t *, void *, size_t, cudaStream_t);
int attend_fwd_run(const LshAttnDims &, const void *, const int32_t *, const uint8_t *, void *, int64_t, int64_t,
                   int64_t, int64_t, float *, const FwdAux *, const AttnKeep *, cudaStream_t);
int combine_fwd_run(const LshAttnDims &, const void *, const float *, void *, float *, cudaStream_t, const BwdPrepOut * = nullptr);
bool attend_bwd_uses_tc(const LshAttnDims &, const AttnKeep *);
int attend_bwd_prep_target(const LshAttnDims &, void *, size_t, const void *, BwdPrepOut *);
int chunk_possort_run(const LshAttnDims &, const int32_t *, int32_t *, int32_t *, cudaStream_t);
size_t attend_bwd_workspace_bytes(const LshAttnDims &);
int attend_bwd_run(const LshAttnDims &, const void *, const int32_t *, const uint8_t *, const void *, const float *,
                   const void *, const float *, const int32_t *, const int32_t *, const AttnKeep *, void *, void *, size_t,
                   cudaStream_t, bool prep_done = false);
int pack_weights_run(const LshAttnDims &, const float *, const float *, const float *, const float *, void *, void *, cudaStream_t);
int pack_all_run(const LshAttnDims &, const float *, const float *, const float *, const float *, void *, void *, void *, void *, cudaStream_t);
int f32_to_bf16_run(const float *, void *, int64_t, cudaStream_t);
int unpack_dwqv_run(const LshAttnDims &, const float *, float *, float *, float *, cudaStream_t);
int pack_heads_run(int, int, int, int, const void *, int, const void *, int, void *, cudaStream_t);
int unpack_heads_run(int, int, int, int, const void *, int, int, int, void *, cudaStream_t);
int make_rotations_run(const LshAttnDims &, const uint32_t *, uint32_t *, float *, cudaStream_t);
int layernorm_fwd_run(int64_t, int, int, const void *, const float *, const float *, void *, float2 *, float, cudaStream_t, bool z_bf16 = false);
int layernorm_bwd_run(int64_t, int, int, const void *, const void *, const void *, const float2 *, const float *, void *, float *,
                      float *, cudaStream_t);
int residual_sub_run(int64_t, int, const void *, const void *, void *, float, cudaStream_t);
int predict_attend_run(const LshAttnDims &, const void *, int32_t *, int64_t, const int32_t *, int, float *, cudaStream_t);
int predict_out_run(const LshAttnDims &, const float *, const float *, void *, cudaStream_t);

// ---- dims ------------------------------------------------------------------------------------------
static int check_dims(const LshAttnDims *dp, bool need_bwd) {
  if (!dp) return set_error("dims == NULL");
  const LshAttnDims &d = *dp;
  if (d.B < 1 || d.H < 1 || d.L < 1 || d.D < 1) return set_error("B, H, L, D must be >= 1");
  if (static_cast<int64_t>(d.B) * d.H * d.L * (d.nh > 0 ? d.nh : 1) >= (1ll << 31))
    return set_error("B*H*n_hashes*seqlen = %lld does not fit 31 bits (row indices of the per-token kernels)",
                     static_cast<long long>(d.B) * d.H * d.L * d.nh);
  if (d.dq != 64 || d.dv != 64)
    return set_error("unsupported head size d_qk=%d d_v=%d: the sm_100a kernels are specialised for 64/64", d.dq, d.dv);
  if (d.C != 32 && d.C != 64 && d.C != 128 && d.C != 256)
    return set_error("unsupported chunk_len=%d (supported: 32, 64, 128, 256)", d.C);
  if (need_bwd && d.C == 32) return set_error("chunk_len=32 has no backward kernel (supported: 64, 128, 256)");
  if (d.nb < 0 || d.na < 0 || 1 + d.nb + d.na > 8) return set_error("n_chunks_before/after out of range");
  if (d.nh < 1 || d.nh > 64) return set_error("n_hashes=%d out of range", d.nh);
  Derived dr = derive(d);
  if (dr.N % d.C != 0) return set_error("n_hashes*seqlen=%d not divisible by chunk_len=%d (EA:210)", dr.N, d.C);
  if (dr.W % 64 != 0) return set_error("window of %d keys must be a multiple of 64", dr.W);
  if (d.n_factors < 1 || d.n_factors > 4) return set_error("n_factors=%d out of range (1..4)", d.n_factors);
  for (int i = 0; i < d.n_factors; ++i)
    if (d.factors[i] < 2 || (d.factors[i] & 1)) return set_error("hash factor %d must be even (EA:80, 87)", d.factors[i]);
  if (d.D % 8 != 0) return set_error("d_model=%d must be a multiple of 8", d.D);
  if (d.act_dtype != LSH_DTYPE_F32 && d.act_dtype != LSH_DTYPE_BF16) return set_error("bad act_dtype");
  if (d.separate_k && d.nh != 1) return set_error("separate_k (SelfAttention(share_qk=False)) has no hashing: n_hashes must be 1");
  // int32 sort key of EA:1947 must not wrap: max key = L*(nh*n_buckets - 1) + L - 1
  const int64_t maxkey = static_cast<int64_t>(d.L) * d.nh * dr.n_buckets - 1;
  if (maxkey >= (1ll << 31))
    return set_error("int32 sort key would wrap: seqlen*n_hashes*n_buckets = %lld >= 2^31 (EA:1947); "
                     "pass a smaller n_buckets", static_cast<long long>(maxkey + 1));
  return 0;
}

// ---- cuBLAS ----------------------------------------------------------------------------------------
static constexpr size_t kCublasWs = 32ull << 20;

// One handle per (thread, device): a thread that moves between devices gets that device's handle back instead of leaking
// the previous one (handles live for the life of the thread, like torch's own).
static constexpr int kMaxDevices = 64;
static cublasHandle_t get_handle() {
  static thread_local cublasHandle_t handles[kMaxDevices] = {};
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess || cur < 0 || cur >= kMaxDevices) return nullptr;
  if (handles[cur] == nullptr && cublasCreate(&handles[cur]) != CUBLAS_STATUS_SUCCESS) handles[cur] = nullptr;
  return handles[cur];
}

// Row-major C[M,N] = op(A)·op(B); A/B bf16, C bf16 or f32; fp32 accumulation.
static int gemm_rm(bool ta, bool tb, int64_t M, int64_t N, int64_t K, const void *A, int64_t lda, const void *B,
                   int64_t ldb, void *C, int64_t ldc, bool c_f32, void *cws, cudaStream_t stream) {
  cublasHandle_t h = get_handle();
  if (!h) return set_error("cublasCreate failed");
  cublasSetStream(h, stream);
  if (cws) cublasSetWorkspace(h, cws, kCublasWs);
  const float alpha = 1.f, beta = 0.f;
  cublasStatus_t st = cublasGemmEx(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, (int)N, (int)M,
                                   (int)K, &alpha, B, CUDA_R_16BF, (int)ldb, A, CUDA_R_16BF, (int)lda, &beta, C,
                                   c_f32 ? CUDA_R_32F : CUDA_R_16BF, (int)ldc, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
  if (st != CUBLAS_STATUS_SUCCESS) return set_error("cublasGemmEx failed with status %d", (int)st);
  return 0;
}

// Activation x weight product C[M, N] = A[M, K] · W: the hand-written tensor-core kernel (gemm_tc.cu) when the weight is
// at hand K-major (`w_k`: (N, K)) and the shape fits it, cuBLAS otherwise (`w_n`: the (K, N) copy, or w_k transposed).
int transpose_bf16_run(const void *src, void *dst, int R, int C, cudaStream_t stream);
static int gemm_aw(int64_t M, int64_t N, int64_t K, const void *A, int64_t lda, const void *w_k, const void *w_n, void *C,
                   int64_t ldc, bool c_f32, void *cws, cudaStream_t stream, const GemmQStats *qs = nullptr, bool *qs_done = nullptr,
                   const GemmResidual *res = nullptr) {
  if (qs_done) *qs_done = false;
  if (res && !res->resid) res = nullptr;
  if (w_k) {
    const int rc = gemm_tc_run(M, N, K, A, lda, w_k, K, C, ldc, c_f32, stream, qs, res);
    if (rc >= 0) {
      if (qs_done) *qs_done = rc == 0 && qs != nullptr;
      return rc;
    }
  }
  int rc;
  if (w_n) rc = gemm_rm(false, false, M, N, K, A, lda, w_n, N, C, ldc, c_f32, cws, stream);
  else rc = gemm_rm(false, true, M, N, K, A, lda, w_k, K, C, ldc, c_f32, cws, stream);
  if (rc || !res) return rc;
  // library-GEMM fallback: the residual as a separate pass, C = resid + sign * C
  if (ldc != N || (M * N) % 8 != 0) return set_error("residual epilogue: unsupported output pitch / size on the fallback path");
  return residual_sub_run(M * N, c_f32 ? LSH_DTYPE_F32 : LSH_DTYPE_BF16, res->resid, C, C, res->acc_sign, stream);
}

// Weight gradient C[M, N] (f32) = A^T · B, A (K, M), B (K, N) bf16: the split-K tensor-core kernel, else cuBLAS.
static int gemm_wgrad(int64_t M, int64_t N, int64_t K, const void *A, const void *B, float *C, void *cws, cudaStream_t stream,
                      const WgradUnpack *up = nullptr, bool *unpacked = nullptr) {
  if (unpacked) *unpacked = false;
  const int rc = gemm_tc_wgrad_run(M, N, K, A, M, B, N, C, cws, cws ? kCublasWs : 0, stream, up, unpacked);
  if (rc >= 0) return rc;
  return gemm_rm(true, false, M, N, K, A, M, B, N, C, N, true, cws, stream);
}

// ---- side stream of lsh_layer_bwd -----------------------------------------------------------------------
// do = dout·w_o^T depends only on the packed weights and on dout, not on the forward recompute: it runs on an internal
// stream (fork / join by events, no host wait; legal inside a stream capture) beside the recompute's latency-bound
// kernels (qscale, sort, position sort), which leave most SMs idle.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream *side_stream() {      // per (thread, device), created once
  static thread_local SideStream per_dev[kMaxDevices];
  int cur = 0;
  if (cudaGetDevice(&cur) != cudaSuccess || cur < 0 || cur >= kMaxDevices) return nullptr;
  SideStream &ss = per_dev[cur];
  if (ss.stream == nullptr) {
    cudaStream_t st = nullptr;
    cudaEvent_t f = nullptr, j = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&j, cudaEventDisableTiming) != cudaSuccess) {
      cudaStreamDestroy(st);
      if (f) cudaEventDestroy(f);
      return nullptr;
    }
    ss.stream = st; ss.fork = f; ss.join = j;
  }
  return &ss;
}
#define LSH_CUDA_OK(call)                                                                          \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return set_error("%s: %s", #call, cudaGetErrorString(e_));              \
  } while (0)

// ---- workspace carving -----------------------------------------------------------------------------
struct Bump {
  char *base; size_t off;
  explicit Bump(void *b) : base(static_cast<char *>(b)), off(0) {}
  void *take(size_t bytes) {
    off = (off + 255) / 256 * 256;
    void *p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

struct LayerWs {
  void *cublas, *cublas2, *xb, *wqv, *wo, *wqv_t, *wo_t, *qv, *o_rounds, *o_comb, *sort_ws, *doutb, *do_comb, *dqv, *bwd_ws, *keep_ws;
  int32_t *sticker;
  float *logits, *lse_tot, *dwqv;
  FwdAux aux;
  size_t sort_bytes, bwd_bytes, total;
};

static LayerWs carve(const LshAttnDims &d, void *ws, bool with_grad) {
  Derived dr = derive(d);
  LayerWs w{};
  Bump b(ws);
  const size_t BL = static_cast<size_t>(d.B) * d.L, rows = static_cast<size_t>(dr.BH) * dr.N;
  w.cublas = b.take(kCublasWs);
  w.xb = d.act_dtype == LSH_DTYPE_F32 ? b.take(BL * d.D * 2) : nullptr;
  w.wqv = b.take(static_cast<size_t>(d.D) * d.H * dr.QV * 2);
  w.wo = b.take(static_cast<size_t>(d.H) * d.dv * d.D * 2);
  w.wqv_t = b.take(static_cast<size_t>(d.D) * d.H * dr.QV * 2);      // K-major copies for the forward projections (gemm_tc.cu)
  w.wo_t = b.take(static_cast<size_t>(d.H) * d.dv * d.D * 2);
  w.qv = b.take(BL * d.H * dr.QV * 2);
  w.aux = fwd_aux_carve(d, b.take(fwd_aux_bytes(d)));
  w.keep_ws = b.take(attn_keep_bytes(d));
  w.sticker = static_cast<int32_t *>(b.take(rows * 4));
  w.sort_bytes = sort_workspace_bytes(d);
  w.sort_ws = b.take(w.sort_bytes);
  w.o_rounds = d.nh > 1 ? b.take(rows * d.dv * 2) : nullptr;
  w.logits = static_cast<float *>(b.take(rows * 4));
  w.o_comb = b.take(BL * d.H * d.dv * 2);
  w.lse_tot = d.nh > 1 ? static_cast<float *>(b.take(static_cast<size_t>(dr.BH) * d.L * 4)) : w.logits;
  if (with_grad) {
    w.cublas2 = b.take(kCublasWs);                       // the side-stream GEMM of lsh_layer_bwd needs its own scratch
    w.doutb = d.act_dtype == LSH_DTYPE_F32 ? b.take(BL * d.D * 2) : nullptr;
    w.do_comb = b.take(BL * d.H * d.dv * 2);
    w.dqv = b.take(BL * d.H * dr.QV * 2);
    w.dwqv = static_cast<float *>(b.take(static_cast<size_t>(d.D) * d.H * dr.QV * 4));
    w.bwd_bytes = attend_bwd_workspace_bytes(d);
    w.bwd_ws = b.take(w.bwd_bytes);
  }
  w.total = b.off + 256;
  return w;
}

// Packed bf16 weights in both orientations: (D, H*QV) / (H*dv, D) for the weight-gradient and cuBLAS paths, and their
// transposes (K-major for x·wqv and o·w_o) for the tensor-core GEMM.
static int pack_layer_weights(const LshAttnDims &d, const LayerWs &w, const float *w_q, const float *w_v, const float *w_o,
                              const float *w_k, cudaStream_t s) {
  Derived dr = derive(d);
  const int rc1 = pack_all_run(d, w_q, w_v, w_o, w_k, w.wqv, w.wqv_t, w.wo, w.wo_t, s);   // one launch for all four layouts
  if (rc1 >= 0) return rc1;
  if (int rc = pack_weights_run(d, w_q, w_v, w_o, w_k, w.wqv, w.wo, s)) return rc;
  if (int rc = transpose_bf16_run(w.wqv, w.wqv_t, d.D, d.H * dr.QV, s)) return rc;
  return transpose_bf16_run(w.wo, w.wo_t, d.H * d.dv, d.D, s);
}

// Forward up to o_comb (EA:1923-1992 for all units).  Returns xb (bf16 view of x).
static int forward_core(const LshAttnDims &d, const LayerWs &w, const void *x, const float *w_q, const float *w_v,
                        const float *w_o, const float *w_k, const float *rotations, const uint8_t *mask, const AttnKeep *keep, int32_t *buckets,
                        int64_t bstride, bool need_lse_tot, const void **xb_out, cudaStream_t s, bool weights_packed = false,
                        const BwdPrepOut *prep = nullptr, cudaEvent_t prep_wait = nullptr) {
  Derived dr = derive(d);
  const int64_t BL = static_cast<int64_t>(d.B) * d.L;
  int rc;
  const void *xb = x;
  if (d.act_dtype == LSH_DTYPE_F32 && !d.x_bf16) {        // (x_bf16: f32 activations whose layer input already is bf16, see lsh_attn.h)
    if ((rc = f32_to_bf16_run(static_cast<const float *>(x), w.xb, BL * d.D, s))) return rc;
    xb = w.xb;
  }
  *xb_out = xb;
  if (!weights_packed && (rc = pack_layer_weights(d, w, w_q, w_v, w_o, w_k, s))) return rc;
  const int64_t NQV = static_cast<int64_t>(d.H) * dr.QV;
  // tcgen05 attention path: the key scales / normalised keys come out of the projection's epilogue (fp32 accumulator ->
  // bf16 q -> statistics, in the registers that already hold the row)
  bool scales_done = false;
  GemmQStats qs = {w.aux.qscale, w.aux.rowmeta, w.aux.qhat, d.L, d.H};
  const bool want_qs = attend_fwd_uses_tc(d) && !d.separate_k;
  if ((rc = gemm_aw(BL, NQV, d.D, xb, d.D, w.wqv_t, w.wqv, w.qv, NQV, false, w.cublas, s, want_qs ? &qs : nullptr, &scales_done))) return rc;
  if (rotations) {
    if (!scales_done && attend_fwd_uses_tc(d) && hash_can_fuse_aux(d)) {
      if ((rc = hash_bf16_qv_aux(d, w.qv, rotations, mask, buckets, bstride, w.aux.qscale, w.aux.rowmeta, w.aux.qhat, s))) return rc;
      scales_done = true;
    } else if ((rc = hash_bf16_qv(d, w.qv, rotations, mask, buckets, bstride, s))) {
      return rc;
    }
  }
  if ((rc = sort_run(d, buckets, bstride, w.sticker, nullptr, w.sort_ws, w.sort_bytes, s))) return rc;
  if ((rc = fwd_aux_prepare(d, w.qv, w.sticker, w.aux, s, scales_done))) return rc;
  if (d.nh > 1) {
    if ((rc = attend_fwd_run(d, w.qv, w.sticker, mask, w.o_rounds, static_cast<int64_t>(d.H) * dr.N * 64,
                             static_cast<int64_t>(dr.N) * 64, static_cast<int64_t>(d.L) * 64, 64, w.logits, &w.aux, keep, s)))
      return rc;
    // backward call: the combine also leaves the gradient kernel's per-token inputs (needs do_comb: wait for the helper
    // stream's GEMM here, long after it has finished, instead of before the first kernel of the recompute)
    if (prep && prep_wait) LSH_CUDA_OK(cudaStreamWaitEvent(s, prep_wait, 0));
    if ((rc = combine_fwd_run(d, w.o_rounds, w.logits, w.o_comb, need_lse_tot ? w.lse_tot : nullptr, s, prep))) return rc;
  } else {
    // single round: rows land directly in the (B, L, H, dv) layout, logits == lse_tot
    if ((rc = attend_fwd_run(d, w.qv, w.sticker, mask, w.o_comb, static_cast<int64_t>(d.L) * d.H * 64, 64, 0,
                             static_cast<int64_t>(d.H) * 64, w.logits, &w.aux, keep, s)))
      return rc;
  }
  return 0;
}

// ---- fast inference (mode='predict'): scratch of one single-token step ---------------------------------------------
struct PredictWs {
  LayerWs lw;               // cublas, xb, wqv, wo, wqv_t, wo_t, qv (the fields pack_layer_weights / gemm_aw use)
  int32_t *hashed;
  float *o;
  size_t total;
};

static PredictWs carve_predict(const LshAttnDims &d, void *ws) {
  Derived dr = derive(d);
  PredictWs p{};
  Bump b(ws);
  const size_t BM = static_cast<size_t>(d.B) * d.L;
  p.lw.cublas = b.take(kCublasWs);
  p.lw.xb = d.act_dtype == LSH_DTYPE_F32 ? b.take(BM * d.D * 2) : nullptr;
  p.lw.wqv = b.take(static_cast<size_t>(d.D) * d.H * dr.QV * 2);
  p.lw.wo = b.take(static_cast<size_t>(d.H) * d.dv * d.D * 2);
  p.lw.wqv_t = b.take(static_cast<size_t>(d.D) * d.H * dr.QV * 2);
  p.lw.wo_t = b.take(static_cast<size_t>(d.H) * d.dv * d.D * 2);
  p.lw.qv = b.take(BM * d.H * dr.QV * 2);
  p.hashed = static_cast<int32_t *>(b.take(static_cast<size_t>(dr.BH) * dr.N * 4));
  p.o = static_cast<float *>(b.take(static_cast<size_t>(dr.BH) * 64 * 4));
  p.total = b.off + 256;
  return p;
}

}  // namespace lsh

using namespace lsh;

extern "C" {

int lsh_attn_abi_version(void) { return LSH_ATTN_ABI_VERSION; }
#ifndef LSH_SRC_HASH
#define LSH_SRC_HASH "unknown"
#endif
const char *lsh_attn_source_hash(void) { return LSH_SRC_HASH; }
const char *lsh_attn_last_error(void) { return g_err; }
int64_t lsh_attn_launch_count(int reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int lsh_attn_check_dims(const LshAttnDims *dims) { return check_dims(dims, false); }

int lsh_pack_weights(const LshAttnDims *dims, const float *w_q, const float *w_v, const float *w_o, const float *w_k,
                     void *wqv, void *wo, void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  return pack_weights_run(*dims, w_q, w_v, w_o, w_k, wqv, wo, static_cast<cudaStream_t>(stream));
}

int lsh_project_qv(const LshAttnDims *dims, const void *x_bf16, const void *wqv, void *qv, void *ws, size_t ws_bytes,
                   void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  const LshAttnDims &d = *dims;
  const int64_t BL = static_cast<int64_t>(d.B) * d.L, NQV = static_cast<int64_t>(d.H) * derive(d).QV;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const void *wt = nullptr;                                          // K-major copy in the workspace tail, if there is room
  const size_t wbytes = static_cast<size_t>(d.D) * NQV * 2;
  if (ws && ws_bytes >= kCublasWs + wbytes + 256) {
    void *t = static_cast<char *>(ws) + kCublasWs;
    if (int rc = transpose_bf16_run(wqv, t, d.D, static_cast<int>(NQV), s)) return rc;
    wt = t;
  }
  return gemm_aw(BL, NQV, d.D, x_bf16, d.D, wt, wqv, qv, NQV, false, ws_bytes >= kCublasWs ? ws : nullptr, s);
}

int lsh_hash(const LshAttnDims *dims, const void *qv, const float *rotations, const uint8_t *mask, int32_t *buckets,
             int64_t buckets_stride, void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  return hash_bf16_qv(*dims, qv, rotations, mask, buckets, buckets_stride, static_cast<cudaStream_t>(stream));
}

int lsh_hash_f32(const LshAttnDims *dims, const float *vecs, const float *rotations, const uint8_t *mask,
                 int32_t *buckets, int64_t buckets_stride, void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  return hash_f32_vecs(*dims, vecs, rotations, mask, buckets, buckets_stride, static_cast<cudaStream_t>(stream));
}

size_t lsh_sort_workspace_bytes(const LshAttnDims *dims) { return dims ? sort_workspace_bytes(*dims) : 0; }

int lsh_sort(const LshAttnDims *dims, const int32_t *buckets, int64_t buckets_stride, int32_t *sticker,
             int32_t *undo_sort, void *ws, size_t ws_bytes, void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  return sort_run(*dims, buckets, buckets_stride, sticker, undo_sort, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

size_t lsh_attend_fwd_workspace_bytes(const LshAttnDims *dims) {
  return dims ? fwd_aux_bytes(*dims) + attn_keep_bytes(*dims) : 0;
}

int lsh_attend_fwd(const LshAttnDims *dims, const void *qv, const int32_t *sticker, const uint8_t *mask,
                   const float *attn_keep, void *o_rounds, float *logits, void *ws, size_t ws_bytes, void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  const LshAttnDims &d = *dims;
  Derived dr = derive(d);
  if (!ws || ws_bytes < lsh_attend_fwd_workspace_bytes(dims)) return set_error("lsh_attend_fwd: workspace too small");
  FwdAux aux = fwd_aux_carve(d, ws);
  AttnKeep keep;
  if (int rc = attn_keep_prepare(d, attn_keep, static_cast<char *>(ws) + fwd_aux_bytes(d), &keep, static_cast<cudaStream_t>(stream))) return rc;
  if (int rc = fwd_aux_prepare(d, qv, sticker, aux, static_cast<cudaStream_t>(stream))) return rc;
  return attend_fwd_run(d, qv, sticker, mask, o_rounds, static_cast<int64_t>(d.H) * dr.N * 64,
                        static_cast<int64_t>(dr.N) * 64, static_cast<int64_t>(d.L) * 64, 64, logits, &aux,
                        attn_keep ? &keep : nullptr, static_cast<cudaStream_t>(stream));
}

int lsh_chunk_possort(const LshAttnDims *dims, const int32_t *sticker, int32_t *sticker2, int32_t *bounds, void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  if (!sticker || !sticker2 || sticker == sticker2) return set_error("lsh_chunk_possort: sticker / sticker2 must be distinct non-NULL buffers");
  return chunk_possort_run(*dims, sticker, sticker2, bounds, static_cast<cudaStream_t>(stream));
}

int lsh_combine_fwd(const LshAttnDims *dims, const void *o_rounds, const float *logits, void *o_comb, float *lse_tot,
                    void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  return combine_fwd_run(*dims, o_rounds, logits, o_comb, lse_tot, static_cast<cudaStream_t>(stream));
}

int lsh_project_out(const LshAttnDims *dims, const void *o_comb, const void *wo, void *out, void *ws, size_t ws_bytes,
                    void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  const LshAttnDims &d = *dims;
  const int64_t BL = static_cast<int64_t>(d.B) * d.L, KO = static_cast<int64_t>(d.H) * d.dv;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const void *wt = nullptr;
  const size_t wbytes = static_cast<size_t>(KO) * d.D * 2;
  if (ws && ws_bytes >= kCublasWs + wbytes + 256) {
    void *t = static_cast<char *>(ws) + kCublasWs;
    if (int rc = transpose_bf16_run(wo, t, static_cast<int>(KO), d.D, s)) return rc;
    wt = t;
  }
  return gemm_aw(BL, d.D, KO, o_comb, KO, wt, wo, out, d.D, d.act_dtype == LSH_DTYPE_F32, ws_bytes >= kCublasWs ? ws : nullptr, s);
}

size_t lsh_attend_bwd_workspace_bytes(const LshAttnDims *dims) {
  return dims ? attend_bwd_workspace_bytes(*dims) + attn_keep_bytes(*dims) : 0;
}

int lsh_attend_bwd(const LshAttnDims *dims, const void *qv, const int32_t *sticker, const uint8_t *mask,
                   const float *attn_keep, const void *o_comb, const float *lse_tot, const void *do_comb, void *dqv, void *ws,
                   size_t ws_bytes, void *stream) {
  if (int rc = check_dims(dims, true)) return rc;
  if (!ws || ws_bytes < lsh_attend_bwd_workspace_bytes(dims)) return set_error("lsh_attend_bwd: workspace too small");
  AttnKeep keep;
  const size_t core = attend_bwd_workspace_bytes(*dims);
  if (int rc = attn_keep_prepare(*dims, attn_keep, static_cast<char *>(ws) + core, &keep, static_cast<cudaStream_t>(stream))) return rc;
  return attend_bwd_run(*dims, qv, sticker, mask, o_comb, lse_tot, do_comb, nullptr, nullptr, nullptr,
                        attn_keep ? &keep : nullptr, dqv, ws, core, static_cast<cudaStream_t>(stream));
}

size_t lsh_layer_workspace_bytes(const LshAttnDims *dims, int with_grad) {
  if (check_dims(dims, with_grad != 0)) return 0;
  return carve(*dims, nullptr, with_grad != 0).total;
}

int lsh_layer_fwd_res(const LshAttnDims *dims, const void *x, const float *w_q, const float *w_v, const float *w_o,
                  const float *w_k, const float *rotations, const uint8_t *mask, const float *attn_keep, int32_t *buckets,
                  int64_t buckets_stride, void *out, const void *residual, float acc_sign, void *ws, size_t ws_bytes, void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  const LshAttnDims &d = *dims;
  if (!x || !w_q || !w_v || !w_o || !buckets || !out || !ws) return set_error("lsh_layer_fwd: NULL argument");
  if ((d.separate_k != 0) != (w_k != nullptr)) return set_error("lsh_layer_fwd: w_k must be given exactly when dims.separate_k is set");
  LayerWs w = carve(d, ws, false);
  if (ws_bytes < w.total) return set_error("lsh_layer_fwd: workspace too small (%zu < %zu)", ws_bytes, w.total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const void *xb;
  AttnKeep keep;
  if (int rc = attn_keep_prepare(d, attn_keep, w.keep_ws, &keep, s)) return rc;
  if (int rc = forward_core(d, w, x, w_q, w_v, w_o, w_k, rotations, mask, attn_keep ? &keep : nullptr, buckets, buckets_stride, false,
                            &xb, s))
    return rc;
  const int64_t BL = static_cast<int64_t>(d.B) * d.L, KO = static_cast<int64_t>(d.H) * d.dv;
  const GemmResidual res = {residual, acc_sign};
  return gemm_aw(BL, d.D, KO, w.o_comb, KO, w.wo_t, w.wo, out, d.D, d.act_dtype == LSH_DTYPE_F32, w.cublas, s, nullptr, nullptr, &res);
}

int lsh_layer_fwd(const LshAttnDims *dims, const void *x, const float *w_q, const float *w_v, const float *w_o,
                  const float *w_k, const float *rotations, const uint8_t *mask, const float *attn_keep, int32_t *buckets,
                  int64_t buckets_stride, void *out, void *ws, size_t ws_bytes, void *stream) {
  return lsh_layer_fwd_res(dims, x, w_q, w_v, w_o, w_k, rotations, mask, attn_keep, buckets, buckets_stride, out, nullptr, 1.f, ws,
                           ws_bytes, stream);
}

int lsh_layer_bwd_res(const LshAttnDims *dims, const void *x, const float *w_q, const float *w_v, const float *w_o,
                  const float *w_k, const uint8_t *mask, const float *attn_keep, const int32_t *buckets, int64_t buckets_stride,
                  const void *dout, void *out, void *dx, float *dw_q, float *dw_v, float *dw_o, float *dw_k, void *ws, size_t ws_bytes,
                  void *ev_dwo_ready, void *ev_dwqv_ready, const void *residual, float acc_sign, void *stream) {
  if (int rc = check_dims(dims, true)) return rc;
  const LshAttnDims &d = *dims;
  if (!x || !w_q || !w_v || !w_o || !buckets || !dout || !dx || !dw_q || !dw_v || !dw_o || !ws)
    return set_error("lsh_layer_bwd: NULL argument");
  if ((d.separate_k != 0) != (w_k != nullptr) || (d.separate_k != 0) != (dw_k != nullptr))
    return set_error("lsh_layer_bwd: w_k / dw_k must be given exactly when dims.separate_k is set");
  Derived dr = derive(d);
  LayerWs w = carve(d, ws, true);
  if (ws_bytes < w.total) return set_error("lsh_layer_bwd: workspace too small (%zu < %zu)", ws_bytes, w.total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const void *xb;
  int rc;
  const int64_t BL = static_cast<int64_t>(d.B) * d.L, KO = static_cast<int64_t>(d.H) * d.dv;
  const int64_t NQV = static_cast<int64_t>(d.H) * dr.QV;
  const bool f32 = d.act_dtype == LSH_DTYPE_F32;
  // B1 (first half) on the side stream: do = dout·w_o^T
  if ((rc = pack_layer_weights(d, w, w_q, w_v, w_o, w_k, s))) return rc;
  SideStream *side = side_stream();
  if (!side) return set_error("lsh_layer_bwd: could not create the internal stream");
  LSH_CUDA_OK(cudaEventRecord(side->fork, s));
  LSH_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
  const void *doutb = dout;
  if (f32) {
    if ((rc = f32_to_bf16_run(static_cast<const float *>(dout), w.doutb, BL * d.D, side->stream))) return rc;
    doutb = w.doutb;
  }
  if ((rc = gemm_aw(BL, KO, d.D, doutb, d.D, w.wo, nullptr, w.do_comb, KO, false, w.cublas2, side->stream))) return rc;
  LSH_CUDA_OK(cudaEventRecord(side->join, side->stream));
  // forward recompute on the caller's stream
  AttnKeep keep;
  if ((rc = attn_keep_prepare(d, attn_keep, w.keep_ws, &keep, s))) return rc;
  const AttnKeep *kp = attn_keep ? &keep : nullptr;
  BwdPrepOut prep;
  const bool fuse_prep = d.nh > 1 && attend_bwd_uses_tc(d, kp);
  if (fuse_prep && (rc = attend_bwd_prep_target(d, w.bwd_ws, w.bwd_bytes, w.do_comb, &prep))) return rc;
  if ((rc = forward_core(d, w, x, w_q, w_v, w_o, w_k, nullptr, mask, kp, const_cast<int32_t *>(buckets), buckets_stride, true,
                         &xb, s, /*weights_packed=*/true, fuse_prep ? &prep : nullptr, side->join)))
    return rc;
  if (out) {
    const GemmResidual res = {residual, acc_sign};
    if ((rc = gemm_aw(BL, d.D, KO, w.o_comb, KO, w.wo_t, w.wo, out, d.D, f32, w.cublas, s, nullptr, nullptr, &res))) return rc;
  }
  LSH_CUDA_OK(cudaStreamWaitEvent(s, side->join, 0));
  // B1 (second half): dW_o = o^T·dout
  if ((rc = gemm_wgrad(KO, d.D, BL, w.o_comb, doutb, dw_o, w.cublas, s))) return rc;
  if (ev_dwo_ready) LSH_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(ev_dwo_ready), s));
  // B2-B6
  if ((rc = attend_bwd_run(d, w.qv, w.sticker, mask, w.o_comb, w.lse_tot, w.do_comb, w.aux.qscale, attend_fwd_uses_tc(d) ? w.aux.sticker2 : nullptr, attend_fwd_uses_tc(d) ? w.aux.bounds : nullptr, kp, w.dqv, w.bwd_ws, w.bwd_bytes, s, fuse_prep)))
    return rc;
  // B7: dW_q|dW_v = x^T·dqv ; dx = dqv·wqv^T
  const WgradUnpack up = {dw_q, dw_v, dw_k, d.H, d.D, d.dq, d.dv};
  bool unpacked = false;
  if ((rc = gemm_wgrad(d.D, NQV, BL, xb, w.dqv, w.dwqv, w.cublas, s, &up, &unpacked))) return rc;
  if (!unpacked && (rc = unpack_dwqv_run(d, w.dwqv, dw_q, dw_v, dw_k, s))) return rc;
  if (ev_dwqv_ready) LSH_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(ev_dwqv_ready), s));
  return gemm_aw(BL, d.D, NQV, w.dqv, NQV, w.wqv, nullptr, dx, d.D, f32, w.cublas, s);
}

int lsh_layer_bwd(const LshAttnDims *dims, const void *x, const float *w_q, const float *w_v, const float *w_o,
                  const float *w_k, const uint8_t *mask, const float *attn_keep, const int32_t *buckets, int64_t buckets_stride,
                  const void *dout, void *out, void *dx, float *dw_q, float *dw_v, float *dw_o, float *dw_k, void *ws, size_t ws_bytes,
                  void *ev_dwo_ready, void *ev_dwqv_ready, void *stream) {
  return lsh_layer_bwd_res(dims, x, w_q, w_v, w_o, w_k, mask, attn_keep, buckets, buckets_stride, dout, out, dx, dw_q, dw_v, dw_o, dw_k,
                           ws, ws_bytes, ev_dwo_ready, ev_dwqv_ready, nullptr, 1.f, stream);
}

int lsh_layernorm_fwd(int64_t rows, int d_model, int act_dtype, const void *x, const float *scale, const float *bias, void *z,
                      float *stats, float epsilon, void *stream) {
  if (!x || !scale || !bias || !z) return set_error("lsh_layernorm_fwd: NULL argument");
  if (act_dtype != LSH_DTYPE_F32 && act_dtype != LSH_DTYPE_BF16) return set_error("lsh_layernorm_fwd: bad act_dtype");
  return layernorm_fwd_run(rows, d_model, act_dtype, x, scale, bias, z, reinterpret_cast<float2 *>(stats), epsilon,
                           static_cast<cudaStream_t>(stream));
}

int lsh_layernorm_fwd_bf16(int64_t rows, int d_model, const float *x, const float *scale, const float *bias, void *z_bf16,
                           float *stats, float epsilon, void *stream) {
  if (!x || !scale || !bias || !z_bf16) return set_error("lsh_layernorm_fwd_bf16: NULL argument");
  return layernorm_fwd_run(rows, d_model, LSH_DTYPE_F32, x, scale, bias, z_bf16, reinterpret_cast<float2 *>(stats), epsilon,
                           static_cast<cudaStream_t>(stream), true);
}

int lsh_layernorm_bwd(int64_t rows, int d_model, int act_dtype, const void *x, const void *dz, const void *ct_in,
                      const float *stats, const float *scale, void *ct_out, float *d_scale, float *d_bias, void *stream) {
  if (!x || !dz || !stats || !scale || !ct_out || !d_scale || !d_bias) return set_error("lsh_layernorm_bwd: NULL argument");
  if (act_dtype != LSH_DTYPE_F32 && act_dtype != LSH_DTYPE_BF16) return set_error("lsh_layernorm_bwd: bad act_dtype");
  return layernorm_bwd_run(rows, d_model, act_dtype, x, dz, ct_in, reinterpret_cast<const float2 *>(stats), scale, ct_out, d_scale,
                           d_bias, static_cast<cudaStream_t>(stream));
}

int lsh_residual_sub(int64_t n, int act_dtype, const void *a, const void *b, void *out, void *stream) {
  if (!a || !b || !out) return set_error("lsh_residual_sub: NULL argument");
  if (act_dtype != LSH_DTYPE_F32 && act_dtype != LSH_DTYPE_BF16) return set_error("lsh_residual_sub: bad act_dtype");
  return residual_sub_run(n, act_dtype, a, b, out, -1.f, static_cast<cudaStream_t>(stream));
}

int lsh_residual_add(int64_t n, int act_dtype, const void *a, const void *b, void *out, void *stream) {
  if (!a || !b || !out) return set_error("lsh_residual_add: NULL argument");
  if (act_dtype != LSH_DTYPE_F32 && act_dtype != LSH_DTYPE_BF16) return set_error("lsh_residual_add: bad act_dtype");
  return residual_sub_run(n, act_dtype, a, b, out, 1.f, static_cast<cudaStream_t>(stream));
}

int lsh_pack_heads(int B, int H, int L, int act_dtype, const void *a, int da, const void *b, int db, void *dst, void *stream) {
  if (!a || !dst) return set_error("lsh_pack_heads: NULL argument");
  if (act_dtype != LSH_DTYPE_F32 && act_dtype != LSH_DTYPE_BF16) return set_error("lsh_pack_heads: bad act_dtype");
  return pack_heads_run(B, H, L, act_dtype, a, da, b, db, dst, static_cast<cudaStream_t>(stream));
}

int lsh_unpack_heads(int B, int H, int L, int act_dtype, const void *src, int d_total, int col0, int d, void *dst, void *stream) {
  if (!src || !dst) return set_error("lsh_unpack_heads: NULL argument");
  if (act_dtype != LSH_DTYPE_F32 && act_dtype != LSH_DTYPE_BF16) return set_error("lsh_unpack_heads: bad act_dtype");
  return unpack_heads_run(B, H, L, act_dtype, src, d_total, col0, d, dst, static_cast<cudaStream_t>(stream));
}

size_t lsh_predict_workspace_bytes(const LshAttnDims *dims) {
  if (check_dims(dims, false)) return 0;
  return carve_predict(*dims, nullptr).total;
}

int lsh_predict_step(const LshAttnDims *dims, const void *mem, const float *w_q, const float *w_v, const float *w_o, const float *w_k,
                     const float *rotations, int32_t *buckets, int64_t buckets_stride, int32_t q_start, void *out, void *ws,
                     size_t ws_bytes, void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  const LshAttnDims &d = *dims;
  if (!mem || !w_q || !w_v || !w_o || !out || !ws) return set_error("lsh_predict_step: NULL argument");
  if ((d.separate_k != 0) != (w_k != nullptr)) return set_error("lsh_predict_step: w_k must be given exactly when dims.separate_k is set");
  if ((rotations != nullptr) != (buckets != nullptr)) return set_error("lsh_predict_step: rotations and buckets go together (LSH) or are both NULL");
  if (rotations && d.separate_k) return set_error("lsh_predict_step: hashing needs shared-QK");
  if (d.masked || d.na != 0) return set_error("lsh_predict_step: masked / n_chunks_after are not part of fast inference (EA:2085)");
  Derived dr = derive(d);
  if (buckets && buckets_stride < dr.N) return set_error("lsh_predict_step: buckets_stride %lld < n_hashes * memory length %d", (long long)buckets_stride, dr.N);
  PredictWs p = carve_predict(d, ws);
  if (ws_bytes < p.total) return set_error("lsh_predict_step: workspace too small (%zu < %zu)", ws_bytes, p.total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t BM = static_cast<int64_t>(d.B) * d.L, NQV = static_cast<int64_t>(d.H) * dr.QV;
  int rc;
  const void *xb = mem;
  if (d.act_dtype == LSH_DTYPE_F32) {
    if ((rc = f32_to_bf16_run(static_cast<const float *>(mem), p.lw.xb, BM * d.D, s))) return rc;
    xb = p.lw.xb;
  }
  if ((rc = pack_layer_weights(d, p.lw, w_q, w_v, w_o, w_k, s))) return rc;
  // EA:2064, 2087-2089: q of the new token, k and v of the attended slots — here: of every memory slot, one GEMM
  if ((rc = gemm_aw(BM, NQV, d.D, xb, d.D, p.lw.wqv_t, p.lw.wqv, p.lw.qv, NQV, false, p.lw.cublas, s))) return rc;
  if (rotations) {
    // EA:2066: the bit-exact hash of the training path over the memory's rows (only the new token's column is consumed)
    if ((rc = hash_bf16_qv(d, p.lw.qv, rotations, nullptr, p.hashed, dr.N, s))) return rc;
  }
  if ((rc = predict_attend_run(d, p.lw.qv, buckets, buckets_stride, rotations ? p.hashed : nullptr, q_start, p.o, s))) return rc;
  return predict_out_run(d, p.o, w_o, out, s);
}

size_t lsh_predict_attend_workspace_bytes(const LshAttnDims *dims) {
  if (check_dims(dims, false)) return 0;
  return static_cast<size_t>(derive(*dims).BH) * derive(*dims).N * 4 + 512;
}

int lsh_predict_attend(const LshAttnDims *dims, const void *qv, const float *rotations, int32_t *buckets, int64_t buckets_stride,
                       int32_t q_start, float *o, void *ws, size_t ws_bytes, void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  const LshAttnDims &d = *dims;
  if (!qv || !rotations || !buckets || !o || !ws) return set_error("lsh_predict_attend: NULL argument");
  if (d.separate_k || d.masked || d.na != 0) return set_error("lsh_predict_attend: shared-QK, unmasked, no look-ahead (EA:2823-2932)");
  Derived dr = derive(d);
  if (buckets_stride < dr.N) return set_error("lsh_predict_attend: buckets_stride %lld < n_hashes * memory length %d", (long long)buckets_stride, dr.N);
  if (ws_bytes < lsh_predict_attend_workspace_bytes(dims)) return set_error("lsh_predict_attend: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int32_t *hashed = reinterpret_cast<int32_t *>((reinterpret_cast<uintptr_t>(ws) + 255) / 256 * 256);
  if (int rc = hash_bf16_qv(d, qv, rotations, nullptr, hashed, dr.N, s)) return rc;      // EA:2890
  return predict_attend_run(d, qv, buckets, buckets_stride, hashed, q_start, o, s);
}

/* Debug aid (not in the public header): device buffer receiving per-phase clock stamps of CTA 0. */
int lsh_debug_set_trace(void *dev_ptr) { lsh::g_fwd_trace = static_cast<long long *>(dev_ptr); return 0; }

/* Measurement hook (not in the public header): see g_bwd_parts in attend_bwd.cu.  Returns the previous mask. */
int lsh_debug_set_bwd_parts(int mask) { const int old = lsh::g_bwd_parts; lsh::g_bwd_parts = mask & 7; return old; }

int lsh_make_rotations(const LshAttnDims *dims, const uint32_t *keys, uint32_t *new_keys, float *rotations,
                       void *stream) {
  if (int rc = check_dims(dims, false)) return rc;
  if (keys == new_keys) return set_error("lsh_make_rotations: new_keys must not alias keys");
  return make_rotations_run(*dims, keys, new_keys, rotations, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ---- TMA descriptor encoder (driver entry point fetched through the runtime: no link-time dependency on libcuda) ------
namespace lsh {
int make_tile_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_cols,
                  uint32_t box_rows) {
  using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (!encode) return set_error("cuTensorMapEncodeTiled is not available from this driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_bytes & 15)) return set_error("TMA: base / pitch must be 16-byte aligned");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return 0;
}
int make_row_gather_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_cols) {
  return make_tile_map(map, base, rows, cols, pitch_bytes, box_cols, 1);
}
}  // namespace lsh

/* Test hook (not in the public header): the tensor-core GEMM of gemm_tc.cu on caller-supplied operands,
 * C[M, N] = A[M, K] · B[N, K]^T (bf16 in, bf16 or f32 out).  Returns -1 when the kernel does not cover the shape. */
extern "C" int lsh_debug_gemm_tc_wgrad(int64_t M, int64_t N, int64_t K, const void *A, const void *B, float *C, void *scratch,
                                       size_t scratch_bytes, void *stream) {
  return lsh::gemm_tc_wgrad_run(M, N, K, A, M, B, N, C, scratch, scratch_bytes, static_cast<cudaStream_t>(stream));
}
extern "C" int lsh_debug_gemm_tc(int64_t M, int64_t N, int64_t K, const void *A, int64_t lda, const void *B, int64_t ldb, void *C,
                                 int64_t ldc, int c_f32, void *stream) {
  return lsh::gemm_tc_run(M, N, K, A, lda, B, ldb, C, ldc, c_f32 != 0, static_cast<cudaStream_t>(stream));
}
