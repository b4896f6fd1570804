// jax.ffi custom-call handlers over the C ABI of include/lsh_attn.h — the binding a Trax maintainer registers so that
// `trax.layers.research.efficient_attention.LSHSelfAttention.forward_and_or_backward` (EA:2261-2561) runs on
// liblsh_attn_b200.so (see INTEGRATION.md section 2 and trax_b200/jax_binding.py for the Python side).
//
// This translation unit needs XLA's FFI headers (`xla/ffi/api/ffi.h`, shipped inside jaxlib:
// `python -c "import jax.ffi; print(jax.ffi.include_dir())"`).  This image has no jaxlib, so the body is compiled out
// here (the file still goes through the compiler in tests/test_abi.py to prove it parses as C++ and that the guard
// works); where jaxlib exists:
//     g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -Iinclude \
//         -I/usr/local/cuda/include trax_b200/csrc/jax_ffi_shim.cc -Ltrax_b200 -llsh_attn_b200 -o liblsh_attn_jax.so
//
// Division of labour: random bits stay on the JAX side — the rotations (EA:91-93), the attention
// dropout keep matrix (EA:258-262) are drawn with jax.random from the layer's keys and passed in as buffers; output
// dropout is folded into w_o by the Python side (EA:271-280 is a column scaling).  Hyper-parameters travel as
// attributes, scratch comes from XLA's allocator, errors become xla::ffi::Error with the library's message.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define LSH_HAVE_XLA_FFI 1
#endif
#endif

#ifdef LSH_HAVE_XLA_FFI

#include <cuda_runtime_api.h>

#include <cstdint>

#include "../../include/lsh_attn.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// LshAttnDims from the operand shapes (x (B, L, D), w_q (H, D, dq), w_v (H, D, dv)) and the attributes.
ffi::ErrorOr<LshAttnDims> MakeDims(const ffi::AnyBuffer &x, const ffi::Buffer<ffi::F32> &w_q, const ffi::Buffer<ffi::F32> &w_v,
                                   int32_t chunk_len, int32_t n_chunks_before, int32_t n_chunks_after, int32_t n_hashes,
                                   ffi::Span<const int32_t> factors, bool causal, bool masked, bool separate_k) {
  if (x.dimensions().size() != 3 || w_q.dimensions().size() != 3 || w_v.dimensions().size() != 3)
    return ffi::Unexpected(ffi::Error(ffi::ErrorCode::kInvalidArgument, "lsh_attn: x must be (B, L, D), weights (H, D, d)"));
  if (factors.size() < 1 || factors.size() > 4)
    return ffi::Unexpected(ffi::Error(ffi::ErrorCode::kInvalidArgument, "lsh_attn: 1..4 bucket factors (EA:1893-1902)"));
  LshAttnDims d = {};
  d.B = static_cast<int32_t>(x.dimensions()[0]);
  d.L = static_cast<int32_t>(x.dimensions()[1]);
  d.D = static_cast<int32_t>(x.dimensions()[2]);
  d.H = static_cast<int32_t>(w_q.dimensions()[0]);
  d.dq = static_cast<int32_t>(w_q.dimensions()[2]);
  d.dv = static_cast<int32_t>(w_v.dimensions()[2]);
  d.C = chunk_len; d.nb = n_chunks_before; d.na = n_chunks_after; d.nh = n_hashes;
  d.n_factors = static_cast<int32_t>(factors.size());
  for (size_t i = 0; i < factors.size(); ++i) d.factors[i] = factors[i];
  d.causal = causal; d.masked = masked; d.separate_k = separate_k;
  switch (x.element_type()) {
    case ffi::F32: d.act_dtype = LSH_DTYPE_F32; break;
    case ffi::BF16: d.act_dtype = LSH_DTYPE_BF16; break;
    default: return ffi::Unexpected(ffi::Error(ffi::ErrorCode::kInvalidArgument, "lsh_attn: activations must be f32 or bf16"));
  }
  if (lsh_attn_check_dims(&d)) return ffi::Unexpected(ffi::Error(ffi::ErrorCode::kInvalidArgument, lsh_attn_last_error()));
  return d;
}

inline ffi::Error Status(int rc) {
  return rc ? ffi::Error(ffi::ErrorCode::kInternal, lsh_attn_last_error()) : ffi::Error::Success();
}

// An optional operand is passed as a zero-element buffer.
template <typename B>
inline auto *OrNull(B &b) { return b.element_count() ? b.typed_data() : nullptr; }

// compute_output=True [, update_state=True when `rotations` is non-empty]: EA:2283-2288 -> (output, buckets).
ffi::Error LayerFwdImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer x, ffi::Buffer<ffi::F32> w_q,
                        ffi::Buffer<ffi::F32> w_v, ffi::Buffer<ffi::F32> w_o, ffi::Buffer<ffi::F32> w_k,
                        ffi::Buffer<ffi::F32> rotations, ffi::Buffer<ffi::U8> mask, ffi::Buffer<ffi::F32> attn_keep,
                        ffi::Buffer<ffi::S32> buckets_in, ffi::Result<ffi::Buffer<ffi::S32>> buckets,
                        ffi::Result<ffi::AnyBuffer> out, int32_t chunk_len, int32_t n_chunks_before, int32_t n_chunks_after,
                        int32_t n_hashes, ffi::Span<const int32_t> factors, bool causal, bool masked, bool separate_k) {
  auto dims = MakeDims(x, w_q, w_v, chunk_len, n_chunks_before, n_chunks_after, n_hashes, factors, causal, masked, separate_k);
  if (dims.has_error()) return dims.error();
  LshAttnDims d = dims.value();
  const size_t ws_bytes = lsh_layer_workspace_bytes(&d, 0);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "lsh_attn: workspace allocation failed");
  const int64_t stride = buckets->dimensions()[1];
  if (!rotations.element_count()) {   // update_state=False: the stored buckets are read (EA:1939-1941); results alias-free
    if (cudaMemcpyAsync(buckets->typed_data(), buckets_in.typed_data(), buckets_in.size_bytes(), cudaMemcpyDeviceToDevice,
                        stream) != cudaSuccess)
      return ffi::Error(ffi::ErrorCode::kInternal, "lsh_attn: bucket copy failed");
  }
  return Status(lsh_layer_fwd(&d, x.untyped_data(), w_q.typed_data(), w_v.typed_data(), w_o.typed_data(), OrNull(w_k),
                              OrNull(rotations), OrNull(mask), OrNull(attn_keep), buckets->typed_data(), stride,
                              out->untyped_data(), *ws, ws_bytes, stream));
}

// output_grad given, update_state=False (the call `backward` and ReversibleHalfResidual.reverse_and_grad make,
// EA:2251-2259, reversible.py:374-378) -> (output, dx, dw_q, dw_v, dw_o, dw_k).
ffi::Error LayerBwdImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer x, ffi::Buffer<ffi::F32> w_q,
                        ffi::Buffer<ffi::F32> w_v, ffi::Buffer<ffi::F32> w_o, ffi::Buffer<ffi::F32> w_k, ffi::Buffer<ffi::U8> mask,
                        ffi::Buffer<ffi::F32> attn_keep, ffi::Buffer<ffi::S32> buckets, ffi::AnyBuffer dout,
                        ffi::Result<ffi::AnyBuffer> out, ffi::Result<ffi::AnyBuffer> dx, ffi::Result<ffi::Buffer<ffi::F32>> dw_q,
                        ffi::Result<ffi::Buffer<ffi::F32>> dw_v, ffi::Result<ffi::Buffer<ffi::F32>> dw_o,
                        ffi::Result<ffi::Buffer<ffi::F32>> dw_k, int32_t chunk_len, int32_t n_chunks_before,
                        int32_t n_chunks_after, int32_t n_hashes, ffi::Span<const int32_t> factors, bool causal, bool masked,
                        bool separate_k, bool compute_output) {
  auto dims = MakeDims(x, w_q, w_v, chunk_len, n_chunks_before, n_chunks_after, n_hashes, factors, causal, masked, separate_k);
  if (dims.has_error()) return dims.error();
  LshAttnDims d = dims.value();
  const size_t ws_bytes = lsh_layer_workspace_bytes(&d, 1);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "lsh_attn: workspace allocation failed");
  return Status(lsh_layer_bwd(&d, x.untyped_data(), w_q.typed_data(), w_v.typed_data(), w_o.typed_data(), OrNull(w_k),
                              OrNull(mask), OrNull(attn_keep), buckets.typed_data(), buckets.dimensions()[1],
                              dout.untyped_data(), compute_output ? out->untyped_data() : nullptr, dx->untyped_data(),
                              dw_q->typed_data(), dw_v->typed_data(), dw_o->typed_data(),
                              separate_k ? dw_k->typed_data() : nullptr, *ws, ws_bytes, /*ev_dwo_ready=*/nullptr,
                              /*ev_dwqv_ready=*/nullptr, stream));
}

// One fast-inference step (mode='predict', EA:2032-2109): `mem` is the input memory with the new token already stored at
// q_start, `buckets_in` the bucket memory after the host-side roll (EA:2036-2053); -> (output (B, 1, D), bucket memory with
// the new token's ids in column q_start).  rotations / buckets_in empty: SelfAttention (EA:1200-1268).
ffi::Error PredictStepImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer mem, ffi::Buffer<ffi::F32> w_q,
                           ffi::Buffer<ffi::F32> w_v, ffi::Buffer<ffi::F32> w_o, ffi::Buffer<ffi::F32> w_k,
                           ffi::Buffer<ffi::F32> rotations, ffi::Buffer<ffi::S32> buckets_in,
                           ffi::Result<ffi::Buffer<ffi::S32>> buckets, ffi::Result<ffi::AnyBuffer> out, int32_t chunk_len,
                           int32_t n_chunks_before, int32_t n_hashes, ffi::Span<const int32_t> factors, bool causal,
                           bool separate_k, int32_t q_start) {
  auto dims = MakeDims(mem, w_q, w_v, chunk_len, n_chunks_before, /*n_chunks_after=*/0, n_hashes, factors, causal, /*masked=*/false,
                       separate_k);
  if (dims.has_error()) return dims.error();
  LshAttnDims d = dims.value();
  const size_t ws_bytes = lsh_predict_workspace_bytes(&d);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "lsh_attn: workspace allocation failed");
  const bool hashed = rotations.element_count() != 0;
  if (hashed && cudaMemcpyAsync(buckets->typed_data(), buckets_in.typed_data(), buckets_in.size_bytes(), cudaMemcpyDeviceToDevice,
                                stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "lsh_attn: bucket copy failed");
  return Status(lsh_predict_step(&d, mem.untyped_data(), w_q.typed_data(), w_v.typed_data(), w_o.typed_data(), OrNull(w_k),
                                 OrNull(rotations), hashed ? buckets->typed_data() : nullptr,
                                 hashed ? buckets->dimensions()[1] : 0, q_start, out->untyped_data(), *ws, ws_bytes, stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    LshPredictStep, PredictStepImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::AnyBuffer>()            // mem (B, M, D)
        .Arg<ffi::Buffer<ffi::F32>>()     // w_q
        .Arg<ffi::Buffer<ffi::F32>>()     // w_v
        .Arg<ffi::Buffer<ffi::F32>>()     // w_o
        .Arg<ffi::Buffer<ffi::F32>>()     // w_k           (empty unless separate_k)
        .Arg<ffi::Buffer<ffi::F32>>()     // rotations     (empty: no hashing, SelfAttention)
        .Arg<ffi::Buffer<ffi::S32>>()     // buckets_in    (bucket memory; empty likewise)
        .Ret<ffi::Buffer<ffi::S32>>()     // buckets
        .Ret<ffi::AnyBuffer>()            // out (B, 1, D)
        .Attr<int32_t>("chunk_len")
        .Attr<int32_t>("n_chunks_before")
        .Attr<int32_t>("n_hashes")
        .Attr<ffi::Span<const int32_t>>("factors")
        .Attr<bool>("causal")
        .Attr<bool>("separate_k")
        .Attr<int32_t>("q_start"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    LshLayerFwd, LayerFwdImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::AnyBuffer>()            // x
        .Arg<ffi::Buffer<ffi::F32>>()     // w_q
        .Arg<ffi::Buffer<ffi::F32>>()     // w_v
        .Arg<ffi::Buffer<ffi::F32>>()     // w_o
        .Arg<ffi::Buffer<ffi::F32>>()     // w_k           (empty unless separate_k)
        .Arg<ffi::Buffer<ffi::F32>>()     // rotations     (empty: update_state=False)
        .Arg<ffi::Buffer<ffi::U8>>()      // mask          (empty unless masked)
        .Arg<ffi::Buffer<ffi::F32>>()     // attn_keep     (empty without attention dropout)
        .Arg<ffi::Buffer<ffi::S32>>()     // buckets_in    (state; read when update_state=False)
        .Ret<ffi::Buffer<ffi::S32>>()     // buckets
        .Ret<ffi::AnyBuffer>()            // out
        .Attr<int32_t>("chunk_len")
        .Attr<int32_t>("n_chunks_before")
        .Attr<int32_t>("n_chunks_after")
        .Attr<int32_t>("n_hashes")
        .Attr<ffi::Span<const int32_t>>("factors")
        .Attr<bool>("causal")
        .Attr<bool>("masked")
        .Attr<bool>("separate_k"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    LshLayerBwd, LayerBwdImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::AnyBuffer>()            // x
        .Arg<ffi::Buffer<ffi::F32>>()     // w_q
        .Arg<ffi::Buffer<ffi::F32>>()     // w_v
        .Arg<ffi::Buffer<ffi::F32>>()     // w_o
        .Arg<ffi::Buffer<ffi::F32>>()     // w_k
        .Arg<ffi::Buffer<ffi::U8>>()      // mask
        .Arg<ffi::Buffer<ffi::F32>>()     // attn_keep
        .Arg<ffi::Buffer<ffi::S32>>()     // buckets (state)
        .Arg<ffi::AnyBuffer>()            // dout
        .Ret<ffi::AnyBuffer>()            // out (written iff compute_output)
        .Ret<ffi::AnyBuffer>()            // dx
        .Ret<ffi::Buffer<ffi::F32>>()     // dw_q
        .Ret<ffi::Buffer<ffi::F32>>()     // dw_v
        .Ret<ffi::Buffer<ffi::F32>>()     // dw_o
        .Ret<ffi::Buffer<ffi::F32>>()     // dw_k
        .Attr<int32_t>("chunk_len")
        .Attr<int32_t>("n_chunks_before")
        .Attr<int32_t>("n_chunks_after")
        .Attr<int32_t>("n_hashes")
        .Attr<ffi::Span<const int32_t>>("factors")
        .Attr<bool>("causal")
        .Attr<bool>("masked")
        .Attr<bool>("separate_k")
        .Attr<bool>("compute_output"));

#else  // !LSH_HAVE_XLA_FFI

// jaxlib's headers are not on the include path: nothing to register.  The symbol below lets a loader (and the test
// suite) tell a shim built without XLA from one built with it.
extern "C" int lsh_attn_jax_ffi_available(void) { return 0; }

#endif
