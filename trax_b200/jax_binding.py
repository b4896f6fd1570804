"""Reference-side binding: `jax.ffi` registration of the handlers in csrc/jax_ffi_shim.cc and the
`forward_and_or_backward` a Trax `LSHSelfAttention` subclass delegates to.

Not importable in this image (no jax / jaxlib: SURVEY.md F2) — importing raises ImportError with that reason; nothing in
`trax_b200` depends on it.  With jaxlib present:

    lib = build_shim()                       # g++ on csrc/jax_ffi_shim.cc against jax.ffi.include_dir()
    register(lib)
    class LSHSelfAttention(trax.layers.research.efficient_attention.LSHSelfAttention):
      def forward_and_or_backward(self, inputs, weights, state, rng, output_grad=None, compute_output=True,
                                  update_state=True):
        return forward_and_or_backward(self, inputs, weights, state, rng, output_grad, compute_output, update_state)

`Layer.pure_fn` / `_do_custom_gradients` (trax/layers/base.py:541-673) and `ReversibleHalfResidual`
(trax/layers/reversible.py:281-286, 373-378) then use the layer unchanged: `has_backward`, `backward` and
`forward_and_or_backward` keep their signatures, and the `(buckets, rng)` state / `(w_q, w_v, w_o)` weights layouts are
the reference's.
"""
import ctypes
import os
import subprocess

try:
  import jax
  import jax.numpy as jnp
except ImportError as e:  # pragma: no cover - this image has no jax
  raise ImportError('trax_b200.jax_binding needs jax / jaxlib (absent from this image); use the torch-hosted layer '
                    '`trax_b200.LSHSelfAttention`, which calls the same C ABI through ctypes') from e

_HERE = os.path.dirname(os.path.abspath(__file__))


def build_shim(out=None):
  """Compiles csrc/jax_ffi_shim.cc against jaxlib's XLA FFI headers and links it to liblsh_attn_b200.so."""
  out = out or os.path.join(_HERE, 'liblsh_attn_jax.so')
  cmd = ['g++', '-std=c++17', '-O2', '-shared', '-fPIC', '-I' + jax.ffi.include_dir(), '-I/usr/local/cuda/include',
         os.path.join(_HERE, 'csrc', 'jax_ffi_shim.cc'), '-L' + _HERE, '-llsh_attn_b200', '-Wl,-rpath,' + _HERE,
         '-L/usr/local/cuda/lib64', '-lcudart', '-o', out]
  subprocess.check_call(cmd)
  return ctypes.CDLL(out)


def register(lib):
  jax.ffi.register_ffi_target('lsh_layer_fwd', jax.ffi.pycapsule(lib.LshLayerFwd), platform='CUDA')
  jax.ffi.register_ffi_target('lsh_layer_bwd', jax.ffi.pycapsule(lib.LshLayerBwd), platform='CUDA')
  jax.ffi.register_ffi_target('lsh_predict_step', jax.ffi.pycapsule(lib.LshPredictStep), platform='CUDA')


def _bucket_factors(n_buckets, seqlen, chunk_len):
  from trax_b200.ops import bucket_factors                       # EA:1890-1902
  return bucket_factors(n_buckets, seqlen, chunk_len)


def forward_and_or_backward(layer, inputs, weights, state, rng, output_grad=None, compute_output=True, update_state=True):
  """EA:2261-2289 on the custom calls.  Random bits stay JAX's: the per-unit rotations (EA:1927-1929, 91-93), the
  attention-dropout keep matrix (EA:255-262) and the output-dropout mask (EA:271-280, folded into w_o) are drawn here."""
  import numpy as np
  x = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
  mask = inputs[1] if isinstance(inputs, (tuple, list)) and len(inputs) > 1 else None
  w_q, w_v, w_o = weights
  buckets, hash_rng = state
  bsz, seqlen, d_model = x.shape
  n_heads, nh, cl = layer._n_heads, layer._n_hashes, layer._chunk_len   # pylint: disable=protected-access
  factors = _bucket_factors(layer._n_buckets, seqlen, cl)              # pylint: disable=protected-access
  rot_cols = sum(factors) // 2
  empty_f = jnp.zeros((0,), jnp.float32)
  attend_rng = output_rng = None
  if rng is not None:
    attend_rng, output_rng = jax.random.split(rng)                                     # EA:1920
  attn_keep = empty_f
  if layer._attention_dropout > 0.0:                                                   # EA:254-262
    keep_prob = 1.0 - layer._attention_dropout
    window = cl * (1 + layer._n_chunks_before + layer._n_chunks_after)
    attn_keep = jax.random.bernoulli(attend_rng, keep_prob, (cl, window)).astype(jnp.float32) / keep_prob
  out_mult = None
  if layer._output_dropout > 0.0:                                                      # EA:271-280
    keep_prob = 1.0 - layer._output_dropout
    out_mult = jax.random.bernoulli(output_rng, keep_prob, (d_model,)).astype(jnp.float32) / keep_prob
    w_o = w_o * out_mult
  attrs = dict(chunk_len=np.int32(cl), n_chunks_before=np.int32(layer._n_chunks_before),
               n_chunks_after=np.int32(layer._n_chunks_after), n_hashes=np.int32(nh),
               factors=np.asarray(factors, np.int32), causal=bool(layer._causal), masked=bool(layer._masked),
               separate_k=False)
  mask_u8 = jnp.zeros((0,), jnp.uint8) if mask is None else mask.astype(jnp.uint8)
  length = nh * (layer._max_length_for_buckets or seqlen)
  out = new_state = dx = dw = None
  if update_state:                                                                      # EA:1926-1937
    keys = jax.vmap(jax.random.split)(hash_rng)                                         # (B*H, 2, key): next state key, draw key
    rotations = jax.vmap(lambda k: jax.random.normal(k, (layer._d_qk, nh, rot_cols)))(keys[:, 1]).astype(jnp.float32)
    buckets, out = jax.ffi.ffi_call(
        'lsh_layer_fwd', (jax.ShapeDtypeStruct((bsz * n_heads, length), jnp.int32), jax.ShapeDtypeStruct(x.shape, x.dtype)))(
            x, w_q, w_v, w_o, empty_f, rotations, mask_u8, attn_keep, jnp.zeros((0,), jnp.int32), **attrs)
    new_state = (buckets, keys[:, 0])
  elif compute_output and output_grad is None:
    _, out = jax.ffi.ffi_call(
        'lsh_layer_fwd', (jax.ShapeDtypeStruct(buckets.shape, jnp.int32), jax.ShapeDtypeStruct(x.shape, x.dtype)))(
            x, w_q, w_v, w_o, empty_f, empty_f, mask_u8, attn_keep, buckets, **attrs)
  if output_grad is not None:
    f32 = lambda w: jax.ShapeDtypeStruct(w.shape, jnp.float32)
    out_b, dx, dw_q, dw_v, dw_o, _ = jax.ffi.ffi_call(
        'lsh_layer_bwd', (jax.ShapeDtypeStruct(x.shape, x.dtype), jax.ShapeDtypeStruct(x.shape, x.dtype), f32(w_q), f32(w_v),
                          f32(w_o), jax.ShapeDtypeStruct((0,), jnp.float32)))(
                              x, w_q, w_v, w_o, empty_f, mask_u8, attn_keep, buckets, output_grad, compute_output=bool(compute_output),
                              **attrs)
    if out_mult is not None:
      dw_o = dw_o * out_mult
    dw = (dw_q, dw_v, dw_o)
    if compute_output:
      out = out_b
    if mask is not None:
      dx = (dx, jnp.zeros_like(mask))
  if not compute_output:
    out = None
  return out, new_state, dx, dw


def predict_step(layer, mem, weights, buckets, hash_rng, q_start):
  """One fast-inference step (EA:2032-2109) on the `lsh_predict_step` custom call: `mem` (B, M, D) is the input memory with the
  new token stored at `q_start` (`_use_predict_mem`, EA:2174-2207, stays JAX code), `buckets` (B*H, n_hashes*M) the bucket
  memory after `roll_buckets` (EA:2036-2053).  The rotations are drawn from the state's key itself — predict mode does not
  split it (EA:2066).  Returns (output (B, 1, D), new bucket memory)."""
  import numpy as np
  w_q, w_v, w_o = weights
  nh, cl = layer._n_hashes, layer._chunk_len                           # pylint: disable=protected-access
  factors = _bucket_factors(layer._n_buckets, 2, cl)                   # pylint: disable=protected-access  (EA:2064: two rows)
  rotations = jax.vmap(lambda k: jax.random.normal(k, (layer._d_qk, nh, sum(factors) // 2)))(hash_rng).astype(jnp.float32)
  new_buckets, out = jax.ffi.ffi_call(
      'lsh_predict_step', (jax.ShapeDtypeStruct(buckets.shape, jnp.int32),
                           jax.ShapeDtypeStruct((mem.shape[0], 1, mem.shape[2]), mem.dtype)))(
                               mem, w_q, w_v, w_o, jnp.zeros((0,), jnp.float32), rotations, buckets,
                               chunk_len=np.int32(cl), n_chunks_before=np.int32(layer._n_chunks_before), n_hashes=np.int32(nh),
                               factors=np.asarray(factors, np.int32), causal=True, separate_k=False, q_start=np.int32(q_start))
  return out, new_buckets
