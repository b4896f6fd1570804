"""ctypes binding of liblsh_attn_b200.so (C ABI: include/lsh_attn.h).

There is no CPU / eager fallback: if the shared library cannot be loaded this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LSH_ATTN_LIB') or os.path.join(_HERE, 'liblsh_attn_b200.so')   # override: kernel experiments

LSH_DTYPE_F32, LSH_DTYPE_BF16 = 0, 1
ABI_VERSION = 5


class LshAttnDims(ctypes.Structure):
  """Mirror of `struct LshAttnDims` (include/lsh_attn.h)."""
  _fields_ = [
      ('B', ctypes.c_int32), ('H', ctypes.c_int32), ('L', ctypes.c_int32), ('D', ctypes.c_int32),
      ('dq', ctypes.c_int32), ('dv', ctypes.c_int32),
      ('C', ctypes.c_int32), ('nb', ctypes.c_int32), ('na', ctypes.c_int32), ('nh', ctypes.c_int32),
      ('n_factors', ctypes.c_int32), ('factors', ctypes.c_int32 * 4),
      ('causal', ctypes.c_int32), ('masked', ctypes.c_int32),
      ('act_dtype', ctypes.c_int32), ('separate_k', ctypes.c_int32), ('x_bf16', ctypes.c_int32), ('reserved', ctypes.c_int32 * 1),
  ]


# name -> (restype, argtypes); the single source of truth checked against include/lsh_attn.h in tests.
_P, _I64, _SZ, _I = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t, ctypes.c_int
_D = ctypes.POINTER(LshAttnDims)
SIGNATURES = {
    'lsh_attn_abi_version': (_I, []),
    'lsh_attn_source_hash': (ctypes.c_char_p, []),
    'lsh_attn_last_error': (ctypes.c_char_p, []),
    'lsh_attn_check_dims': (_I, [_D]),
    'lsh_pack_weights': (_I, [_D, _P, _P, _P, _P, _P, _P, _P]),
    'lsh_project_qv': (_I, [_D, _P, _P, _P, _P, _SZ, _P]),
    'lsh_hash': (_I, [_D, _P, _P, _P, _P, _I64, _P]),
    'lsh_hash_f32': (_I, [_D, _P, _P, _P, _P, _I64, _P]),
    'lsh_sort_workspace_bytes': (_SZ, [_D]),
    'lsh_sort': (_I, [_D, _P, _I64, _P, _P, _P, _SZ, _P]),
    'lsh_attend_fwd_workspace_bytes': (_SZ, [_D]),
    'lsh_attend_fwd': (_I, [_D, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    'lsh_chunk_possort': (_I, [_D, _P, _P, _P, _P]),
    'lsh_combine_fwd': (_I, [_D, _P, _P, _P, _P, _P]),
    'lsh_project_out': (_I, [_D, _P, _P, _P, _P, _SZ, _P]),
    'lsh_attend_bwd_workspace_bytes': (_SZ, [_D]),
    'lsh_attend_bwd': (_I, [_D, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    'lsh_layer_workspace_bytes': (_SZ, [_D, _I]),
    'lsh_layer_fwd': (_I, [_D, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _SZ, _P]),
    'lsh_layer_bwd': (_I, [_D, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P, _P, _P]),
    'lsh_layer_fwd_res': (_I, [_D, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, ctypes.c_float, _P, _SZ, _P]),
    'lsh_layer_bwd_res': (_I, [_D, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P, _P, _P, ctypes.c_float, _P]),
    'lsh_make_rotations': (_I, [_D, _P, _P, _P, _P]),
    'lsh_layernorm_fwd': (_I, [_I64, _I, _I, _P, _P, _P, _P, _P, ctypes.c_float, _P]),
    'lsh_layernorm_fwd_bf16': (_I, [_I64, _I, _P, _P, _P, _P, _P, ctypes.c_float, _P]),
    'lsh_layernorm_bwd': (_I, [_I64, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'lsh_residual_sub': (_I, [_I64, _I, _P, _P, _P, _P]),
    'lsh_residual_add': (_I, [_I64, _I, _P, _P, _P, _P]),
    'lsh_pack_heads': (_I, [_I, _I, _I, _I, _P, _I, _P, _I, _P, _P]),
    'lsh_unpack_heads': (_I, [_I, _I, _I, _I, _P, _I, _I, _I, _P, _P]),
    'lsh_predict_workspace_bytes': (_SZ, [_D]),
    'lsh_predict_step': (_I, [_D, _P, _P, _P, _P, _P, _P, _P, _I64, ctypes.c_int32, _P, _P, _SZ, _P]),
    'lsh_predict_attend_workspace_bytes': (_SZ, [_D]),
    'lsh_predict_attend': (_I, [_D, _P, _P, _P, _I64, ctypes.c_int32, _P, _P, _SZ, _P]),
    'lsh_attn_launch_count': (_I64, [_I]),
}

_lib = None


class LshAttnError(RuntimeError):
  pass


def load():
  """Loads the shared library (building it in-tree with nvcc if it is absent and nvcc exists)."""
  global _lib
  if _lib is not None:
    return _lib
  from trax_b200 import build as _build
  override = bool(os.environ.get('LSH_ATTN_LIB'))      # an explicitly named library (kernel A/B experiments) is taken as is
  if not override and _build.needs_build():
    # missing, or compiled from other sources than the ones beside it (content hash): rebuild, never load a stale library
    try:
      _build.build()
    except Exception as e:  # pylint: disable=broad-except
      raise ImportError(
          'trax_b200: %s is %s and could not be built (%s). There is no CPU fallback.'
          % (LIB_PATH, 'stale' if os.path.exists(LIB_PATH) else 'missing', e)) from e
  try:
    import torch  # noqa: F401  pylint: disable=unused-import  (loads the cuBLAS/cudart the .so links to)
  except ImportError:
    pass
  lib = ctypes.CDLL(LIB_PATH)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)   # AttributeError if the symbol is not exported
    fn.restype, fn.argtypes = res, args
  if lib.lsh_attn_abi_version() != ABI_VERSION:
    raise ImportError('trax_b200: ABI version mismatch (%d != %d)' % (lib.lsh_attn_abi_version(), ABI_VERSION))
  if not override and lib.lsh_attn_source_hash().decode() != _build.source_hash():
    raise ImportError('trax_b200: %s was compiled from different sources (%s != %s)'
                      % (LIB_PATH, lib.lsh_attn_source_hash().decode(), _build.source_hash()))
  _lib = lib
  return lib


def check(rc, what=''):
  if rc != 0:
    msg = load().lsh_attn_last_error().decode('utf-8', 'replace')
    raise LshAttnError('%s: %s' % (what or 'lsh_attn', msg))


def make_dims(B, H, L, D, dq, dv, C, nb, na, nh, factors, causal, masked, act_dtype, separate_k=False):
  d = LshAttnDims()
  d.B, d.H, d.L, d.D, d.dq, d.dv = B, H, L, D, dq, dv
  d.C, d.nb, d.na, d.nh = C, nb, na, nh
  if not 1 <= len(factors) <= 4:
    raise ValueError('n_buckets factor list must have 1..4 entries, got %r' % (factors,))
  d.n_factors = len(factors)
  for i, f in enumerate(factors):
    d.factors[i] = int(f)
  d.causal, d.masked, d.act_dtype = int(bool(causal)), int(bool(masked)), act_dtype
  d.separate_k = int(bool(separate_k))
  return d
