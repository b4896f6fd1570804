"""Stage-level wrappers: torch tensors in, C-ABI calls (include/lsh_attn.h) underneath.

torch is used for device memory and the current CUDA stream only.  Every function requires CUDA
tensors and raises otherwise — there is no CPU path.
"""
import ctypes
import math

import torch

from trax_b200 import _lib


def bucket_factors(n_buckets, seqlen, chunk_len):
  """Resolves `n_buckets` to the factor list hash_vecs receives (EA:1890-1902)."""
  if n_buckets is None:
    n = 2 * max(1, seqlen // chunk_len)
    if n <= 128:
      return [n]
    div = 2 ** math.ceil(math.log2(math.sqrt(n)))
    return [div, 2 * (n // (2 * div))]
  if isinstance(n_buckets, int):
    return [n_buckets]
  return [int(f) for f in n_buckets]


def _ptr(t):
  if t is None:
    return None
  if not t.is_cuda:
    raise _lib.LshAttnError('trax_b200 ops need CUDA tensors (no CPU fallback)')
  if not t.is_contiguous():
    raise _lib.LshAttnError('trax_b200 ops need contiguous tensors')
  return ctypes.c_void_p(t.data_ptr())


def _stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _act_dtype(t):
  if t.dtype == torch.float32:
    return _lib.LSH_DTYPE_F32
  if t.dtype == torch.bfloat16:
    return _lib.LSH_DTYPE_BF16
  raise _lib.LshAttnError('activations must be float32 or bfloat16, got %s' % t.dtype)


_WS = {}


def workspace(device, nbytes):
  """One cached scratch buffer per (device, current stream), grown on demand; the library never allocates.  Keyed by
  the stream as well: two layers running concurrently on different streams must not share scratch memory."""
  key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
  buf = _WS.get(key)
  if buf is None or buf.numel() < nbytes:
    buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
    _WS[key] = buf
  return buf


def _row_width(dims):
  """Columns of one (token, head) row of the packed projections: q | v, plus k with separate keys."""
  return dims.dq + dims.dv + (dims.dq if dims.separate_k else 0)


def pack_weights(dims, w_q, w_v, w_o, w_k=None):
  lib = _lib.load()
  dev = w_q.device
  wqv = torch.empty((dims.D, dims.H, _row_width(dims)), dtype=torch.bfloat16, device=dev)
  wo = torch.empty((dims.H * dims.dv, dims.D), dtype=torch.bfloat16, device=dev)
  _lib.check(lib.lsh_pack_weights(ctypes.byref(dims), _ptr(w_q), _ptr(w_v), _ptr(w_o), _ptr(w_k), _ptr(wqv), _ptr(wo),
                                  _stream()), 'lsh_pack_weights')
  return wqv, wo


def project_qv(dims, x_bf16, wqv):
  lib = _lib.load()
  qv = torch.empty((dims.B, dims.L, dims.H, _row_width(dims)), dtype=torch.bfloat16, device=x_bf16.device)
  ws = workspace(x_bf16.device, 64 << 20)     # cuBLAS scratch + the K-major weight copy of the tensor-core GEMM
  _lib.check(lib.lsh_project_qv(ctypes.byref(dims), _ptr(x_bf16), _ptr(wqv), _ptr(qv), _ptr(ws), ws.numel(),
                                _stream()), 'lsh_project_qv')
  return qv


def hash_qv(dims, qv, rotations, mask=None, buckets=None):
  lib = _lib.load()
  if buckets is None:
    buckets = torch.zeros((dims.B * dims.H, dims.nh * dims.L), dtype=torch.int32, device=qv.device)
  _lib.check(lib.lsh_hash(ctypes.byref(dims), _ptr(qv), _ptr(rotations), _ptr(mask), _ptr(buckets),
                          buckets.stride(0), _stream()), 'lsh_hash')
  return buckets


def hash_f32(dims, vecs, rotations, mask=None):
  lib = _lib.load()
  buckets = torch.zeros((dims.B * dims.H, dims.nh * dims.L), dtype=torch.int32, device=vecs.device)
  _lib.check(lib.lsh_hash_f32(ctypes.byref(dims), _ptr(vecs), _ptr(rotations), _ptr(mask), _ptr(buckets),
                              buckets.stride(0), _stream()), 'lsh_hash_f32')
  return buckets


def sort(dims, buckets, want_undo=True):
  lib = _lib.load()
  n = dims.nh * dims.L
  sticker = torch.empty((dims.B * dims.H, n), dtype=torch.int32, device=buckets.device)
  undo = torch.empty_like(sticker) if want_undo else None
  nbytes = lib.lsh_sort_workspace_bytes(ctypes.byref(dims))
  ws = workspace(buckets.device, nbytes)
  _lib.check(lib.lsh_sort(ctypes.byref(dims), _ptr(buckets), buckets.stride(0), _ptr(sticker), _ptr(undo),
                          _ptr(ws), ws.numel(), _stream()), 'lsh_sort')
  return sticker, undo


def chunk_possort(dims, sticker, with_bounds=False):
  """sticker with every 128-slot chunk re-ordered by position (internal order of the tcgen05 kernels); with_bounds also
  returns the packed neighbour-chunk interval bounds of every row (see include/lsh_attn.h)."""
  lib = _lib.load()
  out = torch.empty_like(sticker)
  bounds = torch.empty_like(sticker) if with_bounds else None
  _lib.check(lib.lsh_chunk_possort(ctypes.byref(dims), _ptr(sticker), _ptr(out), _ptr(bounds), _stream()), 'lsh_chunk_possort')
  return (out, bounds) if with_bounds else out


def attend_fwd(dims, qv, sticker, mask=None, attn_keep=None):
  """attn_keep: optional (chunk_len, window) f32 dropout multiplier of EA:254-262 (values 0 or 1 / (1 - rate))."""
  lib = _lib.load()
  bh, n = dims.B * dims.H, dims.nh * dims.L
  o = torch.empty((bh, n, dims.dv), dtype=torch.bfloat16, device=qv.device)
  logits = torch.empty((bh, n), dtype=torch.float32, device=qv.device)
  ws = workspace(qv.device, lib.lsh_attend_fwd_workspace_bytes(ctypes.byref(dims)))
  _lib.check(lib.lsh_attend_fwd(ctypes.byref(dims), _ptr(qv), _ptr(sticker), _ptr(mask), _ptr(attn_keep), _ptr(o),
                                _ptr(logits), _ptr(ws), ws.numel(), _stream()), 'lsh_attend_fwd')
  return o, logits


def combine_fwd(dims, o_rounds, logits):
  lib = _lib.load()
  o = torch.empty((dims.B, dims.L, dims.H, dims.dv), dtype=torch.bfloat16, device=o_rounds.device)
  lse = torch.empty((dims.B * dims.H, dims.L), dtype=torch.float32, device=o_rounds.device)
  _lib.check(lib.lsh_combine_fwd(ctypes.byref(dims), _ptr(o_rounds), _ptr(logits), _ptr(o), _ptr(lse), _stream()),
             'lsh_combine_fwd')
  return o, lse


def attend_bwd(dims, qv, sticker, o_comb, lse_tot, do_comb, mask=None, attn_keep=None):
  lib = _lib.load()
  dqv = torch.empty_like(qv)
  nbytes = lib.lsh_attend_bwd_workspace_bytes(ctypes.byref(dims))
  ws = workspace(qv.device, nbytes)
  _lib.check(lib.lsh_attend_bwd(ctypes.byref(dims), _ptr(qv), _ptr(sticker), _ptr(mask), _ptr(attn_keep), _ptr(o_comb),
                                _ptr(lse_tot), _ptr(do_comb), _ptr(dqv), _ptr(ws), ws.numel(), _stream()),
             'lsh_attend_bwd')
  return dqv


def make_rotations(dims, keys):
  """keys: (B*H, 2) int32/uint32 bit patterns.  Returns (rotations (BH, dq, nh, R) f32, new_keys)."""
  lib = _lib.load()
  r = sum(int(dims.factors[i]) for i in range(dims.n_factors)) // 2
  rot = torch.empty((dims.B * dims.H, dims.dq, dims.nh, r), dtype=torch.float32, device=keys.device)
  new_keys = torch.empty_like(keys)
  _lib.check(lib.lsh_make_rotations(ctypes.byref(dims), _ptr(keys), _ptr(new_keys), _ptr(rot), _stream()),
             'lsh_make_rotations')
  return rot, new_keys


def launch_count(reset=False):
  return int(_lib.load().lsh_attn_launch_count(1 if reset else 0))


def pack_heads(a, b, n_heads):
  """(B*H, L, da) [, (B*H, L, db)] f32 / bf16 -> (B, L, H, da + db) bf16: the row layout of the core (one kernel, no torch glue)."""
  lib = _lib.load()
  bh, L, da = (int(v) for v in a.shape)
  db = int(b.shape[2]) if b is not None else 0
  if b is not None and (b.dtype != a.dtype or tuple(b.shape[:2]) != (bh, L)):
    raise ValueError('pack_heads: operands must agree in dtype and leading shape')
  B = bh // n_heads
  dst = torch.empty((B, L, n_heads, da + db), dtype=torch.bfloat16, device=a.device)
  _lib.check(lib.lsh_pack_heads(B, n_heads, L, _act_dtype(a), _ptr(a), da, _ptr(b), db, _ptr(dst), _stream()), 'lsh_pack_heads')
  return dst


def unpack_heads(src, col0, d, dtype):
  """(B, L, H, d_total) bf16, columns [col0, col0 + d) -> (B*H, L, d) in `dtype` (f32 or bf16)."""
  lib = _lib.load()
  B, L, H, d_total = (int(v) for v in src.shape)
  if src.dtype != torch.bfloat16:
    raise ValueError('unpack_heads: source must be bf16')
  dst = torch.empty((B * H, L, d), dtype=dtype, device=src.device)
  _lib.check(lib.lsh_unpack_heads(B, H, L, _act_dtype(dst), _ptr(src), d_total, col0, d, _ptr(dst), _stream()), 'lsh_unpack_heads')
  return dst
