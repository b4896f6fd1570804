"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/lsh_attn.h declares; the ctypes struct mirrors the C struct; host-side argument validation works.
No compute calls here (no GPU in this container)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'lsh_attn.h')


def _declared_functions():
  src = open(HEADER).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(lsh_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
  from trax_b200 import _lib
  lib = _lib.load()
  declared = _declared_functions()
  assert len(declared) >= 18
  for name in declared:
    assert hasattr(lib, name), 'symbol %s declared in include/lsh_attn.h is not exported' % name
  assert sorted(_lib.SIGNATURES) == declared, 'ctypes SIGNATURES out of sync with the header'
  out = subprocess.check_output(['nm', '-D', '--defined-only', _lib.LIB_PATH], text=True)
  exported = set(re.findall(r' T (lsh_[a-z0-9_]+)', out))
  assert set(declared) <= exported


def test_header_compiles_as_plain_c_and_struct_matches():
  from trax_b200 import _lib
  prog = r'''
#include "lsh_attn.h"
#include <stdio.h>
#include <stddef.h>
int main(void) { printf("%zu %zu %zu %zu\n", sizeof(LshAttnDims), offsetof(LshAttnDims, factors),
                        offsetof(LshAttnDims, act_dtype), offsetof(LshAttnDims, C)); return 0; }
'''
  d = os.path.join(ROOT, 'tests', '_tmp')
  os.makedirs(d, exist_ok=True)
  c = os.path.join(d, 'abi_probe.c')
  open(c, 'w').write(prog)
  exe = os.path.join(d, 'abi_probe')
  subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
  size, off_f, off_a, off_c = (int(v) for v in subprocess.check_output([exe], text=True).split())
  assert size == ctypes.sizeof(_lib.LshAttnDims)
  assert off_f == _lib.LshAttnDims.factors.offset
  assert off_a == _lib.LshAttnDims.act_dtype.offset
  assert off_c == _lib.LshAttnDims.C.offset


def test_check_dims_messages():
  from trax_b200 import _lib
  lib = _lib.load()
  ok = _lib.make_dims(1, 8, 65536, 1024, 64, 64, 128, 1, 0, 4, [32, 32], True, False, _lib.LSH_DTYPE_BF16)
  assert lib.lsh_attn_check_dims(ctypes.byref(ok)) == 0
  assert lib.lsh_layer_workspace_bytes(ctypes.byref(ok), 1) > lib.lsh_layer_workspace_bytes(ctypes.byref(ok), 0) > 0
  bad = _lib.make_dims(2, 6, 10, 13, 7, 17, 5, 1, 0, 2, [4], True, False, _lib.LSH_DTYPE_F32)
  assert lib.lsh_attn_check_dims(ctypes.byref(bad)) != 0
  assert b'd_qk=7' in lib.lsh_attn_last_error()
  odd = _lib.make_dims(1, 2, 1024, 256, 64, 64, 64, 1, 0, 1, [31], True, False, 0)
  assert lib.lsh_attn_check_dims(ctypes.byref(odd)) != 0 and b'even' in lib.lsh_attn_last_error()
  wrap = _lib.make_dims(1, 16, 1 << 20, 1024, 64, 64, 128, 1, 0, 1, [128, 128], True, False, 1)
  assert lib.lsh_attn_check_dims(ctypes.byref(wrap)) != 0 and b'wrap' in lib.lsh_attn_last_error()
  with pytest.raises(_lib.LshAttnError):
    _lib.check(1, 'x')


def test_layer_constructor_matches_reference_signature():
  """EA:1732-1748 keyword names and defaults (hard-coded: /root/reference is absent on the GPU box)."""
  import inspect
  import trax_b200
  want = dict(n_heads=2, d_qk=64, d_v=64, share_qk='unused', causal=False, masked=False, chunk_len=128,
              n_chunks_before=1, n_chunks_after=0, n_hashes=1, n_buckets=None, mode='train', predict_mem_len=2048,
              predict_drop_len=256, attention_dropout=0.0, output_dropout=0.0, max_length_for_buckets=None, bias=False,
              n_parallel_heads=1, use_python_loop=False, use_reference_code=False)
  sig = inspect.signature(trax_b200.LSHSelfAttention.__init__)
  got = {k: v.default for k, v in sig.parameters.items() if k != 'self'}
  assert got == want and list(got) == list(want)
  fab = inspect.signature(trax_b200.LSHSelfAttention.forward_and_or_backward)
  assert list(fab.parameters) == ['self', 'inputs', 'weights', 'state', 'rng', 'output_grad', 'compute_output', 'update_state']
  bwd = inspect.signature(trax_b200.LSHSelfAttention.backward)
  assert list(bwd.parameters)[:8] == ['self', 'inputs', 'output', 'grad', 'weights', 'state', 'new_state', 'rng']


def test_layer_init_layout_and_no_cpu_fallback():
  import numpy as np
  import torch
  import trax_b200
  from trax_b200 import _lib
  layer = trax_b200.LSHSelfAttention(n_heads=8, d_qk=64, d_v=64, causal=True, chunk_len=128, n_hashes=4,
                                     max_length_for_buckets=4096)
  w, s = layer.init(trax_b200.ShapeDtype((2, 2048, 1024)), rng=np.array([3, 4], np.uint32))
  assert [tuple(t.shape) for t in w] == [(8, 1024, 64), (8, 1024, 64), (8, 64, 1024)]     # EA:1845-1868 + :414-417 of the test
  assert all(t.dtype == torch.float32 for t in w)
  lim = np.sqrt(6.0 / (1024 + 64 * 8))                                                     # EA:1807
  assert float(w[0].abs().max()) <= lim and float(w[0].abs().max()) > 0.9 * lim
  assert tuple(s[0].shape) == (16, 4 * 4096) and s[0].dtype == torch.int32                # EA:1880-1881
  assert tuple(s[1].shape) == (16, 2)
  assert layer.n_in == 1 and layer.n_out == 1 and layer.has_backward
  assert trax_b200.LSHSelfAttention(masked=True).n_in == 2
  with pytest.raises(ValueError):
    layer.foo = 1                                                                          # base.py:679-706
  if not torch.cuda.is_available():
    with pytest.raises(_lib.LshAttnError):
      layer.forward(torch.zeros(2, 2048, 1024))


def test_jax_ffi_shim_translation_unit_compiles_and_is_guarded(tmp_path):
  """csrc/jax_ffi_shim.cc (the `jax.ffi` handlers over the C ABI) is real code behind
  `__has_include("xla/ffi/api/ffi.h")`: without jaxlib's headers it must still compile and say so; its handler bodies call
  the entry points with the argument counts the header declares (checked textually against SIGNATURES)."""
  import re
  import subprocess
  src = os.path.join(ROOT, 'trax_b200', 'csrc', 'jax_ffi_shim.cc')
  obj = str(tmp_path / 'shim.so')
  subprocess.check_call(['g++', '-std=c++17', '-shared', '-fPIC', src, '-o', obj])
  lib = ctypes.CDLL(obj)
  assert lib.lsh_attn_jax_ffi_available() == 0          # no XLA headers in this image
  text = open(src).read()
  assert '__has_include("xla/ffi/api/ffi.h")' in text and 'XLA_FFI_DEFINE_HANDLER_SYMBOL' in text
  from trax_b200 import _lib
  for fn in ('lsh_layer_fwd', 'lsh_layer_bwd', 'lsh_predict_step'):
    m = re.search(fn + r'\((.*?)\)\);', text, re.S)
    assert m, fn
    depth, n_args = 0, 1
    for ch in m.group(1):
      depth += ch in '([{'
      depth -= ch in ')]}'
      n_args += (ch == ',' and depth == 0)
    assert n_args == len(_lib.SIGNATURES[fn][1]), (fn, n_args, len(_lib.SIGNATURES[fn][1]))
  with pytest.raises(ImportError):
    import trax_b200.jax_binding  # noqa: F401  (no jax in this image)


def test_predict_entry_points_validate_their_arguments_without_a_gpu():
  """`lsh_predict_step` / `lsh_predict_attend` (fast inference, ABI v5): workspace queries answer and argument errors come
  back as status + message before any CUDA call; the predict-mode state layout of the layers (EA:1833-1841, 1883-1887,
  2677-2686) is the reference's."""
  import torch
  import trax_b200
  from trax_b200 import _lib
  lib = _lib.load()
  assert lib.lsh_attn_abi_version() == 5
  d = _lib.make_dims(1, 8, 2048, 1024, 64, 64, 128, 1, 0, 4, [64], True, False, _lib.LSH_DTYPE_BF16)
  assert lib.lsh_predict_workspace_bytes(ctypes.byref(d)) > 2048 * 8 * 128 * 2          # at least the projected memory
  assert lib.lsh_predict_attend_workspace_bytes(ctypes.byref(d)) >= 8 * 4 * 2048 * 4    # the step's hash of every slot
  assert lib.lsh_predict_step(ctypes.byref(d), None, None, None, None, None, None, None, 0, 0, None, None, 0, None) != 0
  assert b'NULL' in lib.lsh_attn_last_error()
  assert lib.lsh_predict_attend(ctypes.byref(d), None, None, None, 0, 0, None, None, 0, None) != 0
  assert b'NULL' in lib.lsh_attn_last_error()
  bad = _lib.make_dims(1, 8, 2048, 1024, 7, 17, 128, 1, 0, 4, [64], True, False, _lib.LSH_DTYPE_BF16)
  assert lib.lsh_predict_workspace_bytes(ctypes.byref(bad)) == 0 and b'd_qk=7' in lib.lsh_attn_last_error()
  layer = trax_b200.LSHSelfAttention(n_heads=8, causal=True, n_hashes=4, n_buckets=64, mode='predict')
  layer.init_weights_and_state(trax_b200.ShapeDtype((2, 1, 1024), torch.bfloat16), device='cpu')
  mem_end, (mem,), (buckets, buckets_idx, rng) = layer.state
  assert int(mem_end) == 0 and tuple(mem.shape) == (2, 2048, 1024) and mem.dtype == torch.bfloat16
  assert tuple(buckets.shape) == (16, 4 * 2048) and buckets.dtype == torch.int32 and tuple(buckets_idx.shape) == (16,)
  assert tuple(rng.shape) == (16, 2)
  core = trax_b200.PureLSHSelfAttention(n_heads=8, causal=True, n_hashes=4, n_buckets=64, mode='predict')
  core.init_weights_and_state((trax_b200.ShapeDtype((16, 1, 64)),) * 2, device='cpu')
  assert [tuple(t.shape) for t in core.state[1]] == [(16, 2048, 64)] * 2 and core.weights == ()
  sa = trax_b200.SelfAttention(n_heads=8, causal=True, chunk_len=128, n_chunks_before=1, mode='predict', predict_mem_len=2048,
                               predict_drop_len=256)
  sa.init_weights_and_state(trax_b200.ShapeDtype((2, 1, 1024)), device='cpu')
  assert len(sa.weights) == 4 and sa.state[2] == () and tuple(sa.state[1][0].shape) == (2, 2048, 1024)
  if not torch.cuda.is_available():
    with pytest.raises(_lib.LshAttnError):
      layer.forward(torch.zeros(2, 1, 1024))
