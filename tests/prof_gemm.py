"""Profiling helper (not a test): one launch each of the layer's q|v projection (K = 1024) and output projection (K = 512) GEMM
shapes at config 2 on the own kernel, after a warm-up pair — for `ncu -k regex:gemm_tc -s 2 -c 2`."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trax_b200 import _lib, ops
lib = _lib.load()
f1 = lib.lsh_debug_gemm_tc; f1.restype = ctypes.c_int
f1.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
def run(M, N, K):
  a = torch.randn((M, K), device='cuda').bfloat16(); b = torch.randn((N, K), device='cuda').bfloat16()
  c = torch.empty((M, N), device='cuda', dtype=torch.bfloat16)
  return lambda: f1(M, N, K, a.data_ptr(), K, b.data_ptr(), K, c.data_ptr(), N, 0, ops._stream())
g1, g2 = run(65536, 1024, 1024), run(65536, 1024, 512)
for _ in range(2):
  g1(); g2()
torch.cuda.synchronize()
print('done')
