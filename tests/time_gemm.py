"""Timing helper (not a test): the layer's seven GEMM shapes on the hand-written kernel vs torch.matmul (cuBLAS)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trax_b200 import _lib, ops
lib = _lib.load()
f1 = lib.lsh_debug_gemm_tc; f1.restype = ctypes.c_int
f1.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
f2 = lib.lsh_debug_gemm_tc_wgrad; f2.restype = ctypes.c_int
f2.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_void_p]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
scratch = torch.empty(32 << 20, dtype=torch.uint8, device='cuda')

def timed(fn, n=5):
  tot = 0.0
  for i in range(n + 1):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    if i: tot += e0.elapsed_time(e1)
  return tot / n

def aw(name, M, N, K):
  a = torch.randn((M, K), device='cuda').bfloat16(); b = torch.randn((N, K), device='cuda').bfloat16()
  c = torch.empty((M, N), device='cuda', dtype=torch.bfloat16)
  t_own = timed(lambda: f1(M, N, K, a.data_ptr(), K, b.data_ptr(), K, c.data_ptr(), N, 0, ops._stream()))
  bt = b.t().contiguous()
  t_lib = timed(lambda: torch.matmul(a, bt, out=c))
  fl = 2.0 * M * N * K
  print('%-28s M=%8d N=%5d K=%6d  own %7.3f ms (%6.0f TF/s)  cuBLAS %7.3f ms (%6.0f TF/s)  ratio %.2f' % (name, M, N, K, t_own, fl / t_own / 1e9, t_lib, fl / t_lib / 1e9, t_own / t_lib))

def wg(name, M, N, K):
  a = torch.randn((K, M), device='cuda').bfloat16(); b = torch.randn((K, N), device='cuda').bfloat16()
  c = torch.empty((M, N), device='cuda', dtype=torch.float32)
  t_own = timed(lambda: f2(M, N, K, a.data_ptr(), b.data_ptr(), c.data_ptr(), scratch.data_ptr(), scratch.numel(), ops._stream()))
  af, bf = a, b
  t_lib = timed(lambda: torch.matmul(af.t(), bf))
  fl = 2.0 * M * N * K
  print('%-28s M=%8d N=%5d K=%6d  own %7.3f ms (%6.0f TF/s)  cuBLAS %7.3f ms (%6.0f TF/s)  ratio %.2f' % (name, M, N, K, t_own, fl / t_own / 1e9, t_lib, fl / t_lib / 1e9, t_own / t_lib))

for tag, BL, D, H in (('c2', 65536, 1024, 8), ('c3', 12288, 1024, 8), ('c5-share', 1 << 20, 1024, 2)):
  NQV, KO = H * 128, H * 64
  aw(tag + ' q|v = x wqv', BL, NQV, D)
  aw(tag + ' out = o wo', BL, D, KO)
  aw(tag + ' do = dout wo^T', BL, KO, D)
  aw(tag + ' dx = dqv wqv^T', BL, D, NQV)
  wg(tag + ' dWo = o^T dout', KO, D, BL)
  wg(tag + ' dWqv = x^T dqv', D, NQV, BL)
