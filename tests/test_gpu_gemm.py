"""The hand-written tcgen05 + TMA GEMM of the projections (csrc/gemm_tc.cu; EA:1923-1924, 1995 and their input-gradient
VJPs) against a plain fp32 torch matmul of the same bf16 operands."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(a, b, out_dtype):
  from trax_b200 import _lib, ops
  lib = _lib.load()
  fn = lib.lsh_debug_gemm_tc
  fn.restype = ctypes.c_int
  fn.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
  M, K = a.shape
  N = b.shape[0]
  c = torch.full((M, N), float('nan'), dtype=out_dtype, device=a.device)
  rc = fn(M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), c.data_ptr(), c.stride(0),
          1 if out_dtype == torch.float32 else 0, ops._stream())
  return rc, c


@pytest.mark.parametrize('M,N,K,out_dtype', [
    (256, 256, 64, torch.float32),            # one cluster, one K block
    (128, 128, 128, torch.bfloat16),          # a single tile: no cluster partner
    (1024, 256, 256, torch.bfloat16),         # config 1: x (1024, 256) · wqv
    (4096, 1024, 1024, torch.bfloat16),       # q|v projection shape at d_model 1024, 8 heads
    (4096, 1024, 512, torch.float32),         # output projection (K = H * d_v), f32 activations
    (4096, 512, 1024, torch.bfloat16),        # do = dout · w_o^T
    (1000, 384, 192, torch.bfloat16),         # ragged M (TMA zero fill, predicated stores), N = 3 * 128 (q|v|k rows, 2 heads)
    (384, 128, 64, torch.float32),            # odd number of row blocks: the last cluster has an idle partner
    (33000, 256, 128, torch.bfloat16),        # more tiles than SMs: persistent loop, both accumulators, ring wrap-around
])
def test_gemm_tc_matches_fp32_matmul(M, N, K, out_dtype):
  g = torch.Generator('cuda').manual_seed(M + N + K)
  a = torch.randn((M, K), device='cuda', generator=g).bfloat16()
  b = torch.randn((N, K), device='cuda', generator=g).bfloat16()
  rc, c = _gemm(a, b, out_dtype)
  assert rc == 0
  torch.cuda.synchronize()
  want = a.float() @ b.float().t()
  assert bool(torch.isfinite(c.float()).all())
  if out_dtype == torch.float32:
    torch.testing.assert_close(c, want, rtol=1e-4, atol=1e-3)     # same products, fp32 accumulation in another order
  else:
    torch.testing.assert_close(c.float(), want.bfloat16().float(), rtol=1.6e-2, atol=1e-2)   # one bf16 rounding apart


@pytest.mark.parametrize('M,N,K', [
    (128, 128, 64),              # one tile, one K block, no split
    (256, 256, 4096),            # one cluster of tiles, K split over the machine
    (512, 1024, 8192),           # dW_o shape (H d_v, d_model)
    (1024, 1024, 4096),          # dW_q|v shape
    (384, 384, 1024),            # odd row-block count (no cluster partner), 3 column blocks of 128
])
def test_gemm_tc_wgrad_matches_fp32_matmul(M, N, K):
  """dW = act^T · cotangent (EA:2431 / the VJP of EA:1923-1924, 1995): MN-major operands, K split over the clusters, fp32
  partial sums added in a fixed order."""
  from trax_b200 import _lib, ops
  lib = _lib.load()
  fn = lib.lsh_debug_gemm_tc_wgrad
  fn.restype = ctypes.c_int
  fn.argtypes = [ctypes.c_int64] * 3 + [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_void_p]
  g = torch.Generator('cuda').manual_seed(M + N + K)
  a = torch.randn((K, M), device='cuda', generator=g).bfloat16()
  b = torch.randn((K, N), device='cuda', generator=g).bfloat16()
  c = torch.full((M, N), float('nan'), dtype=torch.float32, device='cuda')
  scratch = torch.empty(32 << 20, dtype=torch.uint8, device='cuda')
  outs = []
  for _ in range(2):
    assert fn(M, N, K, a.data_ptr(), b.data_ptr(), c.data_ptr(), scratch.data_ptr(), scratch.numel(), ops._stream()) == 0
    torch.cuda.synchronize()
    outs.append(c.clone())
  assert torch.equal(outs[0], outs[1])                               # deterministic (no atomics)
  want = a.float().t() @ b.float()
  torch.testing.assert_close(c, want, rtol=2e-4, atol=2e-2 * (K / 4096) ** 0.5)


def test_gemm_tc_declines_shapes_it_does_not_cover():
  a = torch.zeros((128, 72), device='cuda', dtype=torch.bfloat16)
  b = torch.zeros((128, 72), device='cuda', dtype=torch.bfloat16)
  assert _gemm(a, b, torch.float32)[0] == -1                       # K % 64
  a = torch.zeros((128, 64), device='cuda', dtype=torch.bfloat16)
  b = torch.zeros((96, 64), device='cuda', dtype=torch.bfloat16)
  assert _gemm(a, b, torch.float32)[0] == -1                       # N % 128


def test_layer_runs_on_the_hand_written_gemm():
  """The layer's projections take the tcgen05 GEMM (no library kernel between the launches the library counts)."""
  import trax_b200
  from trax_b200 import ops
  layer = trax_b200.LSHSelfAttention(n_heads=2, causal=True, chunk_len=64, n_hashes=2)
  layer.init(trax_b200.ShapeDtype((1, 512, 128)))
  x = torch.randn((1, 512, 128), device='cuda')
  ops.launch_count(reset=True)
  layer.forward(x)
  n_fwd = ops.launch_count(reset=True)
  import os
  assert os.environ.get('LSH_GEMM') == 'cublas' or n_fwd >= 12      # the two projection GEMMs are counted as own launches
