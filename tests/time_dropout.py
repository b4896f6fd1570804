"""Timing helper (not a test): attend_fwd / attend_bwd stages of the C2 workload with and without attention dropout."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trax_b200 import ops, _lib
L = 65536; B, H, D, C, nh = 1, 8, 1024, 128, 4
dims = _lib.make_dims(B, H, L, D, 64, 64, C, 1, 0, nh, ops.bucket_factors(None, L, C), True, False, 1)
g = torch.Generator('cuda').manual_seed(0)
qv = torch.randn((B, L, H, 128), device='cuda', generator=g).bfloat16()
keys = torch.arange(2 * B * H, dtype=torch.int32, device='cuda').reshape(B * H, 2)
rot, _ = ops.make_rotations(dims, keys)
sticker, _ = ops.sort(dims, ops.hash_qv(dims, qv, rot))
keep = (torch.rand((C, 2 * C), device='cuda', generator=g) > 0.2).float() / 0.8
def timed(fn, n=5):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n
for name, kp in (('no dropout', None), ('dropout 0.2', keep)):
  o_r, logits = ops.attend_fwd(dims, qv, sticker, attn_keep=kp)
  o_c, lse = ops.combine_fwd(dims, o_r, logits)
  do = torch.randn_like(o_c)
  tf = timed(lambda: ops.attend_fwd(dims, qv, sticker, attn_keep=kp))
  tb = timed(lambda: ops.attend_bwd(dims, qv, sticker, o_c, lse, do, attn_keep=kp))
  print('%-12s attend_fwd stage %.3f ms   attend_bwd stage %.3f ms' % (name, tf, tb))
