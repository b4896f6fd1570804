"""Timing helper (not a test): attend_fwd stage of the C2 workload (aux + kernel), CUDA events, 10 launches."""
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
for _ in range(3): ops.attend_fwd(dims, qv, sticker)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attend_fwd(dims, qv, sticker)
e1.record(); torch.cuda.synchronize()
print('TIME lib=%s attend_fwd(stage) %.3f ms' % (os.path.basename(_lib.LIB_PATH), e0.elapsed_time(e1) / 10))
