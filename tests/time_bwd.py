import sys, os
sys.path.insert(0, '/root/repo')
import torch
from trax_b200 import ops, _lib
L = 65536; B, H, D, C, nh = 1, 8, 1024, 128, 4
dims = _lib.make_dims(B, H, L, D, 64, 64, C, 1, 0, nh, ops.bucket_factors(None, L, C), True, False, 1)
g = torch.Generator('cuda').manual_seed(0)
qv = torch.randn((B, L, H, 128), device='cuda', generator=g).bfloat16()
keys = torch.arange(2 * B * H, dtype=torch.int32, device='cuda').reshape(B * H, 2)
rot, _ = ops.make_rotations(dims, keys)
sticker, _ = ops.sort(dims, ops.hash_qv(dims, qv, rot))
o_r, logits = ops.attend_fwd(dims, qv, sticker)
o_c, lse = ops.combine_fwd(dims, o_r, logits)
do = torch.randn_like(o_c)
for _ in range(3): ops.attend_bwd(dims, qv, sticker, o_c, lse, do)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attend_bwd(dims, qv, sticker, o_c, lse, do)
e1.record(); torch.cuda.synchronize()
print('TIME lib=%s attend_bwd(stage) %.3f ms' % (os.path.basename(_lib.LIB_PATH), e0.elapsed_time(e1) / 10))
