"""Profiling helper (not a test): runs one stage of the C2 workload a few times so ncu can capture it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trax_b200 import ops, _lib
stage = sys.argv[1] if len(sys.argv) > 1 else 'attend_fwd'
L = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
B, H, D, C, nh = 1, 8, 1024, 128, 4
factors = ops.bucket_factors(None, L, C)
dims = _lib.make_dims(B, H, L, D, 64, 64, C, 1, 0, nh, factors, True, False, 1)
g = torch.Generator('cuda').manual_seed(0)
qv = torch.randn((B, L, H, 128), device='cuda', generator=g).bfloat16()
keys = torch.arange(2 * B * H, dtype=torch.int32, device='cuda').reshape(B * H, 2)
rot, _ = ops.make_rotations(dims, keys)
buckets = ops.hash_qv(dims, qv, rot)
sticker, _ = ops.sort(dims, buckets)
o_r, logits = ops.attend_fwd(dims, qv, sticker)
o_c, lse = ops.combine_fwd(dims, o_r, logits)
do = torch.randn_like(o_c)
torch.cuda.synchronize()
for _ in range(3):
  if stage == 'attend_fwd': ops.attend_fwd(dims, qv, sticker)
  elif stage == 'attend_bwd': ops.attend_bwd(dims, qv, sticker, o_c, lse, do)
  elif stage == 'hash': ops.hash_qv(dims, qv, rot, buckets=buckets)
  elif stage == 'sort': ops.sort(dims, buckets)
  elif stage == 'combine': ops.combine_fwd(dims, o_r, logits)
torch.cuda.synchronize()
print('done', stage)
