import sys, os, ctypes
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
o_r, logits = ops.attend_fwd(dims, qv, sticker)
o_c, lse = ops.combine_fwd(dims, o_r, logits)
do = torch.randn_like(o_c)
ops.attend_bwd(dims, qv, sticker, o_c, lse, do); torch.cuda.synchronize()
tr = torch.zeros(120 * 32, dtype=torch.int64, device='cuda')
lib = ctypes.CDLL(_lib.LIB_PATH); lib.lsh_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
ops.attend_bwd(dims, qv, sticker, o_c, lse, do); torch.cuda.synchronize()
lib.lsh_debug_set_trace(None)
t = tr.cpu().view(120, 32)
if not bool((t > 0).any()):
  print('(library built without -DLSH_TRACE)'); sys.exit(0)
t0 = int(t[t > 0].min())
names = ['pds0', 'pds1', 'done'] + ['p%d' % w for w in range(8)] + ['s%d' % w for w in range(8)] + ['kv0', 'st0', 'kv1', 'st1', 'dq', 'epi', 'kvful', 'prod', 'tiles', 'Wblk', 'Wstw', 'Wfnc', 'Wmid']
print('n   ' + ' '.join('%7s' % n for n in names))
for k in list(range(0, 6)) + list(range(96, 112)):
  print('%3d ' % k + ' '.join('%7d' % (int(v) - t0 if v > 0 else -1) for v in t[k][:len(names)]))
