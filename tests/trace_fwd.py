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
ops.attend_fwd(dims, qv, sticker); torch.cuda.synchronize()
tr = torch.zeros(120 * 16 + 148 + 120 * 12, dtype=torch.int64, device='cuda')
lib = ctypes.CDLL(_lib.LIB_PATH); lib.lsh_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
ops.attend_fwd(dims, qv, sticker); torch.cuda.synchronize()
lib.lsh_debug_set_trace(None)
percta = tr.cpu()[120 * 16:120 * 16 + 148]; t2 = tr.cpu()[120 * 16 + 148:].view(120, 4, 3); t = tr.cpu()[:120 * 16].view(120, 16)
if not bool((t > 0).any()):
  print('(library built without -DLSH_TRACE: timing only)')
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10): ops.attend_fwd(dims, qv, sticker)
  e1.record(); torch.cuda.synchronize()
  print('TIME lib=%s attend_fwd %.3f ms' % (os.path.basename(_lib.LIB_PATH), e0.elapsed_time(e1) / 10))
  sys.exit(0)
t0 = int(t[t > 0].min())
names = ['S_iss', 'PV_iss', 's_full', 'passdone', 'o_full', 'epi_done', 'arr_swait', 'S_start', 'PV0woke', 'PV0iss', 'PV1woke', 'PV1iss', 'S1start', 'w4sfull', 'w4end', '-']
print('k ' + ' '.join('%9s' % n for n in names))
for k in list(range(0, 24)) + list(range(100, 112)):
  print('%3d ' % k + ' '.join('%9d' % (int(v) - t0 if v > 0 else -1) for v in t[k]))

import numpy as np
a = t[8:108].numpy().astype(np.int64)
print('SUMMARY lib=%s  chunk period %.0f  S->s_full %.0f  pass (s_full->passdone) %.0f  passdone->PV_iss %.0f  PV->o_full %.0f  o_full->epi %.0f  S_start->S_iss %.0f' % (
    os.path.basename(_lib.LIB_PATH), (a[-1, 0] - a[0, 0]) / 99.0, (a[:, 2] - a[:, 0]).mean(), (a[:, 3] - a[:, 2]).mean(), (a[:, 1] - a[:, 3]).mean(),
    (a[:, 4] - a[:, 1]).mean(), (a[:, 5] - a[:, 4]).mean(), (a[:, 0] - a[:, 7]).mean()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attend_fwd(dims, qv, sticker)
e1.record(); torch.cuda.synchronize()
print('TIME lib=%s attend_fwd %.3f ms' % (os.path.basename(_lib.LIB_PATH), e0.elapsed_time(e1) / 10))

pc = percta.numpy()
print('PER-CTA cycles: min %d  median %d  max %d  (first 16: %s)' % (pc.min(), np.median(pc), pc.max(), pc[:16].tolist()))
print('slowest CTAs:', np.argsort(-pc)[:12].tolist(), np.sort(pc)[-12:][::-1].tolist())

print('per-warp pass: k | (start-t0, dur, needed blocks, full blocks) x 4 warps')
for k in range(16, 26):
  print('%3d ' % k + '  '.join('%8d %5d (ldwait %5d) %d/%d' % (int(t2[k, w, 0]) - t0, int(t2[k, w, 1] - t2[k, w, 0]), int(t2[k, w, 2]) // 256, (int(t2[k, w, 2]) % 256) // 16, int(t2[k, w, 2]) % 16) for w in range(4)))
d = (t2[8:108, :, 1] - t2[8:108, :, 0]).numpy(); nb = ((t2[8:108, :, 2] % 256) // 16).numpy()
print('mean ld-wait cycles per warp pass:', (t2[8:108, :, 2] // 256).numpy().mean(0))
print('mean pass cycles per warp:', d.mean(0), ' mean needed blocks per warp:', nb.mean(0), ' cycles per needed block: %.0f' % (d.sum() / nb.sum()))
