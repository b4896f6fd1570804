"""Not a test: prints per-stage error statistics of the CUDA path vs the oracle (run on the GPU box)."""
import sys, os, json, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lsh_oracle as O
from tests import util
from tests.test_gpu_stages import _dims, _cuda, _core_case, _oracle_core, CORE_CASES
from trax_b200 import ops

def rep(name, got, want):
  r = util.close_report(got, want)
  print('  %-14s rel_l2=%.2e max_abs=%.3e max_ref=%.3e frac_bad=%.2e worst=%.2f' % (name, r['rel_l2'], r['max_abs'], r['max_ref'], r['frac_bad'], r['worst_ratio']))

for case in CORE_CASES:
  B, H, L, C, nb, na, nh, nbk, causal, masked = case
  print('case', case)
  try:
    cfg, qv, buckets, mask = _core_case(11, B, H, L, C, nb, na, nh, nbk, causal, masked)
    dims = _dims(B, H, L, 128, C, nb, na, nh, [nbk], causal, masked)
    mask_d = None if mask is None else _cuda(mask.astype(np.uint8))
    qv_d = _cuda(qv, torch.bfloat16)
    sticker, undo = ops.sort(dims, _cuda(buckets))
    rng = np.random.default_rng(5)
    do = util.bf16_round(rng.standard_normal((B, L, H, 64)))
    if mask is not None: do = do * mask[:, :, None, None]
    res, grads = _oracle_core(cfg, qv, buckets, mask, B, H, dout=do)
    print('  sticker equal:', all((sticker[u].cpu().numpy() == res[u].sticker).all() for u in range(B*H)))
    o_r, logits = ops.attend_fwd(dims, qv_d, sticker, mask_d)
    o_c, lse_tot = ops.combine_fwd(dims, o_r, logits)
    torch.cuda.synchronize()
    rep('o_rounds', o_r.float().cpu().numpy(), np.stack([r.o_rounds for r in res]))
    rep('logits', logits.cpu().numpy(), np.stack([r.logits for r in res]))
    rep('o_comb', o_c.float().cpu().numpy().transpose(0, 2, 1, 3).reshape(B*H, L, 64), np.stack([r.o for r in res]))
    if C != 32:
      dqv = ops.attend_bwd(dims, qv_d, sticker, o_c, lse_tot, _cuda(do, torch.bfloat16), mask_d)
      torch.cuda.synchronize()
      dqv = dqv.float().cpu().numpy().transpose(0, 2, 1, 3).reshape(B*H, L, 128)
      g = np.stack(grads)
      rep('dq', dqv[..., :64], g[..., :64]); rep('dv', dqv[..., 64:], g[..., 64:])
  except Exception:
    traceback.print_exc()
