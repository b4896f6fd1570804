"""Timing helper (not a test): one fast-inference step of LSHSelfAttention / SelfAttention at the enwik8 decode shape
(memory 2048, d_model 1024, 8 heads, chunk 128, 4 hashes), CUDA events over N steps after warm-up.

    gpurun -- python tests/time_predict.py [steps]

Prints ms per `lsh_predict_step` call (C-ABI level, memory already updated) and per `layer.forward` call (with the memory /
bucket-memory bookkeeping).  NOT yet run on a GPU: round 2's GPU minutes ended before it could be."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import trax_b200
from trax_b200 import predict


def main(steps=200):
  B, H, D, M = 1, 8, 1024, 2048
  for name, cls, kw in (('lsh', trax_b200.LSHSelfAttention, dict(n_hashes=4, n_buckets=64)),
                        ('self', trax_b200.SelfAttention, dict(share_qk=False))):
    layer = cls(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=128, n_chunks_before=1, mode='predict',
                predict_mem_len=M, predict_drop_len=256, **kw)
    weights, state = layer.init(trax_b200.ShapeDtype((B, 1, D), torch.bfloat16))
    x = torch.randn((B, 1, D), device='cuda').to(torch.bfloat16)
    prefix = torch.randn((B, 1024, D), device='cuda').to(torch.bfloat16)
    layer.forward(prefix)
    for _ in range(10):
      layer.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      layer.forward(x)
    e1.record()
    torch.cuda.synchronize()
    api_ms = e0.elapsed_time(e1) / steps
    mem_end, (mem,), inner = layer.state
    rot = predict._step_rotations(layer, B, mem.device, inner[2]) if name == 'lsh' else None
    buckets = inner[0].clone() if name == 'lsh' else None
    q_start = int(mem_end) - 1
    for _ in range(10):
      predict._step(layer, mem, weights, q_start, buckets, rot, True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
      predict._step(layer, mem, weights, q_start, buckets, rot, True)
    e1.record()
    torch.cuda.synchronize()
    print('%s: layer.forward %.3f ms / token, lsh_predict_step %.3f ms / call (memory %d, %d slots filled)'
          % (name, api_ms, e0.elapsed_time(e1) / steps, M, int(mem_end)))


if __name__ == '__main__':
  main(int(sys.argv[1]) if len(sys.argv) > 1 else 200)
  del np
