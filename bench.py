#!/usr/bin/env python
"""bench.py — LSH-attention fwd+bwd tokens/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one forward call of the layer (update_state=True: projection, hash, sort, attention, combine,
output projection) + one backward call (`backward(...)` = forward_and_or_backward(output_grad=g,
compute_output=False, update_state=False): recompute from the stored buckets, then the VJP), i.e. the
unit SURVEY.md §8(d) defines.  N>1: every rank runs the same per-GPU workload on its own examples (weak
scaling over batch) and the timed step ends with the NCCL all-reduce of the weight gradients
(the analogue of trax/optimizers/trainer.py:172-199).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (pinned to the reference's own code by
tests/test_reference_pin.py; the reference's jitted path needs JAX, which this image lacks) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, L, D, H, chunk_len, n_hashes, n_buckets, dtype)
    'c1': dict(B=1, L=1024, D=256, H=2, C=64, nh=1, n_buckets=32, dtype='f32',
               desc='single tl.LSHSelfAttention layer: batch 1, seq 1024, d_model 256, 2 heads, chunk 64, 32 buckets, 1 hash, causal'),
    'c2': dict(B=1, L=65536, D=1024, H=8, C=128, nh=4, n_buckets=None, dtype='bf16',
               desc='ReformerLM enwik8-style LSH layer: seq 65536, d_model 1024, 8 heads, d_qk=d_v=64, chunk 128, 4 hashes, n_buckets auto [32,32], causal'),
    'c3': dict(B=1, L=12288, D=1024, H=8, C=128, nh=2, n_buckets=192, dtype='bf16',
               desc='ReformerLM imagenet64-style LSH layer: seq 12288, d_model 1024, 8 heads, 2 hashes, 192 buckets, 1 example per GPU'),
    'c5': dict(B=1, L=1 << 20, D=1024, H=2, C=128, nh=1, n_buckets=[32, 32], dtype='bf16',
               desc='long-context 1M-token LSH attention, per-GPU share of 16 heads over 8 GPUs (2 heads), 1 hash, n_buckets [32,32] (int32-key safe)'),
}
for _nh in (1, 2, 4, 8):   # BASELINE config 4: n_hashes sweep at seq 16384 (n_buckets None -> [16, 16])
  WORKLOADS['c4-nh%d' % _nh] = dict(B=1, L=16384, D=1024, H=8, C=128, nh=_nh, n_buckets=None, dtype='bf16',
                                    desc='n_hashes sweep member: seq 16384, d_model 1024, 8 heads, chunk 128, %d hashes, n_buckets auto [16,16], causal' % _nh)
WORKLOADS['c4'] = WORKLOADS['c4-nh4']


def _peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    d = json.load(open(p))
    return dict(hbm=d['hbm_gbs'], tc=d['bf16_tflops'], tc_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                src='MEASURED_PEAKS.json')
  return dict(hbm=6650.0, tc=1590.0, tc_sustained=1400.0, src='fallback (B200_PROFILING.md)')


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    self.rows, self.proc, self.thr, self.idx = [], None, None, gpu_index

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits', '-lms', '50'],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
      return
    def pump():
      for line in self.proc.stdout:
        self.rows.append([c.strip() for c in line.split(',')])
    self.thr = threading.Thread(target=pump, daemon=True)
    self.thr.start()

  def stop(self):
    if self.proc is None:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, reasons, pw = [], [], set(), []
    for r in self.rows:
      try:
        sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
      except (ValueError, IndexError):
        continue
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
        if v.lower().startswith('active'):
          reasons.add(name)
    sm.sort()
    return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# ---------------------------------------------------------------------------------------------------
def run_reference(args, wl, name):
  """CPU arm: the oracle restatement of the reference layer (kind "port"), all host threads, bounded sample."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):   # torchrun pins these to 1
    os.environ[var] = str(cores)
  import numpy as np
  from oracle import lsh_oracle as O
  try:
    import threadpoolctl
    threadpoolctl.threadpool_limits(cores)
  except Exception:  # pylint: disable=broad-except
    pass
  L, D, H = wl['L'], wl['D'], wl['H']
  cfg = O.LSHConfig(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=wl['C'], n_hashes=wl['nh'], n_buckets=wl['n_buckets'])
  rng = np.random.default_rng(0)
  x = rng.standard_normal((L, D)).astype(np.float32)
  w = O.init_weights(H, D, 64, 64, seed=1)
  rot = rng.standard_normal(O.rotations_shape(cfg, L)).astype(np.float32)
  dout = rng.standard_normal((L, D)).astype(np.float32)
  units_total = wl['B'] * H
  sample_units = 1 if L >= 8192 else units_total

  def step():
    for h in range(sample_units):
      res = O.forward_unit(cfg, x, w[0][h], w[1][h], w[2][h], rotations=rot, dtype=np.float32)     # forward call
      res2 = O.forward_unit(cfg, x, w[0][h], w[1][h], w[2][h], buckets=res.buckets, dtype=np.float32)  # recompute
      O.backward_unit(cfg, res2, dout)                                                              # backward call
  for _ in range(min(args.warmup, 1) if L >= 8192 else args.warmup):
    step()
  steps = args.steps if L < 8192 else min(args.steps, 2)
  t0 = time.perf_counter()
  for _ in range(steps):
    step()
  dt = (time.perf_counter() - t0) / steps
  tok_s = wl['B'] * L / (dt * units_total / sample_units)
  sample = '%d of %d (example, head) units of the workload per step, fwd call + bwd call (with recompute), fp32 NumPy/BLAS oracle; scaled x%d' % (
      sample_units, units_total, units_total // sample_units)
  line = dict(metric='LSH-attn fwd+bwd tokens/sec', value=tok_s, unit='tokens/s', n_gpus=args.gpus, steps=steps,
              warmup=args.warmup, ms_per_step=dt * 1e3 * units_total / sample_units, higher_is_better=True,
              scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
              config=dict(workload=name + ': ' + wl['desc'], timing='host wall clock (CPU arm)'),
              cpu_baseline=dict(value=tok_s, unit='tokens/s', cores=cores, kind='port', sample=sample),
              e2e=dict(value=tok_s, unit='tokens/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
  print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def run_ours(args, wl, name):
  import numpy as np
  import torch
  import torch.distributed as dist
  import trax_b200
  from trax_b200 import ops, _lib, dp

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  # Libraries (NCCL's version banner) write to fd 1: park stdout on stderr until the ONE JSON line is printed
  sys.stdout.flush()
  saved_stdout = os.dup(1)
  os.dup2(2, 1)
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
      os.environ['NCCL_DEBUG'] = 'WARN'          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    dist.init_process_group('nccl', device_id=dev)
  dtype = torch.bfloat16 if wl['dtype'] == 'bf16' else torch.float32
  B, L, D, H = wl['B'], wl['L'], wl['D'], wl['H']

  layer = trax_b200.LSHSelfAttention(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=wl['C'], n_hashes=wl['nh'],
                                     n_buckets=wl['n_buckets'], mode='train')
  layer.init(trax_b200.ShapeDtype((B, L, D)), rng=np.array([0, 1], np.uint32))   # same weights on every rank
  g = torch.Generator(device=dev).manual_seed(1234 + rank)
  x = torch.randn((B, L, D), generator=g, device=dev, dtype=torch.float32).to(dtype)
  dout = torch.randn((B, L, D), generator=g, device=dev, dtype=torch.float32).to(dtype)
  x_host, dout_host = x.cpu().pin_memory(), dout.cpu().pin_memory()
  weights = layer.weights

  # psum/n of the weight gradients (trainer.py:194-199) runs inside the backward call, on the device (one flat NCCL
  # all-reduce of 6.3 MB on the compute stream) — for host-buffer calls that is before the gradients are downloaded
  trax_b200.set_weight_grad_allreduce(world > 1)

  def step(xi, gi):
    out = layer.forward(xi)                                                                   # forward call
    dx, dw = layer.backward(xi, out, gi, weights, None, layer.state, None)                    # backward call (+ all-reduce)
    return out, dx, dw

  def sync():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps, warmup):
    for _ in range(warmup):
      fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    return ms / steps

  def timed_local(fn, steps, warmup):
    """Rank-local CUDA-event timing (no collectives): used for the per-stage breakdown on rank 0."""
    for _ in range(warmup):
      fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

  # ---- headline: device-resident inputs ----
  sampler = ClockSampler(local)
  sampler.start()                       # samples across warm-up, the timed steps and the post-probe below
  for _ in range(args.warmup):
    step(x, dout)
  sync()
  ops.launch_count(reset=True)
  ms = timed(lambda: step(x, dout), args.steps, 0)
  launches = ops.launch_count(reset=True)
  # the timed region is tens of ms, shorter than nvidia-smi's sampling period: keep the SAME steps running (untimed)
  # until the sampler has seen >= 0.6 s of this load, so that the clock / throttle record describes it
  t_probe = time.perf_counter()
  while time.perf_counter() - t_probe < 0.6:
    step(x, dout)
  torch.cuda.synchronize()
  clocks = sampler.stop()
  clocks['note'] = 'sampled over warm-up + timed steps + 0.6 s of identical untimed steps'
  tok_s = world * B * L / (ms * 1e-3)

  # ---- e2e: host (pinned) buffers through the layer API, copies inside the timed region ----
  # Asynchronous dispatch (JAX-style): each call enqueues its uploads / kernels / downloads and returns; the timed region
  # ends with a full synchronize, so every byte of every step has crossed PCIe inside it.  Bytes are counted from the
  # tensors actually copied (backward(x, ...) re-uses forward's device copy of the same host tensor).
  esz = x.element_size()
  e2e_steps = max(1, min(args.steps, 10))
  trax_b200.set_async_host_io(True)
  step(x_host, dout_host)
  sync()
  trax_b200.host_io_bytes(reset=True)
  ms_e2e = timed(lambda: step(x_host, dout_host), e2e_steps, 2)
  h2d, d2h = trax_b200.host_io_bytes(reset=True)
  trax_b200.set_async_host_io(False)
  e2e = dict(value=world * B * L / (ms_e2e * 1e-3), unit='tokens/s', h2d_bytes_per_step=h2d // (e2e_steps + 2),
             d2h_bytes_per_step=d2h // (e2e_steps + 2), ms_per_step=ms_e2e,
             note='pinned host x/dout in, out/dx/dW to pinned host; uploads, kernels and downloads overlap across calls')

  line = dict(metric='LSH-attn fwd+bwd tokens/sec', value=tok_s, unit='tokens/s', n_gpus=world, steps=args.steps,
              warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None,
              dtype='bf16' if dtype == torch.bfloat16 else 'f32 I/O, bf16 tensor-core operands', data='synthetic',
              config=dict(workload=name + ': ' + wl['desc'], per_gpu_batch=B,
                          parallelism='dp%d over batch, NCCL all-reduce of dW inside the step' % world if world > 1 else 'single GPU',
                          cache='inputs + per-step intermediates exceed the 126 MB L2 (x alone %.0f MB); no explicit flush' % (B * L * D * esz / 1e6)),
              clocks=clocks, e2e=e2e, gpu_launches=launches)

  if rank == 0:
    peaks = _peaks()
    stages = stage_breakdown(layer, x, dout, wl, args, timed_local)
    line['stages_ms'] = stages['ms']
    line['roofline'] = stages['roofline'](peaks)
    line['roofline_layer'] = layer_roofline(wl, ms, peaks)
    if world == 1 and not args.no_cpu_baseline:
      line['cpu_baseline'] = cpu_baseline(wl)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  sys.stdout.flush()
  os.dup2(saved_stdout, 1)
  os.close(saved_stdout)
  if rank == 0:
    print(json.dumps(line))


def layer_roofline(wl, ms, peaks):
  """per-token algorithmic FLOPs and bytes of fwd+bwd -> roofline time; fraction achieved."""
  L, D, H, C, nh = wl['L'], wl['D'], wl['H'], wl['C'], wl['nh']
  W = 2 * C
  from trax_b200.ops import bucket_factors
  R = sum(bucket_factors(wl['n_buckets'], L, C)) // 2
  F = H * (2 * D * 128 + 2 * nh * W * 128 + 2 * 64 * D)
  flops = (3 * F + H * 2 * 64 * nh * R) * L * wl['B']
  e = 2
  fwd_b = D * e + H * 128 * e + 4 * H * nh + 12 * H * nh + (4 * H * nh + H * nh * 128 * e + H * nh * 64 * e + 4 * H * nh) + \
      (H * nh * 64 * e + 8 * H * nh + H * 64 * e) + (H * 64 * e + D * e)
  bwd_b = (D * e + 2 * H * 64 * e) + (H * 64 * e + 2 * H * nh * 64 * e + 12 * H * nh) + \
      (16 * H * nh + H * nh * 128 * e + H * nh * 64 * e + H * nh * 128 * e) + (H * nh * 128 * e + 4 * H * nh + H * 128 * e) + \
      (H * 128 * e + 2 * D * e)
  nbytes = (fwd_b + bwd_b) * L * wl['B']
  t_tc = flops / (peaks['tc_sustained'] * 1e12)
  t_hbm = nbytes / (peaks['hbm'] * 1e9)
  t_roof = max(t_tc, t_hbm)
  return dict(flops=flops, bytes=nbytes, t_tensor_ms=t_tc * 1e3, t_hbm_ms=t_hbm * 1e3, bound='tensor' if t_tc >= t_hbm else 'hbm',
              frac=t_roof / (ms * 1e-3), peaks=peaks['src'] + ' (sustained tensor figure: kernels timed inside a long step)')


def stage_breakdown(layer, x, dout, wl, args, timed):
  """Times each kernel stage alone (CUDA events, same buffers) so the dominant kernel's roofline can be reported."""
  import torch
  from trax_b200 import ops, _lib
  B, L, D, H, C, nh = wl['B'], wl['L'], wl['D'], wl['H'], wl['C'], wl['nh']
  factors = ops.bucket_factors(wl['n_buckets'], L, C)
  dims = _lib.make_dims(B, H, L, D, 64, 64, C, 1, 0, nh, factors, True, False, _lib.LSH_DTYPE_BF16)
  xb = x.to(torch.bfloat16).contiguous()
  wqv, wo = ops.pack_weights(dims, *layer.weights)
  qv = ops.project_qv(dims, xb, wqv)
  keys = torch.arange(2 * B * H, dtype=torch.int32, device=x.device).reshape(B * H, 2)
  rot, _ = ops.make_rotations(dims, keys)
  buckets = ops.hash_qv(dims, qv, rot)
  sticker, _ = ops.sort(dims, buckets)
  o_r, logits = ops.attend_fwd(dims, qv, sticker)
  o_c, lse = ops.combine_fwd(dims, o_r, logits)
  do = torch.randn_like(o_c)
  n = max(3, min(args.steps, 10))
  ms = {}
  ms['project_qv(cublas)'] = timed(lambda: ops.project_qv(dims, xb, wqv), n, 2)
  ms['hash'] = timed(lambda: ops.hash_qv(dims, qv, rot, buckets=buckets), n, 2)
  ms['sort(3 kernels)'] = timed(lambda: ops.sort(dims, buckets, want_undo=False), n, 2)
  ms['attend_fwd'] = timed(lambda: ops.attend_fwd(dims, qv, sticker), n, 2)
  ms['combine_fwd'] = timed(lambda: ops.combine_fwd(dims, o_r, logits), n, 2)
  ms['attend_bwd(prep+bwd+sum_rounds)'] = timed(lambda: ops.attend_bwd(dims, qv, sticker, o_c, lse, do), n, 2)
  N, W = nh * L, 2 * C
  gemm = 2.0 * N * W * 64 * B * H          # one chunked contraction over all units
  algo = {
      'attend_fwd': dict(bound='tensor', work=2 * gemm, unit='TFLOP/s'),
      'attend_bwd(prep+bwd+sum_rounds)': dict(bound='tensor', work=5 * gemm, unit='TFLOP/s'),
      'hash': dict(bound='hbm', work=B * H * L * (128 + 4 * nh), unit='GB/s'),
  }

  def roofline(peaks):
    top = max(algo, key=lambda k: ms[k])
    a = algo[top]
    if a['bound'] == 'tensor':
      ach = a['work'] / (ms[top] * 1e-3) / 1e12
      peak = peaks['tc']
    else:
      ach = a['work'] / (ms[top] * 1e-3) / 1e9
      peak = peaks['hbm']
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    if os.path.exists(tp):
      traffic = json.load(open(tp)).get(top, {}).get('bytes')   # DRAM bytes per launch from the committed ncu capture
    return dict(kernel=top, bound=a['bound'], achieved=ach, peak=peak, unit=a['unit'], frac=ach / peak, traffic=traffic,
                peak_source=peaks['src'] + (' burst' if a['bound'] == 'tensor' else ''), ms_per_launch=ms[top])
  return dict(ms=ms, roofline=roofline)


def cpu_baseline(wl):
  """Oracle ("port") timed on the host cores, bounded sample: one (example, head) unit of the workload."""
  import numpy as np
  from oracle import lsh_oracle as O
  cores = os.cpu_count() or 1
  L, D, H = wl['L'], wl['D'], wl['H']
  cfg = O.LSHConfig(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=wl['C'], n_hashes=wl['nh'], n_buckets=wl['n_buckets'])
  rng = np.random.default_rng(0)
  x = rng.standard_normal((L, D)).astype(np.float32)
  w = O.init_weights(H, D, 64, 64, seed=1)
  rot = rng.standard_normal(O.rotations_shape(cfg, L)).astype(np.float32)
  dout = rng.standard_normal((L, D)).astype(np.float32)
  units = wl['B'] * H
  t0 = time.perf_counter()
  res = O.forward_unit(cfg, x, w[0][0], w[1][0], w[2][0], rotations=rot, dtype=np.float32)
  res2 = O.forward_unit(cfg, x, w[0][0], w[1][0], w[2][0], buckets=res.buckets, dtype=np.float32)
  O.backward_unit(cfg, res2, dout)
  dt = time.perf_counter() - t0
  return dict(value=wl['B'] * L / (dt * units), unit='tokens/s', cores=cores, kind='port',
              sample='1 of %d (example, head) units, fwd call + bwd call, fp32 NumPy/BLAS oracle, %.1f s measured, scaled x%d'
              % (units, dt, units))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  wl = WORKLOADS[args.workload]
  if args.impl == 'reference':
    run_reference(args, wl, args.workload)
    return
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if args.gpus > 1 and world == 1:
    # convenience: re-launch under torchrun when invoked plainly with --gpus N
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
           '--master-addr', '127.0.0.1', '--master-port', '29531', os.path.abspath(__file__)] + sys.argv[1:]
    sys.exit(subprocess.call(cmd))
  args.warmup = max(args.warmup, 3)
  run_ours(args, wl, args.workload)


if __name__ == '__main__':
  main()
