#!/usr/bin/env python
"""bench.py — LSH-attention fwd+bwd tokens/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4-nh{1,2,4,8}|c5|c5-share]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one forward call of the layer (update_state=True: projection, hash, sort, attention, combine,
output projection) + one backward call (`backward(...)` = forward_and_or_backward(output_grad=g,
compute_output=False, update_state=False): recompute from the stored buckets, then the VJP), i.e. the
unit SURVEY.md §8(d) defines.  N>1: every rank runs the same per-GPU workload on its own examples (weak
scaling over batch) and the timed step ends with the NCCL all-reduce of the weight gradients
(the analogue of trax/optimizers/trainer.py:172-199).

`--workload c5` shards the 16 heads of one 1M-token example over the ranks instead (strong scaling) and all-reduces the
head-summed output and input gradient inside the step (trax_b200.dp.HeadShardedLSHSelfAttention).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU port (pinned to the reference's own code by
tests/test_reference_pin.py; the reference's jitted path needs JAX, which this image lacks) on the host cores instead:
every step is one of the workload's independent (example, head) units, `sample_scale` says how many make the workload.
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
    # BASELINE config 5: 16 heads of ONE 1M-token example sharded over the ranks (strong scaling: 16 / N heads per GPU,
    # all-reduce of the head-summed output and input gradient inside the timed step); N = 1 runs all 16 heads on one GPU
    'c5': dict(B=1, L=1 << 20, D=1024, H=16, C=128, nh=1, n_buckets=[32, 32], dtype='bf16', shard='heads',
               desc='long-context 1M-token LSH attention, d_model 1024, 16 heads sharded over the GPUs, chunk 128, 1 hash, n_buckets [32,32] (int32-key safe), causal'),
    'c5-share': dict(B=1, L=1 << 20, D=1024, H=2, C=128, nh=1, n_buckets=[32, 32], dtype='bf16',
                     desc='one GPU\'s share of config 5 at 8 GPUs (2 of 16 heads, 1M tokens), no collective'),
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
def _cpu_port_setup(wl, cores):
  """One (example, head) unit of the workload for the CPU arm: the torch restatement of `forward_unbatched`
  (oracle/lsh_oracle_torch.py, fp32, all host threads; torch.autograd stands in for jax.vjp, EA:2399-2421), buckets from
  the oracle's bit-exact hash + stable sort.  Returns step(): forward call (hash, sort, attention, no tape) followed by
  the backward call (forward recomputed under the tape from the stored buckets, then the VJP) — the same two calls
  our arm times."""
  import numpy as np
  import torch
  from oracle import lsh_oracle as O
  from oracle import lsh_oracle_torch as OT
  torch.set_num_threads(cores)
  L, D, H, C, nh = wl['L'], wl['D'], wl['H'], wl['C'], wl['nh']
  cfg = O.LSHConfig(n_heads=H, d_qk=64, d_v=64, causal=True, chunk_len=C, n_hashes=nh, n_buckets=wl['n_buckets'])
  rng = np.random.default_rng(0)
  x = torch.from_numpy(rng.standard_normal((L, D)).astype(np.float32))
  w = O.init_weights(H, D, 64, 64, seed=1)
  wq, wv, wo = (torch.from_numpy(np.ascontiguousarray(a[0]).astype(np.float32)) for a in w)
  rot = rng.standard_normal(O.rotations_shape(cfg, L)).astype(np.float32)
  dout = torch.from_numpy(rng.standard_normal((L, D)).astype(np.float32))
  kw = dict(seqlen=L, chunk_len=C, n_hashes=nh, n_chunks_before=1, n_chunks_after=0, causal=True)

  def step():
    with torch.no_grad():                                                     # forward call (update_state=True)
      q = (x @ wq).numpy()
      buckets = O.hash_vectors(cfg, q, rot)
      sticker, _ = O.sort_buckets(buckets, L)
      OT.forward_unit_torch(x, wq, wv, wo, sticker, **kw)
    leaves = [t.detach().requires_grad_(True) for t in (x, wq, wv, wo)]      # backward call: recompute + VJP
    sticker, _ = O.sort_buckets(buckets, L)
    out = OT.forward_unit_torch(*leaves, sticker, **kw)
    out.backward(dout)
    return leaves[0].grad
  return step


def run_reference(args, wl, name):
  """CPU arm (kind "port": the reference's jitted path needs JAX, which this image lacks; the port is pinned to the
  reference's own code by tests/test_reference_pin.py).  Every step is ONE of the workload's B*H independent
  (example, head) units — a fixed 1 / (B*H) sample — and all `--steps` are run; `ms_per_step` is what was timed,
  `sample_scale` the factor to the whole workload, `value` = B*L / (ms_per_step * sample_scale)."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):   # torchrun pins these to 1
    os.environ[var] = str(cores)
  units_total = wl['B'] * wl['H']
  step = _cpu_port_setup(wl, cores)
  for _ in range(args.warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    step()
  dt = (time.perf_counter() - t0) / args.steps
  tok_s = wl['B'] * wl['L'] / (dt * units_total)
  sample = '1 of %d (example, head) units per step (units are independent, EA:2402-2432), forward call + backward call with recompute, fp32 torch-CPU restatement, %d threads' % (units_total, cores)
  line = dict(metric='LSH-attn fwd+bwd tokens/sec', value=tok_s, unit='tokens/s', n_gpus=args.gpus, steps=args.steps,
              warmup=args.warmup, ms_per_step=dt * 1e3, sample_scale=units_total, higher_is_better=True,
              scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
              config=_config(name, wl, 1),
              cpu_baseline=dict(value=tok_s, unit='tokens/s', cores=cores, kind='port', sample=sample),
              e2e=dict(value=tok_s, unit='tokens/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
  print(json.dumps(line))


def _config(name, wl, world, head_reduce='all'):
  """Same dict (same keys, same workload string) for both arms."""
  if wl.get('shard') == 'heads':
    par = 'heads sharded over %d GPU(s) (%d per GPU), %s of the head-summed output and input gradient inside the step' % (
        world, wl['H'] // world, 'reduce-scatter (over the sequence)' if head_reduce == 'scatter' else 'all-reduce')
  elif world > 1:
    par = 'dp%d over batch, NCCL all-reduce of dW inside the step' % world
  else:
    par = 'single GPU'
  x_mb = wl['B'] * wl['L'] * wl['D'] * (2 if wl['dtype'] == 'bf16' else 4) / 1e6
  cache = ('inputs + per-step intermediates exceed the 126 MB L2 (x alone %.0f MB)' % x_mb if x_mb > 126 else
           'step working set (x %.1f MB) is smaller than the 126 MB L2: the step time is an L2-warm figure' % x_mb)
  return dict(workload=name + ': ' + wl['desc'], per_gpu_batch=wl['B'], parallelism=par,
              cache=cache + '; stage timings flush L2 before every timed launch')


def _bind_to_gpu_numa_node(local):
  """Pin this rank to the CPUs of its GPU's NUMA node BEFORE it allocates pinned host buffers (first touch places the
  pages there): with 8 ranks on one node the host-buffer path otherwise funnels every rank's PCIe traffic through
  one memory controller."""
  try:
    import torch
    bus = torch.cuda.get_device_properties(local).pci_bus_id
    dom = torch.cuda.get_device_properties(local).pci_domain_id
    dev = torch.cuda.get_device_properties(local).pci_device_id
    path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (dom, bus, dev)
    node = int(open(path).read().strip())
    if node < 0:
      return None
    cpus = set()
    for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
      a, _, b = part.partition('-')
      cpus.update(range(int(a), int(b or a) + 1))
    os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
    return node
  except Exception:  # pylint: disable=broad-except
    return None


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
  numa = _bind_to_gpu_numa_node(local) if world > 1 and not os.environ.get('LSH_BENCH_NO_NUMA') else None
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
      os.environ['NCCL_DEBUG'] = 'WARN'          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    # NCCL's kernels go on a high-priority stream: the weight-gradient all-reduce has to get SMs while the layer's last
    # GEMMs still fill the machine (at default priority it started only when they were done: fully exposed)
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True) if not os.environ.get('LSH_BENCH_NCCL_DEFAULT_PRIO') else None
    dist.init_process_group('nccl', device_id=dev, pg_options=opts)
  dtype = torch.bfloat16 if wl['dtype'] == 'bf16' else torch.float32
  B, L, D, H = wl['B'], wl['L'], wl['D'], wl['H']
  head_sharded = wl.get('shard') == 'heads'
  if head_sharded and H % world != 0:
    raise SystemExit('workload %s: %d heads do not divide over %d GPUs' % (name, H, world))
  H_local = H // world if head_sharded else H

  layer_kw = dict(d_qk=64, d_v=64, causal=True, chunk_len=wl['C'], n_hashes=wl['nh'], n_buckets=wl['n_buckets'], mode='train')
  if args.attention_dropout > 0.0:     # the reference's training configs (reformer_enwik8.gin:94, reformer_imagenet64.gin:69: 0.2)
    layer_kw['attention_dropout'] = args.attention_dropout
  if head_sharded:
    # the full 16-head layer's weights (same rng on every rank; drawn on the host), this rank's heads taken out of them
    full = trax_b200.LSHSelfAttention(n_heads=H, **layer_kw)
    full.init_weights_and_state(trax_b200.ShapeDtype((B, L, D)), device='cpu')
    layer = dp.HeadShardedLSHSelfAttention(trax_b200.LSHSelfAttention(n_heads=H_local, **layer_kw), H, reduce=args.head_reduce)
    layer.load_full(full.weights, full.state)
    layer.local.weights = tuple(w.to(dev) for w in layer.local.weights)
    layer.local.state = tuple(s.to(dev) for s in layer.local.state)
    del full
  else:
    layer = trax_b200.LSHSelfAttention(n_heads=H, **layer_kw)
    layer.init(trax_b200.ShapeDtype((B, L, D)), rng=np.array([0, 1], np.uint32))   # same weights on every rank
  # batch sharding: every rank has its own examples; head sharding: every rank sees the SAME example (replicated input)
  g = torch.Generator(device=dev).manual_seed(1234 + (0 if head_sharded else rank))
  x = torch.randn((B, L, D), generator=g, device=dev, dtype=torch.float32).to(dtype)
  dout = torch.randn((B, L, D), generator=g, device=dev, dtype=torch.float32).to(dtype)
  x_host, dout_host = x.cpu().pin_memory(), dout.cpu().pin_memory()

  # psum/n of the weight gradients (trainer.py:194-199) runs inside the backward call: one in-place all-reduce of the
  # contiguous gradient buffer after the call's last kernel
  # (--grad-allreduce overlap: the slices travel on a communication stream underneath the call's remaining kernels instead;
  # measured 2.56 vs 2.50 ms at 2 GPUs, but 2.9-3.0 vs 2.52 ms at 8 GPUs on one box — DESIGN.md section 7)
  trax_b200.set_weight_grad_allreduce(args.grad_allreduce if (world > 1 and not head_sharded) else False)

  # the dropout masks are a function of the rng: the backward call gets the one the forward call used
  bwd_rng = getattr(layer, 'rng', np.array([0, 0], np.uint32)) if args.attention_dropout > 0.0 else None
  def step(xi, gi):
    out = layer.forward(xi)                                                                   # forward call
    dx, dw = layer.backward(xi, out, gi, layer.weights, None, layer.state, bwd_rng)           # backward call (+ collectives)
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
  n_probe = 0
  while time.perf_counter() - t_probe < 0.6 or n_probe % 4:      # (a multiple of 4 steps: the same count on every rank)
    step(x, dout)
    n_probe += 1
    if world > 1 and n_probe >= 64:
      break
  if world > 1:                          # ranks must issue the same number of collectives
    n_all = torch.tensor([n_probe], device=dev)
    dist.all_reduce(n_all, op=dist.ReduceOp.MAX)
    for _ in range(int(n_all.item()) - n_probe):
      step(x, dout)
  torch.cuda.synchronize()
  clocks = sampler.stop()
  clocks['note'] = 'sampled over warm-up + timed steps + >= 0.6 s of identical untimed steps'
  tokens_per_step = (1 if head_sharded else world) * B * L
  tok_s = tokens_per_step / (ms * 1e-3)

  # ---- e2e: host (pinned) buffers through the layer API, copies inside the timed region ----
  # Asynchronous dispatch (JAX-style): each call enqueues its uploads / kernels / downloads and returns; the timed region
  # ends with a full synchronize, so every byte of every step has crossed PCIe inside it.  Bytes are counted from the
  # tensors actually copied (backward(x, ...) re-uses forward's device copy of the same host tensor).
  e2e_steps = max(1, min(args.steps, 10))
  trax_b200.set_async_host_io(True)
  step(x_host, dout_host)
  sync()
  trax_b200.host_io_bytes(reset=True)
  ms_e2e = timed(lambda: step(x_host, dout_host), e2e_steps, 2)
  h2d, d2h = trax_b200.host_io_bytes(reset=True)
  trax_b200.set_async_host_io(False)
  e2e = dict(value=tokens_per_step / (ms_e2e * 1e-3), unit='tokens/s', h2d_bytes_per_step=h2d // (e2e_steps + 2),
             d2h_bytes_per_step=d2h // (e2e_steps + 2), ms_per_step=ms_e2e,
             note='per rank: pinned host x/dout in, out/dx/dW to pinned host; uploads, kernels and downloads overlap across calls'
                  + ('; rank pinned to NUMA node %d of its GPU' % numa if numa is not None else ''))

  line = dict(metric='LSH-attn fwd+bwd tokens/sec', value=tok_s, unit='tokens/s', n_gpus=world, steps=args.steps,
              warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling='strong' if head_sharded else 'weak',
              vs_baseline=None, dtype='bf16' if dtype == torch.bfloat16 else 'f32 I/O, bf16 tensor-core operands',
              data='synthetic', config=_config(name, wl, world, args.head_reduce), clocks=clocks, e2e=e2e, gpu_launches=launches)
  if args.attention_dropout > 0.0:
    line['config']['attention_dropout'] = args.attention_dropout    # (not a BASELINE config: the reference gins' training setting)
  if head_sharded and world > 1:
    # out and dx, all-reduced once each per step (counted by the layer from the tensors it reduced)
    line['collective_bytes_per_step'] = layer.comm_bytes // max(1, layer.n_calls // 2)

  if rank == 0:
    peaks = _peaks()
    wl_local = dict(wl, H=H_local)
    stages = stage_breakdown(layer.local if head_sharded else layer, x, wl_local, args)
    line['stages_ms'] = stages['ms']
    line['roofline'] = stages['roofline'](peaks)
    line['roofline_hbm'] = stages['roofline_hbm'](peaks)
    line['roofline_layer'] = layer_roofline(wl_local, ms, peaks)
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
              frac=t_roof / (ms * 1e-3), peaks=peaks['src'] + ' (sustained tensor figure: kernels timed inside a long step)',
              note='per GPU: this rank\'s heads; collectives are not part of the model')


def stage_breakdown(layer, x, wl, args):
  """Times every stage alone with CUDA events on the launching stream, L2 flushed (a 256 MB write) before every timed
  launch so that no stage finds its inputs in the 126 MB L2; the attention-gradient kernel, the per-token preparation
  kernels and the sum over rounds are timed separately (lsh_debug_set_bwd_parts)."""
  import ctypes
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
  ops.attend_bwd(dims, qv, sticker, o_c, lse, do)          # fills the stage's workspace (parts are re-timed on it)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=x.device)
  n = max(3, min(args.steps, 10))
  lib = _lib.load()

  def cold(fn):
    tot = 0.0
    for i in range(n + 1):
      flush.zero_()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      fn()
      e1.record()
      torch.cuda.synchronize()
      if i:                                                 # first launch = warm-up
        tot += e0.elapsed_time(e1)
    return tot / n

  def bwd_part(mask):
    def run():
      lib.lsh_debug_set_bwd_parts(mask)
      try:
        ops.attend_bwd(dims, qv, sticker, o_c, lse, do)
      finally:
        lib.lsh_debug_set_bwd_parts(7)
    return run

  ms = {}
  ms['project_qv'] = cold(lambda: ops.project_qv(dims, xb, wqv))
  ms['hash'] = cold(lambda: ops.hash_qv(dims, qv, rot, buckets=buckets))
  ms['sort(3 kernels)'] = cold(lambda: ops.sort(dims, buckets, want_undo=False))
  ms['attend_fwd(aux + kernel)'] = cold(lambda: ops.attend_fwd(dims, qv, sticker))
  ms['combine_fwd'] = cold(lambda: ops.combine_fwd(dims, o_r, logits))
  ms['attend_bwd: prep kernels'] = cold(bwd_part(1))
  ms['attend_bwd: gradient kernel'] = cold(bwd_part(2))
  ms['attend_bwd: sum over rounds'] = cold(bwd_part(4))
  N, W, BH = nh * L, 2 * C, B * H
  gemm = 2.0 * N * W * 64 * BH          # one chunked contraction (QK^T or PV) over all units, SURVEY §8(d)
  # algorithmic bytes of the HBM-bound row kernels (one read / one write of every row they touch)
  hbm = {
      'combine_fwd': BH * (N * 64 * 2 + N * 4 + L * 64 * 2 + L * 4),
      'attend_bwd: sum over rounds': BH * (2 * N * 64 * 2 + L * 128 * 2),
      'attend_bwd: prep kernels': BH * (L * 64 * 2 + L * 4 + 2 * L * 64 * 2 + L * 4 + 3 * L * 4 + 2 * N * 4),
      'sort(3 kernels)': BH * (N * 4 * 2 + N * 4),
  }

  def roofline(peaks):
    # dominant kernel: the attention-gradient kernel ALONE, credited with the 4 gradient GEMMs of §8(d) (dV, dP, dQ, dK;
    # the S recompute it also performs is overhead, not credited)
    k = 'attend_bwd: gradient kernel'
    ach = 4 * gemm / (ms[k] * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    if not os.path.exists(tp):
      tp = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    if os.path.exists(tp):
      tr = json.load(open(tp))
      traffic = (tr.get('attend_bwd_tc_kernel') or tr.get('attend_bwd(prep+bwd+sum_rounds)') or {}).get('bytes')
    return dict(kernel='attend_bwd_tc_kernel' if C == 128 else 'attend_bwd_kernel', bound='tensor', achieved=ach, peak=peaks['tc'],
                unit='TFLOP/s', frac=ach / peaks['tc'], traffic=traffic, traffic_note='ncu capture at config 2 (profiles/)',
                peak_source=peaks['src'] + ' burst (kernel timed alone)', ms_per_launch=ms[k],
                algorithmic_flops=4 * gemm, timing='CUDA events around one launch, L2 flushed before it')

  def roofline_hbm(peaks):
    rows = {}
    for k, b in hbm.items():
      rows[k] = dict(achieved=b / (ms[k] * 1e-3) / 1e9, frac=b / (ms[k] * 1e-3) / 1e9 / peaks['hbm'], algorithmic_bytes=b,
                     ms_per_launch=ms[k])
    worst = min(('combine_fwd', 'attend_bwd: sum over rounds', 'attend_bwd: prep kernels'), key=lambda k: rows[k]['frac'])
    return dict(kernel=worst, bound='hbm', achieved=rows[worst]['achieved'], peak=peaks['hbm'], unit='GB/s',
                frac=rows[worst]['frac'], traffic=None, peak_source=peaks['src'], per_kernel=rows)
  return dict(ms=ms, roofline=roofline, roofline_hbm=roofline_hbm)


def cpu_baseline(wl):
  """The CPU port timed on the host cores, bounded sample: one (example, head) unit of the workload, one step."""
  cores = os.cpu_count() or 1
  units = wl['B'] * wl['H']
  step = _cpu_port_setup(wl, cores)
  t0 = time.perf_counter()
  step()
  dt = time.perf_counter() - t0
  return dict(value=wl['B'] * wl['L'] / (dt * units), unit='tokens/s', cores=cores, kind='port',
              sample='1 of %d (example, head) units, forward call + backward call with recompute, fp32 torch-CPU restatement, %.1f s measured, scaled x%d'
              % (units, dt, units))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--grad-allreduce', default='flat', choices=['flat', 'overlap'],
                  help='data-parallel weight-gradient mean inside the backward call: one flat all-reduce after the last kernel '
                       '(default) or two slices overlapped with the rest of the call')
  ap.add_argument('--attention-dropout', type=float, default=0.0,
                  help='attention dropout rate of the layer (default 0: the BASELINE configs; the reference gins train with 0.2)')
  ap.add_argument('--head-reduce', default='all', choices=['all', 'scatter'],
                  help='config 5: all-reduce the head sums, or reduce-scatter them over the sequence (sequence-parallel consumer)')
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
