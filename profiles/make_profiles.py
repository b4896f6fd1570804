"""Regenerates the committed per-round evidence from one gpurun capture (see the commands in each output's header):
  gpurun_out/kernels_<tag>.csv   ncu --metrics gpu__time_duration.sum,dram__bytes_{read,write}.sum,...pipe/issue... of bench.py
  gpurun_out/prof_<tag>_{attend_bwd,attend_fwd,hash}.ncu-rep   ncu --set full of tests/prof_stage.py <stage>
  gpurun_out/bench_<tag>.json    the bench line of the same build
Usage: python profiles/make_profiles.py <tag> [round prefix, e.g. r2]"""
import collections, csv, io, json, re, shutil, subprocess, sys
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else 'r1'   # file-name prefix of the round
shutil.copy('gpurun_out/bench_%s.json' % tag, 'profiles/%s_bench_line.json' % rnd)
shutil.copy('gpurun_out/kernels_%s.csv' % tag, 'profiles/%s_kernels_metrics.csv' % rnd)
rows = [r for r in csv.reader(open('gpurun_out/kernels_%s.csv' % tag)) if len(r) > 5]
hdr = [r for r in rows if 'Kernel Name' in r][0]; ci = {h: i for i, h in enumerate(hdr)}
byid = collections.OrderedDict()
for r in rows:
  if r is hdr or not r[ci['ID']].isdigit():
    continue
  d = byid.setdefault(int(r[ci['ID']]), {'name': re.sub(r'\(.*', '', r[ci['Kernel Name']])[:60]})
  v = float(r[ci['Metric Value']].replace(',', '')); u = r[ci['Metric Unit']]
  if 'byte' in u: v *= {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[u]
  if u == 'us': v *= 1e3
  if u == 'ms': v *= 1e6
  d[r[ci['Metric Name']]] = v
ids = list(byid)
h = [i for i in ids if 'make_rotations' in byid[i]['name']]
if len(h) >= 3:
  a, b = h[1], h[2]                  # one full step: every forward call starts by drawing its rotations
else:                                # (round-1 builds: rotations came from the host; 5 launches precede the hash kernel)
  h = [i for i in ids if 'hash_' in byid[i]['name']]
  a, b = h[1] - 5, h[2] - 5
sel = [i for i in ids if a <= i < b]
agg = collections.OrderedDict()
for i in sel:
  d = byid[i]; e = agg.setdefault(d['name'], dict(n=0, t=0, rd=0, wr=0, tp=0, fma=0, iss=0))
  e['n'] += 1; e['t'] += d.get('gpu__time_duration.sum', 0); e['rd'] += d.get('dram__bytes_read.sum', 0); e['wr'] += d.get('dram__bytes_write.sum', 0)
  e['tp'] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0)
  e['fma'] += d.get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 0)
  e['iss'] += d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0)
tot = sum(e['t'] for e in agg.values())
out = ['Per-kernel counters of ONE fwd+bwd step of bench.py (workload c2) under ncu (--clock-control none; cold-cache, serialised:',
       'compare shares and per-launch DRAM bytes, not absolute times).  HBM peak measured on this pool: 6546.6 GB/s (MEASURED_PEAKS.json).',
       'Command: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...,',
       '         sm__pipe_fma_cycles_active...,smsp__issue_active... --clock-control none -c 260 python bench.py --steps 2 --warmup 3',
       '%-58s %3s %9s %6s %9s %9s %8s %7s %6s %6s' % ('kernel', 'x', 'us/launch', 'share', 'rd MB', 'wr MB', 'GB/s', 'tensor%', 'fma%', 'issue%')]
for n, e in sorted(agg.items(), key=lambda kv: -kv[1]['t']):
  t = e['t'] / e['n']
  out.append('%-58s %3d %9.1f %5.1f%% %9.1f %9.1f %8.0f %7.1f %6.1f %6.1f' % (
      n, e['n'], t / 1e3, 100 * e['t'] / tot, e['rd'] / e['n'] / 1e6, e['wr'] / e['n'] / 1e6, (e['rd'] + e['wr']) / e['n'] / t,
      e['tp'] / e['n'], e['fma'] / e['n'], e['iss'] / e['n']))
out.append('total %.1f us over %d launches' % (tot / 1e3, len(sel)))
open('profiles/%s_kernel_table.txt' % rnd, 'w').write('\n'.join(out) + '\n')
# launch list in the older format (time only) for summarize_launches.py
with open('profiles/%s_launches.csv' % rnd, 'w') as f:
  w = csv.writer(f, quoting=csv.QUOTE_ALL); w.writerow(hdr)
  for r in rows:
    if r is not hdr and r[ci['Metric Name']] == 'gpu__time_duration.sum':
      w.writerow(r)
keys = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'launch__block_size', 'launch__grid_size',
        'launch__registers_per_thread', 'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
full = []; traffic = {'_comment': 'dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` (profiles/%s_ncu_full_summary.txt)' % rnd + ', workload c2'}
names = {'attend_bwd': 'attend_bwd_tc_kernel', 'attend_fwd': 'attend_fwd_tc_kernel', 'hash': 'hash'}
for st in ['attend_bwd', 'attend_fwd', 'hash']:
  raw = subprocess.run(['ncu', '-i', 'gpurun_out/prof_%s_%s.ncu-rep' % (tag, st), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rr = list(csv.reader(io.StringIO(raw))); hh, uu, r = rr[0], rr[1], rr[2]
  full.append('== prof_%s_%s   (ncu --set full --clock-control none --import-source on, one launch of the C2 workload; tests/prof_stage.py %s)' % (rnd, st, st))
  full.append('%-70s %s' % ('Kernel Name', r[hh.index('Kernel Name')]))
  for k in keys:
    if k in hh: full.append('%-70s %-16s %s' % (k, uu[hh.index(k)], r[hh.index(k)]))
  def val(k):
    return float(r[hh.index(k)].replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(uu[hh.index(k)], 1)
  traffic[names[st]] = {'kernel': r[hh.index('Kernel Name')].split('(')[0], 'bytes': int(val('dram__bytes_read.sum') + val('dram__bytes_write.sum'))}
open('profiles/%s_ncu_full_summary.txt' % rnd, 'w').write('\n'.join(full) + '\n')
json.dump(traffic, open('profiles/%s_traffic.json' % rnd, 'w'), indent=1)
print('\n'.join(out))
