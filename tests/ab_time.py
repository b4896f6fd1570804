"""A/B helper (not a test): per-kernel device times (ncu gpu__time_duration, --clock-control none) of one stage for the
round-1 tree (_r1/, if present) and the working tree on the same box.   usage: python tests/ab_ncu.py <stage> [L] [tag]"""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
stage = sys.argv[1] if len(sys.argv) > 1 else 'attend_bwd'
L = sys.argv[2] if len(sys.argv) > 2 else '65536'
tag = sys.argv[3] if len(sys.argv) > 3 else 'ab'
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
for name, tree in (('r1', os.path.join(ROOT, '_r1')), ('new', ROOT)):
  if not os.path.isdir(tree):
    continue
  log = os.path.join(ROOT, 'gpurun_out', '%s_%s_%s.csv' % (tag, stage, name))
  env = dict(os.environ)
  extra = os.environ.get('AB_ENV_' + name.upper(), '')
  for kv in extra.split():
    k, v = kv.split('=', 1); env[k] = v
  subprocess.run(['ncu', '--metrics', 'gpu__time_duration.sum', '--clock-control', 'none', '--csv', '--log-file', log,
                  sys.executable, 'tests/prof_stage.py', stage, L], cwd=tree, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
  rows = [r for r in csv.reader(open(log)) if len(r) > 5]
  hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
  h = rows[hdr]; ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
  agg = collections.OrderedDict()
  for r in rows[hdr + 1:]:
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] in ('ns', 'nsecond') else (v * 1e3 if r[ui] in ('ms', 'msecond') else v)
    agg.setdefault(r[ki][:70], []).append(v)
  print('== %s %s L=%s' % (name, stage, L))
  for k, v in agg.items():
    tail = v[-3:]
    print('  %-70s n=%3d  last3 avg %8.1f us  min %8.1f us' % (k, len(v), sum(tail) / len(tail), min(v)))
