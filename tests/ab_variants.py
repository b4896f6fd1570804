"""A/B helper (not a test): ncu device time of one kernel for every library under trax_b200/variants/ (built with
LSH_LIB_OUT / LSH_EXTRA_DEFS), same box, same inputs.   usage: python tests/ab_variants.py <stage> <kernel substring> [L]"""
import collections, csv, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
stage = sys.argv[1] if len(sys.argv) > 1 else 'attend_bwd'
pat = sys.argv[2] if len(sys.argv) > 2 else 'attend_bwd_tc'
L = sys.argv[3] if len(sys.argv) > 3 else '65536'
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
libs = sorted(glob.glob(os.path.join(ROOT, 'trax_b200', 'variants', '*.so')))
for rep in range(2):
  for lib in libs:
    name = os.path.basename(lib)[4:-3]
    log = os.path.join(ROOT, 'gpurun_out', 'var_%s_%s.csv' % (stage, name))
    env = dict(os.environ, LSH_ATTN_LIB=lib)
    subprocess.run(['ncu', '--metrics', 'gpu__time_duration.sum', '--clock-control', 'none', '--csv', '--log-file', log,
                    sys.executable, 'tests/prof_stage.py', stage, L], cwd=ROOT, env=env, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    rows = [r for r in csv.reader(open(log)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
    h = rows[hdr]; ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    vals = []
    for r in rows[hdr + 1:]:
      if pat in r[ki]:
        v = float(r[vi].replace(',', ''))
        vals.append(v / 1e3 if r[ui] in ('ns', 'nsecond') else v)
    print('%-12s rep %d  %s' % (name, rep, ' '.join('%.1f' % v for v in vals)), flush=True)
