"""Summarises an `ncu --metrics gpu__time_duration.sum` launch list (profiles/r1_launches.csv) per kernel for
ONE bench step (from one hash_kernel launch to the next).  Times are cold-cache and serialised under ncu:
compare SHARES, not absolutes (B200_PROFILING.md).  Usage: python profiles/summarize_launches.py [csv]"""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else 'profiles/r1_launches.csv'
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = [r for r in rows if 'Kernel Name' in r][0]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if r is not hdr and r[ci['Metric Name']] == 'gpu__time_duration.sum']
names = [re.sub(r'\(.*', '', r[ci['Kernel Name']])[:64] for r in data]
vals = [float(r[ci['Metric Value']].replace(',', '')) / 1e3 for r in data]   # ns -> us
h = [i for i, n in enumerate(names) if 'hash_' in n and 'kernel' in n]
a, b = h[3] - 7, h[4] - 7            # one full step (round 2: a forward call starts 7 launches before its hash kernel; round 1: 5)
agg, cnt = collections.OrderedDict(), collections.Counter()
for n, v in zip(names[a:b], vals[a:b]):
  agg[n] = agg.get(n, 0.0) + v
  cnt[n] += 1
tot = sum(agg.values())
print('one fwd+bwd step of bench.py (workload c2): %d launches, %.1f us under ncu' % (b - a, tot))
for n, v in sorted(agg.items(), key=lambda x: -x[1]):
  print('%9.1f us %5.1f%%  x%d  %s' % (v, 100 * v / tot, cnt[n], n))
