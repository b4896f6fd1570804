"""Writes profiles/<round>_sass_summary.txt (round prefix = first argument, default r2) from the built library (no GPU needed): per kernel, registers / shared memory
(`cuobjdump -res-usage`) and counts of the SASS mnemonics that show which hardware path it uses
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
LDGSTS = cp.async, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = cp.async.bulk, HMMA = mma.sync, MUFU.EX2, FFMA2 = packed fp32).

    python profiles/make_sass_summary.py r2
"""
import collections
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), 'trax_b200', 'liblsh_attn_b200.so')
PATTERNS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'SYNCS', 'LDGSTS', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'MUFU.EX2', 'FFMA2', 'FADD2', 'FMUL2',
            'ELECT', 'NANOSLEEP', 'RED', 'ATOM', 'BAR.SYNC', 'STL', 'LDL']


def demangle(name):
  try:
    return subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip() or name
  except OSError:
    return name


def main():
  sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
  res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True, check=True).stdout
  usage = {}
  cur = None
  for line in res.splitlines():
    m = re.search(r'Function (\S+):', line)
    if m:
      cur = m.group(1)
      continue
    m = re.search(r'REG:(\d+).*?SHARED:(\d+)', line)
    if m and cur:
      usage[cur] = (int(m.group(1)), int(m.group(2)))
      cur = None
  counts, size, cur = {}, collections.Counter(), None
  for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
      cur = m.group(1)
      counts[cur] = collections.Counter()
      continue
    if cur is None:
      continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if not m:
      continue
    size[cur] += 1
    op = m.group(1)
    for p in PATTERNS:
      if op.startswith(p):
        counts[cur][p] += 1
  lines = ['SASS summary of trax_b200/liblsh_attn_b200.so (sm_100a), lsh:: kernels only; static instruction counts',
           '%-58s %5s %7s %6s  %s' % ('kernel', 'regs', 'st.smem', 'insts', 'mnemonic counts (dynamic shared memory is set at launch and not listed)')]
  for fn in sorted(counts, key=lambda f: -size[f]):
    name = demangle(fn)
    if 'lsh::' not in name:
      continue
    short = re.sub(r'\(.*', '', name).replace('void ', '')
    regs, smem = usage.get(fn, (0, 0))
    c = ' '.join('%s=%d' % (p, counts[fn][p]) for p in PATTERNS if counts[fn][p])
    lines.append('%-58s %5d %7d %6d  %s' % (short[:58], regs, smem, size[fn], c))
  import sys
  out = os.path.join(HERE, '%s_sass_summary.txt' % (sys.argv[1] if len(sys.argv) > 1 else 'r2'))
  open(out, 'w').write('\n'.join(lines) + '\n')
  print('\n'.join(lines))


if __name__ == '__main__':
  main()
