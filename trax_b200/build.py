"""Builds trax_b200/liblsh_attn_b200.so from trax_b200/csrc/*.cu with nvcc for sm_100a (in-tree)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'liblsh_attn_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--use_fast_math=false']
for _d in os.environ.get('LSH_EXTRA_DEFS', '').split():
  FLAGS.append('-D' + _d)             # experiment knobs (kernel bring-up only)
if os.environ.get('LSH_LIB_OUT'):
  LIB = os.environ['LSH_LIB_OUT']
if os.environ.get('LSH_DEBUG_SPIN'):
  FLAGS.append('-DLSH_DEBUG_SPIN')   # barrier waits trap instead of hanging (kernel bring-up)


def _cublas_dirs():
  """cuBLAS: prefer the copy torch ships (it is the one already loaded in-process), else the toolkit's."""
  dirs = []
  try:
    import nvidia.cublas  # type: ignore
    for p in nvidia.cublas.__path__:
      dirs.append(os.path.join(p, 'lib'))
  except Exception:  # pylint: disable=broad-except
    pass
  dirs.append('/usr/local/cuda/lib64')
  return [d for d in dirs if os.path.isdir(d)]


def sources():
  return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + [
      os.path.join(HERE, '..', 'include', 'lsh_attn.h')]
  return any(os.path.getmtime(s) > t for s in deps)


def build(force=False, verbose=False):
  if not force and not needs_build():
    return LIB
  objs = []
  odir = os.path.join(HERE, 'build' if not os.environ.get('LSH_LIB_OUT') else os.path.join('build', os.path.basename(LIB)))
  os.makedirs(odir, exist_ok=True)
  procs = []
  for src in sources():
    obj = os.path.join(odir, os.path.basename(src)[:-3] + '.o')
    cmd = [NVCC] + ARCH + [f for f in FLAGS if not f.startswith('--use_fast_math')] + [
        '-Xptxas', '-v' if verbose else '-warn-spills', '-c', src, '-o', obj]
    procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs.append(obj)
  failed = False
  for src, p in procs:
    out, _ = p.communicate()
    if verbose or p.returncode != 0:
      sys.stderr.write('--- %s\n%s\n' % (os.path.basename(src), out))
    failed |= p.returncode != 0
  if failed:
    raise RuntimeError('nvcc failed')
  libdirs = _cublas_dirs()
  link = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs
  for d in libdirs:
    link += ['-L' + d, '-Xlinker', '-rpath', '-Xlinker', d]
  # torch's wheel ships libcublas.so.12 without the dev symlink; link by file name when needed
  cublas = None
  for d in libdirs:
    for name in ('libcublas.so', 'libcublas.so.12'):
      if os.path.exists(os.path.join(d, name)):
        cublas = ('-lcublas' if name == 'libcublas.so' else '-l:libcublas.so.12')
        break
    if cublas:
      break
  link += [cublas or '-lcublas', '-lcudart']
  subprocess.check_call(link)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
