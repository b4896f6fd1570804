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


def source_hash():
  """Content hash of everything the library is compiled from (+ the flags): embedded in the library
  (`lsh_attn_source_hash()`) and stored beside it, so that a stale build is detected by content, not by mtime."""
  import hashlib
  h = hashlib.sha256()
  deps = sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + [os.path.join(HERE, '..', 'include', 'lsh_attn.h')]
  for path in deps:
    h.update(os.path.basename(path).encode())
    with open(path, 'rb') as f:
      h.update(f.read())
  h.update(' '.join(ARCH + FLAGS).encode())
  return h.hexdigest()[:16]


def built_hash():
  try:
    with open(LIB + '.hash') as f:
      return f.read().strip()
  except OSError:
    return None


def needs_build():
  return not os.path.exists(LIB) or built_hash() != source_hash()


def build(force=False, verbose=False):
  if not force and not needs_build():
    return LIB
  objs = []
  src_hash = source_hash()
  odir = os.path.join(HERE, 'build' if not os.environ.get('LSH_LIB_OUT') else os.path.join('build', os.path.basename(LIB)))
  os.makedirs(odir, exist_ok=True)
  procs = []
  for src in sources():
    obj = os.path.join(odir, os.path.basename(src)[:-3] + '.o')
    cmd = [NVCC] + ARCH + [f for f in FLAGS if not f.startswith('--use_fast_math')] + [
        '-Xptxas', '-v' if verbose else '-warn-spills', '-DLSH_SRC_HASH="%s"' % src_hash, '-c', src, '-o', obj]
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
  with open(LIB + '.hash', 'w') as f:
    f.write(src_hash + '\n')
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
