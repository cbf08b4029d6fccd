"""In-tree build of liblanczosnet_b200.so (hand-written CUDA for sm_100a, C ABI in include/).

    python -m lanczosnetwork_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared object lands next to this file so it ships
with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_NAME = 'liblanczosnet_b200.so'
LIB_PATH = os.path.join(HERE, LIB_NAME)
STAMP = LIB_PATH + '.stamp'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr',
]


def sources():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
  h = hashlib.sha256()
  files = sources() + sorted(
      os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh'))
  files.append(os.path.join(HERE, '..', 'include', 'lanczosnet_b200.h'))
  for f in files:
    with open(f, 'rb') as fh:
      h.update(f.encode())
      h.update(fh.read())
  h.update(' '.join(NVCC_FLAGS).encode())
  return h.hexdigest()


def nvcc_path():
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  return 'nvcc'


def build(force=False, verbose=False):
  """Compile if sources changed.  Returns the path of the shared object."""
  digest = _digest()
  if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
    with open(STAMP) as fh:
      if fh.read().strip() == digest:
        return LIB_PATH
  # one nvcc per translation unit, in parallel, then one link (a single nvcc invocation compiles its
  # inputs one after the other: 70 s for the 11 files; the slowest single file takes ~20 s)
  from concurrent.futures import ThreadPoolExecutor
  objdir = os.path.join(HERE, 'build')
  os.makedirs(objdir, exist_ok=True)
  compile_flags = [f for f in NVCC_FLAGS if f != '-shared']

  def compile_one(src):
    obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
    cmd = [nvcc_path()] + compile_flags + ['-c', '-o', obj, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return obj, ' '.join(cmd) + '\n' + proc.stdout + proc.stderr, proc.returncode

  with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
    results = list(pool.map(compile_one, sources()))
  log = ''.join(r[1] for r in results)
  rc = max(r[2] for r in results)
  if rc == 0:
    link = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC',
            '-o', LIB_PATH] + [r[0] for r in results]
    proc = subprocess.run(link, capture_output=True, text=True)
    log += ' '.join(link) + '\n' + proc.stdout + proc.stderr
    rc = proc.returncode
  with open(os.path.join(HERE, 'build.log'), 'w') as fh:
    fh.write(log)
  if rc != 0:
    raise RuntimeError('nvcc failed:\n' + log[-6000:])
  if verbose:
    print(log)
  with open(STAMP, 'w') as fh:
    fh.write(digest)
  return LIB_PATH


if __name__ == '__main__':
  path = build(force='--force' in sys.argv, verbose=True)
  print('built', path)
