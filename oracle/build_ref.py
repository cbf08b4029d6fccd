"""Compile the REFERENCE's own CUDA kernels for the segment-reduction op from the sources where
they lie under /root/reference (operators/src/cuda/segment_reduction.cu -- plain CUDA C with
extern "C" launchers, no torch dependency) into oracle/_ref/libsegment_reduction_ref.so.

TEST INFRASTRUCTURE: used only by tests/ as a second checker next to the numpy restatement
(oracle/segment_oracle.py).  Nothing is copied into the repo; oracle/_ref/ is git-ignored but
travels to the GPU box.  The reference's host-side C++ (segment_reduction.cpp /
segment_reduction_cuda.cpp) needs THC headers that modern torch removed -> unbuildable, see
DESIGN.md.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/operators/src/cuda/segment_reduction.cu'
OUT_DIR = os.path.join(HERE, '_ref')
OUT = os.path.join(OUT_DIR, 'libsegment_reduction_ref.so')


def build(force=False):
  if not os.path.exists(REF_SRC):
    return OUT if os.path.exists(OUT) else None      # GPU box: use the prebuilt file
  if os.path.exists(OUT) and not force and os.path.getmtime(OUT) >= os.path.getmtime(REF_SRC):
    return OUT
  os.makedirs(OUT_DIR, exist_ok=True)
  cmd = ['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O2',
         '-shared', '-Xcompiler', '-fPIC', '-I', os.path.dirname(REF_SRC), '-o', OUT, REF_SRC]
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError('reference kernel build failed:\n' + proc.stdout + proc.stderr)
  return OUT


if __name__ == '__main__':
  print(build(force=True))
