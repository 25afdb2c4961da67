"""Build the C-ABI shared library (hand-written sm_100a CUDA) in-tree with nvcc.

    python -m diffndm_b200.build            # -> diffndm_b200/lib/libdiffndm_b200.so

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libdiffndm_b200.so')


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(PKG), 'include')):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
           '--use_fast_math' if os.environ.get('DNDM_FAST_MATH') else '-DDNDM_NO_FAST_MATH',
           '-Xcompiler', '-fPIC', '-shared', '-cudart', 'static',
           '-o', LIB_PATH, os.path.join(CSRC, 'engine.cu')]
    cmd[1:1] = os.environ.get('DNDM_EXTRA_NVCC_FLAGS', '').split()      # development switches (e.g. -DDNDM_EK_TRACE)
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed building libdiffndm_b200.so')
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
