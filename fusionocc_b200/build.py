"""Build the C-ABI CUDA library in-tree (``fusionocc_b200/lib/libfusionocc_b200.so``).

sm_100a only, plain ``nvcc -shared`` (no torch headers: the boundary is a C ABI,
see ``include/fusionocc_b200.h``).  nvcc cross-compiles without a GPU, so this
also is the "does it build" check of ``__graft_entry__.build()``.

    python -m fusionocc_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libfusionocc_b200.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

SOURCES = ['cabi.cu', 'rank_prepare.cu', 'bev_pool_fwd.cu', 'bev_pool_bwd.cu', 'lift_prepare.cu']
HEADERS = ['common.cuh', 'bucket_sort.cuh']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-shared',
    # no --use_fast_math: the voxel index needs IEEE div.rn / sub.rn
]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, 'fusionocc_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB_PATH + '.tmp'
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
          ['-I', INCLUDE, '-o', tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libfusionocc_b200.so (see stderr)')
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
