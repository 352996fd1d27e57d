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
# FUSIONOCC_LIB_SUFFIX builds a side-by-side variant (A/B measurements with FUSIONOCC_NVCC_EXTRA=-DFO_...=..)
_SUFFIX = os.environ.get('FUSIONOCC_LIB_SUFFIX', '')
LIB_PATH = os.path.join(LIB_DIR, f'libfusionocc_b200{_SUFFIX}.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

SOURCES = ['cabi.cu', 'rank_prepare.cu', 'bev_pool_fwd.cu', 'bev_pool_bwd.cu', 'lift_prepare.cu']
HEADERS = ['common.cuh', 'bucket_sort.cuh', 'tma.cuh', 'rank_chunk.cuh', 'bwd_plan.cuh', 'rank_fast.cuh']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC',
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
    obj_dir = os.path.join(LIB_DIR, 'obj' + _SUFFIX)
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    extra = (['-Xptxas', '-v'] if verbose else []) + os.environ.get('FUSIONOCC_NVCC_EXTRA', '').split()

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + extra + ['-I', INCLUDE, '-c', os.path.join(CSRC, src), '-o', obj]
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    # one nvcc per translation unit, in parallel (the five files are independent; ~15 s instead of ~40 s)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for (obj, res), src in zip(results, SOURCES):
        if verbose or res.returncode != 0:
            sys.stderr.write(f'--- {src}\n' + res.stdout + res.stderr)
    if any(res.returncode != 0 for _, res in results):
        raise RuntimeError('nvcc failed building libfusionocc_b200.so (see stderr)')
    tmp = LIB_PATH + '.tmp'
    res = subprocess.run([nvcc, '-shared', '-Xlinker', '-soname=libfusionocc_b200.so', '-o', tmp] +
                         [obj for obj, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('linking libfusionocc_b200.so failed (see stderr)')
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
