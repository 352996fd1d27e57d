"""ORACLE (test infrastructure) — compile the UNMODIFIED reference CUDA extension.

Sources are compiled where they lie under ``/root/reference`` (never copied):
    mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp
    mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu
with the flags ``setup.py:235-244`` uses (only ``-D__CUDA_NO_HALF_*``), for sm_100a
(``TORCH_CUDA_ARCH_LIST=10.0a``), into ``oracle/_ref/`` (git-ignored, travels to the GPU
box with the snapshot).  The result is the pybind module ``bev_pool_v2_ext`` exporting
``bev_pool_v2_forward`` / ``bev_pool_v2_backward`` (bev_pool.cpp:106-111).

It is used only as a checker / comparison point on the GPU box
(tests/test_gpu_vs_reference_ext.py, bench.py's informational ``ref_cuda`` figure).

    python oracle/build_ref.py [--reference-root /root/reference]
"""
from __future__ import annotations

import argparse
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
NAME = 'bev_pool_v2_ext'


def built_path():
    hits = sorted(glob.glob(os.path.join(OUT, NAME + '*.so')))
    return hits[0] if hits else None


def build(reference_root: str = '/root/reference', force: bool = False):
    src = os.path.join(reference_root, 'mmdet3d', 'ops', 'bev_pool_v2', 'src')
    sources = [os.path.join(src, 'bev_pool.cpp'), os.path.join(src, 'bev_pool_cuda.cu')]
    for s in sources:
        if not os.path.isfile(s):
            raise FileNotFoundError(s)
    if built_path() and not force:
        return built_path()
    os.makedirs(OUT, exist_ok=True)
    os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0a'
    from torch.utils.cpp_extension import load
    flags = ['-D__CUDA_NO_HALF_OPERATORS__', '-D__CUDA_NO_HALF_CONVERSIONS__', '-D__CUDA_NO_HALF2_OPERATORS__']
    load(name=NAME, sources=sources, extra_cuda_cflags=flags, build_directory=OUT, verbose=False,
         is_python_module=True)
    p = built_path()
    if not p:
        raise RuntimeError('reference extension build produced no .so')
    return p


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference-root', default='/root/reference')
    ap.add_argument('--force', action='store_true')
    a = ap.parse_args()
    print(build(a.reference_root, a.force))
