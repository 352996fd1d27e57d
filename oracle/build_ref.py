"""ORACLE (test infrastructure) — compile the UNMODIFIED reference CUDA extensions.

Sources are compiled where they lie under ``/root/reference`` (never copied), with the flags
``setup.py:235-244`` uses (only ``-D__CUDA_NO_HALF_*``), for sm_100a (``TORCH_CUDA_ARCH_LIST=10.0a``), into
git-ignored directories under ``oracle/`` that travel to the GPU box with the snapshot:

``v2``    mmdet3d/ops/bev_pool_v2/src/{bev_pool.cpp, bev_pool_cuda.cu}            -> oracle/_ref/bev_pool_v2_ext*.so
          the pybind module exporting ``bev_pool_v2_forward`` / ``bev_pool_v2_backward`` (bev_pool.cpp:106-111)
``v1``    projects/BEVFusion/bevfusion/ops/bev_pool/src/{bev_pool.cpp, bev_pool_cuda.cu} -> oracle/_ref_v1/bev_pool_ext*.so
          the sibling op ``bev_pool_forward`` / ``bev_pool_backward``
``shim``  the reference's UNMODIFIED mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp + oracle/shim/bev_pool_shim.cpp
          (our three-line forwarding file) linked against fusionocc_b200/lib/libfusionocc_b200.so
          -> oracle/_shim/bev_pool_v2_ext_shim*.so: the L0 drop-in proof — the reference's own binding code
          running on the new kernels.

They are used only as checkers / comparison points on the GPU box (tests/test_gpu_vs_reference_ext.py,
tests/test_gpu_pool_v1.py, tests/test_gpu_l0_shim.py, bench.py's informational ``ref_cuda`` figure).

    python oracle/build_ref.py [--which v2|v1|shim|all] [--reference-root /root/reference] [--force]
"""
from __future__ import annotations

import argparse
import glob
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

TARGETS = {
    'v2': dict(out=os.path.join(HERE, '_ref'), name='bev_pool_v2_ext',
               src=('mmdet3d', 'ops', 'bev_pool_v2', 'src'), files=('bev_pool.cpp', 'bev_pool_cuda.cu')),
    'v1': dict(out=os.path.join(HERE, '_ref_v1'), name='bev_pool_ext',
               src=('projects', 'BEVFusion', 'bevfusion', 'ops', 'bev_pool', 'src'),
               files=('bev_pool.cpp', 'bev_pool_cuda.cu')),
    'shim': dict(out=os.path.join(HERE, '_shim'), name='bev_pool_v2_ext_shim',
                 src=('mmdet3d', 'ops', 'bev_pool_v2', 'src'), files=('bev_pool.cpp',)),
}
# backwards-compatible module attributes (the v2 extension)
OUT = TARGETS['v2']['out']
NAME = TARGETS['v2']['name']


def built_path(which: str = 'v2'):
    t = TARGETS[which]
    hits = sorted(glob.glob(os.path.join(t['out'], t['name'] + '*.so')))
    return hits[0] if hits else None


def build(reference_root: str = '/root/reference', force: bool = False, which: str = 'v2'):
    t = TARGETS[which]
    src = os.path.join(reference_root, *t['src'])
    sources = [os.path.join(src, f) for f in t['files']]
    for s in sources:
        if not os.path.isfile(s):
            raise FileNotFoundError(s)
    if built_path(which) and not force:
        return built_path(which)
    os.makedirs(t['out'], exist_ok=True)
    os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0a'
    from torch.utils.cpp_extension import load
    flags = ['-D__CUDA_NO_HALF_OPERATORS__', '-D__CUDA_NO_HALF_CONVERSIONS__', '-D__CUDA_NO_HALF2_OPERATORS__']
    kw = {}
    if which == 'shim':
        lib_dir = os.path.join(ROOT, 'fusionocc_b200', 'lib')
        if not os.path.isfile(os.path.join(lib_dir, 'libfusionocc_b200.so')):
            raise FileNotFoundError('build libfusionocc_b200.so first (python -m fusionocc_b200.build)')
        sources = sources + [os.path.join(HERE, 'shim', 'bev_pool_shim.cpp')]
        kw = dict(with_cuda=True,            # no .cu among the sources, but bev_pool.cpp includes the CUDA guard headers
                  extra_include_paths=[os.path.join(ROOT, 'include')],
                  extra_ldflags=[f'-L{lib_dir}', '-lfusionocc_b200', f'-Wl,-rpath,{lib_dir}'])
    load(name=t['name'], sources=sources, extra_cuda_cflags=flags, build_directory=t['out'], verbose=False,
         is_python_module=True, **kw)
    p = built_path(which)
    if not p:
        raise RuntimeError(f'reference extension build ({which}) produced no .so')
    return p


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference-root', default='/root/reference')
    ap.add_argument('--which', default='v2', choices=['v2', 'v1', 'shim', 'all'])
    ap.add_argument('--force', action='store_true')
    a = ap.parse_args()
    for w in (['v2', 'v1', 'shim'] if a.which == 'all' else [a.which]):
        print(w, build(a.reference_root, a.force, w))
