"""Import the reference's own view transformer *in this container only*.

Used by ``make_golden.py`` (fixture generation) and by the optional
``test_oracle_vs_reference_live`` tests, both of which skip when
``/root/reference`` is absent (it never exists on the GPU box).

The reference file ``projects/FusionOcc/fusionocc/necks/view_transformer.py``
imports mmcv / mmengine / mmdet / the compiled bev_pool_v2 extension at module
scope (:5-13).  None of those are installed here, and none of them take part in
the functions we need (``create_grid_infos``, ``create_frustum``,
``get_lidar_coor``, ``voxel_pooling_prepare_v2`` are pure torch).  We therefore
register inert stub modules for exactly those imports and load the file by
path, which also avoids executing ``fusionocc/__init__.py`` (it shells out to
pip at import time).  Nothing from the reference is copied.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('FUSIONOCC_REFERENCE_ROOT', '/root/reference')
_VT_PATHS = {
    'fusionocc': 'projects/FusionOcc/fusionocc/necks/view_transformer.py',
    'mmdet3d': 'mmdet3d/models/necks/view_transformer.py',
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _VT_PATHS['fusionocc']))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []          # behave like a package so sub-imports resolve
    return m


class _Registry:
    def register_module(self, *a, **k):
        def deco(cls):
            return cls
        return deco


def load_reference_view_transformer(which: str = 'fusionocc', bev_pool_v2=None):
    """Returns the reference module object (classes usable on CPU)."""
    import contextlib

    import torch.nn as nn

    def _not_available(*a, **k):
        raise RuntimeError('stubbed in golden harness')

    class _autocast(contextlib.ContextDecorator):      # used both as `with` and as decorator
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    stubs = {
        'mmcv': _stub('mmcv'),
        'mmcv.cnn': _stub('mmcv.cnn', build_conv_layer=_not_available),
        'mmcv.runner': _stub('mmcv.runner', BaseModule=nn.Module, force_fp32=lambda *a, **k: (lambda f: f)),
        'mmengine': _stub('mmengine'),
        'mmengine.model': _stub('mmengine.model', BaseModule=nn.Module),
        'mmengine.runner': _stub('mmengine.runner', autocast=_autocast),
        'mmdet': _stub('mmdet'),
        'mmdet.models': _stub('mmdet.models'),
        'mmdet.models.backbones': _stub('mmdet.models.backbones'),
        'mmdet.models.backbones.resnet': _stub('mmdet.models.backbones.resnet', BasicBlock=nn.Module),
        'mmdet3d': _stub('mmdet3d'),
        'mmdet3d.ops': _stub('mmdet3d.ops'),
        'mmdet3d.ops.bev_pool_v2': _stub('mmdet3d.ops.bev_pool_v2'),
        'mmdet3d.ops.bev_pool_v2.bev_pool': _stub('mmdet3d.ops.bev_pool_v2.bev_pool',
                                                  bev_pool_v2=bev_pool_v2 or _not_available),
        'mmdet3d.registry': _stub('mmdet3d.registry', MODELS=_Registry()),
        'mmdet3d.models': _stub('mmdet3d.models'),
        'mmdet3d.models.builder': _stub('mmdet3d.models.builder', NECKS=_Registry()),
    }
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        path = os.path.join(REFERENCE_ROOT, _VT_PATHS[which])
        spec = importlib.util.spec_from_file_location(f'_ref_view_transformer_{which}', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
