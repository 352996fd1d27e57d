"""ctypes binding of ``include/fusionocc_b200.h``.

This is the only place Python touches the native library.  There is no CPU or
pure-torch fallback: if ``libfusionocc_b200.so`` is missing and cannot be built
``load()`` raises, and every op raises on non-CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_int32, c_int64, c_size_t, c_void_p
from typing import Optional

from . import build as _build

FO_OK = 0
FO_LAYOUT_BCZYX = 0
FO_LAYOUT_BZYXC = 1
FO_FWD_ASSUME_SORTED = 1
FO_BWD_PLAN_STRUCTURED = 1
ABI_VERSION = 2

_ERR_NAMES = {1: 'FO_ERR_INVALID_ARG', 2: 'FO_ERR_CUDA', 3: 'FO_ERR_UNSUPPORTED', 4: 'FO_ERR_SCRATCH'}

_f3 = ctypes.c_float * 3

# name -> (restype, argtypes); mirrors include/fusionocc_b200.h one to one
SIGNATURES = {
    'fo_abi_version': (c_int, []),
    'fo_last_error': (c_char_p, []),
    'fo_build_info': (c_char_p, []),
    'fo_fwd_plan_bytes': (c_size_t, [c_int64, c_int64]),
    'fo_fwd_plan_build': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int32,
                                  c_int64, c_void_p, c_size_t]),
    'fo_bev_pool_v2_forward': (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int64, c_void_p, c_int32,
                                       c_int32, c_void_p, c_size_t]),
    'fo_bev_pool_v2_forward_slice': (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int64,
                                             c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_size_t]),
    'fo_bev_pool_v2_backward_slice': (c_int, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                                              c_void_p, c_int64, c_int64, c_int32, c_int64, c_int64, c_int64,
                                              c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p,
                                              c_size_t]),
    'fo_bwd_plan_bytes': (c_size_t, [c_int64, c_int64]),
    'fo_bwd_plan_build': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_int32,
                                  c_void_p, c_size_t, c_int32, c_int64, c_void_p, c_size_t]),
    'fo_bwd_scratch_bytes': (c_size_t, [c_int64, c_int32, c_int32]),
    'fo_bev_pool_v2_backward': (c_int, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int64,
                                        c_int32, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t,
                                        c_void_p, c_size_t, c_void_p, c_size_t]),
    'fo_bev_pool_v2_backward_with_plan': (c_int, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_int64,
                                                  c_void_p, c_int64, c_int32, c_int64, c_int64, c_int64, c_int32,
                                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p,
                                                  c_size_t]),
    'fo_rank_prepare_scratch_bytes': (c_size_t, [c_int64, c_int64]),
    'fo_rank_prepare': (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, _f3, _f3, c_int32,
                                c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p, c_size_t]),
    'fo_rank_prepare_calib': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32,
                                      c_int32, c_int32, c_int32, c_int32, _f3, _f3, c_int32, c_int32, c_int32,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_void_p, c_size_t]),
    'fo_rank_from_keys_scratch_bytes': (c_size_t, [c_int64, c_int64]),
    'fo_rank_from_keys': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_size_t]),
    'fo_lift_prepare_forward': (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p, c_void_p]),
    'fo_lift_prepare_backward': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                         c_int32, c_void_p, c_int32]),
    'fo_compat_bev_pool_v2': (None, [c_int, c_int] + [c_void_p] * 8),
    'fo_compat_bev_pool_v2_grad': (None, [c_int, c_int] + [c_void_p] * 10),
    'fo_view_transform_host_workspace_bytes': (c_size_t, [c_int32] * 10),
    'fo_view_transform_host_calib_workspace_bytes': (c_size_t, [c_int32] * 10),
    'fo_view_transform_host_calib': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p,
                                             c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _f3, _f3,
                                             c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_size_t, c_void_p]),
    'fo_view_transform_host': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                       c_int32, c_int32, c_int32, _f3, _f3, c_int32, c_int32, c_int32, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib: Optional[ctypes.CDLL] = None


class FusionOccNativeError(RuntimeError):
    pass


def lib_path() -> str:
    return os.environ.get('FUSIONOCC_B200_LIB', _build.LIB_PATH)


def _stale() -> bool:
    try:
        return _build.needs_build()
    except OSError:          # sources not shipped next to the library: nothing to compare with
        return False


def load() -> ctypes.CDLL:
    """Load (building in-tree if necessary) the native library; raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    # the in-tree library is rebuilt whenever a source or the header is newer than it (a stale .so would be called
    # with argument lists that no longer match its prototypes); FUSIONOCC_B200_LIB pins an explicit file instead
    stale = 'FUSIONOCC_B200_LIB' not in os.environ and os.path.isfile(path) and _stale()
    if not os.path.isfile(path) or stale:
        try:
            path = _build.build()
        except Exception as e:  # noqa: BLE001
            raise FusionOccNativeError(
                f'native library {path} is missing and could not be built: {e}. '
                'Run `python -m fusionocc_b200.build`. There is no CPU fallback.') from e
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise FusionOccNativeError(f'{path} does not export {name}; rebuild it') from e
        fn.restype = res
        fn.argtypes = args
    if lib.fo_abi_version() != ABI_VERSION:
        raise FusionOccNativeError(f'{path}: ABI version {lib.fo_abi_version()} != expected {ABI_VERSION}')
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != FO_OK:
        msg = load().fo_last_error().decode('utf-8', 'replace')
        raise FusionOccNativeError(f'{what} failed with {_ERR_NAMES.get(rc, rc)}: {msg}')


def f3(values) -> '_f3':
    v = [float(x) for x in values]
    if len(v) != 3:
        raise ValueError('expected 3 floats')
    return _f3(*v)
