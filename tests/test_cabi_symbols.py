"""CPU: the C-ABI library builds, loads without a GPU, and exports every symbol include/*.h declares
(no compute calls).  Also: argument validation that needs no device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'fusionocc_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(fo_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from fusionocc_b200 import _cabi
    lib = _cabi.load()
    names = declared_symbols()
    assert len(names) >= 14, names
    for n in names:
        assert hasattr(lib, n), f'{n} is declared in include/fusionocc_b200.h but not exported'
    # and the Python binding table mirrors the header one to one
    assert sorted(_cabi.SIGNATURES) == names


def test_abi_version_and_build_info():
    from fusionocc_b200 import _cabi
    lib = _cabi.load()
    assert lib.fo_abi_version() == _cabi.ABI_VERSION
    info = lib.fo_build_info().decode()
    assert 'sm_100a' in info


def test_size_queries_are_pure_host_functions():
    from fusionocc_b200 import _cabi
    lib = _cabi.load()
    P, NV, rows = 371712 * 8, 640000 * 8, 4224 * 8
    assert lib.fo_fwd_plan_bytes(NV, P) > 8 * P
    assert lib.fo_rank_prepare_scratch_bytes(P, NV) > 4 * NV
    assert lib.fo_bwd_plan_bytes(P, rows) >= 16 * P
    assert lib.fo_bwd_scratch_bytes(1106717, 32, 0) >= 1106717 * 32 * 4
    assert lib.fo_bwd_scratch_bytes(1106717, 32, 1) == 256
    assert lib.fo_view_transform_host_workspace_bytes(8, 6, 88, 16, 44, 32, 200, 200, 16, 1) > 2 * NV * 32 * 4
    assert lib.fo_fwd_plan_bytes(-1, 0) == 0


def test_argument_errors_are_reported_without_a_device():
    from fusionocc_b200 import _cabi
    lib = _cabi.load()
    rc = lib.fo_bev_pool_v2_forward(None, 0, None, None, None, None, None, None, None, 0, 0, None, 1, 1, None, 0, 0,
                                    None, 0)
    assert rc == 1 and b'channels' in lib.fo_last_error()
    rc = lib.fo_rank_prepare(None, None, 0, 1, 1, 1, 1, _cabi.f3([0, 0, 0]), _cabi.f3([1, 1, 1]), 1, 1, 1, None, None,
                             None, None, None, None, None, 0, None, 0)
    assert rc == 1
    with pytest.raises(_cabi.FusionOccNativeError):
        _cabi.check(rc, 'fo_rank_prepare')


def test_product_does_not_import_the_oracle():
    """The shipped package must not route through oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, 'fusionocc_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, fn)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, os.path.join(dirpath, fn)


def test_binding_argument_counts_match_the_header():
    """Every ctypes signature has as many arguments as the prototype in include/fusionocc_b200.h (a drifted
    binding would corrupt the call silently)."""
    from fusionocc_b200 import _cabi
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = dict(re.findall(r'\b(fo_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S))
    assert sorted(protos) == sorted(_cabi.SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ('', 'void') else len([p for p in params.split(',') if p.strip()])
        assert n == len(_cabi.SIGNATURES[name][1]), f'{name}: header has {n} parameters, binding {len(_cabi.SIGNATURES[name][1])}'
