"""pytest configuration: the ``gpu`` marker and shared helpers.

``-m "not gpu"`` runs here on CPU (oracle vs golden vectors, host logic, C-ABI symbol table, gloo
sharding); ``-m gpu`` runs on a B200 and calls the CUDA path through the C ABI.  Nothing under
``tests/`` reads /root/reference at run time except the optional live cross-checks, which skip
when it is absent.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container (runs under gpurun)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
