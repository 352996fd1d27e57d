"""The C ABI from a plain C++ host (examples/host_demo.cpp): no Python, no torch in the process.  Built with nvcc
against include/fusionocc_b200.h and the in-tree library, run as a subprocess; it checks the whole step
(fo_view_transform_host: rank precompute, forward, backward) against a CPU evaluation of the definition."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_runs_the_whole_step(tmp_path):
    from fusionocc_b200 import build as fo_build
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.isfile(nvcc):
        pytest.skip('nvcc not available on this box')
    lib = fo_build.build()
    exe = str(tmp_path / 'host_demo')
    subprocess.run([nvcc, '-std=c++17', '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'examples', 'host_demo.cpp'),
                    '-L', os.path.dirname(lib), '-lfusionocc_b200', '-Xlinker', '-rpath', '-Xlinker', os.path.dirname(lib),
                    '-o', exe], check=True, capture_output=True, text=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and 'host_demo: OK' in res.stdout, res.stdout + res.stderr
