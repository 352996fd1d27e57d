"""The library's A/B switches must not change a single bit.

Every hot-path variant that ships behind an environment switch — programmatic dependent launch on/off
(``FO_PDL``), the round-1 scan/order passes vs the staged scan + single order launch (``FO_RANK_FAST``
0/1/2), the backward plan inside / in front of / outside the gather launch (``FO_BWD_RIDE``), the gather's
occupancy masks + half boxes vs the iv_vox path (``FO_BWD_HALF``) — runs the whole chain
(rank precompute from the calibration -> forward -> backward incl. its plan) through the C ABI on the same
inputs; rank arrays are compared with the oracle (reference: view_transformer.py:223-281), output and
gradients bit for bit with the default configuration, which tests/test_gpu_parity.py and
tests/test_gpu_vs_reference_ext.py pin against the oracle and the unmodified reference extension.
"""
import os

import numpy as np
import pytest
import torch

from tests.helpers import rig_case

pytestmark = pytest.mark.gpu

SWITCHES = [
    {},
    {'FO_PDL': '0'},
    {'FO_PDL': '5'},
    {'FO_RANK_FAST': '0'},
    {'FO_RANK_FAST': '2'},
    {'FO_RANK_FAST': '2', 'FO_BWD_RIDE': '0'},
    {'FO_BWD_RIDE': '0'},
    {'FO_BWD_RIDE': '1'},
    {'FO_BWD_RIDE': '2'},
    {'FO_BWD_HALF': '0'},
    {'FO_VOX_IMPL': '0'},
    {'FO_RANK_MID': '0'},
    {'FO_RANK_MID': '1'},
]
NAMES = sorted({k for s in SWITCHES for k in s})


def _run(ns):
    ns.step()
    torch.cuda.synchronize()
    nk, ni = (int(v) for v in ns.counts[:2].tolist())
    return (nk, ni, [x.clone() for x in (ns.rb[:nk], ns.rd[:nk], ns.rf[:nk], ns.st[:ni], ns.ln[:ni], ns.out, ns.dg, ns.fg)])


@pytest.mark.parametrize('shape,B', [('small', 2), ('tiny', 3), ('base', 2)])
def test_switches_are_bit_neutral(shape, B):
    from bench import NativeStep, make_inputs
    from fusionocc_b200.rig import SHAPES
    dev = torch.device('cuda:0')
    saved = {k: os.environ.get(k) for k in NAMES}
    try:
        for k in NAMES:
            os.environ.pop(k, None)
        vt, coor, depth, feat, og = make_inputs(SHAPES[shape], B, 0, dev, with_coor=False)
        ns = NativeStep(vt, coor, depth, feat, og)
        ref = None
        for sw in SWITCHES:
            for k in NAMES:
                os.environ.pop(k, None)
            os.environ.update(sw)
            for t in (ns.rb, ns.rd, ns.rf, ns.st, ns.ln):
                t.fill_(-7)
            ns.out.fill_(-1.0); ns.dg.fill_(-1.0); ns.fg.fill_(-1.0)
            got = _run(ns)
            if ref is None:
                ref = got
                # the default configuration against the oracle's ranks (int64-exact restatement of the reference)
                c = rig_case(shape, B)
                rb, rd, rf, st, ln = c['ranks']
                assert got[0] == rb.shape[0] and got[1] == st.shape[0]
                for name, a, w in zip(('ranks_bev', 'ranks_depth', 'ranks_feat', 'interval_starts', 'interval_lengths'),
                                      got[2][:5], (rb, rd, rf, st, ln)):
                    assert np.array_equal(a.cpu().numpy(), w), f'{name} differs from the oracle'
                continue
            assert got[0] == ref[0] and got[1] == ref[1], f'{sw}: counts differ'
            for name, a, w in zip(('ranks_bev', 'ranks_depth', 'ranks_feat', 'interval_starts', 'interval_lengths', 'out',
                                   'depth_grad', 'feat_grad'), got[2], ref[2]):
                assert torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a,
                                   w.view(torch.int32) if w.dtype == torch.float32 else w), f'{sw}: {name} differs'
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
