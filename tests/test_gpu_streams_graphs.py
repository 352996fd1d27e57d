"""Stream and CUDA-graph behaviour of the C ABI.  The reference launches on the legacy default stream
(bev_pool_cuda.cu:127,136) — a latent bug under non-default streams.  Every entry point here takes
the stream explicitly, never synchronises the host and allocates nothing, so the whole step (rank precompute ->
forward -> backward plan -> backward) runs on a side stream and can be captured into a CUDA graph and replayed."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _step():
    import bench
    from fusionocc_b200.rig import SHAPES
    dev = torch.device('cuda:0')
    vt, coor, depth, feat, og = bench.make_inputs(SHAPES['small'], 2, 0, dev)
    return bench.NativeStep(vt, coor, depth, feat, og)


def _results(ns):
    nk, ni = (int(v) for v in ns.counts[:2].tolist())
    return [t.clone() for t in (ns.out, ns.dg, ns.fg, ns.rb[:nk], ns.rd[:nk], ns.rf[:nk], ns.st[:ni], ns.ln[:ni])]


def test_side_stream_and_graph_replay_match_the_default_stream():
    ns = _step()
    ns.step()
    torch.cuda.synchronize()
    want = _results(ns)

    def poison():
        for t in (ns.out, ns.dg, ns.fg):
            t.fill_(float('nan'))
        for t in (ns.rb, ns.rd, ns.rf, ns.st, ns.ln):
            t.fill_(-7)
        torch.cuda.synchronize()

    # (1) a side stream, with unrelated work on the default stream
    poison()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ns.step()
    busy = torch.empty(64 << 20, device='cuda').normal_()           # default stream keeps running meanwhile
    side.synchronize()
    for a, b in zip(_results(ns), want):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    del busy

    # (2) captured once, replayed three times
    poison()
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        ns.step()                                                    # warm-up on the capture stream
        cap.synchronize()
        with torch.cuda.graph(g, stream=cap):
            ns.step()
    for _ in range(3):
        poison()
        g.replay()
        torch.cuda.synchronize()
        for a, b in zip(_results(ns), want):
            assert torch.equal(a.view(torch.int32), b.view(torch.int32))
