"""Probe: run-to-run spread of the e2e leg (host buffers through fo_view_transform_host) inside one process."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fusionocc_b200.rig import SHAPES
dev = torch.device('cuda:0')
vt, coor, depth, feat, og = bench.make_inputs(SHAPES['base'], 8, 0, dev)
ns = bench.NativeStep(vt, coor, depth, feat, og)
for chunks in (int(c) for c in os.environ.get('CHUNKS', '4').split(',')):
    for two in (1, 2, 0):
        hs = bench.HostStep(ns, n_chunks=chunks, two_streams=bool(two))
        if two == 2:                                     # one shared upload stream for all chunks
            hs.up_streams = [hs.up_streams[0]] * len(hs.up_streams)
        for _ in range(2): hs.step()
        torch.cuda.synchronize()
        res = []
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): hs.step()
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 10)
        print(f'chunks={chunks} two_streams={two}: ms/step ' + ' '.join(f'{r:.2f}' for r in res) + f'  -> {8/min(res)*1e3:.0f} samples/s best')
        del hs
