"""Probe: time the forward kernel of alternative library builds (FUSIONOCC_B200_LIB=...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fusionocc_b200.rig import SHAPES
dev = torch.device('cuda:0')
shape = SHAPES[os.environ.get('SHAPE', 'base')]
B = int(os.environ.get('BATCH', '8'))
vt, coor, depth, feat, og = bench.make_inputs(shape, B, 0, dev)
ns = bench.NativeStep(vt, coor, depth, feat, og)
def t(f, it=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
ns.rank_prepare(); ns.forward(); ns.bwd_plan_build()
print(os.environ.get('FUSIONOCC_B200_LIB', 'default'), 'rank %.1f fwd %.1f bwdplan %.1f bwd %.1f us' % (t(ns.rank_prepare), t(ns.forward), t(ns.bwd_plan_build), t(ns.backward)))
