"""Probe: run the rank precompute a few times (for ncu captures of its kernels)."""
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
for _ in range(int(os.environ.get('ITERS', '4'))):
    ns.rank_prepare()
torch.cuda.synchronize()
print('ok', ns.counts.tolist())
