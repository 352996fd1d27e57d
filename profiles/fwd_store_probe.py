"""Probe: how fast is the forward's STORE PATTERN alone?  All frustum points are moved outside the
grid, so every sub-tile takes the zero-streaming path (128-byte lines, 32 channel planes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fusionocc_b200.rig import SHAPES
dev = torch.device('cuda:0')
shape = SHAPES['base']
B = 8
vt, coor, depth, feat, og = bench.make_inputs(shape, B, 0, dev)
for label, c in (('all points outside (pure zero stream)', torch.full_like(coor, 1e6)), ('real geometry', coor)):
    ns = bench.NativeStep(vt, c.contiguous(), depth, feat, og)
    ns.rank_prepare()
    for _ in range(3): ns.forward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ns.forward()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f'{label}: {us:.1f} us  -> {ns.out.numel()*4/us/1e3:.0f} GB/s of output')
    del ns
x = torch.empty(B*32*640000, device=dev)
for _ in range(3): x.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): x.zero_()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1)/20*1e3
print(f'cudaMemset same size: {us:.1f} us -> {x.numel()*4/us/1e3:.0f} GB/s')
