"""Probe (BASELINE.json configs[3]): training shape — batch 8 per GPU, two temporal frames with separate rank sets
(fusion_occ.py:289-316: the adjacent frame runs under no_grad, the key frame fwd+bwd), frames concatenated along
channels (fusion_occ.py:326).  Compares (a) the drop-in op + torch.cat with (b) bev_pool_v2_cat, which writes
both frames into one (B, 2C, Z, Y, X) tensor and reads the key frame's gradient slice in place."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusionocc_b200 import LSSViewTransformer, bev_pool_v2, bev_pool_v2_cat
from fusionocc_b200.rig import SHAPES, make_calibration, make_values
dev = torch.device('cuda:0')
sh = SHAPES['base']; B = 8
vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels, collapse_z=False)
X, Y, Z = vt._grid_xyz(); C = sh.channels
frames = []
for shift in (False, True):                      # key frame, adjacent frame (sensor2keyego shifted)
    cal = [c.to(dev) for c in make_calibration(sh, B, frame_shift=shift)]
    rb, rd, rf, st, ln = vt.voxel_pooling_prepare_v2(vt.get_lidar_coor(*cal))
    d, f = make_values(sh, B)
    frames.append([d.to(dev), f.to(dev).permute(0, 1, 3, 4, 2).contiguous(), rd, rf, rb, st, ln])
og = torch.randn(B, 2 * C, Z, Y, X, device=dev)
shape = (B, Z, Y, X, C)
def run_cat_torch():
    d0 = frames[0][0].clone().requires_grad_(); f0 = frames[0][1].clone().requires_grad_()
    a = bev_pool_v2(d0, f0, *frames[0][2:5], shape, *frames[0][5:])
    with torch.no_grad():
        b = bev_pool_v2(*frames[1][:5], shape, *frames[1][5:])
    torch.cat([a, b], dim=1).backward(og)
    return d0.grad, f0.grad
def run_cat_native():
    d0 = frames[0][0].clone().requires_grad_(); f0 = frames[0][1].clone().requires_grad_()
    out = bev_pool_v2_cat([(d0, f0, *frames[0][2:]), tuple(frames[1])], shape)
    out.backward(og)
    return d0.grad, f0.grad
def t(f, it=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
ga, gb = run_cat_torch(), run_cat_native()
assert all(torch.equal(x.view(torch.int32), y.view(torch.int32)) for x, y in zip(ga, gb)), 'gradients differ'
res = {'shape': 'base', 'batch': B, 'frames': 2, 'op_plus_torch_cat_ms': t(run_cat_torch), 'bev_pool_v2_cat_ms': t(run_cat_native),
       'note': 'two forwards (cached plans) + one backward + concatenation; ranks static across iterations (accelerate-style)'}
print(json.dumps(res))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/c4_training_probe.json', 'w'), indent=1)
