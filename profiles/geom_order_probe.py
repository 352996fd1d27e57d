"""Probe: which fp32 summation order reproduces torch's get_lidar_coor on this GPU, stage by stage."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fusionocc_b200 import LSSViewTransformer
from fusionocc_b200.rig import SHAPES, make_calibration
from fusionocc_b200.view_transformer import pack_calibration, rank_prepare_calib
DEV = 'cuda:0'
sh = SHAPES['base']; B = 2
vt = LSSViewTransformer(sh.grid_cfg(), sh.input_size, sh.downsample, in_channels=8, out_channels=sh.channels, collapse_z=False)
def run(tag, mod):
    cal = [c.clone() for c in make_calibration(sh, B)]
    mod(cal)
    cal = [c.to(DEV) for c in cal]
    want = vt.get_lidar_coor(*cal).contiguous()
    s2e, _, k, pr, pt, bda = cal
    cam, bda12, has_t = pack_calibration(s2e, k, pr, pt, bda)
    res = []
    for mode in range(int(os.environ.get('NMODES', '4'))):
        out = rank_prepare_calib(vt._frustum_on(s2e), cam, bda12, has_t, B, s2e.shape[1], vt.grid_lower_bound.tolist(),
                                 vt.grid_interval.tolist(), vt._grid_xyz(), matvec_mode=mode, return_coor=True)
        res.append(int((out[7].view(torch.int32) != want.view(torch.int32)).sum()))
    print(f'{tag:40s} differing floats per mode: {res} of {want.numel()}')
def rot2(a):
    return torch.tensor([[math.cos(a), -math.sin(a), 0.], [math.sin(a), math.cos(a), 0.], [0., 0., 1.]])
run('rig', lambda c: None)
def bda_rot(c): c[5][:] = (rot2(0.2) * 1.05)
run('bda 3x3 rotation*1.05', bda_rot)
def bda_t(c):
    b4 = torch.eye(4).repeat(B, 1, 1); b4[:, :3, 3] = torch.tensor([0.3, -0.2, 0.1]); c[5] = b4
run('bda 4x4 identity + translation', bda_t)
def bda_rt(c):
    b4 = torch.eye(4).repeat(B, 1, 1); b4[:, :3, :3] = rot2(0.2) * 1.05; b4[:, :3, 3] = torch.tensor([0.3, -0.2, 0.1]); c[5] = b4
run('bda 4x4 rotation + translation', bda_rt)
def pr_rot(c):
    c[3][:] = c[3] @ rot2(0.05)
run('post_rots rotated', pr_rot)
def flip(c): c[5][:] = torch.diag(torch.tensor([-1., 1., 1.]))
run('bda flip x', flip)
