"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S
OWN CODE in this container (CPU, torch), so the oracle is pinned against the
reference and not against itself.

    python tests/golden/make_golden.py

Needs /root/reference (present only in the build container); the produced
*.npz / *.json files are committed and are what travels to the GPU box.

What is pinned
  * a1 create_grid_infos, a2 create_frustum, a3 get_lidar_coor,
    a4 voxel_pooling_prepare_v2 of
    /root/reference/projects/FusionOcc/fusionocc/necks/view_transformer.py
    (functions at :87-103, :105-133, :135-173, :223-281).
  * the op's only known-answer test, mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176
    (values restated as data, the CUDA op itself cannot run here).

Tie order.  The reference calls ``ranks_bev.argsort()`` (view_transformer.py:266),
whose tie order is implementation-defined: on CUDA it is the (stable) radix sort,
on this container's CPU it is NOT stable (about half of all ties come back
inverted — recorded below as ``tie_inversions``).  Fixtures therefore store the
reference output both raw and *tie-canonicalised* (ties put in ascending
``ranks_depth``, which is what the stable device sort yields); ranks_bev,
interval_starts and interval_lengths are unaffected by tie order and are
compared raw.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fusionocc_b200.rig import SHAPES, Shape, make_calibration  # noqa: E402
from tests.golden._ref_import import load_reference_view_transformer, reference_available  # noqa: E402


def canon(rb, rd, rf):
    """Ties (equal ranks_bev) in ascending ranks_depth."""
    order = np.lexsort((rd, rb))
    return rb[order], rd[order], rf[order]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_transformer(mod, shape: Shape, **kw):
    return mod.LSSViewTransformer(grid_config=shape.grid_cfg(), input_size=shape.input_size,
                                  downsample=shape.downsample, in_channels=8,
                                  out_channels=shape.channels, collapse_z=False, **kw)


def run_prepare(vt, coor):
    out = vt.voxel_pooling_prepare_v2(coor)
    if out[0] is None:
        return None
    rb, rd, rf, st, ln = [o.numpy() for o in out]
    same = rb[1:] == rb[:-1]
    inv = int(((rd[1:] < rd[:-1]) & same).sum())
    crb, crd, crf = canon(rb, rd, rf)
    return dict(ranks_bev=rb, ranks_depth_raw=rd, ranks_feat_raw=rf, ranks_depth=crd, ranks_feat=crf,
                interval_starts=st, interval_lengths=ln, tie_inversions=np.int64(inv))


def main():
    if not reference_available():
        raise SystemExit('reference tree not found; golden fixtures can only be generated in the build container')
    torch.manual_seed(0)
    mod = load_reference_view_transformer('fusionocc')

    # ---- (1) tiny rig, B=2, full geometry chain
    sh = SHAPES['tiny']
    vt = ref_transformer(mod, sh)
    cal = make_calibration(sh, 2)
    coor = vt.get_lidar_coor(*cal)
    d = run_prepare(vt, coor)
    np.savez_compressed(os.path.join(HERE, 'geom_tiny.npz'),
                        grid_lower_bound=vt.grid_lower_bound.numpy(), grid_interval=vt.grid_interval.numpy(),
                        grid_size=vt.grid_size.numpy(), frustum=vt.frustum.numpy(),
                        sensor2ego=cal[0].numpy(), ego2global=cal[1].numpy(), cam2img=cal[2].numpy(),
                        post_rots=cal[3].numpy(), post_trans=cal[4].numpy(), bda=cal[5].numpy(),
                        coor=coor.numpy(), **d)

    # ---- (2) SID frustum + non-identity bda / augmentation (rotation+flip-like post_rots)
    vt_sid = ref_transformer(mod, sh, sid=True)
    cal2 = list(make_calibration(sh, 2))
    g = torch.Generator().manual_seed(11)
    ang = 0.05
    rot = torch.tensor([[np.cos(ang), -np.sin(ang), 0.], [np.sin(ang), np.cos(ang), 0.], [0., 0., 1.]],
                       dtype=torch.float32)
    cal2[3] = cal2[3] @ rot                                   # rotated post_rots
    cal2[3][1, :, 0, 0] *= -1.0                               # horizontal flip on sample 1
    cal2[4][1, :, 0] = float(sh.input_size[1])
    bda = torch.eye(3).repeat(2, 1, 1)
    bda[0] = torch.tensor([[0.98, -0.1, 0.], [0.1, 0.98, 0.], [0., 0., 1.02]])
    bda[1, 1, 1] = -1.0                                       # flip_dy
    cal2[5] = bda
    coor2 = vt_sid.get_lidar_coor(*cal2)
    d2 = run_prepare(vt_sid, coor2)
    np.savez_compressed(os.path.join(HERE, 'geom_tiny_sid_aug.npz'),
                        frustum=vt_sid.frustum.numpy(),
                        sensor2ego=cal2[0].numpy(), ego2global=cal2[1].numpy(), cam2img=cal2[2].numpy(),
                        post_rots=cal2[3].numpy(), post_trans=cal2[4].numpy(), bda=cal2[5].numpy(),
                        coor=coor2.numpy(), **d2)

    # ---- (3) hand-made edge coordinates on the full 200x200x16 grid
    base = SHAPES['base']
    vt_b = ref_transformer(mod, base)
    lb, itv = vt_b.grid_lower_bound, vt_b.grid_interval
    edge = []
    for x in (-40.4, -40.39, -40.2, -40.0, -39.999, -0.2, 0.0, 0.39, 39.59, 39.6, 39.99, 40.0, 40.2, 1e9, -1e9):
        for z in (-1.41, -1.39, -1.0, 5.0, 5.39, 5.4, 5.41):
            edge.append((x, 0.1, z))
            edge.append((0.1, x, z))
    edge = torch.tensor(edge, dtype=torch.float32)
    # duplicates -> long intervals; three samples, the last one entirely outside the grid
    n = edge.shape[0]
    reps = 8
    pts = edge.repeat(reps, 1)                                  # (n*reps, 3)
    B, N, D, H, W = 3, 1, reps, 1, n
    coor3 = torch.empty(B, N, D, H, W, 3)
    coor3[0, 0] = pts.view(D, H, W, 3)
    coor3[1, 0] = pts.flip(0).view(D, H, W, 3) + torch.tensor([0.4, 0.0, 0.0])
    coor3[2] = 1000.0
    d3 = run_prepare(vt_b, coor3)
    np.savez_compressed(os.path.join(HERE, 'edge_coords.npz'), coor=coor3.numpy(),
                        grid_lower_bound=lb.numpy(), grid_interval=itv.numpy(), grid_size=vt_b.grid_size.numpy(),
                        **d3)

    # ---- (4) nothing survives -> five Nones
    none_out = vt_b.voxel_pooling_prepare_v2(torch.full((1, 1, 2, 2, 2, 3), 1e6))
    assert all(o is None for o in none_out)

    # ---- (5) fp32 rank hazard: B=27 on the full grid, reference ranks are inexact
    g = torch.Generator().manual_seed(5)
    Bh = 28
    coor5 = torch.empty(Bh, 1, 4, 3, 5, 3)
    coor5[..., 0] = torch.rand(Bh, 1, 4, 3, 5, generator=g) * 80 - 40
    coor5[..., 1] = torch.rand(Bh, 1, 4, 3, 5, generator=g) * 80 - 40
    coor5[..., 2] = torch.rand(Bh, 1, 4, 3, 5, generator=g) * 6.4 - 1
    d5 = run_prepare(vt_b, coor5)
    np.savez_compressed(os.path.join(HERE, 'fp32_hazard_b28.npz'), coor=coor5.numpy(), **d5)

    # ---- (6) full-size digests (arrays too large to commit): base B=1,2 / native / stress
    digests = {}
    for name, Bs in (('base', (1, 2, 8)), ('native', (1, 8)), ('stress', (1, 2))):
        shp = SHAPES[name]
        vtf = ref_transformer(mod, shp)
        for Bn in Bs:
            cal = make_calibration(shp, Bn)
            coor = vtf.get_lidar_coor(*cal)
            dd = run_prepare(vtf, coor)
            digests[f'{name}_B{Bn}'] = dict(
                n_points=int(np.prod(coor.shape[:-1])), n_kept=int(dd['ranks_bev'].shape[0]),
                n_intervals=int(dd['interval_starts'].shape[0]),
                max_interval=int(dd['interval_lengths'].max()),
                tie_inversions=int(dd['tie_inversions']),
                frustum=sha(vtf.frustum.numpy()), coor=sha(coor.numpy()),
                ranks_bev=sha(dd['ranks_bev']), ranks_depth=sha(dd['ranks_depth']),
                ranks_feat=sha(dd['ranks_feat']), interval_starts=sha(dd['interval_starts']),
                interval_lengths=sha(dd['interval_lengths']))
    # adjacent-frame calibration (config C4)
    shp = SHAPES['base']
    vtf = ref_transformer(mod, shp)
    cal = make_calibration(shp, 1, frame_shift=True)
    coor = vtf.get_lidar_coor(*cal)
    dd = run_prepare(vtf, coor)
    digests['base_B1_adjframe'] = dict(n_kept=int(dd['ranks_bev'].shape[0]),
                                       n_intervals=int(dd['interval_starts'].shape[0]),
                                       coor=sha(coor.numpy()), ranks_bev=sha(dd['ranks_bev']),
                                       ranks_depth=sha(dd['ranks_depth']))
    meta = dict(torch=torch.__version__, numpy=np.__version__,
                reference='projects/FusionOcc/fusionocc/necks/view_transformer.py',
                note='ranks_depth/ranks_feat digests are tie-canonicalised (ascending ranks_depth within a voxel)')
    with open(os.path.join(HERE, 'fullsize_digests.json'), 'w') as f:
        json.dump(dict(meta=meta, digests=digests), f, indent=1, sort_keys=True)

    # ---- (6b) occ_pool: the reference's own pure-PyTorch implementation (OCC_Pool.py:39-71), executed here
    import importlib.util
    from tests.golden._ref_import import REFERENCE_ROOT
    op = os.path.join(REFERENCE_ROOT, 'projects', 'CONet', 'mmdet3d_plugin', 'ops', 'occ_pooling', 'OCC_Pool.py')
    spec = importlib.util.spec_from_file_location('_ref_occ_pool', op)
    occ = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(occ)
    go = torch.Generator().manual_seed(11)
    oB, oD, oH, oW, oC, oN = 2, 4, 24, 20, 8, 6000
    ocoords = torch.minimum((torch.rand(oN, 4, generator=go) * torch.tensor([oH, oW, oD, oB])).long(),
                            torch.tensor([oH, oW, oD, oB]) - 1).int()
    ofeats = torch.randn(oN, oC, generator=go)
    oout = occ.occ_pool_pure_pytorch(ofeats, ocoords, oB, oD, oH, oW)
    np.savez_compressed(os.path.join(HERE, 'occ_pool_ref.npz'), feats=ofeats.numpy(), coords=ocoords.numpy(),
                        out=oout.numpy(), dims=np.array([oB, oD, oH, oW]))

    # ---- (7) the reference's known-answer test, as data (bev_pool.py:145-176)
    np.savez(os.path.join(HERE, 'kat_bev_pool_v2.npz'),
             depth=np.array([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9], dtype=np.float32).reshape(1, 1, 2, 2, 2),
             feat=np.ones((1, 1, 2, 2, 2), dtype=np.float32),
             ranks_depth=np.array([0, 4, 1, 6], dtype=np.int32),
             ranks_feat=np.array([0, 0, 1, 2], dtype=np.int32),
             ranks_bev=np.array([0, 0, 1, 1], dtype=np.int32),
             bev_feat_shape=np.array([1, 1, 2, 2, 2]),
             loss=np.float32(4.4),
             grad_depth=np.array([2., 2., 0., 0., 2., 0., 2., 0.], dtype=np.float32).reshape(1, 1, 2, 2, 2),
             grad_feat=np.array([1.0, 1.0, 0.4, 0.4, 0.8, 0.8, 0., 0.], dtype=np.float32).reshape(1, 1, 2, 2, 2))
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(('.npz', '.json')):
            print(f'{fn:32s} {os.path.getsize(os.path.join(HERE, fn)):>9d} B')


if __name__ == '__main__':
    main()
