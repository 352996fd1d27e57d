"""A/B timing of the forward write-out (FO_FWD_TMA = 0: LDS.128 + streaming STG.128 per lane; 1: one bulk tensor
store per sub-tile), same inputs, outputs compared bit for bit.

    python profiles/fwd_variants.py [--shape base] [--batch 8] [--iters 100]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import NativeStep, make_inputs  # noqa: E402
from fusionocc_b200.rig import SHAPES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='base')
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--iters', type=int, default=100)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    vt, coor, depth, feat, og = make_inputs(SHAPES[a.shape], a.batch, 0, dev, with_coor=False)
    ns = NativeStep(vt, None, depth, feat, og)
    ns.rank_prepare_calib()
    torch.cuda.synchronize()
    res, ref = {}, None
    for v in ('0', '1', '0', '1'):
        os.environ['FO_FWD_TMA'] = v
        ns.out.fill_(float('nan'))
        for _ in range(5):
            ns.forward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            ns.forward()
        e1.record()
        torch.cuda.synchronize()
        got = ns.out.clone()
        same = None
        if ref is None:
            ref = got
        else:
            same = bool(torch.equal(ref.view(torch.int32), got.view(torch.int32)))
        res.setdefault(f'tma{v}', []).append({'us': round(e0.elapsed_time(e1) / a.iters * 1e3, 2), 'bit_identical_to_first': same})
    print(json.dumps({'shape': a.shape, 'batch': a.batch, 'forward': res}))


if __name__ == '__main__':
    main()
