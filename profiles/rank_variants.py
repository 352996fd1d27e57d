"""A/B timing of the rank precompute (FO_RANK_IMPL = 0 round-1 global bucket sort, 1 slab sort), from coor and from
the calibration; all five rank arrays and the plan tables compared between the implementations.

    python profiles/rank_variants.py [--shape base] [--batch 8] [--iters 50]
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
    ap.add_argument('--iters', type=int, default=50)
    ap.add_argument('--impls', default='0,1')
    ap.add_argument('--fast', default='0,1', help='FO_RANK_FAST values to compare (with FO_RANK_IMPL=0)')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    vt, coor, depth, feat, og = make_inputs(SHAPES[a.shape], a.batch, 0, dev)
    ns = NativeStep(vt, coor, depth, feat, og)
    ns.setup_calib(vt, vt._bench_cal)
    res, ref = {}, None
    variants = [(impl, '0') for impl in a.impls.split(',') if impl != '0'] + [('0', f) for f in a.fast.split(',')]
    for impl, fast in sorted(variants, key=lambda t: (t[0] != '0', t[1])):
        os.environ['FO_RANK_IMPL'] = impl
        os.environ['FO_RANK_FAST'] = fast
        for mode, fn in (('coor', ns.rank_prepare), ('calib', ns.rank_prepare_calib)):
            for t in (ns.rb, ns.rd, ns.rf, ns.st, ns.ln):
                t.fill_(-7)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            nk, ni = (int(v) for v in ns.counts[:2].tolist())
            ns.forward()
            ns.backward_with_plan()
            got = [ns.rb[:nk].clone(), ns.rd[:nk].clone(), ns.rf[:nk].clone(), ns.st[:ni].clone(), ns.ln[:ni].clone(),
                   ns.out.clone(), ns.dg.clone(), ns.fg.clone()]
            same = None
            if ref is None:
                ref = (nk, ni, got)
            else:
                same = ref[0] == nk and ref[1] == ni and all(torch.equal(x, y) for x, y in zip(ref[2], got))
            if os.environ.get('FO_SLAB_PROF_READ'):
                prof = ns.rank_scratch[64:64 + 96].view(torch.int64).tolist()
                print('slab_sort phases (sum cycles over CTAs, max cycles):', prof)
            res[f'impl{impl}_fast{fast}_{mode}'] = {'us': e0.elapsed_time(e1) / a.iters * 1e3, 'n_kept': nk, 'n_intervals': ni,
                                         'identical_to_first': same}
    print(json.dumps({'shape': a.shape, 'batch': a.batch, 'rank_prepare': res}))


if __name__ == '__main__':
    main()
