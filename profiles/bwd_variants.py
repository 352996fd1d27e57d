"""A/B timing of the backward implementations on one B200 (FO_BWD_IMPL = 0 round-1 kernels, 1 TMA gather +
multi-pixel kernel, 2 fused launch), same inputs, results compared bit for bit.

    python profiles/bwd_variants.py [--shape base] [--batch 8] [--iters 50]
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
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    vt, coor, depth, feat, og = make_inputs(SHAPES[a.shape], a.batch, 0, dev)
    ns = NativeStep(vt, coor, depth, feat, og)
    ns.rank_prepare(); ns.forward(); ns.bwd_plan_build()
    torch.cuda.synchronize()
    res, ref = {}, None
    for impl in a.impls.split(','):
        os.environ['FO_BWD_IMPL'] = impl
        ns.dg.fill_(float('nan')); ns.fg.fill_(float('nan'))
        for _ in range(3):
            ns.backward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            ns.backward()
        e1.record()
        torch.cuda.synchronize()
        got = (ns.dg.clone(), ns.fg.clone())
        same = None
        if ref is None:
            ref = got
        else:
            same = bool(torch.equal(ref[0].view(torch.int32), got[0].view(torch.int32)) and
                        torch.equal(ref[1].view(torch.int32), got[1].view(torch.int32)))
        res[f'impl{impl}'] = {'us': e0.elapsed_time(e1) / a.iters * 1e3, 'bit_identical_to_first': same}
        if impl == '2':
            res[f'impl{impl}']['spin_stats(total,max,waiters)'] = ns.bwd_scratch[:4 * (ns.B + 4)].view(torch.int32)[ns.B + 1:].tolist()
    # plan + backward: two calls vs the one call whose gather kernel carries the plan warps
    os.environ['FO_BWD_IMPL'] = '1'
    combo = {}
    for name, fn, env in (('plan_then_backward', lambda: (ns.bwd_plan_build(), ns.backward()), '1'),
                          ('backward_with_plan_separate_launch', ns.backward_with_plan, '0'),
                          ('backward_with_plan_riding', ns.backward_with_plan, '1')):
        os.environ['FO_BWD_RIDE'] = env
        ns.dg.fill_(float('nan')); ns.fg.fill_(float('nan'))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        same = bool(torch.equal(ref[0].view(torch.int32), ns.dg.view(torch.int32)) and
                    torch.equal(ref[1].view(torch.int32), ns.fg.view(torch.int32)))
        combo[name] = {'us': e0.elapsed_time(e1) / a.iters * 1e3, 'grads_bit_identical': same}
    print(json.dumps({'shape': a.shape, 'batch': a.batch, 'backward': res, 'plan_plus_backward': combo}))


if __name__ == '__main__':
    main()
