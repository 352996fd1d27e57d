"""A/B timing of one whole step (rank precompute from the calibration -> forward -> backward incl. its plan) under
different environment switches of the library, inside ONE process on ONE box; every variant's rank arrays, output and
gradients are compared bit for bit with the first variant's.

    python profiles/step_ab.py --env FO_PDL=0 --env FO_PDL=1 --env FO_PDL=1,FO_RANK_FAST=0 [--shape base] [--batch 8]
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
    ap.add_argument('--iters', type=int, default=300)
    ap.add_argument('--rounds', type=int, default=2)
    ap.add_argument('--env', action='append', default=[], help='comma-separated NAME=VALUE list; one variant per --env')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    vt, coor, depth, feat, og = make_inputs(SHAPES[a.shape], a.batch, 0, dev, with_coor=False)
    ns = NativeStep(vt, coor, depth, feat, og)
    variants = a.env or ['']
    touched = sorted({kv.split('=')[0] for v in variants for kv in v.split(',') if kv})
    res, ref = {}, None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for rnd in range(a.rounds):
        for v in variants:
            for k in touched:
                os.environ.pop(k, None)
            for kv in v.split(','):
                if kv:
                    k, val = kv.split('=')
                    os.environ[k] = val
            for t in (ns.rb, ns.rd, ns.rf, ns.st, ns.ln):
                t.fill_(-7)
            ns.out.fill_(-1.0); ns.dg.fill_(-1.0); ns.fg.fill_(-1.0)
            for _ in range(5):
                ns.step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                ns.step()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / a.iters * 1e3
            ph = [0.0, 0.0, 0.0]
            for _ in range(50):
                ns.step(ev)
                torch.cuda.synchronize()
                for i in range(3):
                    ph[i] += ev[i].elapsed_time(ev[i + 1]) * 1e3 / 50
            nk, ni = (int(x) for x in ns.counts[:2].tolist())
            got = [ns.rb[:nk].clone(), ns.rd[:nk].clone(), ns.rf[:nk].clone(), ns.st[:ni].clone(), ns.ln[:ni].clone(),
                   ns.out.clone(), ns.dg.clone(), ns.fg.clone()]
            same = None
            if ref is None:
                ref = (nk, ni, got)
            else:
                same = ref[0] == nk and ref[1] == ni and all(torch.equal(x, y) for x, y in zip(ref[2], got))
            r = res.setdefault(v or 'default', {'step_us': [], 'rank_fwd_bwd_us': [], 'identical_to_first': True})
            r['step_us'].append(round(us, 2))
            r['rank_fwd_bwd_us'].append([round(x, 1) for x in ph])
            if same is False:
                r['identical_to_first'] = False
    print(json.dumps({'shape': a.shape, 'batch': a.batch, 'n_kept': ref[0], 'n_intervals': ref[1], 'variants': res}))


if __name__ == '__main__':
    main()
