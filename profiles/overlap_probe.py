"""Does the backward plan (issue-bound) hide under the backward's gather (DRAM-bound) when it runs on a second,
high-priority stream?  Times  plan ; backward  on one stream against  plan || backward  on two.

    python profiles/overlap_probe.py [--shape base] [--batch 8]
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
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    vt, coor, depth, feat, og = make_inputs(SHAPES[a.shape], a.batch, 0, dev)
    ns = NativeStep(vt, coor, depth, feat, og)
    ns.rank_prepare(); ns.forward(); ns.bwd_plan_build(); ns.backward()
    torch.cuda.synchronize()
    main_s = torch.cuda.current_stream(dev)
    lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, 'priority_range') else (0, -1)
    res = {}
    for name, prio in (('serial', None), ('side_stream', 0), ('side_stream_high_priority', -1)):
        side = None if prio is None else torch.cuda.Stream(device=dev, priority=prio)
        ev_f, ev_j = torch.cuda.Event(), torch.cuda.Event()

        def once():
            if side is None:
                ns.bwd_plan_build(); ns.backward()
                return
            ev_f.record(main_s)
            side.wait_event(ev_f)
            with torch.cuda.stream(side):
                ns.bwd_plan_build()
                ev_j.record(side)
            ns.backward()                      # (not dependency-correct: measures co-running cost only)
            main_s.wait_event(ev_j)
        for _ in range(5):
            once()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            once()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / a.iters * 1e3
    print(json.dumps({'shape': a.shape, 'batch': a.batch, 'plan_plus_backward_us': res}))


if __name__ == '__main__':
    main()
