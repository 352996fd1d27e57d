"""Probe (BASELINE.json configs[1]): the new op vs the UNMODIFIED reference CUDA extension (oracle/_ref, built from
/root/reference by oracle/build_ref.py) on a B200, same inputs, CUDA events.  The reference side includes what its
Python wrapper does around the kernels (bev_pool.py:17-92: zero fill + permute copy; argsort + interval rebuild +
out_grad un-permute in the backward) and its eager-torch rank precompute (view_transformer.py:223-281, restated in
oracle/torch_cpu_path.py and run on the GPU)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fusionocc_b200.rig import SHAPES
from oracle import ref_ext
from oracle.torch_cpu_path import voxel_pooling_prepare_v2_torch
dev = torch.device('cuda:0')
def t(f, it=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
res = {}
for name, B in (('base', 1), ('base', 8), ('stress', 1)):
    shape = SHAPES[name]
    vt, coor, depth, feat, og = bench.make_inputs(shape, B, 0, dev)
    ns = bench.NativeStep(vt, coor, depth, feat, og)
    ns.rank_prepare(); ns.forward(); ns.bwd_plan_build(); ns.backward(); torch.cuda.synchronize()
    nk, ni = (int(v) for v in ns.counts[:2].tolist())
    rb, rd, rf, st, ln = ns.rb[:nk], ns.rd[:nk], ns.rf[:nk], ns.st[:ni], ns.ln[:ni]
    bshape = (B, ns.Z, ns.Y, ns.X, ns.C)
    lb, itv, gs = vt.grid_lower_bound.to(dev), vt.grid_interval.to(dev), vt.grid_size.to(dev)
    r = {'ours_rank_ms': t(ns.rank_prepare), 'ours_fwd_ms': t(ns.forward),
         'ours_bwd_ms': t(lambda: (ns.bwd_plan_build(), ns.backward())),
         'ref_fwd_ms': t(lambda: ref_ext.forward(ns.depth, ns.feat, rd, rf, rb, bshape, st, ln)),
         'ref_bwd_ms': t(lambda: ref_ext.backward(ns.og, ns.depth, ns.feat, rd, rf, rb)),
         'ref_rank_torch_gpu_ms': t(lambda: voxel_pooling_prepare_v2_torch(coor, lb, itv, gs), it=5),
         'n_kept': nk, 'n_intervals': ni}
    r['speedup_fwd'] = r['ref_fwd_ms'] / r['ours_fwd_ms']
    r['speedup_bwd'] = r['ref_bwd_ms'] / r['ours_bwd_ms']
    r['speedup_rank'] = r['ref_rank_torch_gpu_ms'] / r['ours_rank_ms']
    r['speedup_step'] = (r['ref_fwd_ms'] + r['ref_bwd_ms'] + r['ref_rank_torch_gpu_ms']) / (r['ours_fwd_ms'] + r['ours_bwd_ms'] + r['ours_rank_ms'])
    res[f'{name}_B{B}'] = r
    print(name, B, json.dumps({k: round(v, 4) if isinstance(v, float) else v for k, v in r.items()}))
    del ns
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/ref_ext_compare.json', 'w'), indent=1)
