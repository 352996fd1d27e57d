#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--shape base|native|stress] [--batch B_PER_GPU]

Metric: voxel-pool samples/s (and HBM GB/s of the dominant kernel vs the measured peak) for the
FusionOcc camera->voxel view transformation at 6 cams 256x704 -> 16x44, D=88, C=32, grid 200x200x16.

A "step" is one pass of the whole hot path over one batch of synthetic nuScenes-shaped input:
    rank precompute FROM THE CALIBRATION (fo_rank_prepare_calib: frustum-point -> voxel mapping under the
    camera/ego calibration, sort by voxel rank, interval extraction)  ->  forward splat  ->  backward plan  ->
    backward
with B samples per GPU (default 8: the training shape, BASELINE.json configs[3]/[1]).  Nothing is
cached across steps: every step re-derives ranks, intervals and both plans from the camera matrices.

  value   samples/s over all GPUs with inputs resident in HBM (CUDA-event time, max over ranks)
  e2e     the same step through the host-buffer C-ABI entry (fo_view_transform_host_calib): pinned host
          inputs (calibration, depth, feat, out_grad) H2D and results (voxels, depth_grad, feat_grad) D2H
          inside the timed region
  parity_checked  after the timed region the very buffers that were timed (rank arrays, voxels, both gradients)
          are compared bit for bit with the reference's torch rank precompute run on the GPU + the unmodified
          reference CUDA extension (oracle/_ref), outside any timing
  ref_cuda / python_api / c4_training / stress / c3_inference
          the other BASELINE.json configurations, bounded legs outside the headline timing
  roofline  dominant kernel = fwd_dense_kernel: algorithmic forward bytes (SURVEY.md §8d formula with
          the realised N_k / N_i) / its CUDA-event duration, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference
          the reference-style pure-PyTorch CPU path (oracle/torch_cpu_path.py: eager-torch rank
          precompute + index_add_ scatter + autograd backward) on the box's host cores

Multi-GPU (torchrun, one rank per GPU): samples are batch-sharded, B per rank (weak scaling), no
collective on the data path; the NCCL all_gather of the voxel outputs named by BASELINE.json is timed
separately and reported under "gather".
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KERNELS_PER_STEP = 6 + 1 + 3         # rank_prepare (zero, voxelise, tile sums, scan, place, order), forward, backward (zero, TMA gather with the plan CTAs, pixel kernel)
                                     # (memsets not counted)


# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(B, N, D, H, W, C, V, n_kept, n_iv):
    """fp32 = int32 = 4 bytes."""
    P, rows = B * N * D * H * W, B * N * H * W
    fwd = 4 * (P + rows * C + 3 * n_kept + 2 * n_iv + B * V * C)
    bwd = 4 * (n_iv * C + 2 * P + 2 * rows * C + 3 * n_kept + 2 * rows)
    pre = 4 * (3 * P + 3 * n_kept + 2 * n_iv)
    return dict(fwd=fwd, bwd=bwd, pre=pre, total=fwd + bwd + pre)


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:  # noqa: BLE001
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.gpu), '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def make_inputs(shape, B, first_sample, device, with_coor=True, with_og=True):
    """coor (from the product's own get_lidar_coor on `device`), depth, feat (NHWC fp32), out_grad."""
    import torch
    from fusionocc_b200 import LSSViewTransformer
    from fusionocc_b200.rig import make_calibration
    vt = LSSViewTransformer(shape.grid_cfg(), shape.input_size, shape.downsample, in_channels=8,
                            out_channels=shape.channels, collapse_z=False)
    total = first_sample + B
    cal = [c[first_sample:total].to(device) for c in make_calibration(shape, total)]
    coor = vt.get_lidar_coor(*cal).contiguous() if with_coor else None
    vt._bench_cal = cal
    N, D, C = shape.n_cams, vt.D, shape.channels
    H, W = shape.feat_hw
    X, Y, Z = vt._grid_xyz()
    depth = torch.empty(B, N, D, H, W)
    feat = torch.empty(B, N, H, W, C)
    og = torch.empty(B, C, Z, Y, X) if with_og else None
    for i in range(B):
        b = first_sample + i
        depth[i] = torch.randn(N, D, H, W, generator=torch.Generator().manual_seed(0 + 7919 * b)).softmax(dim=1)
        feat[i] = torch.randn(N, C, H, W, generator=torch.Generator().manual_seed(1 + 7919 * b)).permute(0, 2, 3, 1)
        if with_og:
            og[i] = torch.randn(C, Z, Y, X, generator=torch.Generator().manual_seed(2 + 7919 * b))
    return vt, coor, depth, feat, og


class NativeStep:
    """Static device buffers + the four C-ABI calls of one step (what a C++ host would do).  The rank precompute of
    ``step()`` starts from the calibration (fo_rank_prepare_calib); ``rank_prepare()`` (from a materialised ``coor``)
    is kept for the probes under profiles/."""

    def __init__(self, vt, coor, depth, feat, og, forward_only=False):
        import torch
        from fusionocc_b200 import _cabi
        self.torch, self.cabi, self.lib = torch, _cabi, _cabi.load()
        lib = self.lib
        dev = vt._bench_cal[0].device
        self.dev = dev
        self.B, self.N, self.D, self.H, self.W = depth.shape
        self.C = feat.shape[-1]
        self.X, self.Y, self.Z = vt._grid_xyz()
        self.V = self.X * self.Y * self.Z
        self.P = self.B * self.N * self.D * self.H * self.W
        self.rows = self.B * self.N * self.H * self.W
        NV = self.B * self.V
        self.cap_iv = min(self.P, NV)
        self.lb = _cabi.f3(vt.grid_lower_bound.tolist())
        self.itv = _cabi.f3(vt.grid_interval.tolist())
        i32 = dict(dtype=torch.int32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        self.coor, self.depth, self.feat = coor, depth.to(dev), feat.to(dev).contiguous()
        self.og = og.to(dev) if og is not None else None
        self.rb, self.rd, self.rf = (torch.empty(self.P, **i32) for _ in range(3))
        self.st, self.ln = (torch.empty(self.cap_iv, **i32) for _ in range(2))
        self.counts = torch.zeros(4, **i32)
        self.fwd_plan = torch.empty(lib.fo_fwd_plan_bytes(NV, self.P), **u8)
        self.rank_scratch = torch.empty(lib.fo_rank_prepare_scratch_bytes(self.P, NV), **u8)
        if not forward_only:
            self.bwd_plan = torch.empty(lib.fo_bwd_plan_bytes(self.P, self.rows), **u8)
            self.bwd_scratch = torch.empty(lib.fo_bwd_scratch_bytes(self.cap_iv, self.C, 0), **u8)
            self.out = torch.empty(self.B, self.C, self.Z, self.Y, self.X, device=dev)
            self.dg = torch.empty_like(self.depth)
            self.fg = torch.empty_like(self.feat)
        self.setup_calib(vt, vt._bench_cal)

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    def _s(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def rank_prepare(self):
        p, L = self._p, self.lib
        self.cabi.check(L.fo_rank_prepare(self._s(), p(self.coor), self.B, self.N, self.D, self.H, self.W, self.lb,
                                          self.itv, self.X, self.Y, self.Z, p(self.rb), p(self.rd), p(self.rf),
                                          p(self.st), p(self.ln), p(self.counts), p(self.fwd_plan),
                                          self.fwd_plan.numel(), p(self.rank_scratch), self.rank_scratch.numel()),
                        'fo_rank_prepare')

    def setup_calib(self, vt, cal):
        """Inputs of the fused-geometry rank precompute (fo_rank_prepare_calib)."""
        from fusionocc_b200.view_transformer import DEFAULT_MATVEC_MODE, pack_calibration
        s2e, _e2g, k, pr, pt, bda = cal
        self.cam, self.bda12, self.bda_has_t = pack_calibration(s2e, k, pr, pt, bda)
        self.frustum = vt._frustum_on(s2e).contiguous()
        self.matvec_mode = DEFAULT_MATVEC_MODE

    def rank_prepare_calib(self):
        p, L = self._p, self.lib
        self.cabi.check(L.fo_rank_prepare_calib(self._s(), p(self.frustum), p(self.cam), p(self.bda12),
                                                int(self.bda_has_t), self.matvec_mode, None, self.B, self.N, self.D,
                                                self.H, self.W, self.lb, self.itv, self.X, self.Y, self.Z, p(self.rb),
                                                p(self.rd), p(self.rf), p(self.st), p(self.ln), p(self.counts),
                                                p(self.fwd_plan), self.fwd_plan.numel(), p(self.rank_scratch),
                                                self.rank_scratch.numel()), 'fo_rank_prepare_calib')

    def forward(self):
        p, L = self._p, self.lib
        self.cabi.check(L.fo_bev_pool_v2_forward(
            self._s(), self.C, p(self.depth), p(self.feat), p(self.rd), p(self.rf), p(self.rb), p(self.st),
            p(self.ln), self.P, self.cap_iv, ctypes.c_void_p(self.counts.data_ptr() + 4), self.B, self.V,
            p(self.out), 0, 1, p(self.fwd_plan), self.fwd_plan.numel()), 'fo_bev_pool_v2_forward')

    def bwd_plan_build(self):
        p, L = self._p, self.lib
        self.cabi.check(L.fo_bwd_plan_build(self._s(), p(self.rd), p(self.rf), self.P, p(self.counts), self.P,
                                            self.rows, self.H * self.W, 1, p(self.fwd_plan), self.fwd_plan.numel(),
                                            self.B, self.V, p(self.bwd_plan), self.bwd_plan.numel()),
                        'fo_bwd_plan_build')

    def backward(self):
        p, L = self._p, self.lib
        self.cabi.check(L.fo_bev_pool_v2_backward(
            self._s(), self.C, p(self.og), 0, p(self.depth), p(self.feat), self.P, self.cap_iv, self.B, self.V,
            self.P, self.rows, p(self.dg), p(self.fg), p(self.fwd_plan), self.fwd_plan.numel(), p(self.bwd_plan),
            self.bwd_plan.numel(), p(self.bwd_scratch), self.bwd_scratch.numel()), 'fo_bev_pool_v2_backward')

    def backward_with_plan(self):
        """Backward plan + backward as ONE C-ABI call (the plan rides along the gather kernel)."""
        p, L = self._p, self.lib
        self.cabi.check(L.fo_bev_pool_v2_backward_with_plan(
            self._s(), self.C, p(self.og), 0, p(self.depth), p(self.feat), self.P, p(self.counts), self.cap_iv, self.B,
            self.V, self.P, self.rows, self.H * self.W, p(self.dg), p(self.fg), p(self.fwd_plan), self.fwd_plan.numel(),
            p(self.bwd_plan), self.bwd_plan.numel(), p(self.bwd_scratch), self.bwd_scratch.numel()),
            'fo_bev_pool_v2_backward_with_plan')

    def step(self, events=None):
        """rank precompute (from the calibration) -> forward -> backward incl. its plan, one stream.
        ``events`` brackets the three calls: [0..1] rank, [1..2] forward, [2..3] backward (plan + gather + pixel)."""
        s = self.torch.cuda.current_stream(self.dev)
        if events is None:
            self.rank_prepare_calib(); self.forward(); self.backward_with_plan()
            return
        events[0].record(s); self.rank_prepare_calib()
        events[1].record(s); self.forward()
        events[2].record(s); self.backward_with_plan()
        events[3].record(s)


class HostStep:
    """e2e: pinned host buffers -> fo_view_transform_host_calib -> pinned host buffers, batch split in
    chunks over streams so H2D, kernels and D2H of different chunks overlap."""

    def __init__(self, ns: NativeStep, n_chunks: int, two_streams: bool = True):
        import torch
        self.two_streams = two_streams
        self.torch, self.ns = torch, ns
        lib = ns.lib
        B = ns.B
        n_chunks = max(1, min(n_chunks, B))
        while B % n_chunks:
            n_chunks -= 1
        self.n_chunks, self.cb = n_chunks, B // n_chunks
        pin = lambda x: x.detach().cpu().contiguous().pin_memory()
        self.h_frustum, self.h_cam, self.h_bda = pin(ns.frustum), pin(ns.cam.view(B, ns.N, 24)), pin(ns.bda12)
        self.h_depth, self.h_feat, self.h_og = pin(ns.depth), pin(ns.feat), pin(ns.og)
        self.h_out = torch.empty(ns.out.shape).pin_memory()
        self.h_dg = torch.empty(ns.depth.shape).pin_memory()
        self.h_fg = torch.empty(ns.feat.shape).pin_memory()
        self.h_counts = torch.zeros(n_chunks, 4, dtype=torch.int32).pin_memory()
        wsb = lib.fo_view_transform_host_calib_workspace_bytes(self.cb, ns.N, ns.D, ns.H, ns.W, ns.C, ns.X, ns.Y, ns.Z, 1)
        self.ws = [torch.empty(wsb, dtype=torch.uint8, device=ns.dev) for _ in range(n_chunks)]
        self.streams = [torch.cuda.Stream(device=ns.dev) for _ in range(n_chunks)]
        self.up_streams = [torch.cuda.Stream(device=ns.dev) for _ in range(n_chunks)]
        el = lambda t: t.numel() * t.element_size()
        self.h2d_bytes = (el(self.h_frustum) * n_chunks + el(self.h_cam) + el(self.h_bda) + el(self.h_depth) +
                          el(self.h_feat) + el(self.h_og))
        self.d2h_bytes = el(self.h_out) + el(self.h_dg) + el(self.h_fg) + 16 * n_chunks

    def step(self):
        ns, p = self.ns, NativeStep._p
        cur = self.torch.cuda.current_stream(ns.dev)
        for i, s in enumerate(self.streams):
            s.wait_stream(cur)
            sl = slice(i * self.cb, (i + 1) * self.cb)
            rc = ns.lib.fo_view_transform_host_calib(
                ctypes.c_void_p(s.cuda_stream), p(self.h_frustum), p(self.h_cam[sl]), p(self.h_bda[sl]),
                int(ns.bda_has_t), ns.matvec_mode, p(self.h_depth[sl]), p(self.h_feat[sl]), p(self.h_og[sl]), self.cb,
                ns.N, ns.D, ns.H, ns.W, ns.C, ns.lb, ns.itv, ns.X, ns.Y, ns.Z, p(self.h_out[sl]), p(self.h_dg[sl]),
                p(self.h_fg[sl]), p(self.h_counts[i]), p(self.ws[i]), self.ws[i].numel(),
                ctypes.c_void_p(self.up_streams[i].cuda_stream) if self.two_streams else None)
            ns.cabi.check(rc, 'fo_view_transform_host_calib')
        for s in self.streams:
            cur.wait_stream(s)


def pcie_probe(torch, dev, mb=256):
    """Pinned-memory copy bandwidth of this box, one direction at a time and both at once (GB/s)."""
    n = mb * (1 << 20) // 4
    h0, h1 = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
    d0, d1 = torch.empty(n, device=dev), torch.empty(n, device=dev)
    s0, s1 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(up, down):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s0.wait_stream(torch.cuda.current_stream(dev)); s1.wait_stream(torch.cuda.current_stream(dev))
        for _ in range(3):
            if up:
                with torch.cuda.stream(s0):
                    d0.copy_(h0, non_blocking=True)
            if down:
                with torch.cuda.stream(s1):
                    h1.copy_(d1, non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(s0); torch.cuda.current_stream(dev).wait_stream(s1)
        e1.record()
        torch.cuda.synchronize()
        return 3 * n * 4 * (int(up) + int(down)) / (e0.elapsed_time(e1) * 1e-3) / 1e9
    run(True, True)
    return {'h2d_GBps': run(True, False), 'd2h_GBps': run(False, True), 'both_GBps': run(True, True), 'MiB': mb}


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(shape, steps, warmup, budget_s=None):
    """The reference-style CPU path on host cores, one sample per step (bounded sample of the batch)."""
    import torch
    from fusionocc_b200.rig import make_calibration, make_out_grad, make_values
    from oracle import rank_oracle as ro
    from oracle.torch_cpu_path import bev_pool_v2_pure_torch, voxel_pooling_prepare_v2_torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lb, itv, gs = (torch.from_numpy(a) for a in ro.create_grid_infos(**shape.grid_cfg()))
    fr = ro.create_frustum(shape.depth_cfg, shape.input_size, shape.downsample)
    cal = make_calibration(shape, 1)
    coor = ro.get_lidar_coor(fr, *cal)
    depth, feat_nchw = make_values(shape, 1)
    X, Y, Z = int(gs[0]), int(gs[1]), int(gs[2])
    og = make_out_grad(1, shape.channels, Z, Y, X)
    bshape = (1, Z, Y, X, shape.channels)

    def one():
        rb, rd, rf, st, ln = voxel_pooling_prepare_v2_torch(coor, lb, itv, gs)
        d = depth.detach().requires_grad_(True)
        f = feat_nchw.detach().requires_grad_(True)
        out = bev_pool_v2_pure_torch(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, bshape)
        out.backward(og)
        return out

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        one()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, dt / done * 1e3, done, cores


def run_reference(args):
    from fusionocc_b200.rig import SHAPES
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    shape = SHAPES[args.shape]
    sps, ms, done, cores = cpu_reference_run(shape, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': 'voxel-pool samples/s (rank precompute + bev_pool_v2 fwd + bwd)',
        'value': sps, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': done, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': workload_config(shape, args.batch, 'gpu'),
        'cpu_baseline': {'value': sps, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                         'sample': '1 sample per step (1/8 of the GPU arm\'s per-GPU batch), same shape, '
                                   'eager-torch rank precompute + index_add_ scatter + autograd backward '
                                   '(oracle/torch_cpu_path.py)'},
        'e2e': {'value': sps, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(shape, B, where):
    H, W = shape.feat_hw
    return {'workload': f'FusionOcc view transform, every step from the calibration: rank precompute (frustum -> voxel '
                        f'mapping, sort, intervals) + bev_pool_v2 fwd + bwd, {shape.n_cams} cams '
                        f'{shape.input_size[0]}x{shape.input_size[1]} ({H}x{W} feat), D={shape.D}, C={shape.channels}, '
                        f'grid 200x200x16, batch {B} per {"GPU" if where == "gpu" else "step"}',
            'shape': shape.name, 'batch_per_gpu': B, 'n_cams': shape.n_cams, 'D': shape.D, 'C': shape.channels,
            'feat_hw': [H, W], 'grid_zyx': [16, 200, 200],
            'l2_policy': 'working set >> L2: the dense voxel output alone is 81.92 MB x batch per step '
                         '(655 MB at batch 8) vs 126 MB L2; no explicit flush' if where == 'gpu' else 'n/a'}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from fusionocc_b200.rig import SHAPES

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py (impl=ours) needs a CUDA device: the product has no CPU path')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL prints its version banner to STDOUT at communicator creation; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    shape = SHAPES[args.shape]
    B = args.batch
    vt, coor, depth, feat, og = make_inputs(shape, B, rank * B, dev)
    ns = NativeStep(vt, coor, depth, feat, og)
    K, Wm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):
        ns.step()
    torch.cuda.synchronize()
    n_kept, n_iv = (int(v) for v in ns.counts[:2].tolist())

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.6)
    barrier()
    # ---- timed region: EXACTLY K steps, nothing but the step's own launches on the stream (an event record between two
    #      kernels ends the programmatic launch chain there, so the per-phase events live in a second pass below)
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(K):
        ns.step()
    t_end.record()
    barrier()
    total_ms = t_start.elapsed_time(t_end)
    # ---- second pass over the same K steps with CUDA events between the three C-ABI calls: per-phase times and the
    #      forward kernel's launch duration for the roofline (each event costs the chain a launch boundary:
    #      instrumented_ms_per_step is what this pass takes)
    i_start = torch.cuda.Event(enable_timing=True)
    i_end = torch.cuda.Event(enable_timing=True)
    i_start.record()
    for k in range(K):
        ns.step(ev[k])
    i_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    instrumented_ms_per_step = i_start.elapsed_time(i_end) / K
    phase = [sum(ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(K)) / K for i in range(3)]
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / K
    value = world * B * K / (total_ms * 1e-3)

    def timed(fn, it=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / it

    def leg(fn):
        """Secondary legs never take the headline down with them."""
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            return {'error': f'{type(e).__name__}: {e}'[:300]}

    # ---- parity of the very buffers that were timed (outside any timing)
    parity = leg(lambda: parity_check(torch, ns, vt, coor, n_kept, n_iv))

    # ---- row (f-1): what the fused geometry replaces
    geometry = leg(lambda: {
        'torch_get_lidar_coor_ms': timed(lambda: vt.get_lidar_coor(*vt._bench_cal)),
        'rank_prepare_from_coor_ms': timed(ns.rank_prepare, it=50),
        'rank_prepare_calib_fused_ms': phase[0],
        'note': 'the timed step uses the fused call (fo_rank_prepare_calib: frustum points computed per thread from '
                'the calibration, bit-identical to the torch ops, never stored); the other two are what it replaces'})
    ns.rank_prepare_calib()

    # ---- row (f-2): the step before the splat, one native pass vs the reference's torch ops
    def lift_leg():
        from fusionocc_b200 import lift_prepare
        lift = {}
        for dt, nm in ((torch.float32, 'f32'), (torch.float16, 'f16')):
            xl = torch.randn(B * ns.N, ns.D + ns.C, ns.H, ns.W, device=dev).to(dt)

            def ref_ops():
                d = xl[:, :ns.D].float().softmax(dim=1)                              # view_transformer.py:333-335
                f = xl[:, ns.D:ns.D + ns.C].permute(0, 2, 3, 1).contiguous().float()   # bev_pool.py:20-21
                return d, f
            lift[nm] = {'torch_ops_ms': timed(ref_ops), 'lift_prepare_ms': timed(lambda: lift_prepare(xl, ns.D, ns.C))}
            del xl
        lift['note'] = 'depth softmax + channel split + NCHW->NHWC + fp32 cast of the depth-net output; not part of value'
        return lift
    lift = leg(lift_leg)

    # ---- e2e through the host-buffer entry
    hs = HostStep(ns, n_chunks=args.e2e_chunks)
    for _ in range(4):
        hs.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Ke = max(3, min(K, args.e2e_steps))
    e0.record()
    for _ in range(Ke):
        hs.step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = world * B * Ke / (e2e_ms * 1e-3)
    h2d_bytes, d2h_bytes, e2e_chunks = hs.h2d_bytes, hs.d2h_bytes, hs.n_chunks
    del hs
    pcie = leg(lambda: pcie_probe(torch, dev)) if rank == 0 else None
    barrier()

    # ---- BASELINE.json configs[2]: batch 64 split over the GPUs, forward only, NCCL gather of the voxel outputs
    c3 = leg(lambda: c3_inference(torch, dist, shape, world, rank, dev, args.c3_batch)) if args.c3_batch > 0 else None
    barrier()

    # ---- NCCL gather of this step's voxel outputs, stand-alone (BASELINE.json: "NCCL used only to gather outputs")
    gather = None
    if world > 1:
        gbuf = torch.empty((world,) + tuple(ns.out.shape), device=dev)
        for _ in range(2):
            dist.all_gather_into_tensor(gbuf, ns.out)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            dist.all_gather_into_tensor(gbuf, ns.out)
        g1.record()
        barrier()
        gms = g0.elapsed_time(g1) / 5
        tt = torch.tensor([gms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        gms = float(tt.item())
        recv = (world - 1) * ns.out.numel() * 4
        gather = {'collective': 'all_gather_into_tensor (NCCL)', 'ms': gms, 'recv_GBps_per_rank': recv / gms / 1e6,
                  'bytes_per_rank': ns.out.numel() * 4, 'note': 'timed separately; not part of value'}
        del gbuf

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rank 0 only: the other BASELINE configurations on one GPU (bounded legs)
    ref_cuda = leg(lambda: ref_cuda_leg(torch, shape, dev, timed)) if not args.no_extras else None
    python_api = leg(lambda: python_api_leg(torch, vt, ns, timed)) if not args.no_extras else None
    c4 = leg(lambda: c4_training_leg(torch, shape, B, dev, timed)) if not args.no_extras else None
    del ns.og, ns.out
    torch.cuda.empty_cache()
    stress = leg(lambda: other_shape_leg(torch, 'stress', args.stress_batch, dev)) \
        if (not args.no_extras and args.shape == 'base') else None
    native = leg(lambda: other_shape_leg(torch, 'native', B, dev)) if (not args.no_extras and args.shape == 'base') else None

    ab = algorithmic_bytes(B, ns.N, ns.D, ns.H, ns.W, ns.C, ns.V, n_kept, n_iv)
    peak, peak_src = measured_peak()
    fwd_ms = phase[1]
    achieved = ab['fwd'] / (fwd_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic_fwd_dense.json')
    if os.path.isfile(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(f'{shape.name}_B{B}')
        except Exception:  # noqa: BLE001
            traffic = None

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sps, ms, done, cores = cpu_reference_run(shape, 200, 1, budget_s=args.cpu_budget)
        cpu = {'value': sps, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
               'sample': f'{done} single-sample steps of the same shape (~{args.cpu_budget:.0f} s of CPU work): '
                         'eager-torch rank precompute + index_add_ scatter + autograd backward',
               'ms_per_sample': ms}

    frac = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9 / peak
    line = {
        'metric': 'voxel-pool samples/s (rank precompute + bev_pool_v2 fwd + bwd)',
        'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': K, 'warmup': Wm,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(shape, B, 'gpu'),
        'realised': {'n_points': ns.P, 'n_kept': n_kept, 'n_intervals': n_iv},
        'instrumented_ms_per_step': instrumented_ms_per_step,
        'phases_ms': {'rank_prepare': phase[0], 'forward': phase[1], 'backward_incl_plan': phase[2],
                      'note': 'three C-ABI calls on one stream: fo_rank_prepare_calib (from the calibration), '
                              'fo_bev_pool_v2_forward, fo_bev_pool_v2_backward_with_plan (the backward plan is built by '
                              'extra CTAs of the gather kernel); measured in a second pass over the same K steps with '
                              'CUDA events between the calls (instrumented_ms_per_step): every event ends the '
                              'programmatic launch chain, which the timed region of value / ms_per_step keeps intact'},
        'algorithmic_MB_per_step': {k: v / 1e6 for k, v in ab.items()},
        'step_hbm_frac': frac(ab['total'], ms_per_step),
        'phase_hbm_frac': {'rank_prepare': frac(ab['pre'], phase[0]), 'forward': frac(ab['fwd'], phase[1]),
                           'backward_incl_plan': frac(ab['bwd'], phase[2])},
        'roofline': {'kernel': 'fwd_dense_kernel', 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': ab['fwd'], 'launch_ms': fwd_ms},
        'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d_bytes,
                'd2h_bytes_per_step': d2h_bytes, 'ms_per_step': e2e_ms / Ke, 'steps': Ke,
                'chunks': e2e_chunks, 'api': 'fo_view_transform_host_calib (C ABI, pinned host buffers)',
                'pcie_probe': pcie,
                'wire_floor_samples_per_s_per_gpu': (B / max(h2d_bytes / (pcie['h2d_GBps'] * 1e9), d2h_bytes / (pcie['d2h_GBps'] * 1e9))
                                             if isinstance(pcie, dict) and 'h2d_GBps' in pcie else None),
                # uploads and downloads run at the same time: the box's both-directions throughput is the tighter bound
                'both_ways_floor_samples_per_s_per_gpu': (B / ((h2d_bytes + d2h_bytes) / (pcie['both_GBps'] * 1e9))
                                                          if isinstance(pcie, dict) and 'both_GBps' in pcie else None),
                'note': 'every rank moves its own bytes through the host; the box-wide host<->device throughput, not '
                        'the kernels, bounds this figure (boxes of the pool differ: 89-97 GB/s both ways)'},
        'gpu_launches': KERNELS_PER_STEP * K,
        'clocks': clocks,
        'parity_checked': bool(isinstance(parity, dict) and parity.get('ok')),
        'parity': parity,
    }
    if cpu:
        line['cpu_baseline'] = cpu
    line['geometry'] = geometry
    line['lift_prepare'] = lift
    for k, v in (('ref_cuda', ref_cuda), ('python_api', python_api), ('c4_training', c4), ('stress', stress),
                 ('native', native), ('c3_inference', c3), ('gather', gather)):
        if v is not None:
            line[k] = v
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
# Secondary legs (outside the headline timing)
# ------------------------------------------------------------------------------------------------
def parity_check(torch, ns, vt, coor, n_kept, n_iv):
    """The timed buffers against (a) the reference's eager-torch rank precompute run on this GPU and (b) the
    UNMODIFIED reference CUDA extension (oracle/_ref) fed with THOSE ranks: every array bit for bit."""
    from oracle import ref_ext
    from oracle.torch_cpu_path import voxel_pooling_prepare_v2_torch
    if not ref_ext.available():
        return {'ok': False, 'why': 'oracle/_ref (the reference CUDA extension) is not built on this box'}
    dev = ns.dev
    ns.step()                                      # the buffers exactly as the timed loop leaves them
    torch.cuda.synchronize()
    lb, itv, gs = vt.grid_lower_bound.to(dev), vt.grid_interval.to(dev), vt.grid_size.to(dev)
    want = voxel_pooling_prepare_v2_torch(coor, lb, itv, gs)
    got = (ns.rb[:n_kept], ns.rd[:n_kept], ns.rf[:n_kept], ns.st[:n_iv], ns.ln[:n_iv])
    names = ('ranks_bev', 'ranks_depth', 'ranks_feat', 'interval_starts', 'interval_lengths')
    res = {nm: bool(a.shape == b.shape and torch.equal(a, b)) for nm, a, b in zip(names, got, want)}
    rb, rd, rf, st, ln = want
    bshape = (ns.B, ns.Z, ns.Y, ns.X, ns.C)
    bits = lambda t: t.contiguous().view(torch.int32)
    out = ref_ext.forward(ns.depth, ns.feat, rd, rf, rb, bshape, st, ln)
    res['out'] = bool(torch.equal(bits(out), bits(ns.out)))
    del out
    dg, fg = ref_ext.backward(ns.og, ns.depth, ns.feat, rd, rf, rb)
    res['depth_grad'] = bool(torch.equal(bits(dg), bits(ns.dg)))
    res['feat_grad'] = bool(torch.equal(bits(fg), bits(ns.fg)))
    res['ok'] = all(res.values())
    res['against'] = 'reference torch rank precompute on the GPU + unmodified reference CUDA extension (oracle/_ref), bit for bit'
    return res


def ref_cuda_leg(torch, shape, dev, timed):
    """BASELINE.json configs[1]: forward only, batch 1 and 8, the new op vs the reference CUDA op INCLUDING what its
    Python wrapper does around the kernel (zero fill + permute copy, bev_pool.py:27,91), same precomputed ranks."""
    from oracle import ref_ext
    if not ref_ext.available():
        return {'unavailable': 'oracle/_ref not built'}
    out = {}
    for b in (1, 8):
        vt, coor, depth, feat, og = make_inputs(shape, b, 0, dev)
        ns = NativeStep(vt, coor, depth, feat, og)
        ns.rank_prepare_calib(); ns.forward()
        torch.cuda.synchronize()
        nk, ni = (int(v) for v in ns.counts[:2].tolist())
        rb, rd, rf, st, ln = ns.rb[:nk], ns.rd[:nk], ns.rf[:nk], ns.st[:ni], ns.ln[:ni]
        bshape = (b, ns.Z, ns.Y, ns.X, ns.C)
        ours = timed(ns.forward, it=50)
        ref = timed(lambda: ref_ext.forward(ns.depth, ns.feat, rd, rf, rb, bshape, st, ln), it=10)
        same = bool(torch.equal(ref_ext.forward(ns.depth, ns.feat, rd, rf, rb, bshape, st, ln).view(torch.int32),
                                ns.out.view(torch.int32)))
        out[f'batch{b}'] = {'ours_fwd_ms': ours, 'reference_ext_fwd_ms': ref, 'speedup': ref / ours,
                            'samples_per_s': b / (ours * 1e-3), 'bit_identical': same}
        del ns
    out['note'] = 'forward only, static ranks (the reference op has no rank stage); reference = oracle/_ref built unmodified'
    return out


def python_api_leg(torch, vt, ns, timed):
    """The same step through the drop-in Python surface: LSSViewTransformer.view_transform (fused geometry, rank
    precompute, bev_pool_v2 autograd op, all allocations and ctypes marshalling) + .backward()."""
    B, N, D, H, W, C = ns.B, ns.N, ns.D, ns.H, ns.W, ns.C
    dev = ns.dev
    vt = vt.to(dev)
    cal = vt._bench_cal
    inp = [torch.zeros(B, N, 8, H, W, device=dev)] + list(cal)
    depth = ns.depth.view(B * N, D, H, W).clone().requires_grad_()
    tran = ns.feat.view(B, N, H, W, C).permute(0, 1, 4, 2, 3).reshape(B * N, C, H, W).contiguous().requires_grad_()
    og = ns.og

    def once():
        depth.grad = None; tran.grad = None
        bev, _ = vt.view_transform(inp, depth, tran)
        bev.backward(og)
    ms = timed(once, it=30)
    same = bool(torch.equal(depth.grad.view(-1).view(torch.int32), ns.dg.view(-1).view(torch.int32)))
    return {'ms_per_step': ms, 'samples_per_s': B / (ms * 1e-3), 'depth_grad_bit_identical_to_c_abi_step': same,
            'api': 'LSSViewTransformer.view_transform(...) + autograd backward (fuse_geometry default)'}


def c4_training_leg(torch, shape, B, dev, timed):
    """BASELINE.json configs[3]: two temporal frames with separate rank sets, adjacent frame forward-only under
    no_grad, key frame fwd+bwd, frames concatenated along channels (fusion_occ.py:289-326)."""
    from fusionocc_b200 import LSSViewTransformer, bev_pool_v2, bev_pool_v2_cat
    from fusionocc_b200.rig import make_calibration, make_values
    vt = LSSViewTransformer(shape.grid_cfg(), shape.input_size, shape.downsample, in_channels=8,
                            out_channels=shape.channels, collapse_z=False)
    X, Y, Z = vt._grid_xyz()
    C = shape.channels
    frames = []
    for shift in (False, True):
        cal = [c.to(dev) for c in make_calibration(shape, B, frame_shift=shift)]
        rb, rd, rf, st, ln = vt.voxel_pooling_prepare_v2(vt.get_lidar_coor(*cal))
        d, f = make_values(shape, B)
        frames.append([d.to(dev), f.to(dev).permute(0, 1, 3, 4, 2).contiguous(), rd, rf, rb, st, ln])
    og = torch.randn(B, 2 * C, Z, Y, X, device=dev)
    bshape = (B, Z, Y, X, C)

    def cat_torch():
        d0 = frames[0][0].clone().requires_grad_(); f0 = frames[0][1].clone().requires_grad_()
        a = bev_pool_v2(d0, f0, *frames[0][2:5], bshape, *frames[0][5:])
        with torch.no_grad():
            b = bev_pool_v2(*frames[1][:5], bshape, *frames[1][5:])
        torch.cat([a, b], dim=1).backward(og)
        return d0.grad, f0.grad

    def cat_native():
        d0 = frames[0][0].clone().requires_grad_(); f0 = frames[0][1].clone().requires_grad_()
        out = bev_pool_v2_cat([(d0, f0, *frames[0][2:]), tuple(frames[1])], bshape)
        out.backward(og)
        return d0.grad, f0.grad
    ga, gb = cat_torch(), cat_native()
    same = all(torch.equal(x.view(torch.int32), y.view(torch.int32)) for x, y in zip(ga, gb))
    a_ms, b_ms = timed(cat_torch), timed(cat_native)
    return {'frames': 2, 'batch': B, 'op_plus_torch_cat_ms': a_ms, 'bev_pool_v2_cat_ms': b_ms,
            'samples_per_s': B / (b_ms * 1e-3), 'gradients_bit_identical': bool(same),
            'note': 'two forwards + one backward + concatenation per step; ranks static across iterations'}


def other_shape_leg(torch, name, B, dev, steps=50):
    """Another BASELINE shape through the same C-ABI step (configs[4]: the stress shape; the FusionOcc-native input)."""
    from fusionocc_b200.rig import SHAPES
    shape = SHAPES[name]
    vt, coor, depth, feat, og = make_inputs(shape, B, 0, dev)
    ns = NativeStep(vt, coor, depth, feat, og)
    del coor
    for _ in range(3):
        ns.step()
    torch.cuda.synchronize()
    nk, ni = (int(v) for v in ns.counts[:2].tolist())
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(steps):
        ns.step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    for k in range(steps):
        ns.step(ev[k])
    torch.cuda.synchronize()
    phase = [sum(ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(steps)) / steps for i in range(3)]
    ab = algorithmic_bytes(B, ns.N, ns.D, ns.H, ns.W, ns.C, ns.V, nk, ni)
    peak, _ = measured_peak()
    return {'shape': name, 'batch': B, 'samples_per_s': B / (ms * 1e-3), 'ms_per_step': ms,
            'phases_ms': dict(zip(('rank_prepare', 'forward', 'backward_incl_plan'), phase)),
            'step_hbm_frac': ab['total'] / (ms * 1e-3) / 1e9 / peak,
            'forward_hbm_frac': ab['fwd'] / (phase[1] * 1e-3) / 1e9 / peak,
            'realised': {'n_kept': nk, 'n_intervals': ni}}


def c3_inference(torch, dist, shape, world, rank, dev, total_batch):
    """BASELINE.json configs[2]: `total_batch` samples split over the GPUs, FORWARD ONLY, voxel outputs gathered on
    every rank with NCCL.  The local batch is processed in chunks; with world > 1 every chunk's all_gather is issued
    on a side stream as soon as the chunk's forward is enqueued, so it overlaps the next chunks' kernels."""
    if total_batch % world:
        return {'skipped': f'batch {total_batch} does not split over {world} ranks'}
    lb_ = total_batch // world
    chunk = 4 if lb_ % 4 == 0 else (2 if lb_ % 2 == 0 else 1)
    n_chunks = lb_ // chunk
    steps = []
    for ci in range(n_chunks):
        vt, _coor, depth, feat, _og = make_inputs(shape, chunk, rank * lb_ + ci * chunk, dev, with_coor=False, with_og=False)
        ns = NativeStep(vt, None, depth, feat, None, forward_only=True)
        steps.append(ns)
    out_local = torch.empty(lb_, steps[0].C, steps[0].Z, steps[0].Y, steps[0].X, device=dev)
    for ci, ns in enumerate(steps):
        ns.out = out_local[ci * chunk:(ci + 1) * chunk]
    gathered = torch.empty((world,) + tuple(out_local.shape), device=dev) if world > 1 else None
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    cur = torch.cuda.current_stream(dev)

    def run(with_gather, overlapped):
        evs = []
        for ci, ns in enumerate(steps):
            ns.rank_prepare_calib(); ns.forward()
            if with_gather and overlapped:
                e = torch.cuda.Event(); e.record(cur); evs.append(e)
                side.wait_event(e)
                with torch.cuda.stream(side):
                    dist.all_gather_into_tensor(gathered_chunks[ci], ns.out)
        if with_gather and not overlapped:
            dist.all_gather_into_tensor(gathered, out_local)
        if with_gather and overlapped:
            cur.wait_stream(side)

    # per-chunk gather targets: contiguous (world, chunk, C, Z, Y, X) buffers (all_gather needs contiguous outputs)
    gathered_chunks = [torch.empty((world, chunk) + tuple(out_local.shape[1:]), device=dev) for _ in range(n_chunks)] \
        if world > 1 else None

    def timeit(fn, it=5):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / it
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms
    compute_ms = timeit(lambda: run(False, False))
    res = {'total_batch': total_batch, 'batch_per_gpu': lb_, 'chunk': chunk, 'forward_only': True,
           'compute_ms': compute_ms, 'compute_samples_per_s': total_batch / (compute_ms * 1e-3)}
    if world > 1:
        serial_ms = timeit(lambda: run(True, False))
        over_ms = timeit(lambda: run(True, True))
        recv = (world - 1) * out_local.numel() * 4
        res['gather'] = {'collective': 'all_gather_into_tensor (NCCL)', 'serial_ms': serial_ms, 'overlapped_ms': over_ms,
                         'gather_only_ms': serial_ms - compute_ms,
                         'with_gather_samples_per_s': total_batch / (over_ms * 1e-3),
                         'recv_bytes_per_rank': recv, 'recv_GBps_per_rank_overlapped': recv / over_ms / 1e6,
                         'nvlink_bound_samples_per_s': total_batch / (recv / 770e9),
                         'limiter': 'NVLink ingest of the gathered voxel tensors (81.92 MB per sample lands on every '
                                    'rank; 770 GB/s per direction measured peer-copy peak), not the kernels'}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--shape', default='base', choices=['base', 'native', 'stress'])
    ap.add_argument('--batch', type=int, default=8, help='samples per GPU per step')
    ap.add_argument('--e2e-steps', type=int, default=10)
    ap.add_argument('--e2e-chunks', type=int, default=2,
                    help='batch chunks of the host-buffer leg (profiles/e2e_probe.py: 2 chunks 13.8 ms, 4 chunks 14.0 ms, 8 chunks 14.1 ms per step)')
    ap.add_argument('--cpu-budget', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the secondary legs (ref_cuda, python_api, c4, stress)')
    ap.add_argument('--stress-batch', type=int, default=8)
    ap.add_argument('--c3-batch', type=int, default=64, help='total batch of the configs[2] inference leg (0 = skip)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
