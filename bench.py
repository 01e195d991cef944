#!/usr/bin/env python
"""Benchmark of the camera->BEV pooling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of the R50 workload (BASELINE config 2:
B=8 samples x 6 cameras x D118 x 16x44, C=80 -> 128x128 BEV) per GPU:
    voxel_pooling_prepare_v2  ->  context transpose  ->  bev_pool_v2 forward  ->  backward
exactly what a training step of the reference runs with accelerate=False
(mmdet3d/models/necks/view_transformer.py:180-205, mmdet3d/ops/bev_pool_v2/bev_pool.py).
Weak scaling: every rank pools its own 8 samples, no collective on the data path.

Rank 0 prints ONE JSON line (everything else any library writes to stdout is sent to stderr).
`value` = samples/s with inputs resident in HBM (CUDA events around the K timed steps, max over
ranks; per-stage events on every 8th step only, they cost ~18 us per step); `e2e` = the same step
through the public API from pinned HOST buffers, host<->device copies inside the timed region --
the call a user of the reference makes hands over the CALIBRATION (view_transform's `input[1:7]`,
view_transformer.py:290-294), so the e2e step ships ~5 KB of matrices + depth + context + out_grad
and get_lidar_coor runs fused inside prepare (`e2e_coor_input` = the same with a materialised
`coor` uploaded, round 1's path); `roofline` = the dominant kernel against the measured HBM peak;
`configs` = the other BASELINE configs (temporal fold, hi-res sweep, radar) timed after the
headline region; `reference_gpu` = the reference's own CUDA op (oracle/_ref, when built) and its
torch prepare chain on the same inputs; `cpu_baseline` = the CPU oracle (C, OpenMP) and the
PyTorch restatement timed on this box's host cores (N=1, rank 0).
`--impl reference` times that CPU implementation alone (the reference has no CPU pool kernel;
/root/reference is absent on the GPU box, so the restatement in oracle/ is what runs).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bev_pool_samples_per_s"
UNIT = "samples/s"
WORKLOAD = dict(workload="RCBEVDet R50 256x704 LSS view transform: prepare + bev_pool_v2 fwd+bwd, "
                         "B=8/GPU, 6 cams, D=118, 16x44 features, C=80 -> 128x128 BEV (BASELINE config 2)",
                batch_per_gpu=8, cams=6, D=118, H=16, W=44, C=80, bev=[1, 128, 128])
N_SETS = 3  # rotating input sets
# Per-stage CUDA events are recorded on every 8th timed step: ten event records
# between the kernels cost ~18 us per step (measured: 375.6 vs 357.7 us), so instrumenting every
# step would tax the number the stages are meant to explain.  All steps do identical work.
STAGE_EVENT_EVERY = 8


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# clocks: NVML sampled from a thread during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, torch_device_index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _sample(self):
        try:
            self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            bits = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            for b, name in self.REASONS.items():
                if bits & b and name != "gpu_idle":
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            time.sleep(0.001)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def mark(self):
        """Samples taken so far belong to the pre-spin; what follows is the timed region."""
        self.n_spin = len(self.samples)

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
        if self.ok and not self.samples:
            self._sample()
        n_spin = getattr(self, "n_spin", 0)
        timed = self.samples[n_spin:] or self.samples
        return {"sm_mhz": statistics.median(timed) if timed else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples) - n_spin,
                "samples_pre_spin": n_spin,
                "sm_mhz_pre_spin": statistics.median(self.samples[:n_spin]) if n_spin else None}


def pin_to_gpu_numa_node(torch, local_rank):
    """Run this rank (and therefore the first-touch placement of its pinned staging buffers) on the CPUs
    NVML reports as local to its GPU.  On a single-socket box this is a no-op; the record says which."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"gpu_local_cpus": len(cpus), "pinned_to": len(allowed) or len(os.sched_getaffinity(0)),
                "host_cpus": os.cpu_count()}
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
def make_inputs(torch, rig, device, B, seed, with_calib=False):
    """One input set on `device` (host tensors when device == 'cpu'): the key frame of B samples,
    per-sample ego jitter so the sets differ.  coor comes from the frustum geometry."""
    motion = rig.temporal_motion(B, 2, seed=seed).view(B, 2, 3)[:, 1]
    calib = rig.camera_rig(B, input_size=rig.R50_INPUT, frame_motion=motion)
    coor = rig.lidar_coor(calib, rig.R50_GRID["depth"], rig.R50_INPUT, 16)
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(B, N, D, H, W, WORKLOAD["C"], seed=seed)
    out_grad = torch.randn(B, WORKLOAD["C"], 1, 128, 128, generator=torch.Generator().manual_seed(seed + 2))
    if with_calib:
        return tuple(t.to(device) for t in (coor, depth, feat, out_grad)), calib
    return tuple(t.to(device) for t in (coor, depth, feat, out_grad))


def algorithmic_bytes(K, I, B):
    """bytes each stage must move at minimum (4-byte words)."""
    w = WORKLOAD
    P = B * w["cams"] * w["D"] * w["H"] * w["W"]
    F = B * w["cams"] * w["H"] * w["W"]
    G = B * 128 * 128
    C = w["C"]
    return {"prepare": 12 * P + 12 * K + 8 * I,
            "fwd": 4 * (K + F * C + 2 * K + 3 * I + G * C),
            "bwd": 4 * (I * C + K + F * C + 3 * K + P + F * C)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import _lib, bev_pool as bp, rig
    from rcbevdet_b200.prepare import prepare_async

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: rcbevdet_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL logs (its version banner included) go to stdout by default: keep stdout to the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    affinity = pin_to_gpu_numa_node(torch, local_rank)
    lib = _lib.lib()
    B, C = WORKLOAD["batch_per_gpu"], WORKLOAD["C"]
    lo, iv, sz = rig.grid_tensors(rig.R50_GRID)
    sets = [make_inputs(torch, rig, dev, B, seed=100 * rank + 10 * s + 1) for s in range(N_SETS)]
    stream = torch.cuda.current_stream(dev)

    stage_names = ("prepare", "feat_rows", "fwd", "og_rows", "bwd")
    n_kernels = {"prepare": 5, "feat_rows": 1, "fwd": 1, "og_rows": 1, "bwd": 1}
    stage_events = {s: [] for s in stage_names}

    og_rows_cl = [None] * N_SETS   # out_grad of each set as channels-last rows (the channels-last variant's input)

    def step(i, record, channels_last=False):
        coor, depth, feat, out_grad = sets[i % N_SETS]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if record else None
        if record:
            ev[0].record(stream)
        prepared = prepare_async(coor, lo, iv, sz)                               # row P
        if record:
            ev[1].record(stream)
        rows = bp.feat_rows(feat.permute(0, 1, 3, 4, 2))                         # bev_pool.py:21
        if record:
            ev[2].record(stream)
        d = _lib.PoolDesc()
        d.n_points, d.n_intervals, d.C = prepared.P, 0, C
        d.B, d.Z, d.Y, d.X = B, 1, 128, 128
        d.n_depth, d.n_pixels, d.D, d.HW, d.H = depth.numel(), rows.shape[0], prepared.D, prepared.HW, prepared.H
        d.layout = _lib.LAYOUT_CELLS_C if channels_last else _lib.LAYOUT_B_C_CELLS
        d.feat_dtype, d.flags = _lib.DTYPE_F32, _lib.PLAN_ALL
        out = torch.empty((B, 1, 128, 128, C) if channels_last else (B, C, 1, 128, 128), dtype=torch.float32, device=dev)
        bp.pool_forward(d, depth, rows, prepared.ranks_depth, prepared.ranks_feat, prepared.ranks_bev, None, None,
                        prepared.cell_start, out)                                                  # row F
        if record:
            ev[3].record(stream)
        # backward = out_grad (B,C,cells) -> channels-last rows (bev_pool.py:69), then the gradient kernel
        if channels_last:   # the gradient arrives channels-last (bev_pool_v2(..., channels_last=True)): used in place
            og_rows = og_rows_cl[i % N_SETS]
        else:
            og_rows = torch.empty((B * 128 * 128, C), dtype=torch.float32, device=dev)
            _lib.check(lib.rcb_planes_to_rows(_lib.ptr(out_grad), _lib.ptr(og_rows), B, C, 128 * 128, C * 128 * 128, 4,
                                              dev.index, _lib.stream_ptr(dev)), "og_rows")
        if record:
            ev[4].record(stream)
        depth_grad = torch.empty_like(depth)
        feat_grad = torch.empty_like(rows)
        d.layout = _lib.LAYOUT_CELLS_C
        ws_bytes = lib.rcb_pool_bwd_workspace_bytes(ctypes.byref(d))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.rcb_bev_pool_v2_bwd(ctypes.byref(d), _lib.ptr(og_rows), _lib.ptr(depth), _lib.ptr(rows),
                                           _lib.ptr(prepared.ranks_depth), _lib.ptr(prepared.ranks_feat),
                                           _lib.ptr(prepared.ranks_bev), _lib.ptr(prepared.point_cell),
                                           _lib.ptr(depth_grad), _lib.ptr(feat_grad), _lib.ptr(ws), ws_bytes,
                                           dev.index, _lib.stream_ptr(dev)), "bwd")                # row B
        if record:
            ev[5].record(stream)
            for k, s in enumerate(stage_names):
                stage_events[s].append((ev[k], ev[k + 1]))
        return prepared, out, depth_grad, feat_grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up, then EXACTLY K timed steps ------------------------------------------------
    for i in range(args.warmup):
        last = step(i, False)
    torch.cuda.synchronize(dev)
    prepared = last[0]
    K_pts, I_iv = (int(v) for v in prepared.counts[:2].tolist())
    checksum = float(last[1].double().sum().item())
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    # pre-spin: >= 250 ms of the same steps right before the timed ones, so that a short timed window
    # (the driver's K = 20 is 6 ms) starts at the load clocks and NVML gets enough samples
    spin_t0 = time.perf_counter()
    i_spin = 0
    while time.perf_counter() - spin_t0 < 0.25:
        for _ in range(20):
            step(i_spin, False)
            i_spin += 1
        torch.cuda.synchronize(dev)
    barrier()
    clocks.mark()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    event_phase = min(STAGE_EVENT_EVERY // 2, args.steps - 1)  # not step 0: the launch queue is still empty there
    t0.record(stream)
    host_t0 = time.perf_counter()
    for i in range(args.steps):
        step(i, i % STAGE_EVENT_EVERY == event_phase)
    t1.record(stream)
    host_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps   # launch-side time per step
    barrier()
    clk = clocks.stop()
    total_ms = t0.elapsed_time(t1)
    stage_ms = {s: sum(a.elapsed_time(b) for a, b in stage_events[s]) / len(stage_events[s]) for s in stage_names}

    # ---- optional variant, reported beside the fp32 numbers (not part of `value`): bf16 context rows,
    #      fp32 accumulation (north_star: bf16 context within 1e-2) ------------------------------
    def fwd_only(rows, dtype_code, n=50):
        coor, depth, feat, _ = sets[0]
        prepared = prepare_async(coor, lo, iv, sz)
        d = _lib.PoolDesc()
        d.n_points, d.n_intervals, d.C = prepared.P, 0, C
        d.B, d.Z, d.Y, d.X = B, 1, 128, 128
        d.n_depth, d.n_pixels, d.D, d.HW, d.H = depth.numel(), rows.shape[0], prepared.D, prepared.HW, prepared.H
        d.layout, d.feat_dtype, d.flags = _lib.LAYOUT_B_C_CELLS, dtype_code, _lib.PLAN_ALL
        outs = [torch.empty((B, C, 1, 128, 128), dtype=torch.float32, device=dev) for _ in range(4)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in range(n + 5):
            if k == 5:
                e0.record(stream)
            bp.pool_forward(d, depth, rows, prepared.ranks_depth, prepared.ranks_feat, prepared.ranks_bev, None,
                            None, prepared.cell_start, outs[k % 4])
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    def strips_variant(n=50):
        """The strip kernels (csrc/strips.cu, opt-in: strips.set_mode) on the same inputs: plan build,
        forward and backward with the plan cached -- what cached ranks (the reference's accelerate mode)
        pay per call.  Outputs checked against the cell-/pixel-stationary kernels before timing."""
        from rcbevdet_b200 import strips
        coor, depth, feat, out_grad = sets[0]
        rows = bp.feat_rows(feat.permute(0, 1, 3, 4, 2))
        prepared = prepare_async(coor, lo, iv, sz)
        d = _lib.PoolDesc()
        d.n_points, d.n_intervals, d.C = prepared.P, 0, C
        d.B, d.Z, d.Y, d.X = B, 1, 128, 128
        d.n_depth, d.n_pixels, d.D, d.HW, d.H = depth.numel(), rows.shape[0], prepared.D, prepared.HW, prepared.H
        d.layout, d.feat_dtype, d.flags = _lib.LAYOUT_B_C_CELLS, _lib.DTYPE_F32, _lib.PLAN_ALL
        n_img, W_ = rows.shape[0] // prepared.HW, prepared.HW // prepared.H

        def build():
            return strips.build(prepared.point_cell, prepared.cell_start, n_img, prepared.D, prepared.H, W_, prepared.n_cells)

        sp = build()
        if sp is None or sp.status() != 0:
            return {"status": None if sp is None else sp.status()}
        ref = torch.empty((B, C, 1, 128, 128), dtype=torch.float32, device=dev)
        bp.pool_forward(d, depth, rows, prepared.ranks_depth, prepared.ranks_feat, prepared.ranks_bev, None, None,
                        prepared.cell_start, ref)
        outs = [torch.empty_like(ref) for _ in range(4)]
        dg, fg = torch.empty_like(depth), torch.empty_like(rows)
        strips.forward(sp, d, depth, rows, outs[0])
        err = float((outs[0] - ref).abs().max() / ref.abs().max())

        def timed(fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for k in range(n + 5):
                if k == 5:
                    e0.record(stream)
                fn(k)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            return round(e0.elapsed_time(e1) / n, 5)

        return {"status": 0, "fwd_rel_err_vs_cells": err,
                "plan_ms": timed(lambda k: build()),
                "fwd_ms": timed(lambda k: strips.forward(sp, d, depth, rows, outs[k % 4])),
                "bwd_ms": timed(lambda k: strips.backward(sp, d, out_grad, depth, rows, dg, fg)),
                "note": "strip-stationary kernels with the plan cached (forward = strip pass + cell combine; "
                        "backward = spread + strip pass, takes the (B,C,cells) gradient directly: no transpose); "
                        "compare fwd_f32_rows_ms and stages_ms.og_rows + stages_ms.bwd"}

    def api_chain_variant(n=100):
        """The whole step through the public chain API (rcb.voxel_pooling_v2 + autograd backward: what
        LSSViewTransformer.voxel_pooling_v2 + loss.backward() run) with the default kernels (sorted
        pipeline + cell-/pixel-stationary kernels) and with strips mode "chain" (no sort: frustum cells
        -> strip plan -> strip kernels; the sorted pipeline enqueued behind them, gated on the device).
        Eager and replayed from a CUDA graph (pure device time)."""
        from rcbevdet_b200 import strips
        coor, depth, feat, out_grad = sets[0]
        og4 = out_grad.view(B, C, 128, 128)
        res, outs = {}, {}
        old = strips.MODE
        try:
            for mode in ("off", "chain"):
                strips.set_mode(mode)

                def one():
                    d = depth.detach().requires_grad_(True)
                    f = feat.detach().requires_grad_(True)
                    bev = rcb.voxel_pooling_v2(coor, d, f, lo, iv, sz)
                    bev.backward(og4)
                    return bev.detach(), d.grad, f.grad

                for _ in range(5):
                    outs[mode] = one()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(n):
                    one()
                e1.record(stream)
                torch.cuda.synchronize(dev)
                res[f"{mode}_eager_ms"] = round(e0.elapsed_time(e1) / n, 5)
                side = torch.cuda.Stream(dev)
                with torch.cuda.stream(side):
                    one()
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    one()
                with torch.cuda.stream(side):
                    for _ in range(5):
                        graph.replay()
                    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    g0.record(side)
                    for _ in range(n):
                        graph.replay()
                    g1.record(side)
                torch.cuda.synchronize(dev)
                res[f"{mode}_graph_ms"] = round(g0.elapsed_time(g1) / n, 5)
                del graph
        finally:
            strips.set_mode(old)
        res["rel_err_chain_vs_default"] = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs["chain"], outs["off"])]
        res["samples_per_s_chain_graph"] = round(B / (res["chain_graph_ms"] * 1e-3), 1)
        res["note"] = ("rcb.voxel_pooling_v2(coor, depth, feat) + backward on the device, same inputs every call "
                       "(working set > L2); off = sorted pipeline + cell-/pixel-stationary kernels, chain = strips mode "
                       "'chain' (no sort, no ranks materialised; fallback enqueued gated); *_graph_ms = CUDA-graph replay")
        return res

    variants = None
    if not args.profile:
        # the opt-in channels-last result (no transposed write, gradient consumed in place): same step,
        # out_grad resident as channels-last rows -- what a BEV encoder in torch.channels_last hands back
        for k in range(N_SETS):
            og_rows_cl[k] = sets[k][3].view(B, C, 128 * 128).permute(0, 2, 1).contiguous().view(B * 128 * 128, C)
        n_cl = max(20, min(args.steps, 200))
        for i in range(10):
            step(i, False, channels_last=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for i in range(n_cl):
            step(i, False, channels_last=True)
        c1.record(stream)
        torch.cuda.synchronize(dev)
        step_cl_ms = c0.elapsed_time(c1) / n_cl
        rows32 = bp.feat_rows(sets[0][2].permute(0, 1, 3, 4, 2))
        variants = {"fwd_f32_rows_ms": round(fwd_only(rows32, _lib.DTYPE_F32), 5),
                    "fwd_bf16_rows_ms": round(fwd_only(rows32.bfloat16(), _lib.DTYPE_BF16), 5),
                    "step_channels_last_ms": round(step_cl_ms, 5),
                    "strips": strips_variant(),
                    "api_chain": api_chain_variant(),
                    "note": "fwd_*: forward kernel alone, same plan, 4 rotating outputs (inputs L2-warm); "
                            "step_channels_last_ms: the whole step with bev_pool_v2(..., channels_last=True) "
                            "semantics (rows written directly, channels-last out_grad used in place)"}

    # ---- end to end through the public API from pinned host buffers --------------------------
    from rcbevdet_b200.prepare import frustum_axes, pack_calib
    host_raw = [make_inputs(torch, rig, "cpu", B, seed=100 * rank + 10 * s + 1, with_calib=True) for s in range(2)]
    host_sets = [tuple(t.pin_memory() for t in tensors) for tensors, _ in host_raw]
    host_calib = [calib for _, calib in host_raw]
    axes = frustum_axes(rig.R50_GRID["depth"], rig.R50_INPUT, 16, device=dev)
    host_out = [torch.empty((B, C, 128, 128), dtype=torch.float32).pin_memory(),
                torch.empty(host_sets[0][1].shape, dtype=torch.float32).pin_memory(),
                torch.empty(host_sets[0][2].shape, dtype=torch.float32).pin_memory()]
    d2h = sum(t.numel() * t.element_size() for t in host_out)

    # Three streams, two slots: the upload of step i+1 and the download of step i-1 overlap the
    # kernels of step i (PCIe is full duplex).  Every step's inputs still come from pinned host
    # memory and every step's results (pooled BEV + both gradients) still land in host memory
    # inside the timed region; the host waits for step i-1's results before it issues step i+1.
    s_in, s_comp, s_out = (torch.cuda.Stream(dev) for _ in range(3))
    dev_in = [tuple(torch.empty_like(t, device=dev) for t in host_sets[0]) for _ in range(2)]
    host_res = [[torch.empty_like(t).pin_memory() for t in host_out] for _ in range(2)]
    ev_h2d = [torch.cuda.Event() for _ in range(2)]
    ev_comp = [torch.cuda.Event() for _ in range(2)]
    ev_d2h = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(n, from_calib):
        """from_calib: the step receives what view_transform receives -- calibration matrices (packed on
        the host inside the timed loop, ~5 KB uploaded), depth, context, out_grad.  Otherwise a
        materialised `coor` (47.8 MB) is uploaded too and prepare reads it."""
        for i in range(n):
            slot = i % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_comp[slot])          # slot's previous kernels are done with the buffers
                first = 1 if from_calib else 0
                for dst, src in zip(dev_in[slot][first:], host_sets[slot][first:]):
                    dst.copy_(src, non_blocking=True)
                if from_calib:
                    cam, bda = pack_calib(*host_calib[slot])
                    packed = (cam.pin_memory().to(dev, non_blocking=True), bda.pin_memory().to(dev, non_blocking=True))
                ev_h2d[slot].record(s_in)
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(ev_h2d[slot])
                coor, depth, feat, og = dev_in[slot]
                depth = depth.detach().requires_grad_(True)
                feat = feat.detach().requires_grad_(True)
                if from_calib:                                                   # public API (rows f-1 + V)
                    for t in packed:
                        t.record_stream(s_comp)
                    bev = rcb.voxel_pooling_v2_from_calib(packed, axes, depth, feat, lo, iv, sz)
                else:
                    bev = rcb.voxel_pooling_v2(coor, depth, feat, lo, iv, sz)    # public API (row V)
                bev.backward(og.view(bev.shape))
                results = (bev.detach(), depth.grad, feat.grad)
                ev_comp[slot].record(s_comp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[slot])
                ev_d2h[slot].synchronize()              # host has consumed this slot's previous results
                for dst, src in zip(host_res[slot], results):
                    src.record_stream(s_out)
                    dst.copy_(src, non_blocking=True)
                ev_d2h[slot].record(s_out)
            if i > 0:
                ev_d2h[(i - 1) % 2].synchronize()       # step i-1's results are on the host
        ev_d2h[(n - 1) % 2].synchronize()
        torch.cuda.synchronize(dev)

    e2e_steps = 2 if args.profile else max(4, min(args.steps, 60))

    def e2e_timed(from_calib):
        e2e_run(2 if args.profile else 4, from_calib)
        barrier()
        w0 = time.perf_counter()
        e2e_run(e2e_steps, from_calib)
        barrier()
        return time.perf_counter() - w0, float(host_res[(e2e_steps - 1) % 2][0].double().sum())

    e2e_s, e2e_check = e2e_timed(True)
    e2e_coor_s, e2e_coor_check = e2e_timed(False)
    h2d_coor = sum(t.numel() * t.element_size() for t in host_sets[0])
    h2d = h2d_coor - host_sets[0][0].numel() * 4 + (B * 6 * 24 + B * 9) * 4

    # ---- the other BASELINE configs and the reference's own GPU code, outside the headline region ----
    extras = None
    if not args.profile and rank == 0:
        torch.cuda.empty_cache()
        extras = extra_configs(torch, rcb, rig, _lib, bp, dev, lib)
        extras["reference_gpu"] = reference_gpu(torch, rcb, rig, sets[0], lo, iv, sz, dev)
        torch.cuda.empty_cache()

    # ---- max over ranks ----------------------------------------------------------------------
    times = torch.tensor([total_ms, e2e_s * 1e3, e2e_coor_s * 1e3] + [stage_ms[s] for s in stage_names],
                         dtype=torch.float64, device=dev)
    sums = torch.tensor([checksum], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(sums) for _ in range(world)]
        dist.all_gather(gathered, sums)            # NCCL only for verification/reporting
        checks = [float(g.item()) for g in gathered]
    else:
        checks = [checksum]
    total_ms, e2e_ms, e2e_coor_ms = float(times[0]), float(times[1]), float(times[2])
    stage_ms = {s: float(times[3 + k]) for k, s in enumerate(stage_names)}
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    e2e_coor_value = world * B * e2e_steps / (e2e_coor_ms * 1e-3)

    if rank == 0:
        peak, peak_src = _peaks()
        alg = algorithmic_bytes(K_pts, I_iv, B)
        # the dominant single kernel: forward tile kernel or backward gradient kernel (one launch each)
        kernel_stage = max(("fwd", "bwd"), key=lambda s: stage_ms[s])
        kernel_names = {"fwd": "k_fwd_cells", "bwd": "k_pool_bwd_pixels16"}
        achieved = alg[kernel_stage] / (stage_ms[kernel_stage] * 1e-3) / 1e9
        traffic = kernel_traffic(kernel_stage)
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(WORKLOAD, parallelism=f"sample-sharded x{world}, no data-path collective",
                           l2=f"rotating {N_SETS} input sets (~117 MB each) + ~250 MB of outputs/workspace per "
                              "step: every step's working set exceeds the 126 MB L2",
                           kept_points=K_pts, intervals=I_iv),
            "clocks": clk,
            "host_affinity": affinity,
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "h2d_gbs_per_rank": round(h2d * e2e_steps / (e2e_ms * 1e-3) / 1e9, 2),
                    "d2h_gbs_per_rank": round(d2h * e2e_steps / (e2e_ms * 1e-3) / 1e9, 2),
                    "api": "rcbevdet_b200.voxel_pooling_v2_from_calib (calibration in, get_lidar_coor fused into "
                           "prepare) + autograd backward; pinned host in/out, 3 streams x 2 slots (upload / kernels "
                           "/ download overlapped)", "checksum": e2e_check},
            "e2e_coor_input": {"value": round(e2e_coor_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d_coor,
                               "d2h_bytes_per_step": d2h, "api": "rcbevdet_b200.voxel_pooling_v2 (materialised coor "
                               "uploaded: round 1's path)", "checksum": e2e_coor_check},
            "gpu_launches": args.steps * sum(n_kernels.values()),
            "stages_ms": {s: round(v, 5) for s, v in stage_ms.items()},
            "host_launch_ms_per_step": round(host_ms, 5),
            "stage_events": f"every {STAGE_EVENT_EVERY}th timed step ({len(stage_events['fwd'])} of {args.steps})",
            "stage_frac_of_hbm_peak": {s: round(alg[s] / (stage_ms[s] * 1e-3) / 1e9 / peak, 4)
                                       for s in ("prepare", "fwd", "bwd")},
            "stage_algorithmic_bytes": alg,
            "prepare_plus_fwd": {"samples_per_s": round(world * B / ((stage_ms["prepare"] + stage_ms["feat_rows"] +
                                                                      stage_ms["fwd"]) * 1e-3), 1),
                                 "frac_of_hbm_peak": round((alg["prepare"] + alg["fwd"]) /
                                                           ((stage_ms["prepare"] + stage_ms["feat_rows"] +
                                                             stage_ms["fwd"]) * 1e-3) / 1e9 / peak, 4)},
            "roofline": {"bound": "hbm", "kernel": kernel_names[kernel_stage], "achieved": round(achieved, 1),
                         "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[kernel_stage]},
            "variants": variants,
            "checksums": checks,
        }
        if extras is not None:
            line["reference_gpu"] = extras.pop("reference_gpu")
            line["configs"] = extras
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# DRAM traffic of the dominant kernel: one `ncu --set full` capture per change, kept under profiles/.
# The file is stamped with the sha256 of the kernel sources it was captured from; a stale capture
# (sources changed since) is reported as null instead of a number that no longer describes the code.
# --------------------------------------------------------------------------------------------
# the sources of the kernels the timed step launches (the traffic figures describe these)
STEP_KERNEL_SOURCES = ("common.cuh", "prepare_common.cuh", "prepare.cu", "prepare_lsd.cu", "layout.cu",
                       "pool_fwd_cells.cu", "pool_bwd.cu")


def _kernel_sources_digest():
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "rcbevdet_b200", "csrc")
    for name in STEP_KERNEL_SOURCES:
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def kernel_traffic(stage):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if t.get("csrc_sha256_16") != _kernel_sources_digest():
            return None
        return t.get(stage)
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
# the other BASELINE configs (3: temporal fold, 4: radar, 5: hi-res sweep), device-resident inputs
# --------------------------------------------------------------------------------------------
def _time_cuda(torch, fn, n_warm=3, n=10):
    for _ in range(n_warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def _pool_case(torch, rcb, rig, dev, B, grid, input_size, C, frames=1, backward=True, chain=False):
    """prepare -> pool forward (-> backward) through the fused device chain on B*frames folded samples.
    Returns ms per stage, samples/s and algorithmic GB/s (SURVEY.md 8d formulas on this case's P, K, I, F, G)."""
    from rcbevdet_b200.prepare import prepare_async
    peak, _ = _peaks()
    n_fold = B * frames
    motion = rig.temporal_motion(B, frames, seed=7) if frames > 1 else None
    calib = tuple(t.to(dev) for t in rig.camera_rig(n_fold, input_size=input_size, frame_motion=motion))
    coor = rig.lidar_coor(calib, grid["depth"], input_size, 16)
    _, N, D, H, W, _ = coor.shape
    g = torch.Generator(device=dev).manual_seed(11)
    depth = torch.randn(n_fold, N, D, H, W, device=dev, generator=g).softmax(dim=2)
    feat = torch.randn(n_fold, N, C, H, W, device=dev, generator=g)
    lo, iv, sz = rig.grid_tensors(grid)
    gx, gy = int(sz[0]), int(sz[1])
    og = torch.randn(n_fold, C, gy, gx, device=dev, generator=g)
    t_prep = _time_cuda(torch, lambda: prepare_async(coor, lo, iv, sz))
    t_fwd_all = _time_cuda(torch, lambda: rcb.voxel_pooling_v2(coor, depth, feat, lo, iv, sz))
    out = {"samples_folded": n_fold, "prepare_ms": round(t_prep, 4), "prepare_fwd_ms": round(t_fwd_all, 4)}
    r = prepare_async(coor, lo, iv, sz)
    K, I = (int(v) for v in r.counts[:2].tolist())
    P, F, G = coor.numel() // 3, n_fold * N * H * W, n_fold * gx * gy
    alg_pf = 12 * P + 12 * K + 8 * I + 4 * (K + F * C + 2 * K + 3 * I + G * C)
    out.update(points=P, kept=K, intervals=I,
               prepare_fwd_gbs=round(alg_pf / (t_fwd_all * 1e-3) / 1e9, 1),
               prepare_fwd_frac_of_hbm_peak=round(alg_pf / (t_fwd_all * 1e-3) / 1e9 / peak, 4))
    if backward:
        def full():
            d_ = depth.detach().requires_grad_(True)
            f_ = feat.detach().requires_grad_(True)
            rcb.voxel_pooling_v2(coor, d_, f_, lo, iv, sz).backward(og)
        t_all = _time_cuda(torch, full)
        alg_b = 4 * (I * C + K + F * C + 3 * K + P + F * C)
        out.update(prepare_fwd_bwd_ms=round(t_all, 4), samples_per_s=round(B / (t_all * 1e-3), 1),
                   gbs=round((alg_pf + alg_b) / (t_all * 1e-3) / 1e9, 1),
                   frac_of_hbm_peak=round((alg_pf + alg_b) / (t_all * 1e-3) / 1e9 / peak, 4))
    else:
        out.update(samples_per_s=round(B / (t_fwd_all * 1e-3), 1))
    if chain:   # the same calls in strips mode "chain" (no sort: frustum cells -> strip plan -> strip forward)
        from rcbevdet_b200 import strips
        old = strips.MODE
        strips.set_mode("chain")
        try:
            out["chain_prepare_fwd_ms"] = round(_time_cuda(torch, lambda: rcb.voxel_pooling_v2(coor, depth, feat, lo, iv, sz)), 4)
            if backward:
                out["chain_prepare_fwd_bwd_ms"] = round(_time_cuda(torch, full), 4)
        finally:
            strips.set_mode(old)
    return out


def _radar_case(torch, rcb, rig, dev, B, n):
    peak, _ = _peaks()
    pf, rcs, coors = (t.to(dev) for t in rig.radar_pillars(B, n, n, seed=4))
    t = _time_cuda(torch, lambda: rcb.radar_rcs_scatter(pf, rcs, coors, B, n, n))
    V = pf.shape[0]
    alg = V * (64 + 7 + 4) * 4 + B * n * n * 66 * 4
    return {"pillars": V, "ms": round(t, 4), "samples_per_s": round(B / (t * 1e-3), 1),
            "gbs": round(alg / (t * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(alg / (t * 1e-3) / 1e9 / peak, 4)}


def _graph_case(torch, rcb, rig, dev):
    """BASELINE config 1's geometry (one sample) on the GPU: the calibration-driven chain + backward, eager
    and replayed from a CUDA graph (the chain never touches the host, so it captures as is).  One sample
    is host-launch-bound when launched eagerly."""
    axes = rcb.frustum_axes(rig.R50_GRID["depth"], rig.R50_INPUT, 16, device=dev)
    cam, bda = (t.to(dev) for t in rcb.pack_calib(*rig.camera_rig(1)))
    lo, iv, sz = rig.grid_tensors(rig.R50_GRID)
    depth, feat = (t.to(dev) for t in rig.pooling_inputs(1, 6, 118, 16, 44, WORKLOAD["C"], seed=1))
    og = torch.randn(1, WORKLOAD["C"], 128, 128, device=dev)

    def step():
        d_ = depth.detach().requires_grad_(True)
        f_ = feat.detach().requires_grad_(True)
        rcb.voxel_pooling_v2_from_calib((cam, bda), axes, d_, f_, lo, iv, sz).backward(og)
    t_eager = _time_cuda(torch, step, n_warm=5, n=50)
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        step()
    t_graph = _time_cuda(torch, graph.replay, n_warm=5, n=50)
    return {"eager_ms": round(t_eager, 4), "cuda_graph_ms": round(t_graph, 4),
            "samples_per_s": round(1 / (t_graph * 1e-3), 1),
            "note": "prepare (from calibration) + forward + backward of ONE sample; samples_per_s is the graph replay"}


def extra_configs(torch, rcb, rig, _lib, bp, dev, lib):
    """BASELINE configs 3, 4, 5 on one GPU, each through the public device-side API
    (rcbevdet_b200.voxel_pooling_v2 / radar_rcs_scatter), 10 timed calls after 3 warm-ups."""
    C = WORKLOAD["C"]
    out = {}
    try:
        # config 3: 8 samples x 8 frames folded into the batch dimension (2^20 cells); only the key
        # frame has a backward in the reference (bevdet_rc.py:756-776), i.e. config 2's backward
        t8 = _pool_case(torch, rcb, rig, dev, 8, rig.R50_GRID, rig.R50_INPUT, C, frames=8, backward=False)
        t8["frames_per_s"] = round(t8["samples_folded"] / (t8["prepare_fwd_ms"] * 1e-3), 1)
        t8["note"] = "samples_per_s counts 8-frame samples (prepare + forward of all 64 folded frames)"
        out["temporal_8f"] = t8
        # config 5: 900x1600 -> 56x100 features, 256x256 BEV, batch sweep
        for b in (1, 2, 4, 8):
            out[f"hires_256_B{b}"] = _pool_case(torch, rcb, rig, dev, b, rig.HIRES_GRID, rig.HIRES_INPUT, C, chain=b <= 2)
            torch.cuda.empty_cache()
        out["r50_B1_latency"] = _graph_case(torch, rcb, rig, dev)
        # config 4: RCS-aware radar scatter, B=8, 5 sweeps x 5 radars x 125 points per sample
        out["radar_128"] = _radar_case(torch, rcb, rig, dev, 8, 128)
        out["radar_512"] = _radar_case(torch, rcb, rig, dev, 8, 512)
    except Exception as e:  # the headline line must survive a failing side measurement
        out["error"] = repr(e)
    return out


def reference_gpu(torch, rcb, rig, dev_set, lo, iv, sz, dev):
    """The reference's own GPU implementation on config 2's inputs, as a checker-side timing (outside
    `value`): its CUDA op compiled unmodified into oracle/_ref (absent -> null) driven by its autograd
    host sequence, and its prepare chain of torch ops (oracle/torch_ref.py restates it op for op;
    /root/reference itself does not exist on the GPU box)."""
    try:
        from oracle import ref_cuda, torch_ref
        coor, depth, feat, out_grad = dev_set
        B, C = coor.shape[0], feat.shape[2]
        fview = feat.permute(0, 1, 3, 4, 2)
        lo_d, iv_d, sz_d = (t.to(dev) for t in (lo, iv, sz))
        res = {}
        t_prep = _time_cuda(torch, lambda: torch_ref.prepare(coor, lo_d, iv_d, sz_d), n_warm=2, n=5)
        res["prepare_torch_ms"] = round(t_prep, 3)
        if not ref_cuda.available():
            res["pool"] = None
            res["note"] = "oracle/_ref not built on this box: only the torch prepare chain was timed"
            return res
        rb, rd, rf, st, ln = torch_ref.prepare(coor, lo_d, iv_d, sz_d)
        shape = (B, 1, 128, 128, C)
        og = out_grad.view(B, C, 1, 128, 128).permute(0, 2, 3, 4, 1)

        def fwd_bwd():
            ref_cuda.bev_pool_v2(depth, fview, rd, rf, rb, shape, st, ln)
            ref_cuda.backward(og.contiguous(), depth, fview.contiguous(), rd, rf, rb)
        t_pool = _time_cuda(torch, fwd_bwd, n_warm=2, n=5)
        res.update(fwd_bwd_ms=round(t_pool, 3), with_prepare_ms=round(t_pool + t_prep, 3),
                   samples_per_s=round(B / ((t_pool + t_prep) * 1e-3), 1),
                   note="reference kernels recompiled for sm_100a (bev_pool_cuda.cu, unmodified) + the host-side "
                        "sequence of bev_pool.py; prepare = view_transformer.py:207-265 as torch ops on the GPU")
        return res
    except Exception as e:
        return {"error": repr(e)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's compiled restatement of the reference path, all host threads
# --------------------------------------------------------------------------------------------
def _cpu_step_fn(samples, threads):
    import numpy as np
    import torch

    from oracle import oracle
    from rcbevdet_b200 import rig

    coor, depth, feat, og = make_inputs(torch, rig, "cpu", samples, seed=1)
    lo, iv, sz = (t.numpy() for t in rig.grid_tensors(rig.R50_GRID))
    coor, depth = coor.numpy(), depth.numpy()
    feat_v = feat.permute(0, 1, 3, 4, 2)
    og_v = og.view(samples, WORKLOAD["C"], 1, 128, 128).permute(0, 2, 3, 4, 1)
    shape = (samples, 1, 128, 128, WORKLOAD["C"])

    def step():
        rb, rd, rf, st, ln = oracle.voxel_pooling_prepare_v2_c(coor, lo, iv, sz, threads=threads)
        rows = np.ascontiguousarray(feat_v.numpy())                              # bev_pool.py:21
        out = oracle.bev_pool_v2_forward(depth, rows, rd, rf, rb, shape, st, ln, threads=threads)
        bev = oracle.to_bczyx(out)                                               # bev_pool.py:91
        g_rows = np.ascontiguousarray(og_v.numpy())                              # bev_pool.py:69
        dg, fg = oracle.bev_pool_v2_backward(g_rows, depth, rows, rd, rf, rb, threads=threads)
        return bev, dg, fg

    return step


def cpu_baseline(budget_s=15.0):
    threads = os.cpu_count() or 1
    step = _cpu_step_fn(1, threads)
    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 200:
            break
    res = {"value": round(n / dt, 2), "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{n} single-sample steps (prepare + fwd + bwd, same geometry) in {dt:.1f} s; "
                     "oracle/bevpool_oracle.c with OpenMP"}
    res["torch"] = cpu_torch_baseline(threads, budget_s=min(10.0, budget_s))
    return res


def cpu_torch_baseline(threads, budget_s=10.0):
    """north_star's "reference CPU PyTorch path": the reference's prepare chain as torch ops
    (view_transformer.py:207-265) + the index_add_ restatement of its pool kernel + autograd, on the
    host cores (BASELINE config 1: one sample)."""
    import torch

    from oracle import torch_ref
    from rcbevdet_b200 import rig
    torch.set_num_threads(threads)
    coor, depth, feat, og = make_inputs(torch, rig, "cpu", 1, seed=1)
    lo, iv, sz = rig.grid_tensors(rig.R50_GRID)
    shape = (1, 1, 128, 128, WORKLOAD["C"])

    def step():
        rb, rd, rf, st, ln = torch_ref.prepare(coor, lo, iv, sz)
        d = depth.detach().requires_grad_(True)
        f = feat.detach().requires_grad_(True)
        torch_ref.pool(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, shape).backward(og)
    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 100:
            break
    return {"value": round(n / dt, 2), "unit": UNIT, "threads": torch.get_num_threads(),
            "sample": f"{n} single-sample steps in {dt:.1f} s (torch prepare chain + index_add_ pool + autograd)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bound the run: samples per step chosen so K + W steps end within ~2 minutes
    probe = _cpu_step_fn(1, threads)
    probe()
    t0 = time.perf_counter()
    probe()
    t1 = time.perf_counter() - t0
    budget = 120.0
    per_step = max(1, min(WORKLOAD["batch_per_gpu"], int(budget / max(args.steps + args.warmup, 1) / max(t1, 1e-6))))
    step = _cpu_step_fn(per_step, threads) if per_step != 1 else probe
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = (f"{per_step} sample(s) per step x {args.steps} steps; prepare + fwd + bwd of the same geometry; "
              "CPU restatement of the reference path (oracle/bevpool_oracle.c, OpenMP)")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(WORKLOAD, parallelism="host CPU threads", samples_per_step=per_step),
            "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="timed steps only (no variants, minimal e2e): the command line used under ncu")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line: libraries that print to file descriptor 1 (NCCL's version
    # banner does, whatever NCCL_DEBUG_FILE says) are sent to stderr, Python keeps the real stdout
    sys.stdout.flush()
    sys.stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
