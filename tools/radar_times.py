"""Kernel-level timing of the radar RCS scatter (BASELINE config 4) on the GPU box:
    ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none python tools/radar_times.py
or plain: prints the CUDA-event time of the C-ABI call (no autograd wrapper) at 128^2 and 512^2."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rcbevdet_b200 import _lib, rig  # noqa: E402
from rcbevdet_b200.radar import _desc  # noqa: E402

lib = _lib.lib()
for n in (128, 512):
    pf, rcs, coors = (t.cuda() for t in rig.radar_pillars(8, n, n, seed=4))
    d = _desc(pf, rcs, 8, n, n)
    f = torch.empty((8, 64, n, n), device="cuda")
    h = torch.empty((8, n, n), device="cuda")
    hf = torch.empty((8, 1, n, n), device="cuda")
    ws_bytes = lib.rcb_radar_workspace_bytes(ctypes.byref(d))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    dev = pf.device

    def call():
        _lib.check(lib.rcb_radar_rcs_scatter(ctypes.byref(d), _lib.ptr(pf), _lib.ptr(rcs), _lib.ptr(coors), _lib.ptr(f),
                                             _lib.ptr(h), _lib.ptr(hf), _lib.ptr(ws), ws_bytes, dev.index,
                                             _lib.stream_ptr(dev)), "radar")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        call()
    e1.record()
    torch.cuda.synchronize()
    print(f"radar {n}x{n}: V={pf.shape[0]} C-ABI call {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
