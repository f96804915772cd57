"""Measurement: host<->device copy rates on this box as the end-to-end loop sees them -- by copy size, one direction
and both, with and without kernels running next to the copies (FP32-bound and HBM-bound), pinned host memory.

  python tools/gpu_pcie2.py
"""
import json, sys, time
import numpy as np
import torch

dev = torch.device("cuda", 0)
s_up, s_dn, s_k = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
big_a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
big_b = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
m1 = torch.randn(8192, 8192, device=dev, dtype=torch.float32)
m2 = torch.randn(8192, 8192, device=dev, dtype=torch.float32)


def load(kind, stop):
    """Keeps kernels running on s_k until `stop` is set on the host (a bounded number is queued at a time)."""
    with torch.cuda.stream(s_k):
        if kind == "hbm":
            for _ in range(40):
                big_b.copy_(big_a)
        elif kind == "fp32":
            for _ in range(12):
                torch.mm(m1, m2)


def run(mib, up, down, kind, reps):
    n = mib << 20
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    h_a.zero_(); h_b.zero_()
    torch.cuda.synchronize()
    if kind:
        load(kind, None)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    if up:
        e[0].record(s_up)
        with torch.cuda.stream(s_up):
            for _ in range(reps):
                d_a.copy_(h_a, non_blocking=True)
        e[1].record(s_up)
    if down:
        e[2].record(s_dn)
        with torch.cuda.stream(s_dn):
            for _ in range(reps):
                h_b.copy_(d_b, non_blocking=True)
        e[3].record(s_dn)
    s_up.synchronize(); s_dn.synchronize()
    out = {}
    if up:
        out["h2d_gbs"] = round(reps * n / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9, 1)
    if down:
        out["d2h_gbs"] = round(reps * n / (e[2].elapsed_time(e[3]) * 1e-3) / 1e9, 1)
    torch.cuda.synchronize()
    return out


for kind in (None, "fp32", "hbm"):
    for mib in (1, 4, 16, 64, 256):
        reps = max(2, 512 // mib)
        row = {"kernels": kind or "none", "mib": mib, "reps": reps}
        row["alone"] = {**run(mib, True, False, kind, reps), **run(mib, False, True, kind, reps)}
        row["duplex"] = run(mib, True, True, kind, reps)
        print(json.dumps(row), flush=True)
