"""Dev/measurement: pinned host <-> device copy bandwidth at the bench's transfer sizes (is e2e PCIe-bound?)."""
import torch, subprocess
dev = torch.device("cuda", 0)
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout.strip())
for mb in (0.25, 1.0, 2.6, 2.9, 16.0):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, f in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        print(f"{name} {mb:5.2f} MB: {us:7.1f} us  {n / us / 1e3:6.1f} GB/s")

# ---- do the two directions overlap?  H2D on one stream and D2H on another, at the bench's per-event sizes ----
import time
s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
nu, nd = int(2.6e6), int(2.9e6)
hu, du = torch.empty(nu, dtype=torch.uint8).pin_memory(), torch.empty(nu, dtype=torch.uint8, device=dev)
hd, dd = torch.empty(nd, dtype=torch.uint8).pin_memory(), torch.empty(nd, dtype=torch.uint8, device=dev)


def both(n):
    for _ in range(n):
        with torch.cuda.stream(s_up):
            du.copy_(hu, non_blocking=True)
        with torch.cuda.stream(s_dn):
            hd.copy_(dd, non_blocking=True)


both(5); torch.cuda.synchronize()
t0 = time.perf_counter(); both(200); torch.cuda.synchronize(); us = (time.perf_counter() - t0) * 1e6 / 200
print(f"H2D 2.6 MB + D2H 2.9 MB on two streams: {us:7.1f} us per pair  {(nu + nd) / us / 1e3:6.1f} GB/s combined")

# ---- the same bytes as the bench's per-event copies: 8 uploads (4 of them tiny) and 4 downloads, vs one copy each ----
up_sizes = [128, 8, 36, 48, 950_000, 16_388, 1_630_000, 16_388]
dn_sizes = [262_144, 1_640_000, 410_000, 620_000]
hus = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in up_sizes]
dus = [torch.empty(n, dtype=torch.uint8, device=dev) for n in up_sizes]
hds = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in dn_sizes]
dds = [torch.empty(n, dtype=torch.uint8, device=dev) for n in dn_sizes]


def split(n, up=True, dn=True):
    for _ in range(n):
        if up:
            with torch.cuda.stream(s_up):
                for h, d in zip(hus, dus):
                    d.copy_(h, non_blocking=True)
        if dn:
            with torch.cuda.stream(s_dn):
                for h, d in zip(hds, dds):
                    h.copy_(d, non_blocking=True)


for name, kw in (("8 uploads", dict(dn=False)), ("4 downloads", dict(up=False)), ("8 uploads + 4 downloads on two streams", {})):
    split(5, **kw); torch.cuda.synchronize()
    t0 = time.perf_counter(); split(200, **kw); torch.cuda.synchronize(); us = (time.perf_counter() - t0) * 1e6 / 200
    print(f"{name}: {us:7.1f} us per event")
