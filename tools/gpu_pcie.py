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
