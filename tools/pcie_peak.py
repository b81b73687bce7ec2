"""Measured PCIe ceiling for the host twin (e2e): pinned H2D alone, D2H alone, both directions at once (two streams),
800 MB each way like one 1e8-DOF apply.  python tools/pcie_peak.py"""
import time
import torch

n = 100_160_064
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 8 / 1e9


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("both at once", True, True)):
    run(a, b, 2)
    dt = run(a, b)
    print("%-13s %.2f ms per 801 MB pass -> %.1f GB/s per direction" % (name, dt * 1e3, gb / dt))
