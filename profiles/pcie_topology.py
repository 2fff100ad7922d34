#!/usr/bin/env python3
"""Host <-> device bandwidth of every GPU of the node alone and of all of them at once (one thread per GPU):
what bounds the end-to-end number at N > 1.   python profiles/pcie_topology.py > profiles/rN/pcie_topology.txt"""
import subprocess
import threading
import time

import torch

N = torch.cuda.device_count()
SIZE = 512 << 20
h = [torch.empty(SIZE, dtype=torch.uint8).pin_memory() for _ in range(N)]
d = [torch.empty(SIZE, dtype=torch.uint8, device=f"cuda:{i}") for i in range(N)]


def run(i, h2d, reps, out, both=False):
    torch.cuda.set_device(i)
    s2 = torch.cuda.Stream()
    d2 = torch.empty(SIZE // 2, dtype=torch.uint8, device=f"cuda:{i}") if both else None
    h2 = torch.empty(SIZE // 2, dtype=torch.uint8).pin_memory() if both else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            d[i].copy_(h[i], non_blocking=True)
        else:
            h[i].copy_(d[i], non_blocking=True)
        if both:                                            # the other direction at the same time (JSON text coming back)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    out[i] = SIZE * reps / (time.perf_counter() - t0) / 1e9


def together(ids, h2d, both=False):
    out = {}
    th = [threading.Thread(target=run, args=(i, h2d, 10, out, both)) for i in ids]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return out


print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
for i in range(N):
    together([i], True)
print("H2D alone      GB/s:", {i: round(together([i], True)[i], 1) for i in range(N)})
print("D2H alone      GB/s:", {i: round(together([i], False)[i], 1) for i in range(N)})
if N > 1:
    for k in sorted({2, 4, N} & set(range(2, N + 1))):
        r = together(list(range(k)), True)
        print(f"H2D {k} at once  GB/s:", {i: round(v, 1) for i, v in sorted(r.items())}, "sum", round(sum(r.values()), 1))
    r = together(list(range(N)), False)
    print(f"D2H {N} at once  GB/s:", {i: round(v, 1) for i, v in sorted(r.items())}, "sum", round(sum(r.values()), 1))
    r = together(list(range(N)), True, both=True)
    print(f"H2D {N} at once with D2H of half the bytes beside it  GB/s (H2D):", {i: round(v, 1) for i, v in sorted(r.items())},
          "sum", round(sum(r.values()), 1))
