#!/usr/bin/env python3
"""Measures pinned host<->device copy bandwidth on this box: H2D alone, D2H alone, both at once.
The e2e number of bench.py is bounded by these (every payload byte crosses PCIe once each way)."""
import json
import time

import torch

n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return (int(h2d) + int(d2h)) * n / dt / 1e9


for _ in range(2):
    run(True, True, 1)
print(json.dumps({"h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "both_total_gbs": run(True, True)}))
