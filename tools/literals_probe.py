#!/usr/bin/env python3
"""HPACK string literals at the size of BASELINE config 2 (1M strings of 8-256 B), device pointers:
one JSON line with the rates (profiles/r1_literals.json)."""
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
n = 1_000_000
sampler_t = torch.from_numpy(refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])).to(dev)
lens = bench.string_lengths_torch(bench.SEED_BATCH, 0, n, dev)
off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
off[1:] = torch.cumsum(lens, 0)
total = int(off[-1].item())
raw = bench.symbols_torch(bench.SEED_BATCH, 0, total, sampler_t, dev)
ctx = pkg.BatchContext(pkg.coders_library().coder("hpack"), eos_padding=0xFF, device=0)
framed = torch.empty(total + total // 2 + 64, dtype=torch.uint8, device=dev)
f_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
out = torch.empty(total + 64, dtype=torch.uint8, device=dev)
o_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
status = torch.zeros(n, dtype=torch.int32, device=dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
enc_ms, dec_ms = [], []
framed_size = 0
for rep in range(8):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(stream)
    ctx.hpack_device(True, n, raw, off, total, framed, framed.numel(), f_off, mode=0, stream=stream.cuda_stream)
    ev[1].record(stream)
    framed_size = int(f_off[-1].item())
    ev[2].record(stream)
    ctx.hpack_device(False, n, framed, f_off, framed_size, out, out.numel(), o_off, status=status, stream=stream.cuda_stream)
    ev[3].record(stream)
    torch.cuda.synchronize(dev)
    if rep >= 3:
        enc_ms.append(ev[0].elapsed_time(ev[1]))
        dec_ms.append(ev[2].elapsed_time(ev[3]))
assert int(o_off[-1].item()) == total and torch.equal(out[:total], raw) and int(status.abs().sum().item()) == 0
e, d = float(np.mean(enc_ms)), float(np.mean(dec_ms))
print(json.dumps({"workload": "hpack string literals (RFC 7541 5.2), %d strings of 8-256 B, SMALLEST mode, device pointers" % n,
                  "raw_bytes": total, "framed_bytes": framed_size, "encode_ms": e, "decode_ms": d,
                  "encode_gbs": (total + framed_size) / e / 1e6, "decode_gbs": (total + framed_size) / d / 1e6}))
