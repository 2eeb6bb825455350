#!/usr/bin/env python3
"""Randomised parity run for the HPACK string-literal entry points against oracle/hpack_literals_oracle.py.
    python tests/fuzz_gpu_literals.py [seconds] [seed]"""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
import refcodec  # noqa: E402

spec = importlib.util.spec_from_file_location("lit", os.path.join(ROOT, "oracle", "hpack_literals_oracle.py"))
lit = importlib.util.module_from_spec(spec)
spec.loader.exec_module(lit)

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = graft.load_package()
graft.build_oracle()
oracle = refcodec.OracleLib()
table = oracle.table(*refcodec.table_arrays("hpack"))
lo = lit.LiteralOracle(oracle, table)
ctx = pkg.BatchContext(pkg.coders_library().coder("hpack"), eos_padding=0xFF, device=0)
rng = np.random.default_rng(seed)
sampler = refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])

t0, cases = time.time(), 0
while time.time() - t0 < budget:
    n = int(rng.integers(1, 400))
    items = []
    for _ in range(n):
        kind = rng.random()
        size = int(rng.integers(0, 40)) if kind < 0.5 else int(rng.integers(0, 400)) if kind < 0.95 else int(rng.integers(2000, 9000))
        if rng.random() < 0.25:
            items.append(rng.integers(0, 256, size=size, dtype=np.uint8).tobytes())
        else:
            items.append(sampler[rng.integers(0, 65536, size=size)].tobytes())
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(x) for x in items])
    data = np.frombuffer(b"".join(items), dtype=np.uint8)
    mode = int(rng.integers(0, 3))
    want, want_offs = lo.encode_batch(data, offs, mode)
    got = ctx.hpack_encode_strings(data, offs, out_capacity=len(want) + 64, mode=mode)
    assert np.array_equal(got["out_offsets"], want_offs) and np.array_equal(got["out"][:len(want)], want), "encode case %d" % cases
    lits = [bytearray(want[int(want_offs[i]):int(want_offs[i + 1])]) for i in range(n)]
    for i in range(n):
        r = rng.random()
        if r < 0.06 and len(lits[i]) > 1:
            del lits[i][-1]
        elif r < 0.10:
            lits[i].append(int(rng.integers(0, 256)))
        elif r < 0.16 and len(lits[i]) > 1:
            lits[i][int(rng.integers(1, len(lits[i])))] ^= int(rng.integers(1, 256))
        elif r < 0.18:
            lits[i] = bytearray(rng.integers(0, 256, size=int(rng.integers(0, 12)), dtype=np.uint8).tobytes())
    f_offs = np.zeros(n + 1, dtype=np.uint64)
    f_offs[1:] = np.cumsum([len(x) for x in lits])
    framed = np.frombuffer(b"".join(bytes(x) for x in lits), dtype=np.uint8)
    if os.environ.get("FUZZ_PASSES") and cases % 2:
        os.environ["AWS_HUFFMAN_HPACK_PASSES"] = "1"
    else:
        os.environ.pop("AWS_HUFFMAN_HPACK_PASSES", None)
    w_out, w_offs, w_status = lo.decode_batch(framed, f_offs)
    g = ctx.hpack_decode_strings(framed, f_offs, out_capacity=8 * len(framed) // 5 + 4096)
    assert np.array_equal(g["status"], w_status), "decode case %d status %s" % (cases, np.flatnonzero(g["status"] != w_status)[:5])
    assert np.array_equal(g["out_offsets"], w_offs), "decode case %d offsets" % cases
    assert np.array_equal(g["out"][:len(w_out)], w_out), "decode case %d bytes" % cases
    cases += 1
print("gpu_fuzz_literals: %d cases in %.0f s, all equal to the oracle (seed %d)" % (cases, time.time() - t0, seed))
