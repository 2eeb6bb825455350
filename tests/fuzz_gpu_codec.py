#!/usr/bin/env python3
"""Randomised parity run on the GPU box: random shapes through the packed encode / decode entry points against the
oracle, for a given number of seconds. Prints the number of cases; any mismatch raises.
    python tests/fuzz_gpu_codec.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # (refcodec)
import __graft_entry__ as graft  # noqa: E402
import refcodec  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = graft.load_package()
graft.build_oracle()
oracle = refcodec.OracleLib()
rng = np.random.default_rng(seed)
ctxs = {t: pkg.BatchContext(pkg.coders_library().coder(t), eos_padding=0xFF, device=0) for t in ("hpack", "test")}
tables = {t: oracle.table(*refcodec.table_arrays(t)) for t in ("hpack", "test")}


def lengths(n):
    kind = rng.integers(0, 6)
    if kind == 0:
        return rng.integers(0, 300, size=n)
    if kind == 1:
        return np.minimum(rng.exponential(40, size=n).astype(np.int64), 5000)
    if kind == 2:
        l = rng.integers(0, 40, size=n)
        l[rng.integers(0, n, size=max(1, n // 50))] = rng.integers(1000, 30000, size=max(1, n // 50))
        return l
    if kind == 3:
        return np.where(rng.random(n) < 0.5, 0, rng.integers(1, 20, size=n))
    if kind == 4:
        return rng.integers(180, 1200, size=n)
    return np.full(n, int(rng.integers(1, 200)))


def same(got, want, what):
    for k in ("out_offsets", "status", "out_lens", "consumed"):
        if k in want and k in got:
            assert np.array_equal(got[k], want[k]), "%s: %s differs" % (what, k)
    total = int(want["out_offsets"][-1])
    assert np.array_equal(got["out"][:total], want["out"][:total]), "%s: bytes differ" % what


t0 = time.time()
cases = 0
while time.time() - t0 < budget:
    t = "hpack" if rng.random() < 0.7 else "test"
    ctx, table = ctxs[t], tables[t]
    if rng.random() < 0.2:
        n, lens = 1, np.array([int(rng.integers(1, 3_000_000))])
    else:
        n = int(rng.integers(1, 6000))
        lens = lengths(n)
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    total = int(offs[-1])
    if rng.random() < 0.75:
        data = refcodec.zipf_symbol_sampler(refcodec.table_arrays(t)[1], s=float(rng.choice([0.8, 1.0, 1.5, 2.0])))[
            rng.integers(0, 65536, size=total)]
    else:
        data = rng.integers(0, 256, size=total, dtype=np.uint8)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    cap = 4 * total + 64
    want = oracle.encode_batch(table, 0xFF, data, offs, cap)
    got = ctx.encode(data, offs, cap)
    same(got, want, "encode case %d (%s, n=%d)" % (cases, t, n))
    enc_total = int(want["out_offsets"][-1])
    stream = want["out"][:enc_total].copy()
    if rng.random() < 0.3 and enc_total:
        # damage: flip bits / truncate items by decoding with shifted offsets
        idx = rng.integers(0, enc_total, size=max(1, enc_total // 200))
        stream[idx] ^= rng.integers(1, 256, size=len(idx), dtype=np.uint8)
    dcap = 8 * enc_total // (5 if t == "hpack" else 1) + 64
    want_d = oracle.decode_batch(table, stream, want["out_offsets"], dcap)
    got_d = ctx.decode(stream, want["out_offsets"], dcap)
    same(got_d, want_d, "decode case %d (%s, n=%d)" % (cases, t, n))
    for k in ("leftover_working_bits", "leftover_num_bits"):
        assert np.array_equal(got_d[k], want_d[k]), "decode case %d: %s" % (cases, k)
    cases += 1
print("gpu_fuzz: %d cases in %.0f s, all equal to the oracle (seed %d)" % (cases, time.time() - t0, seed))
