"""bench.py's counter-based synthetic data: the device (torch) and host (numpy) generators agree, and
the workload has the shape SURVEY.md 8(d) states (8..256 B strings, ratio ~0.73 with the HPACK table)."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

import bench
import refcodec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_numpy_and_torch_generators_agree():
    sampler = refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])
    for first in (0, 12345, (3 << 40) + 17):
        a = bench.string_lengths_np(bench.SEED_BATCH, first, 5000)
        b = bench.string_lengths_torch(bench.SEED_BATCH, first, 5000, "cpu").numpy()
        assert np.array_equal(a, b) and a.min() >= 8 and a.max() <= 256
        x = bench.symbols_np(bench.SEED_STREAM, first, 70000, sampler)
        y = bench.symbols_torch(bench.SEED_STREAM, first, 70000, torch.from_numpy(sampler), "cpu", chunk=9999).numpy()
        assert np.array_equal(x, y)


def test_workload_shape(oracle, oracle_tables):
    data, offs, _ = bench.cpu_sample("hpack_batch", 20000)
    lens = np.diff(offs.astype(np.int64))
    assert lens.min() == 8 and lens.max() == 256 and abs(lens.mean() - 132) < 2
    enc = oracle.encode_batch(oracle_tables["hpack"], 0xFF, data, offs, 4 * len(data))
    ratio = int(enc["out_offsets"][-1]) / len(data)
    assert 0.70 < ratio < 0.77, ratio


def test_reference_arm_prints_one_json_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                                   "--steps", "1", "--warmup", "0", "--strings", "4000"], text=True)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
