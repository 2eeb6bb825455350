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


def test_parity_gates_accept_the_oracle_and_reject_a_flipped_bit(oracle, oracle_tables):
    """bench.py's parity gates (the reference on the same inputs) on outputs made by the oracle: equal as they
    are, different after one flipped bit — in a string's bytes, in an offset, and in the middle of a stream whose
    segments start at every bit phase."""
    data, offs, _ = bench.cpu_sample("hpack_batch", 3000)
    enc = oracle.encode_batch(oracle_tables["hpack"], 0xFF, data, offs, 4 * len(data))
    total = int(enc["out_offsets"][-1])
    out = enc["out"][:total].copy()
    ok = bench.parity_batch(data, offs, out, enc["out_offsets"], 3)
    assert ok["encoded_bytes_equal"] and ok["encoded_offsets_equal"] and ok["reference_decodes_gpu_bytes_to_input"]
    assert ok["strings"] == 3000 and ok["encoded_bytes"] == total
    bad = out.copy()
    bad[total // 2] ^= 0x10
    assert not bench.parity_batch(data, offs, bad, enc["out_offsets"], 3)["encoded_bytes_equal"]

    stream, soff, _ = bench.cpu_sample("stream", 300_001)
    senc = oracle.encode_batch(oracle_tables["hpack"], 0xFF, stream, soff, 4 * len(stream))
    sbytes = int(senc["out_offsets"][-1])
    sout = senc["out"][:sbytes + 8].copy()
    good = bench.parity_stream(stream, sout, sbytes, 4, seg_bytes=7001, decode_prefix=5000)
    assert good["encoded_bytes_equal"] and good["reference_decodes_gpu_prefix_to_input"] and good["segments"] == 43
    for pos in (0, sbytes // 3, sbytes - 1):
        for bit in (0x80, 0x01):
            bad = sout.copy()
            bad[pos] ^= bit
            assert not bench.parity_stream(stream, bad, sbytes, 4, seg_bytes=7001, decode_prefix=5000)["encoded_bytes_equal"], (pos, bit)
    assert not bench.parity_stream(stream, sout, sbytes + 1, 2, seg_bytes=7001, decode_prefix=5000)["encoded_bytes_equal"]
