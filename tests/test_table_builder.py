"""Code tables from data: histogram -> optimal length-limited lengths -> canonical codes ->
a working coder. The reference has no table builder, so the checks are the published algorithms written a
second time (oracle/table_builder_oracle.py) and the properties a code table must have.

CPU part: lengths and codes (host code). GPU part: the histogram kernel and the whole chain through the codec."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

import refcodec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("table_builder_oracle", os.path.join(ROOT, "oracle", "table_builder_oracle.py"))
tbo = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(tbo)


@pytest.fixture(scope="module")
def builder(pkg):
    return pkg.TableBuilder()


def _cost(counts, lengths):
    return int(sum(int(c) * int(l) for c, l in zip(counts, lengths)))


def _kraft(lengths, eos=0):
    return sum(2.0 ** -int(l) for l in list(lengths) + [eos] if l)


def test_oracle_package_merge_is_optimal_on_tiny_alphabets():
    rng = np.random.default_rng(1)
    for _ in range(40):
        n = int(rng.integers(2, 6))
        w = [int(x) for x in rng.integers(1, 50, size=n)]
        for limit in range(max(1, int(np.ceil(np.log2(n)))), 5):
            lens = tbo.package_merge([(x, i) for i, x in enumerate(w)], limit)
            assert max(lens.values()) <= limit
            assert sum(w[i] * lens[i] for i in range(n)) == tbo.brute_force_cost(w, limit)
    w = [int(x) for x in rng.integers(1, 1000, size=200)]
    lens = tbo.package_merge([(x, i) for i, x in enumerate(w)], 32)
    assert sum(w[i] * lens[i] for i in range(200)) == tbo.huffman_cost(w)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("max_bits", [8, 9, 12, 16, 32])
def test_lengths_match_package_merge(builder, seed, max_bits):
    rng = np.random.default_rng(seed)
    kind = seed % 3
    if kind == 0:
        counts = rng.integers(0, 1000, size=256)
    elif kind == 1:
        counts = (1e9 / np.arange(1, 257) ** 1.5).astype(np.int64)[rng.permutation(256)]
    else:
        counts = np.zeros(256, dtype=np.int64)
        counts[rng.choice(256, size=40, replace=False)] = rng.integers(1, 1 << 40, size=40)
    counts = counts.astype(np.uint64)
    for cover_all in (True, False):
        lens, _ = builder.lengths(counts, max_bits, cover_all=cover_all)
        present = [s for s in range(256) if counts[s] or cover_all]
        if max_bits == 8 and len(present) > 256:
            continue
        assert all(lens[s] > 0 for s in present) and all(lens[s] == 0 for s in range(256) if s not in present)
        assert lens.max() <= max_bits
        assert _kraft(lens) <= 1.0 + 1e-12
        # the same cost as the independent package-merge on the same (scaled) weights
        scaled = [((int(counts[s]) << 16) | 1, s) for s in present]
        want = tbo.package_merge(scaled, max_bits)
        assert sum(w * int(lens[s]) for w, s in scaled) == sum(w * want[s] for w, s in scaled)
        if max_bits == 32 and not cover_all:
            assert _cost(counts, lens) == tbo.huffman_cost([int(c) for c in counts])


def test_degenerate_alphabets(builder, pkg):
    lens, _ = builder.lengths(np.zeros(256, np.uint64), 16, cover_all=False)
    assert not lens.any()
    one = np.zeros(256, np.uint64)
    one[65] = 7
    lens, _ = builder.lengths(one, 16, cover_all=False)
    assert lens[65] == 1 and lens.sum() == 1
    lens, eos = builder.lengths(one, 16, cover_all=False, reserve_eos=True)
    assert lens[65] == 1 and eos == 1
    with pytest.raises(pkg.CodecError):
        builder.lengths(np.ones(256, np.uint64), 7, cover_all=True)  # 256 symbols need 8 bits
    lens, _ = builder.lengths(np.ones(256, np.uint64), 8, cover_all=True)
    assert (lens == 8).all()


def test_canonical_codes_and_eos(builder):
    counts = (1e8 / np.arange(1, 257) ** 1.3).astype(np.uint64)
    lens, eos_len = builder.lengths(counts, 30, cover_all=True, reserve_eos=True)
    table, eos = builder.canonical(lens, eos_len)
    want, want_eos = tbo.canonical([int(x) for x in lens], eos_len)
    for s in range(256):
        assert (table[s].pattern, table[s].num_bits) == want[s]
    assert (eos.pattern, eos.num_bits) == want_eos
    assert eos.num_bits == lens.max() and eos.pattern == (1 << eos.num_bits) - 1, "EOS is the all-ones code"
    assert abs(_kraft(lens, eos_len) - 1.0) < 1e-12
    # prefix-free: no code is a prefix of another
    codes = sorted((format(table[s].pattern, "0%db" % table[s].num_bits) for s in range(256)))
    assert all(not b.startswith(a) for a, b in zip(codes, codes[1:]))
    # Kraft violations are refused
    bad = np.full(256, 7, dtype=np.uint8)
    with pytest.raises(Exception):
        builder.canonical(bad, 0)


def test_table_coder_and_def_file_round_trip(builder, pkg, product, tmp_path):
    """The built table drives the streaming API (run-time coder) and goes through the generator as a .def."""
    rng = np.random.default_rng(3)
    text = rng.choice(np.frombuffer(b"etaoin shrdlu,.-/0123456789", dtype=np.uint8), size=5000)
    counts = np.bincount(text, minlength=256).astype(np.uint64)
    table, _ = builder.codes(counts, 16, cover_all=True)
    coder = builder.table_coder(table)
    try:
        capi = pkg.capi
        L = product.lib
        enc = capi.aws_huffman_encoder()
        L.aws_huffman_encoder_init(C.byref(enc), C.byref(coder.coder))
        cur = capi.aws_byte_cursor(len(text), text.ctypes.data)
        out = np.zeros(2 * len(text), dtype=np.uint8)
        buf = capi.aws_byte_buf(0, out.ctypes.data, len(out), None)
        assert L.aws_huffman_encode(C.byref(enc), C.byref(cur), C.byref(buf)) == 0
        bits = sum(int(counts[s]) * table[s].num_bits for s in range(256))
        assert buf.len == (bits + 7) // 8 and buf.len < len(text) * 5 // 8
        dec = capi.aws_huffman_decoder()
        L.aws_huffman_decoder_init(C.byref(dec), C.byref(coder.coder))
        cur2 = capi.aws_byte_cursor(buf.len, out.ctypes.data)
        back = np.zeros(len(text) + 8, dtype=np.uint8)
        buf2 = capi.aws_byte_buf(0, back.ctypes.data, len(text), None)
        assert L.aws_huffman_decode(C.byref(dec), C.byref(cur2), C.byref(buf2)) == 0
        assert buf2.len == len(text) and np.array_equal(back[:len(text)], text)
    finally:
        builder.table_coder_clean_up(coder)
    # .def -> generator -> C source that compiles
    path = str(tmp_path / "built.def")
    builder.write_def(table, path)
    lines = [l for l in open(path) if l.startswith("HUFFMAN_CODE(")]
    assert len(lines) == 256
    out_c = str(tmp_path / "built_coder.c")
    subprocess.run([pkg._build.GENERATOR, path, out_c, "built"], check=True)
    assert "built_get_coder" in open(out_c).read()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("size", [0, 1, 15, 16, 17, 4097, 1_000_003])
def test_histogram_matches_bincount(builder, size):
    rng = np.random.default_rng(size)
    sampler = refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])
    data = sampler[rng.integers(0, 65536, size=size)] if size else np.zeros(0, np.uint8)
    got = builder.histogram(data)
    assert np.array_equal(got, np.bincount(data, minlength=256).astype(np.uint64))


@pytest.mark.gpu
def test_histogram_device_unaligned_and_skewed(builder):
    import torch
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(9)
    host = rng.integers(0, 256, size=3_000_000, dtype=np.uint8)
    host[::3] = 0x61  # one value dominates
    whole = torch.from_numpy(host).to(dev)
    counts = torch.zeros(256, dtype=torch.int64, device=dev)
    for lead, cut in ((0, 0), (1, 0), (7, 5), (15, 1)):
        view = whole[lead:len(host) - cut]
        builder.histogram_device(view, counts)
        torch.cuda.synchronize(dev)
        want = np.bincount(host[lead:len(host) - cut], minlength=256)
        assert np.array_equal(counts.cpu().numpy(), want)


@pytest.mark.gpu
def test_data_to_table_to_codec(builder, pkg, oracle):
    """histogram (GPU) -> table -> batched codec context from the table -> encode/decode; compared with the oracle
    driven by the same table; and the built table beats the fixed HPACK table on this data."""
    rng = np.random.default_rng(21)
    data, offs = refcodec.random_batch(rng, 4000, 0, 300, "hpack")
    counts = builder.histogram(data)
    table, _ = builder.codes(counts, 24, cover_all=True)
    ctx = pkg.BatchContext(None, eos_padding=0xFF, device=0, code_table=table)
    try:
        patterns = np.array([table[s].pattern for s in range(256)], dtype=np.uint32)
        num_bits = np.array([table[s].num_bits for s in range(256)], dtype=np.uint8)
        otab = oracle.table(patterns, num_bits)
        cap = 4 * len(data) + 64
        want = oracle.encode_batch(otab, 0xFF, data, offs, cap)
        got = ctx.encode(data, offs, cap)
        total = int(want["out_offsets"][-1])
        assert np.array_equal(got["out_offsets"], want["out_offsets"]) and np.array_equal(got["out"][:total], want["out"][:total])
        back = ctx.decode(got["out"][:total], got["out_offsets"], len(data) + 64)
        assert np.array_equal(back["out"][:len(data)], data) and np.array_equal(back["out_offsets"], offs)
        hp = oracle.table(*refcodec.table_arrays("hpack"))
        hp_total = int(oracle.encode_batch(hp, 0xFF, data, offs, cap)["out_offsets"][-1])
        assert total <= hp_total
    finally:
        ctx.close()
