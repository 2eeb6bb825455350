"""HPACK string literals (RFC 7541 section 5.2).

CPU part (-m "not gpu"): the literal oracle (oracle/hpack_literals_oracle.py) against RFC 7541: the prefix-integer
examples of Appendix C.1 and the string literals of Appendix C.2 / C.4 / C.6 (with their H bit and length).
GPU part: aws_hpack_string_encode_batch / aws_hpack_string_decode_batch through the C ABI against that oracle,
bit for bit, including malformed literals and every padding violation."""
import importlib.util
import os

import numpy as np
import pytest

import refcodec
from test_oracle_pins import RFC7541_C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("hpack_literals_oracle", os.path.join(ROOT, "oracle", "hpack_literals_oracle.py"))
lit = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(lit)

# RFC 7541 Appendix C.2.1 / C.2.2 / C.3.1: literals without Huffman coding
RFC7541_RAW = [
    (b"custom-key", "0a637573746f6d2d6b6579"),
    (b"custom-header", "0d637573746f6d2d686561646572"),
    (b"/sample/path", "0c2f73616d706c652f70617468"),
    (b"www.example.com", "0f7777772e6578616d706c652e636f6d"),
]


@pytest.fixture(scope="module")
def literal_oracle(oracle, oracle_tables):
    return lit.LiteralOracle(oracle, oracle_tables["hpack"])


def test_oracle_prefix_integers_rfc7541_c1():
    assert lit.encode_integer(10, 5) == bytes([0b01010])                              # C.1.1
    assert lit.encode_integer(1337, 5) == bytes([0b11111, 0b10011010, 0b00001010])     # C.1.2
    assert lit.encode_integer(42, 8) == bytes([42])                                    # C.1.3
    assert lit.decode_integer(bytes([0b11111, 0b10011010, 0b00001010]), 5) == (1337, 3)
    for v in (0, 1, 126, 127, 128, 254, 255, 300, 16383 + 127, 16384 + 127, 1 << 40):
        enc = lit.encode_integer(v, 7, 0x80)
        assert lit.decode_integer(enc, 7) == (v, len(enc)) and enc[0] & 0x80
    assert lit.decode_integer(b"", 7) == (None, lit.SHORT_BUFFER)
    assert lit.decode_integer(bytes([0x7f, 0x80]), 7) == (None, lit.SHORT_BUFFER)
    assert lit.decode_integer(bytes([0x7f] + [0xff] * 10 + [0x01]), 7) == (None, lit.INVALID_ARGUMENT)


def test_oracle_string_literals_rfc7541_appendix_c(literal_oracle):
    for text, hexed in RFC7541_C:   # C.4 / C.6: Huffman coded
        payload = bytes.fromhex(hexed)
        want = bytes([0x80 | len(payload)]) + payload
        assert literal_oracle.encode(text, lit.ALWAYS) == want
        # ("307" codes to three octets as well: only a strictly shorter Huffman form is chosen)
        assert literal_oracle.encode(text, lit.SMALLEST) == (want if len(payload) < len(text) else bytes([len(text)]) + text)
        assert literal_oracle.decode(want) == (lit.OK, text)
    for text, hexed in RFC7541_RAW:  # C.2 / C.3: raw
        want = bytes.fromhex(hexed)
        assert literal_oracle.encode(text, lit.NEVER) == want
        assert literal_oracle.decode(want) == (lit.OK, text)


def test_oracle_padding_rule(literal_oracle):
    good = bytes.fromhex("8c" + "f1e3c2e5f23a6ba0ab90f4ff")  # www.example.com, 7 bits of padding... (all ones)
    assert literal_oracle.decode(good) == (lit.OK, b"www.example.com")
    zero_pad = bytearray(good)
    zero_pad[-1] &= 0xF8  # padding bits no longer ones
    assert literal_oracle.decode(bytes(zero_pad))[0] in (lit.INVALID_PADDING, lit.OK)
    long_pad = bytes([0x80 | 13]) + good[1:] + b"\xff"  # a whole byte of padding
    assert literal_oracle.decode(long_pad) == (lit.INVALID_PADDING, b"")
    assert literal_oracle.decode(good[:-1]) == (lit.SHORT_BUFFER, b"")
    assert literal_oracle.decode(good + b"\x00") == (lit.INVALID_ARGUMENT, b"")


# ------------------------------------------------------------------------------------------------ GPU
def _strings(rng, n, max_len, binary_share=0.1):
    """Mostly header-like text (Huffman wins), some binary (raw wins), some empty, some long enough for 2- and
    3-byte length prefixes."""
    data, offs = refcodec.random_batch(rng, n, 0, max_len, "hpack")
    items = [bytearray(data[int(offs[i]):int(offs[i + 1])]) for i in range(n)]
    for i in rng.choice(n, size=max(1, int(n * binary_share)), replace=False):
        items[i] = bytearray(rng.integers(0, 256, size=len(items[i]), dtype=np.uint8).tobytes())
    for i, size in zip(rng.choice(n, size=6, replace=False), (126, 127, 128, 300, 20000, 0)):
        items[i] = bytearray(refcodec.random_batch(rng, 1, size, size, "hpack")[0].tobytes())
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(x) for x in items])
    return np.frombuffer(b"".join(bytes(x) for x in items), dtype=np.uint8), offs


@pytest.fixture(scope="module")
def hpack_ctx(pkg, coders):
    ctx = pkg.BatchContext(coders.coder("hpack"), eos_padding=0xFF, device=0)
    yield ctx
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [lit.SMALLEST, lit.NEVER, lit.ALWAYS])
def test_encode_literals_match_oracle(hpack_ctx, literal_oracle, mode):
    rng = np.random.default_rng(40 + mode)
    data, offs = _strings(rng, 3000, 200)
    want, want_offs = literal_oracle.encode_batch(data, offs, mode)
    got = hpack_ctx.hpack_encode_strings(data, offs, out_capacity=len(want) + 64, mode=mode)
    assert np.array_equal(got["out_offsets"], want_offs)
    assert np.array_equal(got["out"][:len(want)], want)
    if mode == lit.SMALLEST:
        h = np.array([want[int(o)] >> 7 for o in want_offs[:-1] if int(o) < len(want)])
        assert 0 < h.sum() < len(h), "the batch must mix raw and Huffman literals"


@pytest.mark.gpu
def test_rfc7541_literals_through_the_device(hpack_ctx):
    texts = [t for t, _ in RFC7541_C]
    data = np.frombuffer(b"".join(texts), dtype=np.uint8)
    offs = np.cumsum([0] + [len(t) for t in texts]).astype(np.uint64)
    got = hpack_ctx.hpack_encode_strings(data, offs, out_capacity=4096, mode=lit.ALWAYS)
    want = b"".join(bytes([0x80 | len(bytes.fromhex(h))]) + bytes.fromhex(h) for _, h in RFC7541_C)
    assert bytes(got["out"][:int(got["out_offsets"][-1])]) == want
    back = hpack_ctx.hpack_decode_strings(got["out"][:len(want)], got["out_offsets"], out_capacity=4096)
    assert bytes(back["out"][:int(back["out_offsets"][-1])]) == b"".join(texts)
    assert not back["status"].any()


@pytest.mark.gpu
@pytest.mark.parametrize("route", ["one_kernel", "passes"])
def test_decode_literals_match_oracle_including_malformed(hpack_ctx, literal_oracle, monkeypatch, route):
    """one_kernel: the batch decoder parses the literals itself (decode_batch_kernel<true>); passes: the
    parse / gather / decode / finish / move kernels that long single literals and large tables go through."""
    if route == "passes":
        monkeypatch.setenv("AWS_HUFFMAN_HPACK_PASSES", "1")
    rng = np.random.default_rng(77)
    data, offs = _strings(rng, 2500, 160)
    framed, f_offs = literal_oracle.encode_batch(data, offs, lit.SMALLEST)
    items = [bytearray(framed[int(f_offs[i]):int(f_offs[i + 1])]) for i in range(len(f_offs) - 1)]
    # damage every 7th literal in one of several ways
    for k, i in enumerate(range(3, len(items), 7)):
        it = items[i]
        kind = k % 8
        if kind == 0 and len(it) > 1:
            del it[-1]                                  # cut short
        elif kind == 1:
            it.append(0x00)                             # octets left over
        elif kind == 2 and len(it) > 1 and it[0] & 0x80:
            it[-1] &= 0xF0                              # padding bits cleared (may still be a valid code)
        elif kind == 3 and it[0] & 0x80 and (it[0] & 0x7f) < 126:
            it[0] += 1
            it.append(0xFF)                             # a whole byte of padding
        elif kind == 4:
            items[i] = bytearray()                      # nothing at all
        elif kind == 5 and it[0] & 0x80 and (it[0] & 0x7f) < 120:
            it[0] += 4
            it.extend(b"\xff\xff\xff\xff")              # EOS (30 ones) inside the payload
        elif kind == 6:
            items[i] = bytearray([0x7f, 0x80, 0x80])    # unfinished length
        elif kind == 7:
            items[i] = bytearray([0xff] + [0xff] * 10 + [0x01])  # a length beyond 64 bits
    f_offs = np.zeros(len(items) + 1, dtype=np.uint64)
    f_offs[1:] = np.cumsum([len(x) for x in items])
    framed = np.frombuffer(b"".join(bytes(x) for x in items), dtype=np.uint8)
    want, want_offs, want_status = literal_oracle.decode_batch(framed, f_offs)
    got = hpack_ctx.hpack_decode_strings(framed, f_offs, out_capacity=len(data) + 4096)
    assert np.array_equal(got["status"], want_status), np.flatnonzero(got["status"] != want_status)[:10]
    assert np.array_equal(got["out_offsets"], want_offs)
    assert np.array_equal(got["out"][:len(want)], want)
    seen = set(int(s) for s in want_status)
    assert {lit.OK, lit.SHORT_BUFFER, lit.INVALID_ARGUMENT, lit.INVALID_PADDING} <= seen


@pytest.mark.gpu
def test_single_long_literal_and_capacity(hpack_ctx, literal_oracle, pkg):
    rng = np.random.default_rng(5)
    data, offs = refcodec.random_batch(rng, 1, 400_000, 400_000, "hpack")
    want, want_offs = literal_oracle.encode_batch(data, offs, lit.SMALLEST)
    got = hpack_ctx.hpack_encode_strings(data, offs, out_capacity=len(want), mode=lit.SMALLEST)
    assert np.array_equal(got["out"][:len(want)], want) and np.array_equal(got["out_offsets"], want_offs)
    back = hpack_ctx.hpack_decode_strings(want, want_offs, out_capacity=len(data))
    assert np.array_equal(back["out"][:len(data)], data) and back["status"][0] == 0
    with pytest.raises(pkg.CodecError) as err:
        hpack_ctx.hpack_encode_strings(data, offs, out_capacity=len(want) - 1)
    assert err.value.code == pkg.AWS_ERROR_SHORT_BUFFER
    assert int(err.value.result["out_offsets"][-1]) == len(want)
    empty = hpack_ctx.hpack_encode_strings(np.zeros(0, np.uint8), np.zeros(1, np.uint64), out_capacity=8)
    assert int(empty["out_offsets"][0]) == 0


@pytest.mark.gpu
def test_device_pointer_round_trip_at_bench_size(hpack_ctx):
    """200k strings through the device entry points: literals decode back to the strings; prints the rates."""
    import torch
    rng = np.random.default_rng(11)
    data, offs = refcodec.random_batch(rng, 200_000, 8, 256, "hpack")
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(data).to(dev)
    d_off = torch.from_numpy(offs.astype(np.int64)).to(dev)
    n, total = len(offs) - 1, len(data)
    framed = torch.empty(total * 2 + 64, dtype=torch.uint8, device=dev)
    f_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    out = torch.empty(total + 64, dtype=torch.uint8, device=dev)
    o_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    status = torch.ones(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for rep in range(3):
            ev[0].record(stream)
            hpack_ctx.hpack_device(True, n, d_in, d_off, total, framed, framed.numel(), f_off, mode=lit.SMALLEST,
                                   stream=stream.cuda_stream)
            ev[1].record(stream)
            framed_size = int(f_off[-1].item())
            hpack_ctx.hpack_device(False, n, framed, f_off, framed_size, out, out.numel(), o_off, status=status,
                                   stream=stream.cuda_stream)
            ev[2].record(stream)
            torch.cuda.synchronize(dev)
    assert int(o_off[-1].item()) == total and torch.equal(out[:total], d_in) and torch.equal(o_off, d_off)
    assert int(status.abs().sum().item()) == 0
    enc_ms, dec_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print("\nhpack literals, %d strings, %.1f MB raw -> %.1f MB framed: encode %.3f ms (%.0f GB/s in+out), decode %.3f ms (%.0f GB/s)"
          % (n, total / 1e6, framed_size / 1e6, enc_ms, (total + framed_size) / enc_ms / 1e6, dec_ms,
             (total + framed_size) / dec_ms / 1e6))
