"""Host streaming codec (aws-c-compression_b200/host/huffman.c) through the reference's own C API.

Mirrors the reference's 15 test cases (tests/huffman_test.c) and its three fuzz properties
(tests/fuzz/*.c), then drives OUR library and the UNMODIFIED reference build through the very same
ctypes calls on random inputs and compares every observable: return code, aws_last_error, cursor,
buffer length, bytes, overflow_bits / working_bits state."""
import ctypes as C

import numpy as np
import pytest

import refcodec
from refcodec import OK, SHORT_BUFFER, UNKNOWN_SYMBOL


@pytest.fixture(scope="module")
def capi(pkg):
    return pkg.capi


@pytest.fixture(scope="module")
def G():
    g = refcodec.golden("reference_vectors.json")
    kats = [(bytes.fromhex(k["input_hex"]), bytes.fromhex(k["encoded_hex"])) for k in g["encode_kats"]]
    return {"url": kats[0], "all": kats[1], "exact": kats[2:], "steps": g["step_sizes"], "table": g["test_table"]}


class Driver:
    """Drives one library (ours or the reference) through aws_huffman_* with real structs."""

    def __init__(self, capi, lib, coder):
        self.capi, self.lib, self.coder = capi, lib, coder

    def encoder(self, eos=0xFF):
        e = self.capi.aws_huffman_encoder()
        self.lib.aws_huffman_encoder_init(C.byref(e), self.coder)
        e.eos_padding = eos
        return e

    def decoder(self):
        d = self.capi.aws_huffman_decoder()
        self.lib.aws_huffman_decoder_init(C.byref(d), self.coder)
        return d

    def call(self, fn, state, src, src_pos, out, out_len, capacity):
        """One call on src[src_pos:]; returns (rc, err, new_src_pos, new_out_len)."""
        cur = self.capi.aws_byte_cursor(len(src) - src_pos, src.ctypes.data + src_pos if len(src) else None)
        buf = self.capi.aws_byte_buf(out_len, out.ctypes.data, capacity, None)
        self.lib.aws_reset_error()
        rc = fn(C.byref(state), C.byref(cur), C.byref(buf))
        err = self.lib.aws_last_error() if rc else 0
        return rc, err, len(src) - cur.len, buf.len

    def encode(self, *a):
        return self.call(self.lib.aws_huffman_encode, *a)

    def decode(self, *a):
        return self.call(self.lib.aws_huffman_decode, *a)

    def encoded_length(self, data):
        data = np.frombuffer(bytes(data), dtype=np.uint8)
        e = self.encoder()
        cur = self.capi.aws_byte_cursor(len(data), data.ctypes.data if len(data) else None)
        return self.lib.aws_huffman_get_encoded_length(C.byref(e), cur)


@pytest.fixture(scope="module")
def ours(capi, product, coders):
    return Driver(capi, product.lib, coders.coder("test"))


def arr(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


# ---- the reference's test cases, on our library -------------------------------------------------

def test_symbol_coder_rows(coders, G):
    # huffman_symbol_encoder / huffman_symbol_decoder (huffman_test.c:42-60,199-220) on OUR generated coder
    coder = coders.coder("test").contents
    for sym in range(256):
        code = coder.encode(sym, None)
        assert (code.pattern, code.num_bits) == (G["table"]["patterns"][sym], G["table"]["num_bits"][sym])
        out = C.c_uint8(0)
        used = coder.decode((code.pattern << (32 - code.num_bits)) & 0xFFFFFFFF, C.byref(out), None)
        assert (used, out.value) == (code.num_bits, sym)


@pytest.mark.parametrize("which", ["url", "all"])
def test_encoder_and_decoder_golden(ours, G, which):
    # huffman_encoder, huffman_encoder_all_code_points, huffman_decoder, huffman_decoder_all_code_points
    text, want = G[which]
    assert ours.encoded_length(text) == len(want)
    out = np.zeros(len(want) + 1, dtype=np.uint8)
    rc, err, used, n = ours.encode(ours.encoder(), arr(text), 0, out, 0, len(want))
    assert (rc, used, n, out[len(want)]) == (0, len(text), len(want), 0)
    assert bytes(out[:n]) == want
    back = np.zeros(len(text) + 1, dtype=np.uint8)
    rc, err, used, n = ours.decode(ours.decoder(), arr(want), 0, back, 0, len(text))
    assert (rc, used, n, back[len(text)]) == (0, len(want), len(text), 0)
    assert bytes(back[:n]) == text


def test_encoder_partial_output(ours, G):
    # huffman_encoder_partial_output (huffman_test.c:117-165)
    text, want = G["all"]
    src = arr(text)
    enc = ours.encoder()
    for step in G["steps"]:
        ours.lib.aws_huffman_encoder_reset(C.byref(enc))
        out = np.zeros(len(want), dtype=np.uint8)
        pos = n = cap = 0
        while n < len(want):
            cap = min(cap + step, len(want))
            before = n
            rc, err, pos, n = ours.encode(enc, src, pos, out, n, cap)
            assert n > before and bytes(out[:n]) == want[:n]
            if n == len(want):
                assert rc == 0
            else:
                assert (rc, err) == (-1, SHORT_BUFFER)


def test_encoder_exact_output(ours, G):
    # huffman_encoder_exact_output (huffman_test.c:167-197): one encoder reused without reset
    enc = ours.encoder()
    for text, want in G["exact"]:
        out = np.zeros(2, dtype=np.uint8)
        rc, err, used, n = ours.encode(enc, arr(text), 0, out, 0, len(want))
        assert (rc, n, bytes(out[:n])) == (0, len(want), want)


def test_decoder_partial_input(ours, G):
    # huffman_decoder_partial_input (huffman_test.c:275-314)
    text, enc_bytes = G["all"]
    src = arr(enc_bytes)
    dec = ours.decoder()
    for step in G["steps"]:
        ours.lib.aws_huffman_decoder_reset(C.byref(dec))
        out = np.zeros(150, dtype=np.uint8)
        n = pos = 0
        while n < len(text):
            chunk = src[pos:pos + step].copy()
            rc, err, used, n = ours.decode(dec, chunk, 0, out, n, len(text))
            assert used == len(chunk) and bytes(out[:n]) == text[:n]
            pos += len(chunk)
            if n == len(text):
                assert rc == 0
        assert n == len(text)


def test_decoder_partial_output(ours, G):
    # huffman_decoder_partial_output (huffman_test.c:316-363)
    text, enc_bytes = G["all"]
    src = arr(enc_bytes)
    dec = ours.decoder()
    for step in G["steps"]:
        ours.lib.aws_huffman_decoder_reset(C.byref(dec))
        out = np.zeros(150, dtype=np.uint8)
        pos = n = cap = 0
        while n < len(text):
            cap = min(cap + step, len(text))
            before = n
            rc, err, pos, n = ours.decode(dec, src, pos, out, n, cap)
            assert n > before and bytes(out[:n]) == text[:n]
            if n == len(text):
                assert rc == 0
            else:
                assert (rc, err) == (-1, SHORT_BUFFER)


def test_decoder_allow_growth(capi, product, ours, G):
    # huffman_decoder_allow_growth (huffman_test.c:365-385)
    text, enc_bytes = G["url"]
    lib = product.lib
    dec = ours.decoder()
    lib.aws_huffman_decoder_allow_growth(C.byref(dec), True)
    buf = capi.aws_byte_buf()
    assert lib.aws_byte_buf_init(C.byref(buf), lib.aws_default_allocator(), 1) == 0
    src = arr(enc_bytes)
    cur = capi.aws_byte_cursor(len(src), src.ctypes.data)
    assert lib.aws_huffman_decode(C.byref(dec), C.byref(cur), C.byref(buf)) == 0
    assert cur.len == 0 and C.string_at(buf.buffer, buf.len) == text
    lib.aws_byte_buf_clean_up(C.byref(buf))


def test_transitive_helpers(product, coders, G):
    # huffman_transitive* (huffman_test.c:387-446) through the exported helpers
    lib, coder = product.lib, coders.coder("test")
    msg = C.c_char_p()
    for text, size in [(G["url"][0], len(G["url"][1])), (b"cdfh", 3), (G["all"][0], len(G["all"][1]))]:
        assert lib.huffman_test_transitive(coder, text, len(text), size, C.byref(msg)) == 0, msg.value
    for step in G["steps"]:
        text, want = G["all"]
        assert lib.huffman_test_transitive_chunked(coder, text, len(text), len(want), step, C.byref(msg)) == 0, msg.value
    assert lib.huffman_test_transitive(coder, b"cdfh", 4, 99, C.byref(msg)) == -1
    assert msg.value == b"encoded length is incorrect"


def test_library_init_registers_error_strings(product):
    # library_init (tests/library_test.c:6-22), plus the two codes the B200 build adds
    lib = product.lib
    lib.aws_compression_library_init(lib.aws_default_allocator())
    assert lib.aws_error_name(3072) == b"AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL"
    assert lib.aws_error_name(3073) == b"AWS_ERROR_COMPRESSION_DEVICE_FAILURE"
    assert lib.aws_error_name(3074) == b"AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE"
    lib.aws_compression_library_init(lib.aws_default_allocator())  # idempotent
    lib.aws_compression_library_clean_up()
    assert lib.aws_error_name(3072) == b"Unknown Error Code"
    lib.aws_compression_library_clean_up()


# ---- fuzz properties (tests/fuzz/*.c) as seeded random tests ------------------------------------

def test_fuzz_transitive_and_chunked(product, coders):
    lib = product.lib
    rng = np.random.default_rng(0xF0F0)
    msg = C.c_char_p()
    for table in ("test", "hpack"):
        coder = coders.coder(table)
        sampler = refcodec.zipf_symbol_sampler(refcodec.table_arrays(table)[1], s=1.0)
        for _ in range(150):
            size = int(rng.integers(1, 300))
            if table == "hpack":
                # hpack codes reach 30 bits; the helpers' scratch is 2x the input (same in the reference),
                # so feed text-like symbols
                data = bytes(sampler[rng.integers(0, 65536, size=size)])
            else:
                data = bytes(rng.integers(0, 256, size=size, dtype=np.uint8))
            assert lib.huffman_test_transitive(coder, data, size, 0, C.byref(msg)) == 0, msg.value
            for step in (1, 2, 4, 8, 16, 32, 64, 128):
                assert lib.huffman_test_transitive_chunked(coder, data, size, 0, step, C.byref(msg)) == 0, msg.value


# ---- differential: our library vs the unmodified reference, call for call ------------------------

@pytest.mark.parametrize("table", ["test", "hpack"])
def test_streaming_calls_match_reference(capi, product, coders, ref, table):
    rng = np.random.default_rng(0xD1FF + len(table))
    reflib = capi.bind_streaming_api(ref.lib)
    holes = [3, 50, 97, 200]
    ours_plain = Driver(capi, product.lib, coders.coder(table))
    ref_plain = Driver(capi, reflib, C.cast(ref.coder(table), C.POINTER(capi.aws_huffman_symbol_coder)))
    ref_masked = Driver(capi, reflib, C.cast(ref.masked_coder(table, holes), C.POINTER(capi.aws_huffman_symbol_coder)))
    # our side of the masked coder: python callbacks around our generated coder
    inner = coders.coder(table).contents

    def enc_py(sym):
        code = inner.encode(sym, None)
        return (0, 0) if sym in holes else (code.pattern, code.num_bits)

    def dec_py(bits):
        tmp = C.c_uint8(0)
        used = inner.decode(bits, C.byref(tmp), None)
        return None if used == 0 or tmp.value in holes else (tmp.value, used)

    masked = capi.python_coder(enc_py, dec_py)
    ours_masked = Driver(capi, product.lib, C.pointer(masked))
    sampler = refcodec.zipf_symbol_sampler(refcodec.table_arrays(table)[1])

    def state_of(x):
        return bytes(x)[8:]  # skip the coder pointer

    for trial in range(400):
        a, b = (ours_masked, ref_masked) if trial % 4 == 0 else (ours_plain, ref_plain)
        size = int(rng.integers(0, 120))
        data = sampler[rng.integers(0, 65536, size=size)] if trial % 3 else rng.integers(0, 256, size=size, dtype=np.uint8)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        eos = int(rng.integers(0, 256))
        # encode with a randomly opening output window until it finishes or errors for good
        ea, eb = a.encoder(eos), b.encoder(eos)
        room = 4 * size + 8
        oa, ob = np.zeros(room + 1, dtype=np.uint8), np.zeros(room + 1, dtype=np.uint8)
        pa = pb = na = nb = cap = 0
        for _ in range(200):
            cap = min(room, cap + int(rng.integers(0, 9)) if trial % 2 else room)
            ra = a.encode(ea, data, pa, oa, na, cap)
            rb = b.encode(eb, data, pb, ob, nb, cap)
            assert ra == rb, (trial, ra, rb)
            _, err, pa, na = ra
            _, _, pb, nb = rb
            assert np.array_equal(oa, ob) and state_of(ea) == state_of(eb)
            if ra[0] == 0 or err != SHORT_BUFFER:
                break
        if ra[0] != 0:
            continue
        # decode it back with random input chunking and a randomly opening output window
        enc_bytes = oa[:na].copy()
        if trial % 5 == 0 and na:
            enc_bytes[int(rng.integers(0, na))] ^= 1 << int(rng.integers(0, 8))
        da, db = a.decoder(), b.decoder()
        ba, bb = np.zeros(8 * na + 9, dtype=np.uint8), np.zeros(8 * na + 9, dtype=np.uint8)
        pos = na2 = nb2 = cap = 0
        while True:
            take = int(rng.integers(0, 7)) if trial % 2 else len(enc_bytes) - pos
            chunk = enc_bytes[pos:pos + take].copy()
            cap = min(len(ba) - 1, cap + int(rng.integers(0, 12)))
            ra = a.decode(da, chunk, 0, ba, na2, cap)
            rb = b.decode(db, chunk, 0, bb, nb2, cap)
            assert ra == rb, (trial, ra, rb)
            assert np.array_equal(ba, bb) and state_of(da) == state_of(db)
            _, err, used, na2 = ra
            nb2 = na2
            pos += used
            if ra[0] != 0 and err != SHORT_BUFFER:
                break
            if ra[0] == 0 and pos >= len(enc_bytes):
                break
