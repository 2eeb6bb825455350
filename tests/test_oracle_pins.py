"""Pins the oracle (oracle/huffman_oracle.c) before anything is checked against it:
golden vectors of the reference's own tests, RFC 7541 Appendix C for HPACK, outputs of the
unmodified reference committed under tests/golden/, and (when oracle/_ref is present) a live
differential run against the unmodified reference."""
import numpy as np
import pytest

import refcodec
from refcodec import OK, SHORT_BUFFER, UNKNOWN_SYMBOL

RFC7541_C = [  # RFC 7541 Appendix C.4 / C.6 Huffman-coded string literals
    (b"www.example.com", "f1e3c2e5f23a6ba0ab90f4ff"),
    (b"no-cache", "a8eb10649cbf"),
    (b"custom-key", "25a849e95ba97d7f"),
    (b"custom-value", "25a849e95bb8e8b4bf"),
    (b"302", "6402"),
    (b"private", "aec3771a4b"),
    (b"Mon, 21 Oct 2013 20:13:21 GMT", "d07abe941054d444a8200595040b8166e082a62d1bff"),
    (b"Mon, 21 Oct 2013 20:13:22 GMT", "d07abe941054d444a8200595040b8166e084a62d1bff"),
    (b"https://www.example.com", "9d29ad171863c78f0b97c8e9ae82ae43d3"),
    (b"307", "640eff"),
    (b"gzip", "9bd9ab"),
    (b"foo=ASDJKHQKBZXOQWEOPIUAXQWEOIU; max-age=3600; version=1",
     "94e7821dd7f2e6c7b335dfdfcd5b3960d5af27087f3672c1ab270fb5291f9587316065c003ed4ee5b1063d5007"),
]


def encode_one(oracle, table, data, cap=None, eos=0xFF):
    data = np.frombuffer(bytes(data), dtype=np.uint8)
    cap = 4 * len(data) + 8 if cap is None else cap
    r = oracle.encode_batch(table, eos, data, [0, len(data)], max(cap, 1), out_offsets=[0], out_caps=[cap])
    return r, bytes(r["out"][:int(r["out_lens"][0])])


def decode_one(oracle, table, enc, cap=None):
    enc = np.frombuffer(bytes(enc), dtype=np.uint8)
    cap = 8 * len(enc) + 8 if cap is None else cap
    r = oracle.decode_batch(table, enc, [0, len(enc)], max(cap, 1), out_offsets=[0], out_caps=[cap])
    return r, bytes(r["out"][:int(r["out_lens"][0])])


def test_symbol_coder_matches_every_def_row(oracle, oracle_tables):
    # reference tests huffman_symbol_encoder / huffman_symbol_decoder (huffman_test.c:42-60,199-220)
    t = refcodec.golden("reference_vectors.json")["test_table"]
    table = oracle_tables["test"]
    for sym in range(256):
        pattern, n = t["patterns"][sym], t["num_bits"][sym]
        assert oracle.encode_symbol(table, sym) == (pattern, n)
        assert oracle.decode_symbol(table, pattern << (32 - n)) == (n, sym)


def test_reference_golden_encodings(oracle, oracle_tables):
    g = refcodec.golden("reference_vectors.json")
    table = oracle_tables["test"]
    for kat in g["encode_kats"]:
        data, want = bytes.fromhex(kat["input_hex"]), bytes.fromhex(kat["encoded_hex"])
        assert oracle.encoded_length(table, np.frombuffer(data, dtype=np.uint8)) == len(want), kat["cite"]
        # exact-size buffer, like huffman_test.c:72,175-194
        r, got = encode_one(oracle, table, data, cap=len(want))
        assert (int(r["status"][0]), got) == (OK, want), kat["cite"]
        r, back = decode_one(oracle, table, want, cap=len(data))
        assert (int(r["status"][0]), back, int(r["consumed"][0])) == (OK, data, len(want)), kat["cite"]
    for kat in g["encoded_length_kats"]:
        data = bytes.fromhex(kat["input_hex"])
        r, got = encode_one(oracle, table, data)
        assert len(got) == kat["encoded_len"] and got == bytes.fromhex("8218a3")


def test_rfc7541_appendix_c(oracle, oracle_tables):
    table = oracle_tables["hpack"]
    for text, hexed in RFC7541_C:
        r, got = encode_one(oracle, table, text)
        assert got.hex() == hexed
        r, back = decode_one(oracle, table, bytes.fromhex(hexed))
        assert back == text and int(r["status"][0]) == OK
        # padding check of the reference README (README.md:176-183): leftover bits are all ones
        nb = int(r["leftover_num_bits"][0])
        assert nb < 8 and (int(r["leftover_working_bits"][0]) >> (64 - nb) if nb else 0) == (1 << nb) - 1
    allbytes = bytes(range(256))
    assert len(encode_one(oracle, table, allbytes)[1]) == 583


def test_eos_padding_uses_low_bits(oracle, oracle_tables):
    # SURVEY.md App. B.3 (reference huffman.c:178-182), test table: 'a' = 00101
    table = oracle_tables["test"]
    for eos, one, two in [(0xFF, "2f", "297f"), (0x55, "2d", "2955"), (0x00, "28", "2940"), (0xAA, "2a", "296a"),
                          (0x0F, "2f", "294f")]:
        assert encode_one(oracle, table, b"a", eos=eos)[1].hex() == one
        assert encode_one(oracle, table, b"aa", eos=eos)[1].hex() == two


def test_short_buffer_closed_forms(oracle, oracle_tables):
    # "www.example.com" with capacities 0..4
    table = oracle_tables["test"]
    want = [(0, 0, 0), (2, 0x7, 4), (3, 0x3, 2), (4, 0x1, 1), (6, 0x1, 6)]
    for cap, (consumed, pattern, bits) in enumerate(want):
        r, got = encode_one(oracle, table, b"www.example.com", cap=cap)
        assert int(r["status"][0]) == SHORT_BUFFER and len(got) == cap
        assert (int(r["consumed"][0]), int(r["overflow_pattern"][0]), int(r["overflow_num_bits"][0])) == (
            consumed, pattern, bits)


def test_decode_termination_rules(oracle, oracle_tables):
    # 0x29 then zeros; all-ones
    table = oracle_tables["test"]
    for nbytes, (status, nsym, left) in {1: (OK, 1, 3), 2: (OK, 2, 6), 3: (OK, 2, 14), 5: (OK, 2, 30),
                                         6: (UNKNOWN_SYMBOL, 2, None)}.items():
        r, got = decode_one(oracle, table, bytes([0x29] + [0] * (nbytes - 1)))
        assert (int(r["status"][0]), len(got)) == (status, nsym)
        if left is not None:
            assert int(r["leftover_num_bits"][0]) == left
    for nbytes, status in [(1, OK), (3, OK), (4, UNKNOWN_SYMBOL)]:
        r, got = decode_one(oracle, table, b"\xff" * nbytes)
        assert (int(r["status"][0]), len(got)) == (status, 0)
    # B.7 cursor over-read: www.example.com golden, output capacities 0/3/6/9/12
    enc = bytes.fromhex(refcodec.golden("reference_vectors.json")["encode_kats"][0]["encoded_hex"])
    for cap, (left_bytes, nb) in zip([0, 3, 6, 9, 12], [(8, 32), (5, 38), (3, 34), (1, 32), (0, 22)]):
        r, got = decode_one(oracle, table, enc, cap=cap)
        assert int(r["status"][0]) == SHORT_BUFFER and len(got) == cap
        assert (len(enc) - int(r["consumed"][0]), int(r["leftover_num_bits"][0])) == (left_bytes, nb)


@pytest.mark.parametrize("table_name", ["test", "hpack"])
def test_committed_reference_outputs(oracle, oracle_tables, table_name):
    """tests/golden/differential_*.json were produced by the UNMODIFIED reference (tools/make_golden.py)."""
    cases = refcodec.golden("differential_%s.json" % table_name)
    patterns, num_bits = refcodec.table_arrays(table_name)
    masked_bits = num_bits.copy()
    masked_bits[cases["unknown_symbols"]] = 0
    masked = oracle.table(patterns, masked_bits)
    table = oracle_tables[table_name]
    for c in cases["encode"]:
        r, got = encode_one(oracle, masked if c["masked"] else table, bytes.fromhex(c["in"]), cap=c["cap"], eos=c["eos"])
        assert got.hex() == c["out"]
        assert (int(r["status"][0]), int(r["consumed"][0]), int(r["overflow_pattern"][0]),
                int(r["overflow_num_bits"][0])) == (c["status"], c["consumed"], c["ovf_pattern"], c["ovf_bits"])
    for c in cases["decode"]:
        r, got = decode_one(oracle, table, bytes.fromhex(c["in"]), cap=c["cap"])
        assert got.hex() == c["out"]
        assert (int(r["status"][0]), int(r["consumed"][0]), int(r["leftover_working_bits"][0]),
                int(r["leftover_num_bits"][0])) == (c["status"], c["consumed"], c["left_bits"], c["left_num"])


@pytest.mark.parametrize("table_name", ["test", "hpack"])
def test_live_differential_against_reference(oracle, oracle_tables, ref, table_name):
    """Random batches, packed and capacity-limited, masked and unmasked: oracle == unmodified reference."""
    rng = np.random.default_rng(0xC0FFEE + len(table_name))
    patterns, num_bits = refcodec.table_arrays(table_name)
    holes = [int(x) for x in rng.choice(256, size=5, replace=False)]
    masked_bits = num_bits.copy()
    masked_bits[holes] = 0
    pairs = [(oracle_tables[table_name], ref.coder(table_name)),
             (oracle.table(patterns, masked_bits), ref.masked_coder(table_name, holes))]
    for o_table, r_coder in pairs:
        for zipf in (True, False):
            data, offs = refcodec.random_batch(rng, 3000, 0, 90, table_name, zipf=zipf)
            n = len(offs) - 1
            cap_total = 4 * len(data) + 16
            a = oracle.encode_batch(o_table, 0xFF, data, offs, cap_total)
            b = ref.encode_batch(r_coder, 0xFF, data, offs, cap_total)
            for k in a:
                assert np.array_equal(a[k], b[k]), k
            # slotted with random capacities around the needed size
            need = a["out_lens"].astype(np.int64)
            caps = np.maximum(0, need + rng.integers(-6, 3, size=n)).astype(np.uint64)
            slots = np.zeros(n, dtype=np.uint64)
            slots[1:] = np.cumsum(caps)[:-1]
            total = int(caps.sum()) + 1
            eos = int(rng.integers(0, 256))
            a = oracle.encode_batch(o_table, eos, data, offs, total, out_offsets=slots, out_caps=caps)
            b = ref.encode_batch(r_coder, eos, data, offs, total, out_offsets=slots, out_caps=caps)
            for k in a:
                assert np.array_equal(a[k], b[k]), k
            assert (a["status"] == SHORT_BUFFER).any()
    # decode: valid streams, corrupted streams, random bytes; packed and capacity-limited
    o_table, r_coder = pairs[0]
    data, offs = refcodec.random_batch(rng, 3000, 0, 90, table_name)
    enc = oracle.encode_batch(o_table, 0xFF, data, offs, 4 * len(data) + 16)
    stream = enc["out"][:int(enc["out_offsets"][-1])].copy()
    noisy = stream.copy()
    flips = rng.integers(0, len(noisy), size=len(noisy) // 50)
    noisy[flips] ^= (1 << rng.integers(0, 8, size=len(flips))).astype(np.uint8)
    rand = rng.integers(0, 256, size=len(stream), dtype=np.uint8)
    for payload in (stream, noisy, rand):
        a = oracle.decode_batch(o_table, payload, enc["out_offsets"], 8 * len(payload) + 16)
        b = ref.decode_batch(r_coder, payload, enc["out_offsets"], 8 * len(payload) + 16)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        n = len(offs) - 1
        caps = np.maximum(0, a["out_lens"].astype(np.int64) + rng.integers(-5, 2, size=n)).astype(np.uint64)
        slots = np.zeros(n, dtype=np.uint64)
        slots[1:] = np.cumsum(caps)[:-1]
        a = oracle.decode_batch(o_table, payload, enc["out_offsets"], int(caps.sum()) + 1, out_offsets=slots, out_caps=caps)
        b = ref.decode_batch(r_coder, payload, enc["out_offsets"], int(caps.sum()) + 1, out_offsets=slots, out_caps=caps)
        for k in a:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("table_name", ["test", "hpack"])
def test_streaming_state_against_reference(oracle, oracle_tables, ref, pkg, table_name):
    """The oracle's state ACROSS calls (encoder overflow bits, decoder register) against the unmodified
    reference driven the same way: output in small pieces, input in small chunks (the reference's own
    chunked tests, tests/huffman_test.c:117-165,275-363). This is what the *_resume GPU tests lean on."""
    import ctypes as C
    capi = pkg.capi
    L = capi.bind_streaming_api(C.CDLL(refcodec.REF_SO))
    coder = C.cast(ref.coder(table_name), C.POINTER(capi.aws_huffman_symbol_coder))
    table = oracle_tables[table_name]
    rng = np.random.default_rng(0xC0DE)

    def ref_call(fn, state, data, cap):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        cur = capi.aws_byte_cursor(len(data), data.ctypes.data if len(data) else None)
        buf = capi.aws_byte_buf(0, out.ctypes.data, cap, None)
        L.aws_reset_error()
        rc = fn(C.byref(state), C.byref(cur), C.byref(buf))
        code = 0 if rc == 0 else L.aws_last_error()
        return code, len(data) - cur.len, bytes(out[:buf.len])

    for case in range(60):
        text, _ = refcodec.random_batch(rng, 1, 0, 120, table_name)
        cap = int(rng.integers(1, 9))
        enc_r = capi.aws_huffman_encoder()
        L.aws_huffman_encoder_init(C.byref(enc_r), coder)
        enc_o = oracle.new_encoder(table, 0xFF)
        pos, encoded = 0, b""
        for _ in range(2000):
            code, used, out_r = ref_call(L.aws_huffman_encode, enc_r, text[pos:], cap)
            out_o = np.zeros(cap, dtype=np.uint8)
            rc_o, used_o, len_o = oracle.encode_call(enc_o, text[pos:], out_o, 0, cap)
            assert (rc_o, used_o, bytes(out_o[:len_o])) == (code, used, out_r), "encode case %d" % case
            assert enc_o.overflow_bits.num_bits == enc_r.overflow_bits.num_bits
            if enc_r.overflow_bits.num_bits:
                assert enc_o.overflow_bits.pattern == enc_r.overflow_bits.pattern
            pos += used
            encoded += out_r
            if code == OK:
                break
            assert code == SHORT_BUFFER
        # decode: input in chunks of `step` bytes, output in pieces of `cap` bytes
        step = int(rng.integers(1, 7))
        stream = np.frombuffer(encoded, dtype=np.uint8)
        dec_r = capi.aws_huffman_decoder()
        L.aws_huffman_decoder_init(C.byref(dec_r), coder)
        dec_o = oracle.new_decoder(table)
        pos = fed = 0
        for _ in range(5000):
            if pos == fed:
                fed = min(len(stream), fed + step)
            code, used, out_r = ref_call(L.aws_huffman_decode, dec_r, stream[pos:fed], cap)
            out_o = np.zeros(cap, dtype=np.uint8)
            rc_o, used_o, len_o = oracle.decode_call(dec_o, stream[pos:fed], out_o, 0, cap)
            assert (rc_o, used_o, bytes(out_o[:len_o])) == (code, used, out_r), "decode case %d" % case
            assert (dec_o.working_bits, dec_o.num_bits) == (dec_r.working_bits, dec_r.num_bits)
            pos += used
            if code == OK and fed == len(stream) and pos == fed:
                break
