"""GPU parity for the streaming continuation entry points:
aws_huffman_encode_batch_resume / aws_huffman_decode_batch_resume against the oracle's streaming
encoder / decoder driven call by call on the same chunks, the way the reference's own chunked tests do
(tests/huffman_test.c:117-165 output in pieces, :275-363 and source/huffman_testing.c transitive chunked).
Every call is compared: bytes, lengths, cursor advance, status and the carried state."""
import numpy as np
import pytest

import refcodec
from refcodec import OK, SHORT_BUFFER, UNKNOWN_SYMBOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def contexts(pkg, coders):
    made = {}

    def get(table, eos=0xFF):
        key = (table, eos)
        if key not in made:
            made[key] = pkg.BatchContext(coders.coder(table), eos_padding=eos, device=0)
        return made[key]

    yield get
    for ctx in made.values():
        ctx.close()


def _csr(chunks):
    lens = np.array([len(c) for c in chunks], dtype=np.uint64)
    offs = np.zeros(len(chunks) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    data = np.concatenate([np.asarray(c, dtype=np.uint8) for c in chunks]) if len(chunks) and offs[-1] else np.zeros(0, np.uint8)
    return data, offs


@pytest.mark.parametrize("table_name", ["test", "hpack"])
@pytest.mark.parametrize("cap", [1, 2, 3, 5, 16])
def test_encode_with_output_in_pieces(contexts, oracle, oracle_tables, table_name, cap):
    """Each stream is encoded into output buffers of `cap` bytes, call after call, until its input is used up."""
    rng = np.random.default_rng(1000 + cap)
    table = oracle_tables[table_name]
    ctx = contexts(table_name)
    n = 257
    texts, offs = refcodec.random_batch(rng, n=n, min_len=0, max_len=90, table=table_name)
    streams = [texts[int(offs[i]):int(offs[i + 1])] for i in range(n)]
    encoders = [oracle.new_encoder(table, 0xFF) for _ in range(n)]
    pos = np.zeros(n, dtype=np.int64)
    done = np.zeros(n, dtype=bool)
    state = (np.zeros(n, np.uint32), np.zeros(n, np.uint8))
    rounds = 0
    while not done.all():
        rounds += 1
        assert rounds < 2000
        live = np.flatnonzero(~done)
        chunks = [streams[i][pos[i]:] for i in live]
        data, in_off = _csr(chunks)
        slots = np.arange(len(live), dtype=np.uint64) * cap
        caps = np.full(len(live), cap, dtype=np.uint64)
        got = ctx.encode(data, in_off, out_capacity=cap * len(live) + 8, out_offsets=slots, out_caps=caps,
                         state=(state[0][live], state[1][live]))
        for j, i in enumerate(live):
            out = np.zeros(cap, dtype=np.uint8)
            rc, used, olen = oracle.encode_call(encoders[i], chunks[j], out, 0, cap)
            assert int(got["status"][j]) == rc, "stream %d round %d status" % (i, rounds)
            assert int(got["consumed"][j]) == used, "stream %d round %d consumed" % (i, rounds)
            assert int(got["out_lens"][j]) == olen
            assert np.array_equal(got["out"][j * cap:j * cap + olen], out[:olen]), "stream %d round %d bytes" % (i, rounds)
            assert int(got["overflow_num_bits"][j]) == encoders[i].overflow_bits.num_bits
            if encoders[i].overflow_bits.num_bits:
                assert int(got["overflow_pattern"][j]) == encoders[i].overflow_bits.pattern
            pos[i] += used
            state[0][i] = got["overflow_pattern"][j]
            state[1][i] = got["overflow_num_bits"][j]
            assert rc in (OK, SHORT_BUFFER)
            if rc == OK:
                done[i] = True
    assert rounds > 1


@pytest.mark.parametrize("table_name", ["test", "hpack"])
@pytest.mark.parametrize("in_chunk,out_cap", [(1, 64), (2, 3), (5, 1), (7, 400), (64, 5)])
def test_decode_fed_in_chunks(contexts, oracle, oracle_tables, table_name, in_chunk, out_cap):
    """Each encoded stream reaches its decoder `in_chunk` bytes at a time and is decoded into buffers of
    `out_cap` bytes; what a call leaves unread is offered again (with the decoder's carried register)."""
    rng = np.random.default_rng(77 + in_chunk * 31 + out_cap)
    table = oracle_tables[table_name]
    ctx = contexts(table_name)
    n = 193
    texts, offs = refcodec.random_batch(rng, n=n, min_len=0, max_len=120, table=table_name)
    enc = oracle.encode_batch(table, 0xFF, texts, offs, out_capacity=4 * len(texts) + 16)
    streams = [enc["out"][int(enc["out_offsets"][i]):int(enc["out_offsets"][i + 1])] for i in range(n)]
    decoders = [oracle.new_decoder(table) for _ in range(n)]
    pos = np.zeros(n, dtype=np.int64)      # bytes of the stream the decoder has pulled in
    fed = np.zeros(n, dtype=np.int64)      # bytes offered so far
    done = np.zeros(n, dtype=bool)
    outputs = [[] for _ in range(n)]
    state = (np.zeros(n, np.uint64), np.zeros(n, np.uint8))
    rounds = 0
    while not done.all():
        rounds += 1
        assert rounds < 5000
        live = np.flatnonzero(~done)
        for i in live:
            if pos[i] == fed[i]:
                fed[i] = min(len(streams[i]), fed[i] + in_chunk)
        chunks = [streams[i][pos[i]:fed[i]] for i in live]
        data, in_off = _csr(chunks)
        slots = np.arange(len(live), dtype=np.uint64) * out_cap
        caps = np.full(len(live), out_cap, dtype=np.uint64)
        got = ctx.decode(data, in_off, out_capacity=out_cap * len(live) + 8, out_offsets=slots, out_caps=caps,
                         state=(state[0][live], state[1][live]))
        for j, i in enumerate(live):
            out = np.zeros(out_cap, dtype=np.uint8)
            rc, used, olen = oracle.decode_call(decoders[i], chunks[j], out, 0, out_cap)
            assert int(got["status"][j]) == rc, "stream %d round %d status" % (i, rounds)
            assert int(got["consumed"][j]) == used
            assert int(got["out_lens"][j]) == olen
            assert np.array_equal(got["out"][j * out_cap:j * out_cap + olen], out[:olen])
            assert int(got["leftover_num_bits"][j]) == decoders[i].num_bits
            assert int(got["leftover_working_bits"][j]) == decoders[i].working_bits
            outputs[i].append(out[:olen].copy())
            pos[i] += used
            state[0][i] = got["leftover_working_bits"][j]
            state[1][i] = got["leftover_num_bits"][j]
            assert rc in (OK, SHORT_BUFFER)
            if rc == OK and fed[i] == len(streams[i]) and pos[i] == fed[i]:
                done[i] = True
    for i in range(n):
        want = texts[int(offs[i]):int(offs[i + 1])]
        have = np.concatenate(outputs[i]) if outputs[i] else np.zeros(0, np.uint8)
        # (a stream cut at a byte boundary may decode padding bits as symbols only if they form a code: the
        # oracle does the same, and the whole-stream result must still start with the text)
        assert np.array_equal(have[:len(want)], want), "stream %d" % i


def test_packed_layout_with_carried_decoder_state(contexts, oracle, oracle_tables):
    """Packed layout (no capacities): chunked input only; the library scans the lengths itself."""
    rng = np.random.default_rng(5)
    table = oracle_tables["hpack"]
    ctx = contexts("hpack")
    n = 300
    texts, offs = refcodec.random_batch(rng, n=n, min_len=1, max_len=200, table="hpack")
    enc = oracle.encode_batch(table, 0xFF, texts, offs, out_capacity=4 * len(texts) + 16)
    streams = [enc["out"][int(enc["out_offsets"][i]):int(enc["out_offsets"][i + 1])] for i in range(n)]
    decoders = [oracle.new_decoder(table) for _ in range(n)]
    state = (np.zeros(n, np.uint64), np.zeros(n, np.uint8))
    cut = [len(s) // 2 for s in streams]
    for part in range(2):
        chunks = [s[:c] if part == 0 else s[c:] for s, c in zip(streams, cut)]
        data, in_off = _csr(chunks)
        got = ctx.decode(data, in_off, out_capacity=2 * len(texts) + 64, state=state)
        for i in range(n):
            out = np.zeros(600, dtype=np.uint8)
            rc, used, olen = oracle.decode_call(decoders[i], chunks[i], out, 0, 600)
            a, b = int(got["out_offsets"][i]), int(got["out_offsets"][i + 1])
            assert b - a == olen and int(got["status"][i]) == rc and int(got["consumed"][i]) == used
            assert np.array_equal(got["out"][a:b], out[:olen])
            assert int(got["leftover_num_bits"][i]) == decoders[i].num_bits
            assert int(got["leftover_working_bits"][i]) == decoders[i].working_bits
        state = (got["leftover_working_bits"], got["leftover_num_bits"])


def test_unknown_symbol_after_pending_bits(pkg, oracle, ref_free_masked_coder):
    """Pending overflow bits are written, then a symbol without a code stops the item (huffman.c:62-64)."""
    coder, patterns, num_bits = ref_free_masked_coder
    table = oracle.table(patterns, num_bits)
    ctx = pkg.BatchContext(coder, eos_padding=0xFF, device=0)
    try:
        known = [s for s in range(256) if num_bits[s]]
        unknown = [s for s in range(256) if not num_bits[s]]
        rng = np.random.default_rng(9)
        n = 64
        streams = []
        for i in range(n):
            body = rng.choice(known, size=int(rng.integers(3, 40))).astype(np.uint8)
            if i % 2:
                body[int(rng.integers(1, len(body)))] = unknown[i % len(unknown)]
            streams.append(body)
        encoders = [oracle.new_encoder(table, 0xFF) for _ in range(n)]
        # first call: one byte of room, so most items come back short with pending bits
        state = (np.zeros(n, np.uint32), np.zeros(n, np.uint8))
        pos = np.zeros(n, dtype=np.int64)
        for cap in (1, 64):
            chunks = [streams[i][pos[i]:] for i in range(n)]
            data, in_off = _csr(chunks)
            slots = np.arange(n, dtype=np.uint64) * cap
            got = ctx.encode(data, in_off, out_capacity=cap * n + 8, out_offsets=slots,
                             out_caps=np.full(n, cap, dtype=np.uint64), state=state)
            for i in range(n):
                out = np.zeros(cap, dtype=np.uint8)
                rc, used, olen = oracle.encode_call(encoders[i], chunks[i], out, 0, cap)
                assert int(got["status"][i]) == rc and int(got["consumed"][i]) == used and int(got["out_lens"][i]) == olen
                assert np.array_equal(got["out"][i * cap:i * cap + olen], out[:olen])
                assert int(got["overflow_num_bits"][i]) == encoders[i].overflow_bits.num_bits
                pos[i] += used
            state = (got["overflow_pattern"], got["overflow_num_bits"])
        assert UNKNOWN_SYMBOL in set(int(s) for s in got["status"])
    finally:
        ctx.close()


def test_resume_needs_its_state_arrays(contexts, pkg):
    ctx = contexts("hpack")
    data = np.frombuffer(b"abc", dtype=np.uint8)
    offs = np.array([0, 3], dtype=np.uint64)
    arrays = {"in_": data, "in_offsets": offs, "out": np.zeros(16, np.uint8), "out_offsets": np.zeros(2, np.uint64)}
    with pytest.raises(pkg.CodecError):
        ctx._call("aws_huffman_encode_batch_resume", 1, arrays, 16)
    with pytest.raises(pkg.CodecError):
        ctx._call("aws_huffman_decode_batch_resume", 1, arrays, 16)
