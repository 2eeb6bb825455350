"""Test infrastructure: ctypes drivers for the two CPU checkers and seeded input generators.

  OracleLib  oracle/_build/liboracle_huffman.so   the restatement (oracle/huffman_oracle.c)
  RefLib     oracle/_ref/libref_huffman.so        the UNMODIFIED reference compiled in place

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module. The product never does.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle_huffman.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_huffman.so")

OK, SHORT_BUFFER, UNKNOWN_SYMBOL = 0, 4, 3072


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a):
    # always a c_void_p: a bare Python int would be passed as a 32-bit C int
    return C.c_void_p(None if a is None else a.ctypes.data)


def _batch_result(n, out_capacity, encode, out_offsets, out):
    res = {
        "out": out if out is not None else np.zeros(max(int(out_capacity), 1), dtype=np.uint8),
        "out_offsets": out_offsets,
        "out_lens": np.zeros(n, dtype=np.uint64),
        "status": np.zeros(n, dtype=np.int32),
        "consumed": np.zeros(n, dtype=np.uint64),
    }
    if encode:
        res["overflow_pattern"] = np.zeros(n, dtype=np.uint32)
        res["overflow_num_bits"] = np.zeros(n, dtype=np.uint8)
    else:
        res["leftover_working_bits"] = np.zeros(n, dtype=np.uint64)
        res["leftover_num_bits"] = np.zeros(n, dtype=np.uint8)
    return res


class _BatchDriver:
    """Shared marshalling for {oracle,ref}_{encode,decode}_batch (same array contract)."""

    def _run(self, fn, head_args, encode, data, in_offsets, out_capacity, out_offsets=None, out_caps=None, out=None):
        data, in_offsets = _u8(data), _u64(in_offsets)
        n = len(in_offsets) - 1
        slotted = out_caps is not None
        offs = _u64(out_offsets).copy() if slotted else np.zeros(n + 1, dtype=np.uint64)
        caps = _u64(out_caps) if slotted else None
        res = _batch_result(n, out_capacity, encode, offs, out)
        tail = ([res["overflow_pattern"], res["overflow_num_bits"]] if encode
                else [res["leftover_working_bits"], res["leftover_num_bits"]])
        fn(*head_args, _p(data), _p(in_offsets), C.c_size_t(n), _p(res["out"]), C.c_uint64(int(out_capacity)),
           _p(offs), _p(caps), _p(res["out_lens"]), _p(res["status"]), _p(res["consumed"]), _p(tail[0]), _p(tail[1]))
        return res


class OracleLib(_BatchDriver):
    class Table(C.Structure):
        _fields_ = [("enc", C.c_uint8 * (8 * 256)), ("nodes", C.c_void_p), ("num_nodes", C.c_int32),
                    ("cap_nodes", C.c_int32)]

    class Code(C.Structure):
        _fields_ = [("pattern", C.c_uint32), ("num_bits", C.c_uint8)]

    class Encoder(C.Structure):
        pass

    class Decoder(C.Structure):
        pass

    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            raise FileNotFoundError("%s missing: run `make -C oracle oracle` (or __graft_entry__.build())" % path)
        self.lib = L = C.CDLL(path)
        OracleLib.Encoder._fields_ = [("table", C.c_void_p), ("eos_padding", C.c_uint8),
                                      ("overflow_bits", OracleLib.Code)]
        OracleLib.Decoder._fields_ = [("table", C.c_void_p), ("working_bits", C.c_uint64), ("num_bits", C.c_uint8)]
        L.oracle_table_init.restype = C.c_int
        L.oracle_decode_symbol.restype = C.c_uint8
        L.oracle_decode_symbol.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint8)]
        L.oracle_encode_symbol.restype = OracleLib.Code
        L.oracle_encode_symbol.argtypes = [C.c_void_p, C.c_uint8]
        L.oracle_get_encoded_length.restype = C.c_size_t
        L.oracle_get_encoded_length.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        for name in ("oracle_encode", "oracle_decode"):
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t,
                           C.POINTER(C.c_size_t)]
        for name in ("oracle_encode_batch", "oracle_decode_batch"):
            getattr(L, name).restype = None

    def table(self, patterns, num_bits):
        t = OracleLib.Table()
        patterns = np.ascontiguousarray(patterns, dtype=np.uint32)
        num_bits = np.ascontiguousarray(num_bits, dtype=np.uint8)
        rc = self.lib.oracle_table_init(C.byref(t), _p(patterns), _p(num_bits))
        if rc != 0:
            raise ValueError("not a prefix code")
        return t

    def encode_symbol(self, table, sym):
        c = self.lib.oracle_encode_symbol(C.byref(table), sym)
        return c.pattern, c.num_bits

    def decode_symbol(self, table, window):
        sym = C.c_uint8(0)
        n = self.lib.oracle_decode_symbol(C.byref(table), window & 0xFFFFFFFF, C.byref(sym))
        return n, sym.value

    def new_encoder(self, table, eos_padding=0xFF):
        e = OracleLib.Encoder()
        self.lib.oracle_encoder_init(C.byref(e), C.byref(table))
        e.eos_padding = eos_padding
        return e

    def new_decoder(self, table):
        d = OracleLib.Decoder()
        self.lib.oracle_decoder_init(C.byref(d), C.byref(table))
        return d

    def encoded_length(self, table, data):
        data = _u8(data)
        e = self.new_encoder(table)
        return self.lib.oracle_get_encoded_length(C.byref(e), _p(data), len(data))

    def encode_call(self, encoder, data, out, out_len, capacity):
        """One aws_huffman_encode call. Returns (rc, consumed, new_out_len)."""
        data = _u8(data)
        used, olen = C.c_size_t(0), C.c_size_t(out_len)
        rc = self.lib.oracle_encode(C.byref(encoder), _p(data), len(data), C.byref(used), _p(out), capacity,
                                    C.byref(olen))
        return rc, used.value, olen.value

    def decode_call(self, decoder, data, out, out_len, capacity):
        data = _u8(data)
        used, olen = C.c_size_t(0), C.c_size_t(out_len)
        rc = self.lib.oracle_decode(C.byref(decoder), _p(data), len(data), C.byref(used), _p(out), capacity,
                                    C.byref(olen))
        return rc, used.value, olen.value

    def encode_batch(self, table, eos_padding, data, in_offsets, out_capacity, **kw):
        return self._run(self.lib.oracle_encode_batch, [C.byref(table), C.c_uint8(eos_padding)], True, data,
                         in_offsets, out_capacity, **kw)

    def decode_batch(self, table, data, in_offsets, out_capacity, **kw):
        return self._run(self.lib.oracle_decode_batch, [C.byref(table)], False, data, in_offsets, out_capacity, **kw)


class RefLib(_BatchDriver):
    """The unmodified reference (plus oracle/ref_extras.c drivers)."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError("%s missing (needs /root/reference at build time)" % path)
        self.lib = L = C.CDLL(path)
        for name in ("test_get_coder", "hpack_get_coder", "ref_masked_coder_new"):
            getattr(L, name).restype = C.c_void_p
        L.ref_masked_coder_new.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_masked_coder_free.argtypes = [C.c_void_p]
        for name in ("ref_encode_batch", "ref_decode_batch"):
            getattr(L, name).restype = None

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def coder(self, name):
        return C.c_void_p(getattr(self.lib, name + "_get_coder")())

    def masked_coder(self, name, unknown_symbols):
        mask = np.zeros(256, dtype=np.uint8)
        mask[list(unknown_symbols)] = 1
        return C.c_void_p(self.lib.ref_masked_coder_new(self.coder(name), _p(mask)))

    def encode_batch(self, coder, eos_padding, data, in_offsets, out_capacity, **kw):
        return self._run(self.lib.ref_encode_batch, [coder, C.c_uint8(eos_padding)], True, data, in_offsets,
                         out_capacity, **kw)

    def decode_batch(self, coder, data, in_offsets, out_capacity, **kw):
        return self._run(self.lib.ref_decode_batch, [coder], False, data, in_offsets, out_capacity, **kw)


# ---------------------------------------------------------------------------------------------
# seeded inputs
# ---------------------------------------------------------------------------------------------

def golden(name):
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", name)))


def table_arrays(table):
    """(patterns, num_bits) of a named table WITHOUT going through product code."""
    if table == "test":
        t = golden("reference_vectors.json")["test_table"]
        return np.array(t["patterns"], dtype=np.uint32), np.array(t["num_bits"], dtype=np.uint8)
    if table == "hpack":
        import importlib.util
        spec = importlib.util.spec_from_file_location("make_hpack_def", os.path.join(ROOT, "tools", "make_hpack_def.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        codes = m.canonical_codes(m.LENGTHS)
        return (np.array([codes[s] for s in range(256)], dtype=np.uint32),
                np.array(m.LENGTHS[:256], dtype=np.uint8))
    raise KeyError(table)


def zipf_symbol_sampler(num_bits, s=1.5):
    """Symbols ranked by (code length, value); P(rank r) ~ 1/(r+1)^s. Returns a 65536-entry table so
    sampling is a gather on uniform 16-bit integers."""
    order = sorted(range(256), key=lambda x: (int(num_bits[x]) if num_bits[x] else 99, x))
    w = 1.0 / np.power(np.arange(1, 257, dtype=np.float64), s)
    cdf = np.cumsum(w / w.sum())
    ranks = np.searchsorted(cdf, (np.arange(65536) + 0.5) / 65536.0)
    return np.array(order, dtype=np.uint8)[np.minimum(ranks, 255)]


def random_batch(rng, n, min_len, max_len, table="hpack", zipf=True):
    lens = rng.integers(min_len, max_len + 1, size=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    if zipf:
        sampler = zipf_symbol_sampler(table_arrays(table)[1])
        data = sampler[rng.integers(0, 65536, size=total)]
    else:
        data = rng.integers(0, 256, size=total, dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8), offsets


def make_differential_cases(ref, table, seed, n_encode=160, n_decode=160):
    """Small seeded cases run through the UNMODIFIED reference; kept as hex in tests/golden/."""
    rng = np.random.default_rng(seed)
    coder = ref.coder(table)
    holes = sorted(int(x) for x in rng.choice(256, size=6, replace=False))
    masked = ref.masked_coder(table, holes)
    cases = {"table": table, "seed": seed, "unknown_symbols": holes, "encode": [], "decode": []}
    for i in range(n_encode):
        length = int(rng.integers(0, 70))
        data = (zipf_symbol_sampler(table_arrays(table)[1])[rng.integers(0, 65536, size=length)]
                if i % 3 else rng.integers(0, 256, size=length, dtype=np.uint8))
        eos = int(rng.choice([0xFF, 0x00, 0x55, 0xAA, 0x0F]))
        use_mask = i % 4 == 3
        full = ref.encode_batch(coder, eos, data, [0, length], 4 * length + 8)
        need = int(full["out_lens"][0])
        cap = int(rng.integers(0, need + 3)) if i % 2 else 4 * length + 8
        r = ref.encode_batch(masked if use_mask else coder, eos, data, [0, length], max(cap, 1),
                             out_offsets=[0], out_caps=[cap])
        cases["encode"].append({
            "in": bytes(data).hex(), "eos": eos, "cap": cap, "masked": use_mask,
            "out": bytes(r["out"][:int(r["out_lens"][0])]).hex(), "status": int(r["status"][0]),
            "consumed": int(r["consumed"][0]), "ovf_pattern": int(r["overflow_pattern"][0]),
            "ovf_bits": int(r["overflow_num_bits"][0])})
    for i in range(n_decode):
        if i % 3 == 0:
            enc = rng.integers(0, 256, size=int(rng.integers(0, 40)), dtype=np.uint8)
        else:
            length = int(rng.integers(0, 60))
            data = zipf_symbol_sampler(table_arrays(table)[1])[rng.integers(0, 65536, size=length)]
            full = ref.encode_batch(coder, 0xFF, data, [0, length], 4 * length + 8)
            enc = full["out"][:int(full["out_lens"][0])].copy()
            if i % 3 == 2 and len(enc):
                enc[int(rng.integers(0, len(enc)))] ^= 1 << int(rng.integers(0, 8))
        full = ref.decode_batch(coder, enc, [0, len(enc)], 8 * len(enc) + 8)
        need = int(full["out_lens"][0])
        cap = int(rng.integers(0, need + 2)) if i % 2 else 8 * len(enc) + 8
        r = ref.decode_batch(coder, enc, [0, len(enc)], max(cap, 1), out_offsets=[0], out_caps=[cap])
        cases["decode"].append({
            "in": bytes(enc).hex(), "cap": cap, "out": bytes(r["out"][:int(r["out_lens"][0])]).hex(),
            "status": int(r["status"][0]), "consumed": int(r["consumed"][0]),
            "left_bits": int(r["leftover_working_bits"][0]), "left_num": int(r["leftover_num_bits"][0])})
    return cases
