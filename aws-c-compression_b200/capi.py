"""ctypes bindings for include/aws/compression/*.h. Layouts mirror the C structs one for one."""
import ctypes as C
import os

import numpy as np

from . import build as _build

AWS_OP_SUCCESS = 0
AWS_OP_ERR = -1
AWS_ERROR_SHORT_BUFFER = 4
AWS_ERROR_INVALID_ARGUMENT = 34
AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL = 3072
AWS_ERROR_COMPRESSION_DEVICE_FAILURE = 3073
AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE = 3074
AWS_ERROR_COMPRESSION_INVALID_PADDING = 3075
HPACK_HUFFMAN_SMALLEST, HPACK_HUFFMAN_NEVER, HPACK_HUFFMAN_ALWAYS = 0, 1, 2


class CodecError(RuntimeError):
    def __init__(self, code, what):
        super().__init__("%s failed: aws error %d" % (what, code))
        self.code = code


class aws_huffman_code(C.Structure):
    _fields_ = [("pattern", C.c_uint32), ("num_bits", C.c_uint8)]


ENCODE_FN = C.CFUNCTYPE(aws_huffman_code, C.c_uint8, C.c_void_p)
DECODE_FN = C.CFUNCTYPE(C.c_uint8, C.c_uint32, C.POINTER(C.c_uint8), C.c_void_p)


# ctypes cannot RETURN a struct from a Python callback. On x86-64 SysV an 8-byte integer-class struct
# comes back in RAX, so a Python encode callback is declared as returning uint64 = pattern | bits << 32.
ENCODE_FN_PY = C.CFUNCTYPE(C.c_uint64, C.c_uint8, C.c_void_p)


class aws_huffman_symbol_coder(C.Structure):
    _fields_ = [("encode", ENCODE_FN), ("decode", DECODE_FN), ("userdata", C.c_void_p)]


class aws_huffman_encoder(C.Structure):
    _fields_ = [("coder", C.POINTER(aws_huffman_symbol_coder)), ("eos_padding", C.c_uint8),
                ("overflow_bits", aws_huffman_code)]


class aws_huffman_decoder(C.Structure):
    _fields_ = [("coder", C.POINTER(aws_huffman_symbol_coder)), ("allow_growth", C.c_bool),
                ("working_bits", C.c_uint64), ("num_bits", C.c_uint8)]


class aws_byte_cursor(C.Structure):
    _fields_ = [("len", C.c_size_t), ("ptr", C.c_void_p)]


class aws_byte_buf(C.Structure):
    _fields_ = [("len", C.c_size_t), ("buffer", C.c_void_p), ("capacity", C.c_size_t), ("allocator", C.c_void_p)]


class aws_huffman_batch(C.Structure):
    _fields_ = [
        ("n", C.c_size_t),
        ("in_", C.c_void_p),
        ("in_offsets", C.c_void_p),
        ("in_size", C.c_uint64),
        ("out", C.c_void_p),
        ("out_capacity", C.c_uint64),
        ("out_offsets", C.c_void_p),
        ("out_caps", C.c_void_p),
        ("out_lens", C.c_void_p),
        ("status", C.c_void_p),
        ("consumed", C.c_void_p),
        ("overflow_pattern", C.c_void_p),
        ("overflow_num_bits", C.c_void_p),
        ("leftover_working_bits", C.c_void_p),
        ("leftover_num_bits", C.c_void_p),
    ]


class aws_huffman_table_coder(C.Structure):
    _fields_ = [("coder", aws_huffman_symbol_coder), ("codes", aws_huffman_code * 256),
                ("lut_entries", C.c_void_p), ("lut_count", C.c_uint32), ("lut_root_bits", C.c_uint8)]


def python_coder(encode, decode=None):
    """A symbol coder backed by Python callables (tests only).
    encode(sym) -> (pattern, num_bits); decode(bits) -> (symbol, num_bits) or None."""
    def enc_cb(sym, _):
        pattern, nbits = encode(sym)
        return (int(pattern) & 0xFFFFFFFF) | (int(nbits) << 32)

    def dec_cb(bits, out_sym, _):
        hit = decode(bits)
        if not hit:
            return 0
        out_sym[0] = hit[0]
        return hit[1]

    enc_c = ENCODE_FN_PY(enc_cb)
    dec_c = DECODE_FN(dec_cb) if decode else C.cast(None, DECODE_FN)
    coder = aws_huffman_symbol_coder(C.cast(enc_c, ENCODE_FN), dec_c, None)
    coder._keepalive = (enc_c, dec_c)
    return coder


assert C.sizeof(aws_huffman_code) == 8 and C.sizeof(aws_huffman_symbol_coder) == 24
assert C.sizeof(aws_huffman_encoder) == 24 and C.sizeof(aws_huffman_decoder) == 32

# every symbol include/aws/compression/*.h declares (tests check the library exports them all)
EXPORTED_SYMBOLS = [
    "aws_compression_library_init", "aws_compression_library_clean_up",
    "aws_huffman_encoder_init", "aws_huffman_encoder_reset", "aws_huffman_decoder_init",
    "aws_huffman_decoder_reset", "aws_huffman_get_encoded_length", "aws_huffman_encode",
    "aws_huffman_decode", "aws_huffman_decoder_allow_growth",
    "huffman_test_transitive", "huffman_test_transitive_chunked",
    "aws_huffman_batch_ctx_new", "aws_huffman_batch_ctx_new_from_code_table", "aws_huffman_batch_ctx_destroy",
    "aws_huffman_encode_batch",
    "aws_huffman_decode_batch", "aws_huffman_encode_batch_device", "aws_huffman_decode_batch_device",
    "aws_huffman_encode_batch_resume", "aws_huffman_decode_batch_resume",
    "aws_huffman_encode_batch_resume_device", "aws_huffman_decode_batch_resume_device",
    "aws_huffman_histogram", "aws_huffman_histogram_device", "aws_huffman_code_lengths_from_counts",
    "aws_huffman_canonical_codes", "aws_huffman_code_table_from_counts", "aws_huffman_code_table_write_def",
    "aws_huffman_table_coder_init", "aws_huffman_table_coder_clean_up",
    "aws_hpack_string_encode_batch", "aws_hpack_string_decode_batch",
    "aws_hpack_string_encode_batch_device", "aws_hpack_string_decode_batch_device",
    "aws_huffman_get_encoded_length_batch", "aws_huffman_get_encoded_length_batch_device",
    "aws_huffman_encode_batch_multi", "aws_huffman_decode_batch_multi", "aws_huffman_batch_ctx_synchronize",
    "aws_huffman_batch_ctx_stream", "aws_huffman_batch_ctx_device", "aws_huffman_batch_ctx_launch_count",
    "aws_huffman_batch_plan_shards", "aws_huffman_batch_concat_offsets",
]


def bind_streaming_api(lib):
    """Declares the reference-compatible streaming API on a CDLL (ours or the reference build)."""
    P = C.POINTER
    lib.aws_huffman_encoder_init.argtypes = [P(aws_huffman_encoder), P(aws_huffman_symbol_coder)]
    lib.aws_huffman_encoder_init.restype = None
    lib.aws_huffman_encoder_reset.argtypes = [P(aws_huffman_encoder)]
    lib.aws_huffman_encoder_reset.restype = None
    lib.aws_huffman_decoder_init.argtypes = [P(aws_huffman_decoder), P(aws_huffman_symbol_coder)]
    lib.aws_huffman_decoder_init.restype = None
    lib.aws_huffman_decoder_reset.argtypes = [P(aws_huffman_decoder)]
    lib.aws_huffman_decoder_reset.restype = None
    lib.aws_huffman_decoder_allow_growth.argtypes = [P(aws_huffman_decoder), C.c_bool]
    lib.aws_huffman_decoder_allow_growth.restype = None
    lib.aws_huffman_get_encoded_length.argtypes = [P(aws_huffman_encoder), aws_byte_cursor]
    lib.aws_huffman_get_encoded_length.restype = C.c_size_t
    lib.aws_huffman_encode.argtypes = [P(aws_huffman_encoder), P(aws_byte_cursor), P(aws_byte_buf)]
    lib.aws_huffman_encode.restype = C.c_int
    lib.aws_huffman_decode.argtypes = [P(aws_huffman_decoder), P(aws_byte_cursor), P(aws_byte_buf)]
    lib.aws_huffman_decode.restype = C.c_int
    lib.huffman_test_transitive.argtypes = [P(aws_huffman_symbol_coder), C.c_char_p, C.c_size_t, C.c_size_t,
                                            P(C.c_char_p)]
    lib.huffman_test_transitive.restype = C.c_int
    lib.huffman_test_transitive_chunked.argtypes = [P(aws_huffman_symbol_coder), C.c_char_p, C.c_size_t, C.c_size_t,
                                                    C.c_size_t, P(C.c_char_p)]
    lib.huffman_test_transitive_chunked.restype = C.c_int
    lib.aws_last_error.restype = C.c_int
    lib.aws_reset_error.restype = None
    lib.aws_error_name.argtypes = [C.c_int]
    lib.aws_error_name.restype = C.c_char_p
    lib.aws_compression_library_init.argtypes = [C.c_void_p]
    lib.aws_compression_library_init.restype = None
    lib.aws_compression_library_clean_up.restype = None
    lib.aws_default_allocator.restype = C.c_void_p
    lib.aws_byte_buf_init.argtypes = [P(aws_byte_buf), C.c_void_p, C.c_size_t]
    lib.aws_byte_buf_init.restype = C.c_int
    lib.aws_byte_buf_clean_up.argtypes = [P(aws_byte_buf)]
    lib.aws_byte_buf_clean_up.restype = None
    return lib


class Library:
    """The product library."""

    def __init__(self, path=None):
        # AWS_HUFFMAN_B200_LIB: another build of the same library (kernel A/B runs, tools/build_variant.sh)
        path = path or os.environ.get("AWS_HUFFMAN_B200_LIB") or _build.PRODUCT_LIB
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no Python or CPU fallback for the batched codec." % path)
        self.path = path
        self.lib = bind_streaming_api(C.CDLL(path))
        L, P = self.lib, C.POINTER
        L.aws_huffman_batch_ctx_new.argtypes = [P(C.c_void_p), P(aws_huffman_symbol_coder), C.c_uint8, C.c_int]
        L.aws_huffman_batch_ctx_new.restype = C.c_int
        L.aws_huffman_batch_ctx_new_from_code_table.argtypes = [P(C.c_void_p), C.c_void_p, C.c_uint8, C.c_int]
        L.aws_huffman_batch_ctx_new_from_code_table.restype = C.c_int
        L.aws_huffman_batch_ctx_destroy.argtypes = [C.c_void_p]
        L.aws_huffman_batch_ctx_destroy.restype = None
        for name in ("aws_huffman_encode_batch", "aws_huffman_decode_batch",
                     "aws_huffman_encode_batch_resume", "aws_huffman_decode_batch_resume"):
            getattr(L, name).argtypes = [C.c_void_p, P(aws_huffman_batch)]
            getattr(L, name).restype = C.c_int
        for name in ("aws_huffman_encode_batch_device", "aws_huffman_decode_batch_device",
                     "aws_huffman_encode_batch_resume_device", "aws_huffman_decode_batch_resume_device"):
            getattr(L, name).argtypes = [C.c_void_p, P(aws_huffman_batch), C.c_void_p]
            getattr(L, name).restype = C.c_int
        L.aws_huffman_get_encoded_length_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.aws_huffman_get_encoded_length_batch.restype = C.c_int
        L.aws_huffman_get_encoded_length_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                                                 C.c_void_p]
        L.aws_huffman_get_encoded_length_batch_device.restype = C.c_int
        for name in ("aws_huffman_encode_batch_multi", "aws_huffman_decode_batch_multi"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_size_t, P(aws_huffman_batch)]
            getattr(L, name).restype = C.c_int
        L.aws_huffman_batch_ctx_synchronize.argtypes = [C.c_void_p]
        L.aws_huffman_batch_ctx_synchronize.restype = C.c_int
        L.aws_huffman_batch_ctx_stream.argtypes = [C.c_void_p]
        L.aws_huffman_batch_ctx_stream.restype = C.c_void_p
        L.aws_huffman_batch_ctx_device.argtypes = [C.c_void_p]
        L.aws_huffman_batch_ctx_device.restype = C.c_int
        L.aws_huffman_batch_ctx_launch_count.argtypes = [C.c_void_p]
        L.aws_huffman_batch_ctx_launch_count.restype = C.c_uint64
        L.aws_huffman_batch_plan_shards.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.aws_huffman_batch_plan_shards.restype = C.c_int
        L.aws_huffman_batch_concat_offsets.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.aws_huffman_batch_concat_offsets.restype = C.c_int

    def last_error(self):
        return self.lib.aws_last_error()

    def plan_shards(self, in_offsets, num_shards):
        in_offsets = np.ascontiguousarray(in_offsets, dtype=np.uint64)
        n = len(in_offsets) - 1
        begin = np.zeros(num_shards + 1, dtype=np.uint64)  # size_t
        rc = self.lib.aws_huffman_batch_plan_shards(in_offsets.ctypes.data, n, num_shards, begin.ctypes.data)
        if rc != 0:
            raise CodecError(self.last_error(), "aws_huffman_batch_plan_shards")
        return begin.astype(np.int64)

    def concat_offsets(self, shard_offsets):
        arrays = [np.ascontiguousarray(a, dtype=np.uint64) for a in shard_offsets]
        ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        items = np.array([len(a) - 1 for a in arrays], dtype=np.uint64)
        out = np.zeros(int(items.sum()) + 1, dtype=np.uint64)
        rc = self.lib.aws_huffman_batch_concat_offsets(ptrs, items.ctypes.data, len(arrays), out.ctypes.data)
        if rc != 0:
            raise CodecError(self.last_error(), "aws_huffman_batch_concat_offsets")
        return out


_product = None
_coders = None


class TableBuilder:
    """include/aws/compression/huffman_table_builder.h through ctypes (tests and tools)."""

    def __init__(self, library=None):
        self.library = library or product_library()
        L, P = self.library.lib, C.POINTER
        L.aws_huffman_histogram.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.aws_huffman_histogram_device.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.aws_huffman_code_lengths_from_counts.argtypes = [C.c_void_p, C.c_uint, C.c_bool, C.c_bool, C.c_void_p, P(C.c_uint8)]
        L.aws_huffman_canonical_codes.argtypes = [C.c_void_p, C.c_uint8, C.c_void_p, P(aws_huffman_code)]
        L.aws_huffman_code_table_from_counts.argtypes = [C.c_void_p, C.c_uint, C.c_bool, C.c_bool, C.c_void_p, P(aws_huffman_code)]
        L.aws_huffman_code_table_write_def.argtypes = [C.c_void_p, C.c_char_p]
        L.aws_huffman_table_coder_init.argtypes = [P(aws_huffman_table_coder), C.c_void_p]
        L.aws_huffman_table_coder_clean_up.argtypes = [P(aws_huffman_table_coder)]
        L.aws_huffman_table_coder_clean_up.restype = None

    def _check(self, rc, what):
        if rc != 0:
            raise CodecError(self.library.last_error(), what)

    def histogram(self, data, device=0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        counts = np.zeros(256, dtype=np.uint64)
        self._check(self.library.lib.aws_huffman_histogram(device, data.ctypes.data, len(data), counts.ctypes.data),
                    "aws_huffman_histogram")
        return counts

    def histogram_device(self, tensor, counts_tensor, stream=None):
        self._check(self.library.lib.aws_huffman_histogram_device(tensor.data_ptr(), tensor.numel(), counts_tensor.data_ptr(),
                                                                  stream or 0), "aws_huffman_histogram_device")

    def lengths(self, counts, max_bits=32, cover_all=True, reserve_eos=False):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        lens = np.zeros(256, dtype=np.uint8)
        eos = C.c_uint8(0)
        self._check(self.library.lib.aws_huffman_code_lengths_from_counts(counts.ctypes.data, max_bits, cover_all, reserve_eos,
                                                                          lens.ctypes.data, C.byref(eos)),
                    "aws_huffman_code_lengths_from_counts")
        return lens, eos.value

    def codes(self, counts, max_bits=32, cover_all=True, reserve_eos=False):
        """Returns (aws_huffman_code[256] ctypes array, eos aws_huffman_code)."""
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        table = (aws_huffman_code * 256)()
        eos = aws_huffman_code()
        self._check(self.library.lib.aws_huffman_code_table_from_counts(counts.ctypes.data, max_bits, cover_all, reserve_eos,
                                                                        C.cast(table, C.c_void_p), C.byref(eos)),
                    "aws_huffman_code_table_from_counts")
        return table, eos

    def canonical(self, lengths, eos_length=0):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
        table = (aws_huffman_code * 256)()
        eos = aws_huffman_code()
        self._check(self.library.lib.aws_huffman_canonical_codes(lengths.ctypes.data, eos_length, C.cast(table, C.c_void_p),
                                                                 C.byref(eos)), "aws_huffman_canonical_codes")
        return table, eos

    def write_def(self, table, path):
        self._check(self.library.lib.aws_huffman_code_table_write_def(C.cast(table, C.c_void_p), path.encode()),
                    "aws_huffman_code_table_write_def")

    def table_coder(self, table):
        coder = aws_huffman_table_coder()
        self._check(self.library.lib.aws_huffman_table_coder_init(C.byref(coder), C.cast(table, C.c_void_p)),
                    "aws_huffman_table_coder_init")
        return coder

    def table_coder_clean_up(self, coder):
        self.library.lib.aws_huffman_table_coder_clean_up(C.byref(coder))


def product_library():
    global _product
    if _product is None:
        _product = Library()
    return _product


class CodersLibrary:
    """Coders emitted by OUR generator: hpack and (from the golden fixture) the reference test table."""

    def __init__(self, path=None):
        path = path or _build.CODERS_LIB
        if not os.path.exists(path):
            raise FileNotFoundError("%s is missing: run __graft_entry__.build()" % path)
        self.lib = C.CDLL(path)

    def coder(self, name):
        fn = getattr(self.lib, name + "_get_coder")
        fn.restype = C.POINTER(aws_huffman_symbol_coder)
        return fn()

    def code_table_pointer(self, name):
        fn = getattr(self.lib, name + "_get_code_table")
        fn.restype = C.POINTER(aws_huffman_code * 256)
        return fn()

    def code_table(self, name):
        fn = getattr(self.lib, name + "_get_code_table")
        fn.restype = C.POINTER(aws_huffman_code * 256)
        table = fn().contents
        return (np.array([c.pattern for c in table], dtype=np.uint32),
                np.array([c.num_bits for c in table], dtype=np.uint8))


def coders_library():
    global _coders
    if _coders is None:
        _coders = CodersLibrary()
    return _coders


def _ptr(x):
    """Address of a numpy array / torch tensor / None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch tensor (host or device)


class BatchContext:
    """aws_huffman_batch_ctx for one (coder, eos_padding, device)."""

    def __init__(self, coder, eos_padding=0xFF, device=0, library=None, code_table=None):
        """`coder`: a pointer to aws_huffman_symbol_coder, or None with `code_table` = pointer to 256
        aws_huffman_code entries (aws_huffman_batch_ctx_new_from_code_table: no callbacks)."""
        self.library = library or product_library()
        self._keep = coder  # callbacks must stay alive while the C side probes them
        handle = C.c_void_p()
        if code_table is not None:
            rc = self.library.lib.aws_huffman_batch_ctx_new_from_code_table(
                C.byref(handle), C.cast(code_table, C.c_void_p), eos_padding, device)
            if rc != 0:
                raise CodecError(self.library.last_error(), "aws_huffman_batch_ctx_new_from_code_table")
        else:
            coder_ptr = coder if not isinstance(coder, aws_huffman_symbol_coder) else C.pointer(coder)
            rc = self.library.lib.aws_huffman_batch_ctx_new(C.byref(handle), coder_ptr, eos_padding, device)
            if rc != 0:
                raise CodecError(self.library.last_error(), "aws_huffman_batch_ctx_new")
        self.handle = handle
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.library.lib.aws_huffman_batch_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return self.library.lib.aws_huffman_batch_ctx_stream(self.handle)

    @property
    def launch_count(self):
        return int(self.library.lib.aws_huffman_batch_ctx_launch_count(self.handle))

    def synchronize(self):
        if self.library.lib.aws_huffman_batch_ctx_synchronize(self.handle) != 0:
            raise CodecError(self.library.last_error(), "aws_huffman_batch_ctx_synchronize")

    def _call(self, fn_name, n, arrays, out_capacity, stream=None, in_size=0):
        b = aws_huffman_batch()
        b.n = n
        b.in_size = in_size
        b.out_capacity = out_capacity
        for key, value in arrays.items():
            setattr(b, key, _ptr(value))
        fn = getattr(self.library.lib, fn_name)
        rc = fn(self.handle, C.byref(b)) if stream is None else fn(self.handle, C.byref(b), stream)
        if rc != 0:
            raise CodecError(self.library.last_error(), fn_name)

    # ---- host-buffer entry points (numpy in, numpy out) ----
    def _host(self, encode, data, in_offsets, out_capacity, out_offsets=None, out_caps=None, out=None, extras=True,
              state=None):
        """state = (overflow_pattern, overflow_num_bits) resp. (leftover_working_bits, leftover_num_bits) of the
        previous call: the *_resume entry points (the arrays come back updated in the result)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        in_offsets = np.ascontiguousarray(in_offsets, dtype=np.uint64)
        n = len(in_offsets) - 1
        slotted = out_caps is not None
        res = {
            "out": out if out is not None else np.zeros(max(int(out_capacity), 1), dtype=np.uint8),
            "out_offsets": (np.ascontiguousarray(out_offsets, dtype=np.uint64) if slotted
                            else np.zeros(n + 1, dtype=np.uint64)),
            "out_lens": np.zeros(n, dtype=np.uint64),
            "status": np.zeros(n, dtype=np.int32),
            "consumed": np.zeros(n, dtype=np.uint64),
        }
        if not extras:  # only payload, offsets and status (what a throughput-minded caller asks for)
            del res["out_lens"], res["consumed"]
        elif encode:
            res["overflow_pattern"] = np.zeros(n, dtype=np.uint32)
            res["overflow_num_bits"] = np.zeros(n, dtype=np.uint8)
        else:
            res["leftover_working_bits"] = np.zeros(n, dtype=np.uint64)
            res["leftover_num_bits"] = np.zeros(n, dtype=np.uint8)
        if state is not None:
            if encode:
                res["overflow_pattern"] = np.array(state[0], dtype=np.uint32)
                res["overflow_num_bits"] = np.array(state[1], dtype=np.uint8)
            else:
                res["leftover_working_bits"] = np.array(state[0], dtype=np.uint64)
                res["leftover_num_bits"] = np.array(state[1], dtype=np.uint8)
        arrays = dict(res)
        arrays["in_"] = data
        arrays["in_offsets"] = in_offsets
        if slotted:
            arrays["out_caps"] = np.ascontiguousarray(out_caps, dtype=np.uint64)
        name = "aws_huffman_encode_batch" if encode else "aws_huffman_decode_batch"
        self._call(name + ("_resume" if state is not None else ""), n, arrays, int(out_capacity))
        return res

    # ---- HPACK string literals (hpack_string_batch.h), host buffers ----
    def hpack_encode_strings(self, data, in_offsets, out_capacity, mode=HPACK_HUFFMAN_SMALLEST):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        in_offsets = np.ascontiguousarray(in_offsets, dtype=np.uint64)
        n = len(in_offsets) - 1
        res = {"out": np.zeros(max(int(out_capacity), 1), dtype=np.uint8), "out_offsets": np.zeros(n + 1, dtype=np.uint64)}
        fn = self.library.lib.aws_hpack_string_encode_batch
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        fn.restype = C.c_int
        rc = fn(self.handle, n, data.ctypes.data, in_offsets.ctypes.data, mode, res["out"].ctypes.data, int(out_capacity),
                res["out_offsets"].ctypes.data)
        if rc != 0:
            err = CodecError(self.library.last_error(), "aws_hpack_string_encode_batch")
            err.result = res
            raise err
        return res

    def hpack_decode_strings(self, data, in_offsets, out_capacity):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        in_offsets = np.ascontiguousarray(in_offsets, dtype=np.uint64)
        n = len(in_offsets) - 1
        res = {"out": np.zeros(max(int(out_capacity), 1), dtype=np.uint8), "out_offsets": np.zeros(n + 1, dtype=np.uint64),
               "status": np.zeros(n, dtype=np.int32)}
        fn = self.library.lib.aws_hpack_string_decode_batch
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        rc = fn(self.handle, n, data.ctypes.data, in_offsets.ctypes.data, res["out"].ctypes.data, int(out_capacity),
                res["out_offsets"].ctypes.data, res["status"].ctypes.data)
        if rc != 0:
            err = CodecError(self.library.last_error(), "aws_hpack_string_decode_batch")
            err.result = res
            raise err
        return res

    def hpack_device(self, encode, n, in_, in_offsets, in_size, out, out_capacity, out_offsets, status=None, mode=0, stream=None):
        L = self.library.lib
        if encode:
            fn = L.aws_hpack_string_encode_batch_device
            fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64,
                           C.c_void_p, C.c_void_p]
            rc = fn(self.handle, n, _ptr(in_), _ptr(in_offsets), int(in_size), mode, _ptr(out), int(out_capacity),
                    _ptr(out_offsets), stream or 0)
        else:
            fn = L.aws_hpack_string_decode_batch_device
            fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                           C.c_void_p, C.c_void_p]
            rc = fn(self.handle, n, _ptr(in_), _ptr(in_offsets), int(in_size), _ptr(out), int(out_capacity),
                    _ptr(out_offsets), _ptr(status), stream or 0)
        if rc != 0:
            raise CodecError(self.library.last_error(), "aws_hpack_string_%s_batch_device" % ("encode" if encode else "decode"))

    def encode(self, data, in_offsets, out_capacity, **kw):
        return self._host(True, data, in_offsets, out_capacity, **kw)

    def decode(self, data, in_offsets, out_capacity, **kw):
        return self._host(False, data, in_offsets, out_capacity, **kw)

    def encoded_lengths(self, data, in_offsets):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        in_offsets = np.ascontiguousarray(in_offsets, dtype=np.uint64)
        n = len(in_offsets) - 1
        lens = np.zeros(n, dtype=np.uint64)
        rc = self.library.lib.aws_huffman_get_encoded_length_batch(
            self.handle, data.ctypes.data, in_offsets.ctypes.data, n, lens.ctypes.data)
        if rc != 0:
            raise CodecError(self.library.last_error(), "aws_huffman_get_encoded_length_batch")
        return lens

    def encoded_lengths_device(self, n, in_, in_offsets, lens, stream=None):
        rc = self.library.lib.aws_huffman_get_encoded_length_batch_device(
            self.handle, _ptr(in_), _ptr(in_offsets), n, _ptr(lens), stream or 0)
        if rc != 0:
            raise CodecError(self.library.last_error(), "aws_huffman_get_encoded_length_batch_device")

    # ---- device-buffer entry points (torch CUDA tensors; enqueued on `stream`) ----
    def encode_device(self, n, arrays, in_size, out_capacity, stream=None):
        self._call("aws_huffman_encode_batch_device", n, arrays, int(out_capacity), stream=stream or 0,
                   in_size=int(in_size))

    def decode_device(self, n, arrays, in_size, out_capacity, stream=None):
        self._call("aws_huffman_decode_batch_device", n, arrays, int(out_capacity), stream=stream or 0,
                   in_size=int(in_size))


def run_multi(contexts, encode, data, in_offsets, out_capacity):
    """aws_huffman_encode_batch_multi / aws_huffman_decode_batch_multi: one packed batch over several contexts
    (host buffers, numpy in / numpy out)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    in_offsets = np.ascontiguousarray(in_offsets, dtype=np.uint64)
    n = len(in_offsets) - 1
    res = {"out": np.zeros(max(int(out_capacity), 1), dtype=np.uint8), "out_offsets": np.zeros(n + 1, dtype=np.uint64),
           "out_lens": np.zeros(n, dtype=np.uint64), "status": np.zeros(n, dtype=np.int32),
           "consumed": np.zeros(n, dtype=np.uint64)}
    b = aws_huffman_batch()
    b.n = n
    b.in_size = int(in_offsets[n]) if n else 0
    b.out_capacity = int(out_capacity)
    b.in_ = _ptr(data)
    b.in_offsets = _ptr(in_offsets)
    for key, value in res.items():
        setattr(b, key, _ptr(value))
    lib = contexts[0].library
    handles = (C.c_void_p * len(contexts))(*[c.handle for c in contexts])
    fn = lib.lib.aws_huffman_encode_batch_multi if encode else lib.lib.aws_huffman_decode_batch_multi
    rc = fn(handles, len(contexts), C.byref(b))
    if rc != 0:
        raise CodecError(lib.last_error(), "aws_huffman_%s_batch_multi" % ("encode" if encode else "decode"))
    return res
