"""B200-native Huffman codec behind the aws-c-compression C API.

The product is the C-ABI shared library lib/libaws-c-compression.so (host streaming codec in
host/, batched CUDA codec in csrc/). This Python package is only the thin ctypes view of that
ABI that tests/, bench.py and __graft_entry__.py drive it through; it contains no codec logic and
no fallback: if the library is missing it raises.

The directory name has a hyphen (it is named after the reference); import it through
`__graft_entry__.load_package()` which registers it as `aws_c_compression_b200`.
"""
from . import build as _build  # noqa: F401
from .capi import (  # noqa: F401
    AWS_ERROR_COMPRESSION_DEVICE_FAILURE,
    AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE,
    AWS_ERROR_COMPRESSION_INVALID_PADDING,
    HPACK_HUFFMAN_ALWAYS,
    HPACK_HUFFMAN_NEVER,
    HPACK_HUFFMAN_SMALLEST,
    AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL,
    AWS_ERROR_SHORT_BUFFER,
    BatchContext,
    CodecError,
    Library,
    TableBuilder,
    coders_library,
    product_library,
)
