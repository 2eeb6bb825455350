import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package with its native libraries built (idempotent; seconds when up to date)."""
    p = graft.load_package()
    p._build.build()
    graft.build_oracle()
    return p


@pytest.fixture(scope="session")
def product(pkg):
    return pkg.product_library()


@pytest.fixture(scope="session")
def coders(pkg):
    return pkg.coders_library()


@pytest.fixture(scope="session")
def oracle(pkg):
    import refcodec
    return refcodec.OracleLib()


@pytest.fixture(scope="session")
def ref(pkg):
    """The unmodified reference build; absent when oracle/_ref was never built."""
    import refcodec
    if not refcodec.RefLib.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return refcodec.RefLib()


@pytest.fixture(scope="session")
def oracle_tables(oracle):
    import refcodec
    return {name: oracle.table(*refcodec.table_arrays(name)) for name in ("test", "hpack")}


@pytest.fixture(scope="session")
def ref_free_masked_coder(pkg, coders):
    """A Python-callback coder (no reference needed): the test table with 8 symbols removed from
    its ENCODE side. Returns (coder struct, patterns, num_bits)."""
    import numpy as np
    import refcodec
    capi = pkg.capi
    patterns, num_bits = refcodec.table_arrays("test")
    num_bits = num_bits.copy()
    num_bits[[0, 7, 65, 97, 101, 128, 200, 255]] = 0
    inner = coders.coder("test").contents

    def decode(bits):
        import ctypes as C
        tmp = C.c_uint8(0)
        used = inner.decode(bits, C.byref(tmp), None)
        return None if used == 0 or num_bits[tmp.value] == 0 else (tmp.value, used)

    coder = capi.python_coder(lambda sym: (patterns[sym], num_bits[sym]), decode)
    return coder, patterns, num_bits
