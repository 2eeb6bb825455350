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
