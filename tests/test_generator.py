"""The .def -> C coder generator (aws-c-compression_b200/generator/huffman_generator.c): same CLI and
grammar as the reference tool (source/huffman_generator/generator.c:216-226, :42-105), emitted coder
equivalent to the reference's goto-tree coder on every window that matters."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import refcodec


@pytest.fixture(scope="module")
def gen(pkg):
    return pkg._build.GENERATOR


def run(gen, *args):
    return subprocess.run([gen, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def test_usage_error_matches_reference_cli(gen):
    r = run(gen)
    assert r.returncode == 1 and "generator expects 3 arguments" in r.stderr
    assert "struct aws_huffman_symbol_coder *[encoding name]_get_coder()" in r.stderr
    r = run(gen, "/nonexistent.def", "/tmp/x.c", "x")
    assert r.returncode == 1 and "Failed to open file '/nonexistent.def' for read." in r.stdout


def test_grammar_comments_and_preprocessor_lines(gen, pkg, tmp_path):
    src = tmp_path / "t.def"
    src.write_text('#ifndef HUFFMAN_CODE\n#error "x"\n#endif\n'
                   '/* HUFFMAN_CODE(9, "1", 0x1, 1) inside a comment\n   spanning lines */\n'
                   '/*           sym          bits   code len */\n'
                   'HUFFMAN_CODE(  0, "0", 0x0, 1)\nHUFFMAN_CODE(65,"10",0x2,2) HUFFMAN_CODE( 66 , "110" , 0x6 , 3 )\n')
    out = tmp_path / "t.c"
    r = run(gen, str(src), str(out), "tiny")
    assert r.returncode == 0, r.stderr
    text = out.read_text()
    assert "struct aws_huffman_symbol_coder *tiny_get_coder(void)" in text
    so = tmp_path / "libtiny.so"
    inc = [a for d in pkg._build.INCLUDES for a in ("-I", d)]
    subprocess.check_call(["gcc", "-std=gnu99", "-shared", "-fPIC", *inc, "-o", str(so), str(out)])
    lib = C.CDLL(str(so))
    lib.tiny_get_coder.restype = C.POINTER(pkg.capi.aws_huffman_symbol_coder)
    coder = lib.tiny_get_coder().contents
    got = {s: (coder.encode(s, None).pattern, coder.encode(s, None).num_bits) for s in (0, 9, 65, 66, 67)}
    assert got == {0: (0, 1), 9: (0, 0), 65: (2, 2), 66: (6, 3), 67: (0, 0)}
    sym = C.c_uint8(0)
    assert coder.decode(0xC0000000, C.byref(sym), None) == 3 and sym.value == 66
    assert coder.decode(0xE0000000, C.byref(sym), None) == 0  # 111... is a hole


@pytest.mark.parametrize("bad,why", [
    ('HUFFMAN_CODE(0, "0", 0x0, 1)\nHUFFMAN_CODE(1, "01", 0x1, 2)\n', "not a prefix code"),
    ('HUFFMAN_CODE(0, "0", 0x0, 1)\nHUFFMAN_CODE(0, "1", 0x1, 1)\n', "defined twice"),
    ('HUFFMAN_CODE(300, "0", 0x0, 1)\n', "bad symbol"),
    ('HUFFMAN_CODE(1, "0", 0x0, 33)\n', "bad length"),
    ('HUFFMAN_CODE(1, "111", 0x7, 2)\n', "does not fit"),
    ('/* nothing here */\n', "no HUFFMAN_CODE entries"),
])
def test_rejects_broken_tables(gen, tmp_path, bad, why):
    src = tmp_path / "bad.def"
    src.write_text(bad)
    r = run(gen, str(src), str(tmp_path / "bad.c"), "bad")
    assert r.returncode == 1 and why in r.stderr


@pytest.mark.parametrize("table", ["test", "hpack"])
def test_emitted_decode_equals_tree_walk_on_many_windows(coders, oracle, oracle_tables, table):
    """The emitted table-driven decode_symbol == the bit-at-a-time tree walk the reference generator
    emits (restated in oracle_decode_symbol), on all 2^17 17-bit prefixes x two fills plus every code
    followed by random bits."""
    coder = coders.coder(table).contents
    otable = oracle_tables[table]
    rng = np.random.default_rng(3)
    patterns, num_bits = refcodec.table_arrays(table)
    windows = [(p << 15) | fill for p in range(0, 1 << 17, 7) for fill in (0, 0x7FFF)]
    for s in range(256):
        n = int(num_bits[s])
        for _ in range(8):
            tail = int(rng.integers(0, 1 << (32 - n))) if n < 32 else 0
            windows.append(((int(patterns[s]) << (32 - n)) | tail) & 0xFFFFFFFF)
    sym = C.c_uint8(0)
    for w in windows:
        used = coder.decode(w, C.byref(sym), None)
        want_used, want_sym = oracle.decode_symbol(otable, w)
        assert used == want_used and (used == 0 or sym.value == want_sym), hex(w)


def test_reference_def_file_in_place(gen, pkg, tmp_path):
    """When the reference tree is on this machine, our generator eats its .def file directly."""
    path = "/root/reference/tests/test_huffman_static_table.def"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    out = tmp_path / "ref_table.c"
    assert run(gen, path, str(out), "reftable").returncode == 0
    so = tmp_path / "libreftable.so"
    inc = [a for d in pkg._build.INCLUDES for a in ("-I", d)]
    subprocess.check_call(["gcc", "-std=gnu99", "-shared", "-fPIC", *inc, "-o", str(so), str(out)])
    lib = C.CDLL(str(so))
    lib.reftable_get_code_table.restype = C.POINTER(pkg.capi.aws_huffman_code * 256)
    table = lib.reftable_get_code_table().contents
    patterns, num_bits = refcodec.table_arrays("test")
    assert [c.pattern for c in table] == list(patterns) and [c.num_bits for c in table] == list(num_bits)
