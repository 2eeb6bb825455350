#!/usr/bin/env python3
"""Writes tests/golden/*.json from the reference tree (run HERE, where /root/reference exists).

  reference_vectors.json  the golden vectors the reference's own tests hold for the path:
                          the 256 rows of tests/test_huffman_static_table.def and the literal
                          strings / byte arrays of tests/huffman_test.c:20-37,175-194,408.
  differential_*.json     outputs of the UNMODIFIED reference (oracle/_ref/libref_huffman.so, built
                          in place by oracle/Makefile) on seeded random inputs, including the
                          short-buffer and unknown-symbol paths, so the GPU box (which has no
                          /root/reference) can still check against the reference itself.
"""
import ctypes
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFERENCE_DIR", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))


def parse_def(path):
    patterns, lens = [0] * 256, [0] * 256
    rx = re.compile(r'HUFFMAN_CODE\(\s*(\d+)\s*,\s*"([01]+)"\s*,\s*(0x[0-9a-fA-F]+)\s*,\s*(\d+)\s*\)')
    for m in rx.finditer(open(path).read()):
        sym, bits, code, n = int(m.group(1)), m.group(2), int(m.group(3), 16), int(m.group(4))
        assert len(bits) == n and int(bits, 2) == code
        patterns[sym], lens[sym] = code, n
    return patterns, lens


def c_string_literals(src, name):
    """Concatenated C string literal assigned to `name[]`."""
    m = re.search(r'%s\[\]\s*=\s*((?:\s*"(?:[^"\\]|\\.)*")+)\s*;' % re.escape(name), src)
    parts = re.findall(r'"((?:[^"\\]|\\.)*)"', m.group(1))
    return bytes("".join(parts), "latin-1").decode("unicode_escape").encode("latin-1")


def c_byte_array(src, name):
    m = re.search(r'%s\[\]\s*=\s*\{([^}]*)\}' % re.escape(name), src)
    return bytes(int(x, 16) for x in re.findall(r'0x[0-9a-fA-F]+', m.group(1)))


def main():
    os.makedirs(GOLD, exist_ok=True)
    patterns, lens = parse_def(os.path.join(REF, "tests", "test_huffman_static_table.def"))
    src = open(os.path.join(REF, "tests", "huffman_test.c")).read()
    url = c_string_literals(src, "s_url_string")
    allc = c_string_literals(src, "s_all_codes")
    vectors = {
        "source": "awslabs/aws-c-compression tests/test_huffman_static_table.def + tests/huffman_test.c",
        "test_table": {"patterns": patterns, "num_bits": lens},
        "encode_kats": [
            {"cite": "huffman_test.c:20-24", "input_hex": url.hex(), "encoded_hex": c_byte_array(src, "s_encoded_url").hex()},
            {"cite": "huffman_test.c:26-37", "input_hex": allc.hex(), "encoded_hex": c_byte_array(src, "s_encoded_codes").hex()},
            {"cite": "huffman_test.c:179-183", "input_hex": b"?".hex(), "encoded_hex": "ba"},
            {"cite": "huffman_test.c:189-194", "input_hex": b"yz".hex(), "encoded_hex": "a379"},
        ],
        "encoded_length_kats": [{"cite": "huffman_test.c:408", "input_hex": b"cdfh".hex(), "encoded_len": 3}],
        "step_sizes": [1, 2, 4, 8, 16, 32, 64, 128],
    }
    assert url == b"www.example.com" and len(allc) == 95
    with open(os.path.join(GOLD, "reference_vectors.json"), "w") as f:
        json.dump(vectors, f, indent=1)
    print("wrote reference_vectors.json")

    # ---- differential vectors from the unmodified reference ----
    import refcodec  # tests/refcodec.py: ctypes driver shared with the test-suite

    ref = refcodec.RefLib(os.path.join(ROOT, "oracle", "_ref", "libref_huffman.so"))
    for table in ("test", "hpack"):
        cases = refcodec.make_differential_cases(ref, table, seed=0xD1FF0000 + (1 if table == "hpack" else 0))
        with open(os.path.join(GOLD, "differential_%s.json" % table), "w") as f:
            json.dump(cases, f, indent=0, separators=(",", ":"))
        print("wrote differential_%s.json:" % table, len(cases["encode"]), "encode +", len(cases["decode"]), "decode cases")


if __name__ == "__main__":
    main()
