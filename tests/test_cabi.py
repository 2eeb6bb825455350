"""The C-ABI library loads on a machine without a GPU, exports every symbol the headers declare, keeps
the reference's struct layouts, and fails loudly (never falls back) when no device is usable."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    inc = os.path.join(ROOT, "include", "aws", "compression")
    for d, _, files in os.walk(inc):
        for f in files:
            text = open(os.path.join(d, f)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            for m in re.finditer(r"AWS_COMPRESSION_API\s+[\w\s\*]+?\b(\w+)\s*\(", text):
                names.add(m.group(1))
    return names


def test_every_declared_symbol_is_exported(pkg, product):
    declared = declared_functions()
    assert len(declared) >= 25
    assert declared == set(pkg.capi.EXPORTED_SYMBOLS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", product.path], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert not (declared - exported), "missing exports: %s" % sorted(declared - exported)


def test_struct_layouts_match_the_reference_abi(pkg, tmp_path):
    # aws_huffman_code 8 B {0,4}; symbol_coder 24 B; encoder 24 B {0,8,12}; decoder 32 B {0,8,16,24}
    src = tmp_path / "layout.c"
    src.write_text('#include <aws/compression/huffman_batch.h>\n#include <stdio.h>\n#include <stddef.h>\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %d\\n",'
                   'sizeof(struct aws_huffman_code),offsetof(struct aws_huffman_code,num_bits),'
                   'sizeof(struct aws_huffman_symbol_coder),sizeof(struct aws_huffman_encoder),'
                   'offsetof(struct aws_huffman_encoder,eos_padding),offsetof(struct aws_huffman_encoder,overflow_bits),'
                   'sizeof(struct aws_huffman_decoder),offsetof(struct aws_huffman_decoder,allow_growth),'
                   'offsetof(struct aws_huffman_decoder,working_bits),offsetof(struct aws_huffman_decoder,num_bits),'
                   'sizeof(struct aws_huffman_batch),offsetof(struct aws_huffman_batch,leftover_num_bits),'
                   '(int)AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL);return 0;}\n')
    exe = tmp_path / "layout"
    inc = [a for d in pkg._build.INCLUDES for a in ("-I", d)]
    subprocess.check_call(["gcc", "-std=gnu99", *inc, "-o", str(exe), str(src)])
    got = subprocess.check_output([str(exe)], text=True).split()
    assert got == ["8", "4", "24", "24", "8", "12", "32", "8", "16", "24", str(C.sizeof(pkg.capi.aws_huffman_batch)),
                   str(pkg.capi.aws_huffman_batch.leftover_num_bits.offset), "3072"]


def test_headers_compile_as_c_and_cxx(pkg, tmp_path):
    inc = [a for d in pkg._build.INCLUDES for a in ("-I", d)]
    for compiler, ext, std in (("gcc", "c", "-std=c99"), ("g++", "cpp", "-std=c++14")):
        src = tmp_path / ("inc." + ext)
        src.write_text("#include <aws/compression/huffman_batch.h>\n#include <aws/compression/private/huffman_testing.h>\n"
                       "int main(void){struct aws_huffman_batch b; (void)b; return 0;}\n")
        subprocess.check_call([compiler, std, "-Wall", "-Werror", *inc, "-c", "-o", str(tmp_path / "o.o"), str(src)])


def test_no_device_means_device_failure_not_a_fallback(pkg, coders, product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.CodecError) as err:
        pkg.BatchContext(coders.coder("hpack"), device=0)
    assert err.value.code == pkg.AWS_ERROR_COMPRESSION_DEVICE_FAILURE


def test_invalid_code_tables_are_rejected_before_touching_the_device(pkg):
    capi = pkg.capi
    clash = capi.python_coder(lambda s: (0, 1) if s < 2 else (0, 0))       # two symbols share code '0'
    prefix = capi.python_coder(lambda s: {0: (0, 1), 1: (1, 2)}.get(s, (0, 0)))  # '0' prefixes '01'
    toolong = capi.python_coder(lambda s: (1, 33) if s == 0 else (0, 0))
    for coder in (clash, prefix, toolong):
        with pytest.raises(pkg.CodecError) as err:
            pkg.BatchContext(coder, device=0)
        assert err.value.code == pkg.AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE


def test_context_from_a_raw_code_table_validates_like_the_callback_route(pkg, coders):
    """aws_huffman_batch_ctx_new_from_code_table: same checks, no callbacks."""
    import ctypes as C
    import torch
    capi = pkg.capi
    bad = (capi.aws_huffman_code * 256)()
    bad[0].pattern, bad[0].num_bits = 0, 1
    bad[1].pattern, bad[1].num_bits = 0, 1  # two symbols share code '0'
    with pytest.raises(pkg.CodecError) as err:
        pkg.BatchContext(None, device=0, code_table=C.pointer(bad))
    assert err.value.code == pkg.AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE
    if not torch.cuda.is_available():
        # a valid table reaches the device step and fails there (no CPU fallback)
        with pytest.raises(pkg.CodecError) as err:
            pkg.BatchContext(None, device=0, code_table=coders.code_table_pointer("hpack"))
        assert err.value.code == pkg.AWS_ERROR_COMPRESSION_DEVICE_FAILURE


def test_product_never_references_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg_dir = os.path.join(ROOT, "aws-c-compression_b200")
    for d, _, files in os.walk(pkg_dir):
        if os.path.basename(d) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".so", ".o", ".pyc")):
                continue
            text = open(os.path.join(d, f), errors="ignore").read()
            assert "oracle" not in text.lower().replace("nothing here touches oracle/", ""), os.path.join(d, f)
    for d, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            assert "oracle" not in open(os.path.join(d, f)).read().lower()


def test_shard_planning_and_offset_concatenation(product):
    rng = np.random.default_rng(9)
    lens = rng.integers(8, 257, size=10007)
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    for shards in (1, 2, 3, 4, 8):
        begin = product.plan_shards(offs, shards)
        assert begin[0] == 0 and begin[-1] == len(lens) and (np.diff(begin) >= 0).all()
        per = np.array([int(offs[begin[s + 1]] - offs[begin[s]]) for s in range(shards)])
        assert per.max() - per.min() <= 2 * 256, "shards are balanced by bytes"
        local = []
        for s in range(shards):
            chunk = lens[begin[s]:begin[s + 1]]
            lo = np.zeros(len(chunk) + 1, dtype=np.uint64)
            lo[1:] = np.cumsum(chunk)
            local.append(lo)
        assert np.array_equal(product.concat_offsets(local), offs)
    # degenerate shapes
    assert list(product.plan_shards(np.zeros(1, dtype=np.uint64), 4)) == [0, 0, 0, 0, 0]
    assert list(product.plan_shards(np.zeros(9, dtype=np.uint64), 4)) == [0, 2, 4, 6, 8]
    one_big = np.array([0, 1 << 30], dtype=np.uint64)
    assert list(product.plan_shards(one_big, 4))[-1] == 1
