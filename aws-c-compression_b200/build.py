"""Builds the native artefacts in-tree (so they travel to the GPU box with the snapshot):

    lib/libaws-c-compression.so   the product: host streaming codec + batched CUDA codec (sm_100a)
    lib/huffman_generator         the .def -> C coder generator
    lib/libhuffman_coders.so      coders emitted by OUR generator (hpack + the reference's test
                                  table restored from tests/golden/reference_vectors.json); used by
                                  tests, smoke() and bench.py, not part of the product library

nvcc cross-compiles for sm_100a without a GPU. Nothing here touches oracle/.
"""
import json
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "lib")
OBJ = os.path.join(PKG, "build")
INCLUDES = [os.path.join(ROOT, "include"), os.path.join(ROOT, "shim", "aws-c-common", "include")]

PRODUCT_LIB = os.path.join(LIB, "libaws-c-compression.so")
GENERATOR = os.path.join(LIB, "huffman_generator")
CODERS_LIB = os.path.join(LIB, "libhuffman_coders.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CC = os.environ.get("CC", "gcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
] + (["-D" + d for d in os.environ.get("HB_EXTRA_DEFINES", "").split()] if os.environ.get("HB_EXTRA_DEFINES") else [])
C_FLAGS = ["-std=gnu99", "-O2", "-DNDEBUG", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter"]

C_SOURCES = [
    os.path.join(PKG, "host", "huffman.c"),
    os.path.join(PKG, "host", "compression.c"),
    os.path.join(PKG, "host", "huffman_testing.c"),
    os.path.join(PKG, "host", "huffman_lut.c"),
    os.path.join(PKG, "host", "huffman_table_builder.c"),
    os.path.join(ROOT, "shim", "aws-c-common", "source", "common_shim.c"),
]
CU_SOURCES = [os.path.join(PKG, "csrc", "huffman_batch.cu")]


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("build step failed: %s\n%s" % (" ".join(cmd), proc.stdout))
    if verbose and proc.stdout.strip():
        print(proc.stdout)
    return proc.stdout


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _all_deps():
    deps = []
    for base in (os.path.join(PKG, "csrc"), os.path.join(PKG, "host"), os.path.join(ROOT, "include"),
                 os.path.join(ROOT, "shim")):
        for d, _, files in os.walk(base):
            deps += [os.path.join(d, f) for f in files]
    return deps


def write_def(path, patterns, num_bits):
    """Writes a code table in the .def grammar (HUFFMAN_CODE(sym, "bits", 0xcode, len))."""
    with open(path, "w") as f:
        f.write("#ifndef HUFFMAN_CODE\n#error \"define HUFFMAN_CODE first\"\n#endif\n")
        for sym in range(256):
            n = num_bits[sym]
            if n:
                f.write('HUFFMAN_CODE(%d, "%s", 0x%x, %d)\n' % (sym, format(patterns[sym], "0%db" % n), patterns[sym], n))


def build(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    inc = [a for d in INCLUDES for a in ("-I", d)]

    # 1. generator
    gen_srcs = [os.path.join(PKG, "generator", "huffman_generator.c"), os.path.join(PKG, "host", "huffman_lut.c")]
    if force or _newer(GENERATOR, gen_srcs + [os.path.join(PKG, "host", "huffman_lut.h")]):
        _run([CC, "-std=gnu99", "-O2", "-Wall", "-o", GENERATOR] + gen_srcs, verbose)

    # 2. product library
    if force or _newer(PRODUCT_LIB, _all_deps()):
        if not (os.path.exists(NVCC) or shutil.which(NVCC)):
            raise RuntimeError("nvcc not found (%s): the batched codec has no CPU fallback" % NVCC)
        objs = []
        for src in C_SOURCES:
            obj = os.path.join(OBJ, os.path.basename(src) + ".o")
            _run([CC] + C_FLAGS + inc + ["-c", src, "-o", obj], verbose)
            objs.append(obj)
        for src in CU_SOURCES:
            obj = os.path.join(OBJ, os.path.basename(src) + ".o")
            _run([NVCC] + NVCC_FLAGS + inc + ["-Xptxas", "-v", "-c", src, "-o", obj], verbose)
            objs.append(obj)
        # (linked next to its final place and renamed: a snapshot of the tree never sees half a library)
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-Bsymbolic",
              "-o", PRODUCT_LIB + ".tmp"] + objs, verbose)
        os.replace(PRODUCT_LIB + ".tmp", PRODUCT_LIB)

    # 3. coders emitted by our generator
    hpack_def = os.path.join(PKG, "tables", "hpack.def")
    golden = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")
    if force or _newer(CODERS_LIB, [GENERATOR, hpack_def, golden]):
        gen_dir = os.path.join(OBJ, "gen")
        os.makedirs(gen_dir, exist_ok=True)
        coder_srcs = []
        _run([GENERATOR, hpack_def, os.path.join(gen_dir, "hpack_coder.c"), "hpack"], verbose)
        coder_srcs.append(os.path.join(gen_dir, "hpack_coder.c"))
        if os.path.exists(golden):
            table = json.load(open(golden))["test_table"]
            test_def = os.path.join(gen_dir, "test_table.def")
            write_def(test_def, table["patterns"], table["num_bits"])
            _run([GENERATOR, test_def, os.path.join(gen_dir, "test_coder.c"), "test"], verbose)
            coder_srcs.append(os.path.join(gen_dir, "test_coder.c"))
        _run([CC] + C_FLAGS + inc + ["-shared", "-o", CODERS_LIB] + coder_srcs, verbose)
    return PRODUCT_LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", PRODUCT_LIB)
