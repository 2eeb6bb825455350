"""Debug helper: encode a small batch through the C ABI and report where it differs from the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft
import refcodec
pkg = graft.load_package(); pkg._build.build(); graft.build_oracle()
oracle = refcodec.OracleLib()
coders = pkg.coders_library()
name = os.environ.get("DBG_TABLE", "hpack")
table = oracle.table(*refcodec.table_arrays(name))
ctx = pkg.BatchContext(coders.coder(name), eos_padding=0xFF, device=0)
sampler = refcodec.zipf_symbol_sampler(refcodec.table_arrays(name)[1])
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
nitems = int(sys.argv[1]) if len(sys.argv) > 1 else 20
maxlen = int(sys.argv[2]) if len(sys.argv) > 2 else 256
if os.environ.get("DBG_RANDOM_BATCH"):
    data, offs = refcodec.random_batch(np.random.default_rng(0xB200 + 1), nitems, 0, maxlen, name, zipf=True)
    lens = np.diff(offs.astype(np.int64))
else:
    lens = rng.integers(24, maxlen + 1, size=nitems)
    offs = np.zeros(nitems + 1, dtype=np.uint64); offs[1:] = np.cumsum(lens)
    data = np.ascontiguousarray(sampler[rng.integers(0, 65536, size=int(offs[-1]))])
cap = 4 * len(data) + 64
want = oracle.encode_batch(table, 0xFF, data, offs, cap)
got = ctx.encode(data, offs, cap)
tot = int(want["out_offsets"][-1])
print("items", nitems, "raw", len(data), "enc", tot, "offsets equal:", np.array_equal(got["out_offsets"], want["out_offsets"]))
bad = np.flatnonzero(got["out"][:tot] != want["out"][:tot])
print("mismatching bytes:", len(bad), "first", bad[:10], "last", bad[-5:] if len(bad) else None)
if len(bad):
    b = int(bad[0])
    it = int(np.searchsorted(want["out_offsets"], b, side="right") - 1)
    print("first bad byte", b, "in item", it, "item out range", int(want["out_offsets"][it]), int(want["out_offsets"][it + 1]),
          "in range", int(offs[it]), int(offs[it + 1]))
    print("got ", got["out"][max(0, b - 4):b + 12]); print("want", want["out"][max(0, b - 4):b + 12])
    # run lengths of bad regions
    runs = np.split(bad, np.flatnonzero(np.diff(bad) > 1) + 1)
    print("bad runs (start,len):", [(int(r[0]), len(r)) for r in runs[:12]])
    for r in runs[:6]:
        b = int(r[0]); it = int(np.searchsorted(want["out_offsets"], b, side="right") - 1)
        slot_base = np.concatenate([[0], np.cumsum((lens + 31) // 32)])
        print("  bad byte", b, "item", it, "len", int(lens[it]), "byte-in-item", b - int(want["out_offsets"][it]), "of", int(want["out_offsets"][it+1]-want["out_offsets"][it]),
              "slots", int(slot_base[it]), "..", int(slot_base[it+1]), "tile", int(slot_base[it]) // 256, "slot-in-tile", int(slot_base[it]) % 256,
              "neighbour lens", lens[max(0,it-2):it+3].tolist(), "got", got["out"][b], "want", want["out"][b])
