#!/usr/bin/env python3
"""Cycles per phase of decode_batch_kernel (a -DHB_PHASE_TIMING build named by AWS_HUFFMAN_B200_LIB)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench, refcodec
import __graft_entry__ as graft
pkg = graft.load_package()
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sampler_t = torch.from_numpy(refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])).to(dev)
ctx = pkg.BatchContext(pkg.coders_library().coder("hpack"), eos_padding=0xFF, device=0)
lens = bench.string_lengths_torch(bench.SEED_BATCH, 0, n, dev)
in_off = torch.zeros(n + 1, dtype=torch.int64, device=dev); in_off[1:] = torch.cumsum(lens, 0)
raw_bytes = int(in_off[-1].item())
raw = bench.symbols_torch(bench.SEED_BATCH, 0, raw_bytes, sampler_t, dev)
enc = torch.empty(raw_bytes * 3 // 2 + 1024, dtype=torch.uint8, device=dev); enc_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
dec = torch.empty(raw_bytes + 1024, dtype=torch.uint8, device=dev); dec_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev)
ctx.encode_device(n, {"in_": raw, "in_offsets": in_off, "out": enc, "out_offsets": enc_off, "status": st}, raw_bytes, enc.numel())
torch.cuda.synchronize()
eb = int(enc_off[-1].item())
lib = ctypes.CDLL(os.environ["AWS_HUFFMAN_B200_LIB"])
buf = (ctypes.c_ulonglong * 16)()
def run():
    ctx.decode_device(n, {"in_": enc, "in_offsets": enc_off, "out": dec, "out_offsets": dec_off, "status": st}, eb, dec.numel())
    torch.cuda.synchronize()
run(); lib.aws_huffman_batch_debug_phase_cycles(buf, 1)
reps = 5
for _ in range(reps): run()
lib.aws_huffman_batch_debug_phase_cycles(buf, 1)
if not os.environ.get("AWS_HUFFMAN_BATCH_EXPERIMENT"): assert torch.equal(dec[:raw_bytes], raw)
names = ["prep (ticket, table, stage, sort)", "decode", "scan+publish", "rows->image", "wait for the position", "image -> global", ""]
tot = sum(buf[i] for i in range(8))
print("scout: %.0f cycles per resolve, %.1f retries and %.1f rounds per resolve" % (buf[8] / max(1, buf[10]) * (buf[10] / max(1.0, reps * n / 224.0)), buf[9] / (reps * n / 224.0), buf[10] / (reps * n / 224.0)))
print("scout resolve durations: <4K %d  <16K %d  <64K %d  <256K %d  more %d   max %d cycles" % (buf[11], buf[12], buf[13], buf[14], buf[15], buf[7]))
print("fast loop: %.1f cycles per round of 6 steps (longest lane of each warp)" % (buf[14] / max(1, buf[15])))
print("slow resolves by hand-off number: h=0 %d  h=1 %d  later %d" % (buf[5], buf[6], buf[13]))
for i in range(6):
    print("phase %d %-55s %6.1f %%   %8.0f cycles/block/call" % (i, names[i], 100.0 * buf[i] / max(tot, 1), buf[i] / reps / 296.0))

tt = (ctypes.c_ulonglong * (4 * 8192))()
lib.aws_huffman_batch_debug_tile_times(tt)
T = np.frombuffer(tt, dtype=np.uint64).reshape(4, 8192).astype(np.int64)
nt = (n + 223) // 224
T = T[:, :nt]
t0 = T[0].min()
tk, pb, rs, re = [(x - t0) / 1000.0 for x in T]
print("tiles %d; kernel span %.1f us" % (nt, (re.max() - 0)))
late = pb - np.maximum.accumulate(pb)  # <0: published before some lower tile... use running max of LOWER tiles
runmax = np.concatenate([[0], np.maximum.accumulate(pb)[:-1]])
dur = re - rs
slow = np.argsort(-dur)[:12]
for t in sorted(slow):
    w = max(0, t - 128)
    j = w + int(np.argmax(pb[w:t])) if t > 0 else 0
    print("tile %4d: ticket %7.1f publish %7.1f resolve %7.1f..%7.1f (%.1f us)  latest predecessor in window: tile %d ticket %.1f publish %.1f" % (
        t, tk[t], pb[t], rs[t], re[t], dur[t], j, tk[j], pb[j]))
print("decode+prep time per tile (ticket->publish): mean %.1f  p50 %.1f  p99 %.1f  max %.1f us" % (
    (pb - tk).mean(), np.percentile(pb - tk, 50), np.percentile(pb - tk, 99), (pb - tk).max()))
