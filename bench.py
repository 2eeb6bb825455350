#!/usr/bin/env python3
"""Headline benchmark: Huffman encode + decode GB/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|hpack_batch|stream] [--impl reference]

One "step" = one encode pass + one decode pass of the workload through the batched C ABI.
  value      device-resident: inputs already in HBM, kernels timed with CUDA events on the launching
             stream; algorithmic bytes (payload in + payload out, each direction) / time.
  e2e        the same step through the host-pointer entry points (aws_huffman_encode_batch /
             aws_huffman_decode_batch) from pinned host buffers, H2D and D2H copies inside the timing;
             `copies_alone` = the same copies with no kernels (what the host side of the box allows).
  roofline   the dominant kernel family (encode or decode, whichever takes longer) against the measured
             HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's CPU path (oracle/_ref if it was built, else the oracle port) on this
             host, 1 thread, on a stated sample of the same workload.
  parity_checked  before the line is printed the reference (oracle/_ref) encodes the SAME inputs on the host:
             every encoded byte and offset of the GPU must equal its output (all strings of the batch; the
             whole stream, bit offsets beyond 2^32 included), and it decodes the GPU's bytes back to the input.
             A difference prints an error object and exits 1: no number without parity.

Workloads; the headline (top-level keys) is hpack_batch:
  hpack_batch    BASELINE configs[1]: HPACK table, 1,000,000 strings of 8..256 B per GPU, Zipf(1.5) bytes.
                 Under torchrun every rank runs its own such batch: weak scaling, no collective; time is the max
                 over ranks, value the sum of bytes / that time.
  stream         BASELINE configs[2]+[3]: one 2^30-byte Zipf(1.5) stream, encode then decode. One GPU only
                 (the single stream does not shard): the `stream` object of the 1-GPU line.
  sharded_batch  BASELINE configs[4]: ONE batch of 64 x 2^20 strings cut by aws_huffman_batch_plan_shards into
                 world_size byte-balanced shards, rank r runs shard r, offsets rebuilt by
                 aws_huffman_batch_concat_offsets and verified; whole-box GB/s, strong scaling: the
                 `sharded_batch` object of every line (N = 1 included).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MASK64 = (1 << 64) - 1
SEED_BATCH = 0x5EED0002
SEED_STREAM = 0x5EED0003
SEED_SHARDED = 0x5EED0005


# ------------------------------------------------------------------------------------------------
# Counter-based synthetic data: splitmix64(seed ^ index), identical on numpy (host) and torch (device)
# ------------------------------------------------------------------------------------------------
def _splitmix64_np(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _i64(v):
    v &= MASK64
    return v - (1 << 64) if v >= (1 << 63) else v


def _splitmix64_torch(x):
    import torch  # int64 arithmetic wraps; logical shifts are emulated with masks

    def lsr(v, k):
        return (v >> k) & ((1 << (64 - k)) - 1)

    x = x + _i64(0x9E3779B97F4A7C15)
    x = (x ^ lsr(x, 30)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ lsr(x, 27)) * _i64(0x94D049BB133111EB)
    return x ^ lsr(x, 31)


def string_lengths_np(seed, first, count):
    with np.errstate(over="ignore"):
        r = _splitmix64_np(np.uint64(seed) ^ np.arange(first, first + count, dtype=np.uint64))
    return (np.uint64(8) + r % np.uint64(249)).astype(np.int64)


def symbols_np(seed, first, count, sampler):
    with np.errstate(over="ignore"):
        r = _splitmix64_np(np.uint64(seed ^ 0xA5A5A5A5) ^ np.arange(first, first + count, dtype=np.uint64))
    return sampler[(r & np.uint64(0xFFFF)).astype(np.int64)]


def string_lengths_torch(seed, first, count, device):
    import torch
    idx = torch.arange(first, first + count, dtype=torch.int64, device=device)
    r = _splitmix64_torch(idx ^ _i64(seed))
    # r mod 249 on the unsigned value: split into high/low halves to stay in int64
    hi = (r >> 32) & 0xFFFFFFFF
    lo = r & 0xFFFFFFFF
    return 8 + ((hi % 249) * ((1 << 32) % 249) + lo % 249) % 249


def symbols_torch(seed, first, count, sampler_t, device, chunk=1 << 27):
    import torch
    out = torch.empty(count, dtype=torch.uint8, device=device)
    for a in range(0, count, chunk):
        b = min(count, a + chunk)
        idx = torch.arange(first + a, first + b, dtype=torch.int64, device=device)
        r = _splitmix64_torch(idx ^ _i64(seed ^ 0xA5A5A5A5))
        out[a:b] = sampler_t[r & 0xFFFF]
    return out


# ------------------------------------------------------------------------------------------------
def clocks_sampler(path, device_index):
    cmd = ["nvidia-smi", "-i", str(device_index),
           "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
           "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
           "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
           "--format=csv,noheader,nounits", "-lms", "100"]
    try:
        return subprocess.Popen(cmd, stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except OSError:
        return None


def summarize_clocks(path):
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, flag in zip(names, f[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
    except OSError:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
            "samples": len(sm)}


def measured_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref) or the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_codec():
    import refcodec
    patterns, num_bits = refcodec.table_arrays("hpack")
    if refcodec.RefLib.available():
        ref = refcodec.RefLib()
        coder = ref.coder("hpack")
        return ("reference", lambda d, o, cap: ref.encode_batch(coder, 0xFF, d, o, cap),
                lambda d, o, cap: ref.decode_batch(coder, d, o, cap))
    oracle = refcodec.OracleLib()
    table = oracle.table(patterns, num_bits)
    return ("port", lambda d, o, cap: oracle.encode_batch(table, 0xFF, d, o, cap),
            lambda d, o, cap: oracle.decode_batch(table, d, o, cap))


def cpu_sample(workload, strings):
    """The host copy of (a prefix of) the workload."""
    import refcodec
    sampler = refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])
    if workload == "hpack_batch":
        lens = string_lengths_np(SEED_BATCH, 0, strings)
        offs = np.zeros(strings + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(lens)
        data = symbols_np(SEED_BATCH, 0, int(offs[-1]), sampler)
        return data, offs, "first %d strings of the batch (%d B raw)" % (strings, len(data))
    nbytes = strings
    data = symbols_np(SEED_STREAM, 0, nbytes, sampler)
    return data, np.array([0, nbytes], dtype=np.uint64), "first %d B of the stream" % nbytes


def run_cpu(workload, sample_units, threads, repeats=1):
    """Returns (GB/s over encode+decode, seconds, description). threads > 1 splits the items over
    Python threads (ctypes releases the GIL); a single stream cannot be split."""
    kind, enc_fn, dec_fn = cpu_codec()
    data, offs, what = cpu_sample(workload, sample_units)
    n = len(offs) - 1
    threads = max(1, min(threads, n))
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)
    shards = []
    for t in range(threads):
        a, b = int(bounds[t]), int(bounds[t + 1])
        o = (offs[a:b + 1] - offs[a]).astype(np.uint64)
        shards.append((data[int(offs[a]):int(offs[b])], o))
    results = [None] * threads
    best = None
    for _ in range(repeats):
        def work(t):
            d, o = shards[t]
            enc = enc_fn(d, o, 4 * len(d) + 16)
            total = int(enc["out_offsets"][-1])
            dec = dec_fn(enc["out"][:total], enc["out_offsets"], len(d) + 16)
            results[t] = (len(d), total, int(dec["out_offsets"][-1]))
        t0 = time.perf_counter()
        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    raw = sum(r[0] for r in results)
    encd = sum(r[1] for r in results)
    assert sum(r[2] for r in results) == raw, "CPU round trip lost bytes"
    gbs = 2.0 * (raw + encd) / best / 1e9
    return kind, gbs, best, what, raw, encd


# ------------------------------------------------------------------------------------------------
# Parity gate ("encoded bytes identical to the CPU path on the same inputs before any number
# is reported"): the CPU arm above encodes the SAME inputs on the host; its bytes must equal the GPU's.
# ------------------------------------------------------------------------------------------------
def host_threads(world=1):
    return max(1, min(32, (os.cpu_count() or 1) // max(1, world)))


def parity_batch(h_raw, h_in_off, h_enc, h_enc_off, threads):
    """Every string: reference-encoded bytes and offsets == the GPU's; the reference decodes the GPU's bytes back
    to the input. Items are split over host threads (ctypes releases the GIL)."""
    kind, enc_fn, dec_fn = cpu_codec()
    n = len(h_in_off) - 1
    threads = max(1, min(threads, n))
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)
    res = [None] * threads

    def work(t):
        a, b = int(bounds[t]), int(bounds[t + 1])
        o = (h_in_off[a:b + 1] - h_in_off[a]).astype(np.uint64)
        d = h_raw[int(h_in_off[a]):int(h_in_off[b])]
        ref = enc_fn(d, o, 4 * len(d) + 16)
        tot = int(ref["out_offsets"][-1])
        g0, g1 = int(h_enc_off[a]), int(h_enc_off[b])
        g_off = (h_enc_off[a:b + 1] - h_enc_off[a]).astype(np.uint64)
        same_off = bool(np.array_equal(g_off, ref["out_offsets"]))
        same_bytes = (g1 - g0) == tot and bool(np.array_equal(h_enc[g0:g1], ref["out"][:tot]))
        back = dec_fn(h_enc[g0:g1], g_off, len(d) + 16)
        same_dec = (int(back["out_offsets"][-1]) == len(d) and bool(np.array_equal(back["out"][:len(d)], d))
                    and not back["status"].any() and not ref["status"].any())
        res[t] = (same_off, same_bytes, same_dec, b - a, len(d), tot)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return {"checker": "oracle/_ref (unmodified reference)" if kind == "reference" else "oracle port",
            "strings": int(sum(r[3] for r in res)), "raw_bytes": int(sum(r[4] for r in res)),
            "encoded_bytes": int(sum(r[5] for r in res)),
            "encoded_offsets_equal": all(r[0] for r in res), "encoded_bytes_equal": all(r[1] for r in res),
            "reference_decodes_gpu_bytes_to_input": all(r[2] for r in res),
            "host_threads": threads, "seconds": round(time.perf_counter() - t0, 2)}


def parity_stream(h_raw, h_enc, enc_bytes, threads, seg_bytes=8 << 20, decode_prefix=32 << 20):
    """The whole stream: the reference encodes it in segments (each a stream of its own, on its own host
    thread); the encoding of the whole is the bit-concatenation of the segments' encodings, so segment k must
    equal the GPU's bits [B_k, B_k + L_k) — B_k the running sum of code lengths — whatever B_k mod 8 is. The
    bit offsets pass 2^32 after 730 MB of input. The reference then decodes a prefix of the GPU's bytes."""
    import refcodec
    kind, enc_fn, dec_fn = cpu_codec()
    num_bits = refcodec.table_arrays("hpack")[1].astype(np.int64)
    size = len(h_raw)
    nseg = max(1, (size + seg_bytes - 1) // seg_bytes)
    seg_bits = [0] * nseg
    enc_seg = [None] * nseg
    t0 = time.perf_counter()

    def encode_seg(k):
        d = h_raw[k * seg_bytes:min(size, (k + 1) * seg_bytes)]
        seg_bits[k] = int(np.bincount(d, minlength=256).astype(np.int64) @ num_bits)
        r = enc_fn(d, np.array([0, len(d)], dtype=np.uint64), 4 * len(d) + 16)
        enc_seg[k] = r["out"][:int(r["out_offsets"][-1])]

    def pool(fn, count):
        nxt = [0]
        lock = threading.Lock()

        def loop():
            while True:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= count:
                    return
                fn(k)
        ths = [threading.Thread(target=loop) for _ in range(max(1, min(threads, count)))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    pool(encode_seg, nseg)
    start = np.concatenate([[0], np.cumsum(np.array(seg_bits, dtype=np.int64))])
    total_bits = int(start[-1])
    ok = [False] * nseg

    def compare_seg(k):
        B, L, e = int(start[k]), seg_bits[k], enc_seg[k]
        s, byte0, full, rem = B & 7, B >> 3, L >> 3, L & 7
        g = np.zeros(full + 2, dtype=np.uint8)
        got = h_enc[byte0:min(enc_bytes, byte0 + full + 2)]
        g[:len(got)] = got
        x = g[:-1] if s == 0 else ((g[:-1] << np.uint8(s)) | (g[1:] >> np.uint8(8 - s)))
        good = len(e) == full + (1 if rem else 0) and bool(np.array_equal(x[:full], e[:full]))
        if good and rem:
            mask = (0xFF << (8 - rem)) & 0xFF
            good = (int(x[full]) & mask) == (int(e[full]) & mask)
        ok[k] = good

    pool(compare_seg, nseg)
    length_ok = enc_bytes == (total_bits + 7) // 8
    pad = (8 - (total_bits & 7)) & 7
    padding_ok = pad == 0 or (int(h_enc[enc_bytes - 1]) & ((1 << pad) - 1)) == (1 << pad) - 1
    P = min(decode_prefix, enc_bytes)
    back = dec_fn(h_enc[:P], np.array([0, P], dtype=np.uint64), 2 * P + 64)
    m = int(back["out_offsets"][-1])
    dec_ok = m > 0 and bool(np.array_equal(back["out"][:m], h_raw[:m]))
    return {"checker": "oracle/_ref (unmodified reference)" if kind == "reference" else "oracle port",
            "raw_bytes": int(size), "encoded_bytes": int(enc_bytes), "encoded_bits": total_bits,
            "bit_offsets_beyond_2^32": total_bits > (1 << 32), "segments": nseg,
            "encoded_bytes_equal": all(ok) and length_ok and padding_ok,
            "reference_decodes_gpu_prefix_to_input": dec_ok, "decoded_prefix_symbols": m,
            "host_threads": threads, "seconds": round(time.perf_counter() - t0, 2)}


def host_codec_leg(total_bytes=64 << 20, piece=256 << 10):
    """The reference-compatible STREAMING API on the host (what a one-string-per-call caller links): this
    repository's host/huffman.c + its table-driven generated coder against the unmodified reference + the
    reference generator's goto-tree coder, one core each, same bytes, through the helper both libraries export
    with the same signature (huffman_test_transitive: encode, decode, compare; reference
    source/huffman_testing.c:15-80). Pieces of 256 KiB: the reference keeps its scratch on the stack."""
    import ctypes as C
    import refcodec
    import __graft_entry__ as graft
    pkg = graft.load_package()
    sampler = refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])
    data = symbols_np(SEED_STREAM, 0, total_bytes, sampler)
    pieces = [data[a:a + piece].tobytes() for a in range(0, total_bytes, piece)]
    out = {"bytes": total_bytes, "piece_bytes": piece, "cores": 1,
           "what": "huffman_test_transitive (encode + decode + compare) over the first %d B of the stream workload" % total_bytes}
    ours = pkg.product_library().lib
    legs = [("b200_host_library", ours, pkg.coders_library().coder("hpack"))]
    if refcodec.RefLib.available():
        ref = refcodec.RefLib()
        ref.lib.huffman_test_transitive.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_char_p)]
        ref.lib.huffman_test_transitive.restype = C.c_int
        legs.append(("reference", ref.lib, ref.coder("hpack")))
    for name, lib, coder in legs:
        msg = C.c_char_p()
        lib.huffman_test_transitive(coder, pieces[0], len(pieces[0]), 0, C.byref(msg))  # warm
        t0 = time.perf_counter()
        for p in pieces:
            if lib.huffman_test_transitive(coder, p, len(p), 0, C.byref(msg)) != 0:
                raise RuntimeError("%s round trip failed: %s" % (name, msg.value))
        dt = time.perf_counter() - t0
        out[name + "_raw_gbs"] = total_bytes / dt / 1e9
    if "reference_raw_gbs" in out:
        out["speedup"] = out["b200_host_library_raw_gbs"] / out["reference_raw_gbs"]
    return out



# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class Arm:
    def __init__(self, pkg, local_rank, rank, world, dist):
        import torch
        import refcodec
        self.torch, self.pkg, self.dist = torch, pkg, dist
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.device = torch.device("cuda", local_rank)
        self.sampler_t = torch.from_numpy(refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])).to(self.device)
        self.ctx = pkg.BatchContext(pkg.coders_library().coder("hpack"), eos_padding=0xFF, device=local_rank)
        # a real (non-default) stream: the C ABI reads a NULL stream as "the context's own stream", and
        # torch's default stream handle is 0. Kernels and timing events must sit on the same stream.
        self.stream = torch.cuda.Stream(self.device)
        torch.cuda.set_stream(self.stream)
        self.sptr = self.stream.cuda_stream
        assert self.sptr != 0
        self.check = not os.environ.get("AWS_HUFFMAN_BATCH_EXPERIMENT")

    def barrier(self):
        self.torch.cuda.synchronize(self.device)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t.cpu()]

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def device_pass(arm, n, in_off, raw, raw_bytes, steps, warmup):
    """Device-resident encode + decode of one packed batch (n == 1: one stream): a checked pass, `warmup`
    untimed steps, a barrier, `steps` timed steps with CUDA events on the launching stream."""
    torch, ctx = arm.torch, arm.ctx
    dev = arm.device
    enc_cap = raw_bytes + raw_bytes // 2 + 1024
    t = {"n": n, "raw_bytes": raw_bytes, "enc_cap": enc_cap, "in_off": in_off, "raw": raw}
    t["enc"] = torch.empty(enc_cap, dtype=torch.uint8, device=dev)
    t["enc_off"] = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    t["dec"] = torch.empty(raw_bytes + 1024, dtype=torch.uint8, device=dev)
    t["dec_off"] = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    t["enc_status"] = torch.zeros(n, dtype=torch.int32, device=dev)
    t["dec_status"] = torch.zeros(n, dtype=torch.int32, device=dev)
    known = [0]

    def encode_step():
        ctx.encode_device(n, {"in_": raw, "in_offsets": in_off, "out": t["enc"], "out_offsets": t["enc_off"],
                              "status": t["enc_status"]}, raw_bytes, enc_cap, stream=arm.sptr)

    def decode_step():
        ctx.decode_device(n, {"in_": t["enc"], "in_offsets": t["enc_off"], "out": t["dec"], "out_offsets": t["dec_off"],
                              "status": t["dec_status"]}, known[0], raw_bytes + 1024, stream=arm.sptr)

    encode_step()
    torch.cuda.synchronize(dev)
    enc_bytes = known[0] = int(t["enc_off"][-1].item())
    decode_step()
    torch.cuda.synchronize(dev)
    if arm.check:  # the round trip must be exact before anything is timed
        assert int(t["dec_off"][-1].item()) == raw_bytes and torch.equal(t["dec"][:raw_bytes], raw), "round trip mismatch"
        assert int(t["enc_status"].abs().sum().item()) == 0 and int(t["dec_status"].abs().sum().item()) == 0
    for _ in range(warmup):
        encode_step()
        decode_step()
    arm.barrier()
    launches_before = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    ev[0].record(arm.stream)
    for s in range(steps):
        encode_step()
        ev[2 * s + 1].record(arm.stream)
        decode_step()
        ev[2 * s + 2].record(arm.stream)
    torch.cuda.synchronize(dev)
    t["launches"] = ctx.launch_count - launches_before
    t["enc_ms"] = float(np.mean([ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(steps)]))
    t["dec_ms"] = float(np.mean([ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(steps)]))
    t["step_ms"] = ev[0].elapsed_time(ev[-1]) / steps
    t["enc_bytes"] = enc_bytes
    return t


def e2e_pass(arm, t, steps):
    """The same step through the host-pointer C ABI (aws_huffman_encode_batch / aws_huffman_decode_batch) on
    pinned host buffers: every H2D and D2H copy is inside the timed region. Also times the step's copies alone
    (both directions at once, no kernels): the ceiling the host side of the box allows this rank right now."""
    torch, ctx = arm.torch, arm.ctx
    n, raw_bytes, enc_cap, enc_bytes = t["n"], t["raw_bytes"], t["enc_cap"], t["enc_bytes"]

    def pinned(count, dtype):
        return torch.empty(count, dtype=dtype).pin_memory()

    p_raw = t["raw"].cpu().pin_memory()
    p_enc, p_dec = pinned(enc_cap, torch.uint8), pinned(raw_bytes + 1024, torch.uint8)
    h = {"raw": p_raw.numpy(), "in_off": t["in_off"].cpu().pin_memory().numpy().view(np.uint64),
         "enc": p_enc.numpy(), "enc_off": pinned(n + 1, torch.int64).numpy().view(np.uint64),
         "dec": p_dec.numpy(), "dec_off": pinned(n + 1, torch.int64).numpy().view(np.uint64),
         "status": pinned(n, torch.int32).numpy()}

    def step():
        ctx._call("aws_huffman_encode_batch", n, {"in_": h["raw"], "in_offsets": h["in_off"], "out": h["enc"],
                                                  "out_offsets": h["enc_off"], "status": h["status"]}, enc_cap)
        total = int(h["enc_off"][n])
        ctx._call("aws_huffman_decode_batch", n, {"in_": h["enc"], "in_offsets": h["enc_off"], "out": h["dec"],
                                                  "out_offsets": h["dec_off"], "status": h["status"]}, raw_bytes + 1024)
        return total, int(h["dec_off"][n])

    step()
    arm.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        total, back = step()
    torch.cuda.synchronize(arm.device)
    e2e_s = (time.perf_counter() - t0) / steps
    if arm.check:
        assert total == enc_bytes and back == raw_bytes
        assert np.array_equal(h["dec"][:raw_bytes], h["raw"]), "e2e round trip mismatch"
        assert not h["status"].any()

    # copies of one step, nothing else: up = raw + encoded, down = encoded + raw, on two streams at once
    up, down = torch.cuda.Stream(arm.device), torch.cuda.Stream(arm.device)
    arm.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        with torch.cuda.stream(up):
            t["raw"].copy_(p_raw, non_blocking=True)
            t["enc"][:enc_bytes].copy_(p_enc[:enc_bytes], non_blocking=True)
        with torch.cuda.stream(down):
            p_enc[:enc_bytes].copy_(t["enc"][:enc_bytes], non_blocking=True)
            p_dec[:raw_bytes].copy_(t["dec"][:raw_bytes], non_blocking=True)
    up.synchronize()
    down.synchronize()
    copy_s = (time.perf_counter() - t0) / steps
    return {"e2e_s": e2e_s, "copy_s": copy_s, "host": h,
            "h2d_bytes": int(raw_bytes + 8 * (n + 1) + enc_bytes + 8 * (n + 1)),
            "d2h_bytes": int(enc_bytes + raw_bytes + 2 * 8 * (n + 1) + 2 * 4 * n)}


KERNELS = {"hpack_batch": {"encode": "str_pack_kernel (+ str_prep / str_bits / str_scan launches)",
                           "decode": "decode_batch_kernel"},
           "stream": {"encode": "encode_tiled_kernel<false>",
                      "decode": "stream_fused_kernel (+ verify and gated fallback launches)"}}


def roofline(kind, enc_ms, dec_ms, raw_bytes, enc_bytes, traffic_key=None):
    """The dominant call (encode or decode, whichever takes longer) against the measured HBM copy rate:
    algorithmic bytes (payload in + payload out of that direction) / its CUDA-event time."""
    peak, peak_src = measured_peak()
    one_way = raw_bytes + enc_bytes
    enc_gbs, dec_gbs = one_way / (enc_ms * 1e-3) / 1e9, one_way / (dec_ms * 1e-3) / 1e9
    dominant = "decode" if dec_ms >= enc_ms else "encode"
    achieved = dec_gbs if dominant == "decode" else enc_gbs
    return {"bound": "hbm", "kernel": KERNELS[kind][dominant] + " (per GPU; CUDA events around the " + dominant + " call)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": profiled_traffic(traffic_key, dominant) if traffic_key else None,
            "algorithmic_bytes_per_launch": int(one_way), "peak_source": peak_src,
            "frac_of_8000_nominal": achieved / 8000.0, "encode_frac": enc_gbs / peak, "decode_frac": dec_gbs / peak}


def profiled_traffic(key, dominant):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of the dominant kernel from the
    committed ncu --set full capture of exactly this workload shape (profiles/r2_kernels.json, made by
    tools/ncu_summary.py) — None when this run's shape has no capture (nothing is scaled or guessed)."""
    try:
        group = json.load(open(os.path.join(ROOT, "profiles", "r2_kernels.json")))[key]
        want = {"hpack_batch": {"encode": "str_pack", "decode": "decode_batch"},
                "stream": {"encode": "encode_tiled", "decode": "stream_fused_kernel"}}[key.split(":")[0]][dominant]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        for k in group:
            if want in k["kernel"]:
                return int(k["dram_read"] * scale[k["dram_read_unit"]] + k["dram_write"] * scale[k["dram_write_unit"]])
    except Exception:
        pass
    return None


def run_sharded(arm, args, peak):
    """BASELINE configs[4]: ONE batch of `--sharded-strings` strings cut by aws_huffman_batch_plan_shards (balanced
    by bytes) into world_size contiguous shards; rank r encodes + decodes shard r on its GPU (no collective on the
    data path); the shard-local out_offsets are rebuilt into the batch's with aws_huffman_batch_concat_offsets and
    checked against the running sum of the per-item encoded lengths. Strong scaling: the batch is fixed."""
    torch, dist, pkg = arm.torch, arm.dist, arm.pkg
    dev, rank, world = arm.device, arm.rank, arm.world
    n_total = args.sharded_strings
    lens = string_lengths_torch(SEED_SHARDED, 0, n_total, dev)
    off_g = torch.zeros(n_total + 1, dtype=torch.int64, device=dev)
    off_g[1:] = torch.cumsum(lens, 0)
    del lens
    h_off_g = off_g.cpu().numpy().view(np.uint64)
    begin = pkg.product_library().plan_shards(h_off_g, world)  # the C ABI's host-side planner
    a, b = int(begin[rank]), int(begin[rank + 1])
    n = b - a
    byte0, raw_bytes = int(h_off_g[a]), int(h_off_g[b]) - int(h_off_g[a])
    in_off = (off_g[a:b + 1] - byte0).contiguous()
    del off_g
    raw = symbols_torch(SEED_SHARDED, byte0, raw_bytes, arm.sampler_t, dev)  # global byte index: any shard is reproducible
    steps = max(1, min(args.steps, 5))
    t = device_pass(arm, n, in_off, raw, raw_bytes, steps, min(args.warmup, 3))
    enc_bytes = t["enc_bytes"]

    # per-item encoded lengths from their own kernel; shard-local offsets must be their running sum
    item_len = torch.zeros(n, dtype=torch.int64, device=dev)
    arm.ctx.encoded_lengths_device(n, raw, in_off, item_len, stream=arm.sptr)
    torch.cuda.synchronize(dev)
    local_ok = bool(torch.equal(t["enc_off"][1:] - t["enc_off"][:-1], item_len)) and int(t["enc_off"][0].item()) == 0

    # a sample of every shard against the reference on the host: its first and its last strings (the last ones sit
    # beyond bit offset 2^32 of the shard's output when the shard is larger than 512 MiB encoded)
    k = min(n, args.sharded_parity_strings)
    sample = {"encoded_offsets_equal": True, "encoded_bytes_equal": True, "reference_decodes_gpu_bytes_to_input": True,
              "strings": 0}
    if arm.check and k:
        for lo in sorted({0, n - k}):
            io = t["in_off"][lo:lo + k + 1].cpu().numpy().view(np.uint64)
            eo = t["enc_off"][lo:lo + k + 1].cpu().numpy().view(np.uint64)
            r0, r1, e0, e1 = int(io[0]), int(io[-1]), int(eo[0]), int(eo[-1])
            p = parity_batch(raw[r0:r1].cpu().numpy(), io - io[0], t["enc"][e0:e1].cpu().numpy(), eo - eo[0],
                             max(1, host_threads(world) // 2))
            for key in ("encoded_offsets_equal", "encoded_bytes_equal", "reference_decodes_gpu_bytes_to_input"):
                sample[key] = sample[key] and p[key]
            sample["strings"] += p["strings"]
            sample["checker"] = p["checker"]
        sample["last_string_bit_offset_in_shard"] = int(t["enc_off"][n - 1].item()) * 8

    # concatenation on rank 0: gather the shard-local offsets and the per-item lengths
    n_max = int(arm.reduce([n], "MAX")[0])
    concat_ok = None
    if world > 1:
        pad_off = torch.zeros(n_max + 1, dtype=torch.int64, device=dev)
        pad_off[:n + 1] = t["enc_off"]
        pad_len = torch.zeros(n_max + 1, dtype=torch.int64, device=dev)
        pad_len[:n] = item_len
        got_off = [torch.empty_like(pad_off) for _ in range(world)] if rank == 0 else None
        got_len = [torch.empty_like(pad_len) for _ in range(world)] if rank == 0 else None
        dist.gather(pad_off, got_off, dst=0)
        dist.gather(pad_len, got_len, dst=0)
        if rank == 0:
            counts = [int(begin[r + 1] - begin[r]) for r in range(world)]
            shard_offs = [got_off[r][:counts[r] + 1].cpu().numpy().view(np.uint64) for r in range(world)]
            glob = pkg.product_library().concat_offsets(shard_offs)
            want = np.zeros(n_total + 1, dtype=np.uint64)
            np.cumsum(np.concatenate([got_len[r][:counts[r]].cpu().numpy() for r in range(world)]).view(np.uint64), out=want[1:])
            concat_ok = bool(np.array_equal(glob, want))
            del got_off, got_len
    else:
        glob = pkg.product_library().concat_offsets([t["enc_off"].cpu().numpy().view(np.uint64)])
        want = np.zeros(n_total + 1, dtype=np.uint64)
        np.cumsum(item_len.cpu().numpy().view(np.uint64), out=want[1:])
        concat_ok = bool(np.array_equal(glob, want))

    step_ms, enc_ms, dec_ms = arm.reduce([t["step_ms"], t["enc_ms"], t["dec_ms"]], "MAX")
    all_raw, all_enc, all_launches, all_ok = arm.reduce([raw_bytes, enc_bytes, t["launches"], float(local_ok)], "SUM")
    shards = arm.gather_objects({"rank": rank, "strings": n, "raw_bytes": raw_bytes, "encoded_bytes": enc_bytes,
                                 "encode_ms": t["enc_ms"], "decode_ms": t["dec_ms"], "parity_sample": sample})
    if rank != 0:
        return None
    whole = 2.0 * (all_raw + all_enc)
    one_way_max = max(s["raw_bytes"] + s["encoded_bytes"] for s in shards)
    return {
        "workload": "sharded_batch: ONE batch of %d HPACK strings of 8-256 B (seed 0x%X) cut by "
                    "aws_huffman_batch_plan_shards into %d byte-balanced shards, one GPU each (BASELINE configs[4])"
                    % (n_total, SEED_SHARDED, world),
        "value": whole / (step_ms * 1e-3) / 1e9, "unit": "GB/s", "scaling": "strong", "n_gpus": world,
        "ms_per_step": step_ms, "steps": steps, "encode_ms": enc_ms, "decode_ms": dec_ms,
        "raw_bytes": int(all_raw), "encoded_bytes": int(all_enc), "gpu_launches": int(all_launches),
        "roofline_frac_per_gpu": {"encode": one_way_max / (enc_ms * 1e-3) / 1e9 / peak,
                                  "decode": one_way_max / (dec_ms * 1e-3) / 1e9 / peak},
        "shard_offsets_are_running_sums_of_item_lengths": all_ok == world,
        "concat_offsets_equal_global_running_sum": concat_ok,
        "parity_sample": {"strings": sum(s["parity_sample"]["strings"] for s in shards),
                          "all_equal": all(s["parity_sample"][k] for s in shards for k in
                                           ("encoded_offsets_equal", "encoded_bytes_equal", "reference_decodes_gpu_bytes_to_input")),
                          "what": "first and last %d strings of every shard vs the reference on the host" % k},
        "shards": [{k2: v for k2, v in s.items() if k2 != "parity_sample"} for s in shards],
    }


def run_stream(arm, args, peak, with_cpu):
    """BASELINE configs[2]+[3]: one `--stream-bytes` Zipf stream, encode then decode, one GPU."""
    torch = arm.torch
    raw_bytes = args.stream_bytes
    in_off = torch.tensor([0, raw_bytes], dtype=torch.int64, device=arm.device)
    raw = symbols_torch(SEED_STREAM, 0, raw_bytes, arm.sampler_t, arm.device)
    t = device_pass(arm, 1, in_off, raw, raw_bytes, args.steps, args.warmup)
    e = e2e_pass(arm, t, max(1, min(args.steps, 3)))
    return t, e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "hpack_batch", "stream"],
                    help="all: the hpack_batch headline plus the stream (1 GPU) and sharded-batch objects")
    ap.add_argument("--strings", type=int, default=1_000_000, help="strings per GPU (hpack_batch)")
    ap.add_argument("--stream-bytes", type=int, default=1 << 30)
    ap.add_argument("--sharded-strings", type=int, default=64 << 20, help="strings of the ONE sharded batch (configs[4])")
    ap.add_argument("--sharded-parity-strings", type=int, default=32768)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "huffman_encode_decode_throughput"
    headline = "stream" if args.workload == "stream" else "hpack_batch"
    config = {
        "workload": ("hpack_batch: HPACK table, %d strings of 8-256 B per GPU, Zipf(1.5) symbols "
                     "(BASELINE configs[1])" % args.strings) if headline == "hpack_batch" else
                    ("stream: one %d-byte Zipf(1.5) stream, HPACK table, encode then decode "
                     "(BASELINE configs[2]+[3])" % args.stream_bytes),
        "step": "one encode pass + one decode pass",
        "bytes_counted": "payload in + payload out per direction (offset arrays not counted)",
        "l2_policy": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)",
        "parallelism": "independent shards per GPU, no collective" if args.gpus > 1 else "single GPU",
    }
    if args.workload == "all":
        config["also_in_this_line"] = (
            ("`stream`: one %d-byte stream on one GPU (configs[2]+[3]); " % args.stream_bytes if world == 1 else
             "(the single stream stays on one GPU: reported by the 1-GPU run); ") +
            "`sharded_batch`: ONE batch of %d strings sharded over %d GPU(s), strong scaling (configs[4])"
            % (args.sharded_strings, world))

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        if headline == "hpack_batch":
            sample = min(args.strings, 250_000 * max(1, min(threads, 32)) // 4)
            threads = min(threads, 64)
        else:
            sample, threads = min(args.stream_bytes, 1 << 26), 1  # one stream: inherently serial
        for _ in range(min(args.warmup, 1)):
            run_cpu(headline, max(1, sample // 8), threads)
        times, gbs_all = [], []
        kind = what = None
        for _ in range(args.steps):
            kind, gbs, dt, what, raw, encd = run_cpu(headline, sample, threads)
            times.append(dt)
            gbs_all.append(gbs)
        value = float(np.mean(gbs_all))
        line = {
            "impl": "reference", "metric": metric, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": kind,
                             "sample": what + ("; items split over %d host threads" % threads if threads > 1 else "")},
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft

    pkg = graft.load_package()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    arm = Arm(pkg, local_rank, rank, world, dist)
    device = arm.device
    peak, _ = measured_peak()

    if headline == "hpack_batch":
        n = args.strings
        lens = string_lengths_torch(SEED_BATCH, rank * n, n, device)
        in_off = torch.zeros(n + 1, dtype=torch.int64, device=device)
        in_off[1:] = torch.cumsum(lens, 0)
        raw_bytes = int(in_off[-1].item())
        # byte index space is per-rank (rank-major) so any shard is reproducible on the host
        raw = symbols_torch(SEED_BATCH, (rank << 40), raw_bytes, arm.sampler_t, device)
    else:
        n = 1
        raw_bytes = args.stream_bytes
        in_off = torch.tensor([0, raw_bytes], dtype=torch.int64, device=device)
        raw = symbols_torch(SEED_STREAM, 0, raw_bytes, arm.sampler_t, device)

    clock_file = tempfile.NamedTemporaryFile(prefix="clocks_r%d_" % rank, suffix=".csv", delete=False).name
    sampler_proc = clocks_sampler(clock_file, local_rank)  # every rank watches its own GPU
    t = device_pass(arm, n, in_off, raw, raw_bytes, args.steps, args.warmup)
    e2e_steps = max(1, min(args.steps, 5))
    e = e2e_pass(arm, t, e2e_steps)
    if sampler_proc is not None:
        sampler_proc.terminate()
        sampler_proc.wait()
    clocks = summarize_clocks(clock_file)
    enc_bytes = t["enc_bytes"]

    # ---- parity gate: the reference's bytes on the same inputs (every rank checks its own batch)
    parity = {}
    if arm.check and not args.no_parity:
        h = e["host"]
        if headline == "hpack_batch":
            parity["hpack_batch"] = parity_batch(h["raw"], h["in_off"], h["enc"], h["enc_off"], host_threads(world))
        else:
            parity["stream"] = parity_stream(h["raw"], h["enc"], enc_bytes, host_threads(world))
        pk = parity[headline]
        good = all(v for k, v in pk.items() if isinstance(v, bool) and k != "bit_offsets_beyond_2^32")
        all_good = arm.reduce([float(good)], "MIN")[0] == 1.0
        pk["ranks_checked"] = world
        if not all_good:
            print(json.dumps({"error": "GPU output differs from the reference's on the same inputs", "rank": rank,
                              "parity_checked": parity}), flush=True)
            return 1

    # ---- aggregate over ranks: max time, sum bytes
    step_bytes = 2.0 * (raw_bytes + enc_bytes)
    ms_per_step, e2e_ms, enc_ms_mean, dec_ms_mean, copy_ms = arm.reduce(
        [t["step_ms"], e["e2e_s"] * 1e3, t["enc_ms"], t["dec_ms"], e["copy_s"] * 1e3], "MAX")
    all_bytes, all_raw, all_enc, all_launches = arm.reduce([step_bytes, raw_bytes, enc_bytes, t["launches"]], "SUM")
    per_rank = arm.gather_objects({
        "rank": rank, "gpu": local_rank, "encode_ms": t["enc_ms"], "decode_ms": t["dec_ms"], "step_ms": t["step_ms"],
        "e2e_ms": e["e2e_s"] * 1e3, "e2e_h2d_gbs": e["h2d_bytes"] / e["e2e_s"] / 1e9, "e2e_d2h_gbs": e["d2h_bytes"] / e["e2e_s"] / 1e9,
        "copies_alone_ms": e["copy_s"] * 1e3,
        "copies_alone_h2d_gbs": (raw_bytes + enc_bytes) / e["copy_s"] / 1e9, "copies_alone_d2h_gbs": (raw_bytes + enc_bytes) / e["copy_s"] / 1e9,
        "clocks": clocks})
    del e
    line = None
    if rank == 0:
        line = {
            "metric": metric, "value": all_bytes / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config,
            "encode_gbs_per_gpu": (raw_bytes + enc_bytes) / (enc_ms_mean * 1e-3) / 1e9,
            "decode_gbs_per_gpu": (raw_bytes + enc_bytes) / (dec_ms_mean * 1e-3) / 1e9,
            "encode_ms": enc_ms_mean, "decode_ms": dec_ms_mean,
            "raw_bytes_per_gpu": raw_bytes, "encoded_bytes_per_gpu": enc_bytes,
            "roofline": roofline(headline, enc_ms_mean, dec_ms_mean, raw_bytes, enc_bytes,
                                 "hpack_batch" if (headline == "hpack_batch" and args.strings == 1_000_000) else
                                 ("stream:%d" % args.stream_bytes if headline == "stream" else None)),
            "e2e": {"value": all_bytes / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": int(raw_bytes + 8 * (n + 1) + enc_bytes + 8 * (n + 1)),
                    "d2h_bytes_per_step": int(enc_bytes + raw_bytes + 2 * 8 * (n + 1) + 2 * 4 * n),
                    "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "copies_alone": {"value": all_bytes / (copy_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": copy_ms,
                                     "what": "the step's H2D and D2H copies from/to the same pinned buffers on two streams, "
                                             "all ranks at once, no kernels: what the host side of the box allows"}},
            "gpu_launches": int(all_launches),
            "clocks": clocks,
            "per_rank": per_rank,
            "parity_checked": parity,
        }
    del t, raw, in_off
    torch.cuda.empty_cache()

    # ---- configs[2]+[3] beside the headline (one GPU only: the single stream does not shard)
    if args.workload == "all" and world == 1:
        st, se = run_stream(arm, args, peak, not args.no_cpu_baseline)
        sb = 2.0 * (st["raw_bytes"] + st["enc_bytes"])
        stream_obj = {
            "workload": "stream: one %d-byte Zipf(1.5) stream, HPACK table, encode then decode (BASELINE configs[2]+[3])" % args.stream_bytes,
            "value": sb / (st["step_ms"] * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": st["step_ms"],
            "encode_ms": st["enc_ms"], "decode_ms": st["dec_ms"],
            "encode_gbs": (st["raw_bytes"] + st["enc_bytes"]) / (st["enc_ms"] * 1e-3) / 1e9,
            "decode_gbs": (st["raw_bytes"] + st["enc_bytes"]) / (st["dec_ms"] * 1e-3) / 1e9,
            "raw_bytes": st["raw_bytes"], "encoded_bytes": st["enc_bytes"], "gpu_launches": int(st["launches"]),
            "roofline": roofline("stream", st["enc_ms"], st["dec_ms"], st["raw_bytes"], st["enc_bytes"], "stream:%d" % args.stream_bytes),
            "e2e": {"value": sb / se["e2e_s"] / 1e9, "unit": "GB/s", "ms_per_step": se["e2e_s"] * 1e3,
                    "h2d_bytes_per_step": se["h2d_bytes"], "d2h_bytes_per_step": se["d2h_bytes"],
                    "copies_alone_gbs": sb / se["copy_s"] / 1e9},
        }
        if arm.check and not args.no_parity:
            ps = parity_stream(se["host"]["raw"], se["host"]["enc"], st["enc_bytes"], host_threads(1))
            parity["stream"] = ps
            if not (ps["encoded_bytes_equal"] and ps["reference_decodes_gpu_prefix_to_input"]):
                print(json.dumps({"error": "GPU stream output differs from the reference's", "parity_checked": parity}), flush=True)
                return 1
        if not args.no_cpu_baseline:
            kind, gbs, dt, what, _, _ = run_cpu("stream", min(args.stream_bytes, 1 << 27), 1)
            stream_obj["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": 1, "kind": kind,
                                          "sample": what + ", encode + decode, one pass of %.1f s" % dt}
        line["stream"] = stream_obj
        del st, se
        torch.cuda.empty_cache()

    # ---- configs[4]: one batch, sharded (every N, N = 1 included)
    if args.workload == "all":
        sharded = run_sharded(arm, args, peak)
        if rank == 0:
            line["sharded_batch"] = sharded
            if arm.check and not (sharded["concat_offsets_equal_global_running_sum"] and
                                  sharded["shard_offsets_are_running_sums_of_item_lengths"] and sharded["parity_sample"]["all_equal"]):
                print(json.dumps({"error": "sharded batch failed its checks", "sharded_batch": sharded}), flush=True)
                return 1

    if rank == 0:
        if not args.no_cpu_baseline and args.gpus == 1:
            # a bounded sample worth ~10 s of single-core work: the whole 1M-string batch, best of 2 passes
            sample = min(args.strings, 1_000_000) if headline == "hpack_batch" else min(args.stream_bytes, 1 << 28)
            repeats = 2
            kind, gbs, dt, what, _, _ = run_cpu(headline, sample, 1, repeats=repeats)
            line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": 1, "kind": kind,
                                    "sample": what + ", encode + decode, best of %d passes of %.1f s" % (repeats, dt),
                                    "host_cpus": os.cpu_count()}
            line["host_codec"] = host_codec_leg()
        print(json.dumps(line))
    arm.ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
