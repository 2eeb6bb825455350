#!/usr/bin/env python3
"""Headline benchmark: Huffman encode + decode GB/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hpack_batch|stream] [--impl reference]

One "step" = one encode pass + one decode pass of the workload through the batched C ABI.
  value      device-resident: inputs already in HBM, kernels timed with CUDA events on the launching
             stream; algorithmic bytes (payload in + payload out, each direction) / time.
  e2e        the same step through the host-pointer entry points (aws_huffman_encode_batch /
             aws_huffman_decode_batch) from pinned host buffers, H2D and D2H copies inside the timing.
  roofline   the dominant kernel family (encode or decode, whichever takes longer) against the measured
             HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's CPU path (oracle/_ref if it was built, else the oracle port) on this
             host, 1 thread, on a stated sample of the same workload.

Workloads:
  hpack_batch  BASELINE configs[1]: HPACK table, 1,000,000 strings of 8..256 B per GPU, Zipf(1.5) bytes.
  stream       BASELINE configs[2]+[3]: one 2^30-byte Zipf(1.5) stream, encode then decode (1 GPU).

Under torchrun (N > 1) every rank runs the same per-GPU workload on its own shard (independent
strings, no collective: weak scaling); time is the max over ranks, value the sum of bytes / that time.
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


def profiled_traffic(workload, dominant, raw_bytes):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r1_kernels.json, made by tools/ncu_summary.py), or None. The stream capture was taken on a
    256 MiB stream; traffic is proportional to the stream size, so it is scaled to this run's."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r1_kernels.json")))
        group = prof["hpack_batch" if workload == "hpack_batch" else "stream_256MiB"]
        want = {"encode": "encode_slots" if workload == "hpack_batch" else "encode_tiled",
                "decode": "decode_batch" if workload == "hpack_batch" else "stream_fused_kernel"}[dominant]
        for k in group:
            if want in k["kernel"]:
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
                total = k["dram_read"] * scale[k["dram_read_unit"]] + k["dram_write"] * scale[k["dram_write_unit"]]
                if workload != "hpack_batch":
                    total *= raw_bytes / float(1 << 28)
                return int(total)
    except Exception:
        pass
    return None


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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hpack_batch", choices=["hpack_batch", "stream"])
    ap.add_argument("--strings", type=int, default=1_000_000, help="strings per GPU (hpack_batch)")
    ap.add_argument("--stream-bytes", type=int, default=1 << 30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "huffman_encode_decode_throughput"
    config = {
        "workload": ("hpack_batch: HPACK table, %d strings of 8-256 B per GPU, Zipf(1.5) symbols "
                     "(BASELINE configs[1])" % args.strings) if args.workload == "hpack_batch" else
                    ("stream: one %d-byte Zipf(1.5) stream, HPACK table, encode then decode "
                     "(BASELINE configs[2]+[3])" % args.stream_bytes),
        "step": "one encode pass + one decode pass",
        "bytes_counted": "payload in + payload out per direction (offset arrays not counted)",
        "l2_policy": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)",
        "parallelism": "independent shards per GPU, no collective" if args.gpus > 1 else "single GPU",
    }

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        if args.workload == "hpack_batch":
            sample = min(args.strings, 250_000 * max(1, min(threads, 32)) // 4)
            threads = min(threads, 64)
        else:
            sample, threads = min(args.stream_bytes, 1 << 26), 1  # one stream: inherently serial
        for _ in range(min(args.warmup, 1)):
            run_cpu(args.workload, max(1, sample // 8), threads)
        times, gbs_all = [], []
        kind = what = None
        for _ in range(args.steps):
            kind, gbs, dt, what, raw, encd = run_cpu(args.workload, sample, threads)
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
    import refcodec

    pkg = graft.load_package()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    sampler_np = refcodec.zipf_symbol_sampler(refcodec.table_arrays("hpack")[1])
    sampler_t = torch.from_numpy(sampler_np).to(device)
    ctx = pkg.BatchContext(pkg.coders_library().coder("hpack"), eos_padding=0xFF, device=local_rank)
    # a real (non-default) stream: the C ABI reads a NULL stream as "the context's own stream", and
    # torch's default stream handle is 0. Kernels and timing events must sit on the same stream.
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    if args.workload == "hpack_batch":
        n = args.strings
        first_string = rank * n
        lens = string_lengths_torch(SEED_BATCH, first_string, n, device)
        in_off = torch.zeros(n + 1, dtype=torch.int64, device=device)
        in_off[1:] = torch.cumsum(lens, 0)
        raw_bytes = int(in_off[-1].item())
        # byte index space is per-rank (rank-major) so any shard is reproducible on the host
        raw = symbols_torch(SEED_BATCH, (rank << 40), raw_bytes, sampler_t, device)
    else:
        n = 1
        raw_bytes = args.stream_bytes
        in_off = torch.tensor([0, raw_bytes], dtype=torch.int64, device=device)
        raw = symbols_torch(SEED_STREAM, 0, raw_bytes, sampler_t, device)

    enc_cap = raw_bytes + raw_bytes // 2 + 1024
    enc = torch.empty(enc_cap, dtype=torch.uint8, device=device)
    enc_off = torch.zeros(n + 1, dtype=torch.int64, device=device)
    dec = torch.empty(raw_bytes + 1024, dtype=torch.uint8, device=device)
    dec_off = torch.zeros(n + 1, dtype=torch.int64, device=device)
    enc_status = torch.zeros(n, dtype=torch.int32, device=device)
    dec_status = torch.zeros(n, dtype=torch.int32, device=device)

    def encode_step():
        ctx.encode_device(n, {"in_": raw, "in_offsets": in_off, "out": enc, "out_offsets": enc_off,
                              "status": enc_status}, raw_bytes, enc_cap, stream=sptr)

    def decode_step():
        ctx.decode_device(n, {"in_": enc, "in_offsets": enc_off, "out": dec, "out_offsets": dec_off,
                              "status": dec_status}, enc_bytes_known[0], raw_bytes + 1024, stream=sptr)

    # one checked pass before any timing: round trip must be exact
    enc_bytes_known = [0]
    encode_step()
    torch.cuda.synchronize(device)
    enc_bytes = int(enc_off[-1].item())
    enc_bytes_known[0] = enc_bytes
    decode_step()
    torch.cuda.synchronize(device)
    if not os.environ.get("AWS_HUFFMAN_BATCH_EXPERIMENT"):
        assert int(dec_off[-1].item()) == raw_bytes and torch.equal(dec[:raw_bytes], raw), "round trip mismatch"
        assert int(enc_status.abs().sum().item()) == 0 and int(dec_status.abs().sum().item()) == 0

    for _ in range(args.warmup):
        encode_step()
        decode_step()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)

    clock_file = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
    sampler_proc = clocks_sampler(clock_file, local_rank) if rank == 0 else None

    launches_before = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    ev[0].record(stream)
    for s in range(args.steps):
        encode_step()
        ev[2 * s + 1].record(stream)
        decode_step()
        ev[2 * s + 2].record(stream)
    torch.cuda.synchronize(device)
    launches = ctx.launch_count - launches_before
    enc_ms = [ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(args.steps)]
    dec_ms = [ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])

    # ---- e2e through the host-pointer C ABI, pinned host buffers, copies inside the timed region
    e2e_steps = max(1, min(args.steps, 5))

    def pinned(count, dtype):
        return torch.empty(count, dtype=dtype).pin_memory().numpy()

    h_raw = raw.cpu().pin_memory().numpy()
    h_in_off = in_off.cpu().to(torch.int64).pin_memory().numpy().view(np.uint64)
    h_enc = pinned(enc_cap, torch.uint8)
    h_enc_off = pinned(n + 1, torch.int64).view(np.uint64)
    h_dec = pinned(raw_bytes + 1024, torch.uint8)
    h_dec_off = pinned(n + 1, torch.int64).view(np.uint64)
    h_status = pinned(n, torch.int32)

    def e2e_step():
        # the call a user of the C ABI makes: aws_huffman_encode_batch / aws_huffman_decode_batch on host memory
        ctx._call("aws_huffman_encode_batch", n, {"in_": h_raw, "in_offsets": h_in_off, "out": h_enc,
                                                  "out_offsets": h_enc_off, "status": h_status}, enc_cap)
        total = int(h_enc_off[n])
        ctx._call("aws_huffman_decode_batch", n, {"in_": h_enc, "in_offsets": h_enc_off, "out": h_dec,
                                                  "out_offsets": h_dec_off, "status": h_status}, raw_bytes + 1024)
        return total, int(h_dec_off[n])

    e2e_step()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        total, back = e2e_step()
    torch.cuda.synchronize(device)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if not os.environ.get("AWS_HUFFMAN_BATCH_EXPERIMENT"):
        assert total == enc_bytes and back == raw_bytes
        assert np.array_equal(h_dec[:raw_bytes], h_raw), "e2e round trip mismatch"
        assert not h_status.any()

    if sampler_proc is not None:
        sampler_proc.terminate()
        sampler_proc.wait()

    # ---- aggregate over ranks: max time, sum bytes
    step_bytes = 2.0 * (raw_bytes + enc_bytes)
    stats = torch.tensor([total_ms / args.steps, e2e_s * 1e3, float(np.mean(enc_ms)), float(np.mean(dec_ms))],
                         dtype=torch.float64, device=device)
    sums = torch.tensor([step_bytes, float(raw_bytes), float(enc_bytes), float(launches)],
                        dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_per_step, e2e_ms, enc_ms_mean, dec_ms_mean = [float(x) for x in stats.cpu()]
    all_bytes, all_raw, all_enc, all_launches = [float(x) for x in sums.cpu()]

    if rank == 0:
        peak, peak_src = measured_peak()
        one_way = (raw_bytes + enc_bytes)  # per GPU, per direction
        enc_gbs = one_way / (enc_ms_mean * 1e-3) / 1e9
        dec_gbs = one_way / (dec_ms_mean * 1e-3) / 1e9
        dominant = "decode" if dec_ms_mean >= enc_ms_mean else "encode"
        achieved = dec_gbs if dominant == "decode" else enc_gbs
        line = {
            "metric": metric, "value": all_bytes / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config,
            "encode_gbs_per_gpu": enc_gbs, "decode_gbs_per_gpu": dec_gbs,
            "encode_ms": enc_ms_mean, "decode_ms": dec_ms_mean,
            "raw_bytes_per_gpu": raw_bytes, "encoded_bytes_per_gpu": enc_bytes,
            "roofline": {"bound": "hbm",
                         "kernel": {"hpack_batch": {"encode": "encode_slots_kernel (+ its slot scan and tile index launches)",
                                                    "decode": "decode_batch_kernel"},
                                    "stream": {"encode": "encode_tiled_kernel<false>",
                                               "decode": "stream_fused_kernel (+ verify and gated fallback launches)"}}
                                   [args.workload][dominant] + " (per GPU; CUDA events around the " + dominant + " call)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profiled_traffic(args.workload, dominant, raw_bytes),
                         "algorithmic_bytes_per_launch": int(one_way),
                         "peak_source": peak_src, "frac_of_8000_nominal": achieved / 8000.0,
                         "encode_frac": enc_gbs / peak, "decode_frac": dec_gbs / peak},
            "e2e": {"value": all_bytes / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": int(raw_bytes + 8 * (n + 1) + enc_bytes + 8 * (n + 1)),
                    "d2h_bytes_per_step": int(enc_bytes + raw_bytes + 2 * 8 * (n + 1) + 2 * 4 * n),
                    "ms_per_step": e2e_ms, "steps": e2e_steps},
            "gpu_launches": int(all_launches),
            "clocks": summarize_clocks(clock_file),
        }
        if not args.no_cpu_baseline and args.gpus == 1:
            # a bounded sample worth ~10 s of single-core work: the whole 1M-string batch, best of 4 passes
            sample = min(args.strings, 1_000_000) if args.workload == "hpack_batch" else min(args.stream_bytes, 1 << 28)
            repeats = 4 if args.workload == "hpack_batch" else 2
            kind, gbs, dt, what, _, _ = run_cpu(args.workload, sample, 1, repeats=repeats)
            line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": 1, "kind": kind,
                                    "sample": what + ", encode + decode, best of %d passes of %.1f s" % (repeats, dt),
                                    "host_cpus": os.cpu_count()}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
