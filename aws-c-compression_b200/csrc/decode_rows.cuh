// Batch decoder variant of round 2: TWO kernels, no worst-case rows in shared memory. OPT-IN (AWS_HUFFMAN_BATCH_ROWS=1
// at context creation): MEASURED SLOWER than decode_batch_kernel (decode_fast.cuh) — kept, tested, as the recorded A/B
// of what round 1's review proposed ("decode straight into a worst-case-spaced global scratch, then one HBM-bound
// compaction pass: spend bandwidth to buy occupancy").
//
//   decode_rows_kernel    every string is decoded ONCE, from global memory, into a row of a global scratch at
//                         worst-case spacing (row i starts at 16 floor(in_offset_i / (2 min_len)) + 32 i: known
//                         without knowing any symbol count). Shared memory holds the table and, per WARP, two
//                         small rings (input words, output words; [word][lane]: a lane's words sit in its own bank,
//                         so ring traffic never conflicts). No rows, no stage, no teams: 24 warps per SM, and a
//                         warp never waits for another one inside a pool of strings. Pools of 1024 strings are
//                         counting-sorted by length; warps pull groups of 32 similar lengths, longest first.
//   compact_rows_kernel   symbol counts -> offsets (block scan + single-pass decoupled look-back) and rows -> dense
//                         output: the tile's stretch of the scratch comes in coalesced, threads move their rows into
//                         a shared-memory image of the tile's output, which leaves with coalesced 128-bit stores.
//
// What the measurement says (1M strings of 8..256 B, profiles/README.md): decode_rows_kernel alone takes 230 us
// against 312 us for ALL of decode_batch_kernel. It does what it was built for — no barrier stall, 24 balanced warps
// per SM, issue slots 59 % busy instead of 42 % — but a lane that streams its string through per-lane rings pays
// for them in instructions (40 per step instead of 30: ring addresses, upkeep once per round, per-string set-up), and
// the decode step is bound by the ALU pipe (60 % busy; shifts, logic ops, selects and compares all issue there, one
// warp instruction every two cycles), not by latency: more resident warps do not buy back more instructions. The
// compaction pass (87..240 us depending on its shape: every version was a chain of latencies with one or two blocks per
// SM) comes on top. The staged, team-based kernel stays the product path; its balance problem is attacked
// inside it (larger tiles for the same shared memory).
//
// The decode step is the lean 64-bit-entry step of decode_fast.cuh (one lookup per step whatever the table, two
// symbols per root entry); the last < 32 bits of a string run the exact one-symbol loop with the reference's
// end-of-stream rules (huffman.c:196-211, 240-255) on a 64-bit register. Results are bit-identical to
// decode_batch_kernel (tests/test_gpu_multi.py: the rows decoder test).
#pragma once

#include "decode_fast.cuh"

namespace hb {

#ifndef HB_ROWS_POOL
#define HB_ROWS_POOL 1024
#endif
#ifndef HB_ROWS_BLOCKS
#define HB_ROWS_BLOCKS 3
#endif
constexpr int kRowsThreads = 256;
constexpr int kRowsWarps = kRowsThreads / 32;
constexpr int kRowsPool = HB_ROWS_POOL;     // strings sorted and dealt out together
constexpr uint32_t kRowsSlack = 32;          // bytes between rows beyond the worst case (16-byte flushes, alignment)
constexpr uint32_t kRowsRingBytes = 2048;    // per warp: input ring [8 words][32 lanes] + output ring [8][32]

struct DecRowsArgs {
    BatchView b;
    uint64_t total_in;
    const uint2 *lut2;
    uint32_t lut2_count, lut2_trap, root_bits, min_len;
    uint8_t *rows;       // scratch: worst-case spaced rows, 16-byte aligned
    uint64_t *row_pos;   // n: where string i's row starts (bytes from `rows`)
    uint32_t *cnt;       // n: symbols of string i
    uint32_t *ticket;    // zeroed
    uint32_t num_pools;
};

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// address of byte `o` of a lane's output ring (base: the lane's column, 1 KiB-aligned ring)
__device__ __forceinline__ uint32_t rows_out_addr(uint32_t obase, uint32_t o) { return obase + ((o << 5) & 0x380u) + (o & 3u); }

__global__ void __launch_bounds__(kRowsThreads, HB_ROWS_BLOCKS) decode_rows_kernel(DecRowsArgs a) {
    extern __shared__ __align__(128) uint8_t s_rows_dyn[];  // [LUT2][pad to 1 KiB][rings: kRowsWarps x 2 KiB]
    __shared__ uint32_t s_len[kRowsPool];
    __shared__ uint16_t s_perm[kRowsPool];
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_wsum[kRowsWarps];
    __shared__ uint32_t s_pool, s_next;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Lut2 t = lut2_load(reinterpret_cast<uint2 *>(s_rows_dyn), a.lut2, a.lut2_count, a.root_bits, a.lut2_trap);
    const uint32_t ring0 = ((t.addr + 8u * a.lut2_count + 1023u) & ~1023u) + warp * kRowsRingBytes;
    const uint32_t ibase = ring0 + 4u * lane, obase = ring0 + 1024u + 4u * lane;
    const BatchView &b = a.b;
    const uint8_t *const in_end = b.in + a.total_in;  // (the strings are packed: the batch's input is in[0, total_in))
    const uint32_t root_tb = t.addr;

    for (;;) {
        __syncthreads();  // the previous pool is done (first trip: the table is in place)
        if (tid == 0) {
            s_pool = atomicAdd(a.ticket, 1u);
            s_next = kRowsWarps;
        }
        s_hist[tid] = 0;
        __syncthreads();
        const uint32_t pool = s_pool;
        if (pool >= a.num_pools) break;
        const uint64_t item0 = (uint64_t)pool * kRowsPool;
        const uint32_t nitems = (uint32_t)min((uint64_t)kRowsPool, b.n - item0);
        const uint32_t ngroups = (nitems + 31u) >> 5;
        // ---- counting sort of the pool by decreasing length (4-byte buckets) ------------------------------------
        for (uint32_t i = tid; i < nitems; i += kRowsThreads) {
            const uint64_t len = b.in_offsets[item0 + i + 1] - b.in_offsets[item0 + i];
            const uint32_t l32 = (uint32_t)min(len, (uint64_t)0xffffffffu);
            s_len[i] = l32;
            atomicAdd(&s_hist[255u - min((l32 + 3u) >> 2, 255u)], 1u);
        }
        __syncthreads();
        {
            const uint32_t v = s_hist[tid];
            const uint32_t incl = warp_inclusive_scan(v);
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t before = 0;
#pragma unroll
            for (int w = 0; w < kRowsWarps; ++w)
                if ((uint32_t)w < warp) before += s_wsum[w];
            s_hist[tid] = before + incl - v;
        }
        __syncthreads();
        for (uint32_t i = tid; i < nitems; i += kRowsThreads)
            s_perm[atomicAdd(&s_hist[255u - min((s_len[i] + 3u) >> 2, 255u)], 1u)] = (uint16_t)i;
        __syncthreads();

        // ---- decode: warps pull groups of 32 strings of similar length, longest first -----------------------------
        for (uint32_t g = warp; g < ngroups;) {
            const uint32_t slot = g * 32u + lane;
            if (slot < nitems) {
                const uint32_t it = s_perm[slot];
                const uint64_t item = item0 + it;
                const uint64_t in0 = b.in_offsets[item];
                const uint64_t len = b.in_offsets[item + 1] - in0;
                const uint64_t row = 16ull * (in0 / (2ull * a.min_len)) + (uint64_t)kRowsSlack * item;
                uint8_t *const rowp = a.rows + row;
                a.row_pos[item] = row;
                const uint8_t *const payload = b.in + in0;
                uint64_t cbits = 0;
                uint32_t nsym = 0, term = kTermEnd;
                const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(payload) & 15);
                const uint8_t *const v0 = payload - lead;
                // (bit positions are 32-bit in the fast path, and it reads whole 16-byte vectors: a string whose last
                // vector crosses the end of the input buffer — the batch's last few bytes — takes the exact loop too)
                const bool slow = len >= (1ull << 28) || v0 + ((lead + len + 15) & ~15ull) > in_end;
                bool trapped = false;
                if (len && !slow) {
                    const uint4 *const vp = reinterpret_cast<const uint4 *>(v0);
                    const uint32_t jmax = ((lead + (uint32_t)len + 15u) >> 4) - 1u;  // the string's last vector
                    // four big-endian stream words into the input ring at word0 .. word0 + 3 (word0 % 4 == 0)
                    auto ring_put = [&](uint32_t word0, const uint4 &v) {
                        const uint32_t p = ibase + ((word0 & 4u) << 7);
                        sts_u32(p, __byte_perm(v.x, 0, 0x0123));
                        sts_u32(p + 128u, __byte_perm(v.y, 0, 0x0123));
                        sts_u32(p + 256u, __byte_perm(v.z, 0, 0x0123));
                        sts_u32(p + 384u, __byte_perm(v.w, 0, 0x0123));
                    };
                    uint32_t pos = 8u * lead;
                    const uint32_t pos0 = pos, end = pos + 8u * (uint32_t)len;
                    // (vectors past the string's last are never needed: the index is clamped instead of tested)
                    ring_put(0, __ldg(vp));
                    ring_put(4, __ldg(vp + min(1u, jmax)));
                    uint4 nv = __ldg(vp + min(2u, jmax));  // prefetched, still little-endian
                    uint32_t wtop = 8;  // words of the stream that have been put into the ring
                    const uint32_t wi = pos >> 5;
                    uint32_t w0 = lds_u32(ibase + (wi << 7)), w1 = lds_u32(ibase + ((wi + 1u) << 7)), w2 = lds_u32(ibase + ((wi + 2u) << 7));
                    uint32_t wa = ibase + (((wi + 3u) & 7u) << 7);
                    int limit = (int)(32u * (wi + 1u));
                    // A step is safe while 32 real bits follow (no code is longer): both symbols of a root entry then
                    // lie inside the string. (The chunked stream decoder stops root_bits earlier because its spans
                    // end before the stream does; a string's span ends with the string.)
                    const uint32_t pair_end = end >= 32u ? end - 31u : 0u;
                    uint32_t ns = t.root_ns, tb = root_tb, acc = 0, out = 0, flushed = 0;
                    uint8_t *rowcur = rowp;  // rowp + flushed
                    while (pos < pair_end || tb != root_tb) {
                        // Ring upkeep, once per round and branch-free (lanes need it at different times, so a branch
                        // would be taken by the warp in nearly every round anyway). Input: a round pulls at most three
                        // words; when at most four are left, the prefetched vector goes in (the slots it overwrites
                        // were pulled long ago) and the next one is requested. Output: a round adds at most 12 bytes;
                        // a complete 16-byte piece of the row leaves the ring.
                        asm volatile(
                            "{\n\t"
                            ".reg .pred q;\n\t"
                            ".reg .b32 a, p, t0, t1, t2, t3, j;\n\t"
                            ".reg .b64 off, adr;\n\t"
                            "shr.u32 a, %5, 5;\n\t"
                            "sub.u32 a, %0, a;\n\t"
                            "setp.le.u32 q, a, 6;\n\t"
                            "and.b32 p, %0, 4;\n\t"
                            "shl.b32 p, p, 7;\n\t"
                            "add.u32 p, p, %6;\n\t"
                            "prmt.b32 t0, %1, 0, 0x0123;\n\t"
                            "prmt.b32 t1, %2, 0, 0x0123;\n\t"
                            "prmt.b32 t2, %3, 0, 0x0123;\n\t"
                            "prmt.b32 t3, %4, 0, 0x0123;\n\t"
                            "@q st.shared.u32 [p], t0;\n\t"
                            "@q st.shared.u32 [p+128], t1;\n\t"
                            "@q st.shared.u32 [p+256], t2;\n\t"
                            "@q st.shared.u32 [p+384], t3;\n\t"
                            "@q add.u32 %0, %0, 4;\n\t"
                            "shr.u32 j, %0, 2;\n\t"
                            "min.u32 j, j, %7;\n\t"
                            "mul.wide.u32 off, j, 16;\n\t"
                            "add.u64 adr, %8, off;\n\t"
                            "@q ld.global.nc.v4.u32 {%1, %2, %3, %4}, [adr];\n\t"
                            "}"
                            : "+r"(wtop), "+r"(nv.x), "+r"(nv.y), "+r"(nv.z), "+r"(nv.w)
                            : "r"(limit), "r"(ibase), "r"(jmax), "l"(vp)
                            : "memory");
                        asm volatile(
                            "{\n\t"
                            ".reg .pred q;\n\t"
                            ".reg .b32 a, p, t0, t1, t2, t3;\n\t"
                            "sub.u32 a, %2, %0;\n\t"
                            "setp.ge.u32 q, a, 16;\n\t"
                            "and.b32 p, %0, 16;\n\t"
                            "shl.b32 p, p, 5;\n\t"
                            "add.u32 p, p, %3;\n\t"
                            "@q ld.shared.u32 t0, [p];\n\t"
                            "@q ld.shared.u32 t1, [p+128];\n\t"
                            "@q ld.shared.u32 t2, [p+256];\n\t"
                            "@q ld.shared.u32 t3, [p+384];\n\t"
                            "@q st.global.v4.u32 [%1], {t0, t1, t2, t3};\n\t"
                            "@q add.u32 %0, %0, 16;\n\t"
                            "@q add.u64 %1, %1, 16;\n\t"
                            "}"
                            : "+r"(flushed), "+l"(rowcur)
                            : "r"(out), "r"(obase)
                            : "memory");
#pragma unroll
                        for (int step = 0; step < kUnifiedSteps; ++step) {
                            asm volatile(
                                "{\n\t"
                                ".reg .pred p, c, w;\n\t"
                                ".reg .b32 win, idx, adr, x, u, t, sh, lo, sp, no, wadr;\n\t"
                                "setp.lt.u32 p, %0, %10;\n\t"
                                "setp.ne.or.u32 p, %8, %11, p;\n\t"
                                "shf.l.wrap.b32 win, %2, %1, %0;\n\t"
                                "shf.r.wrap.b32 idx, win, 0, %7;\n\t"
                                "mad.lo.u32 adr, idx, 8, %8;\n\t"
                                "mov.b32 x, 0;\n\t"
                                "@p ld.shared.v2.u32 {x, %7}, [adr];\n\t"
                                "and.b32 %8, %7, 0x00ffffe0;\n\t"
                                "shr.u32 u, %7, 24;\n\t"
                                "@p add.u32 %0, %0, u;\n\t"
                                "and.b32 t, x, 0xffff;\n\t"
                                "shl.b32 sh, %6, 3;\n\t"
                                "shf.l.wrap.b32 lo, 0, t, sh;\n\t"
                                "shf.l.wrap.b32 sp, t, 0, sh;\n\t"
                                "or.b32 lo, lo, %9;\n\t"
                                "shr.u32 u, x, 30;\n\t"
                                "add.u32 no, %6, u;\n\t"
                                "xor.b32 u, no, %6;\n\t"
                                "and.b32 u, u, 4;\n\t"
                                "setp.ne.u32 w, u, 0;\n\t"
                                "shl.b32 wadr, %6, 5;\n\t"
                                "lop3.b32 wadr, wadr, 0x380, %12, 0xEA;\n\t"
                                "@w st.shared.u32 [wadr], lo;\n\t"
                                "selp.b32 %9, sp, lo, w;\n\t"
                                "mov.b32 %6, no;\n\t"
                                "setp.ge.s32 c, %0, %5;\n\t"
                                "@c mov.b32 %1, %2;\n\t"
                                "@c mov.b32 %2, %3;\n\t"
                                "@c ld.shared.u32 %3, [%4];\n\t"
                                "@c add.u32 u, %4, 128;\n\t"
                                "@c lop3.b32 %4, %4, u, 0x380, 0xD8;\n\t"
                                "@c add.s32 %5, %5, 32;\n\t"
                                "}"
                                : "+r"(pos), "+r"(w0), "+r"(w1), "+r"(w2), "+r"(wa), "+r"(limit), "+r"(out), "+r"(ns), "+r"(tb), "+r"(acc)
                                : "r"(pair_end), "r"(root_tb), "r"(obase)
                                : "memory");
                        }
                        if (tb == t.trap_tb) {  // no code matches: the exact loop over global memory redoes the string
                            trapped = true;
                            break;
                        }
                    }
                    if (!trapped) {
                        if (out & 3u) sts_u32(rows_out_addr(obase, out & ~3u), acc);  // the symbols still in `acc`
                        // the rest of the string: one lookup at a time with the end-of-stream rules; the window is the
                        // stream zero-extended (huffman.c:196-211), a code that does not fit ends the stream (:240-255).
                        // At most 31 + root_bits bits are left, all of them in w0..w2: they move into ONE 64-bit
                        // register, left-aligned and zero-extended, so a window is its upper half and consuming is a shift.
                        auto flush = [&]() {
                            const uint32_t p = obase + ((flushed & 16u) << 5);
                            const uint4 v = make_uint4(lds_u32(p), lds_u32(p + 128u), lds_u32(p + 256u), lds_u32(p + 384u));
                            *reinterpret_cast<uint4 *>(rowp + flushed) = v;
                            flushed += 16u;
                        };
                        uint32_t left = end - pos;  // <= 31 + root_bits (a string shorter than that never entered the loop)
                        const uint32_t rel = pos - ((uint32_t)limit - 32u);  // < 32
                        uint64_t rest = ((uint64_t)__funnelshift_l(w1, w0, rel) << 32) | __funnelshift_l(w2, w1, rel);
                        rest = left >= 64u ? rest : left ? rest & (~0ull << (64u - left)) : 0ull;
                        term = kTermStop;
                        while (left) {
                            if (out - flushed >= 16u) flush();  // (the ring holds 32 bytes)
                            const Lut2Hit h = lut2_lookup(t, (uint32_t)(rest >> 32));
                            if (h.x == 0) {
                                term = left < 32u ? kTermEnd : kTermUnknown;  // fewer than 32 bits left: padding
                                break;
                            }
                            const uint32_t len1 = lut2_len1(h.x);
                            if (len1 > left) {
                                term = kTermEnd;  // a code cut short by the end of the stream
                                break;
                            }
                            const bool two = (h.x >> 30) == 2u && h.total <= left;
                            sts_u8(rows_out_addr(obase, out), h.x);
                            if (two) sts_u8(rows_out_addr(obase, out + 1u), h.x >> 8);
                            out += two ? 2u : 1u;
                            const uint32_t used = two ? h.total : len1;
                            rest <<= used;
                            left -= used;
                            pos += used;
                        }
                        if (term == kTermStop) term = kTermEnd;
                        while (flushed < out) flush();
                        cbits = pos - pos0;
                        nsym = out;
                    }
                }
                if (len && (slow || trapped)) {  // one symbol at a time, straight from global memory to the row
                    ByteWriter wr;
                    wr.init(rowp, ~0ull);
                    const DecodeSpan r = decode_span_lut2<true>(t, payload, 0, len * 8, len, &wr);
                    wr.finish();
                    cbits = r.pos;
                    nsym = (uint32_t)r.nsym;
                    term = r.term;
                }
                int32_t st = term == kTermUnknown ? kStatusUnknownSymbol : kStatusOk;
                a.cnt[item] = nsym;
                if (b.status) b.status[item] = st;
                if (b.consumed || b.leftover_working_bits || b.leftover_num_bits)
                    leftover_state(
                        payload, len, cbits, term == kTermUnknown, b.consumed ? b.consumed + item : nullptr,
                        b.leftover_working_bits ? b.leftover_working_bits + item : nullptr,
                        b.leftover_num_bits ? b.leftover_num_bits + item : nullptr);
            }
            uint32_t next = 0;
            if (lane == 0) next = atomicAdd(&s_next, 1u);
            g = __shfl_sync(0xffffffffu, next, 0);
        }
    }
}

// ---- rows -> dense output ------------------------------------------------------------------------------------
constexpr int kCompactThreads = 256;
constexpr int kCompactMaxItems = 2 * kCompactThreads;  // strings per tile (two per thread in the block scan)

struct CompactArgs {
    BatchView b;
    const uint8_t *rows;
    const uint64_t *row_pos;
    const uint32_t *cnt;
    uint64_t *tile_state;  // look-back descriptors (32-byte aligned, zeroed)
    uint32_t *ticket;      // zeroed
    uint32_t num_tiles;
    uint32_t items_per_tile;  // <= kCompactMaxItems, even
    uint32_t stage_bytes;     // shared memory: [stage: the tile's stretch of the row scratch][image: its dense output]
    uint32_t image_bytes;
};

// A tile's rows lie in ONE stretch of the scratch (rows follow each other at worst-case spacing), so the stretch
// comes in with coalesced 128-bit loads, all in flight at once (first version: every thread read its own rows, a
// chain of L2 / HBM latencies: 174 us for 1M strings). Threads then move their rows (shared to shared) into a dense
// image of the tile's output, laid out with the alignment of its global address, which leaves with coalesced
// 128-bit stores.
__global__ void __launch_bounds__(kCompactThreads) compact_rows_kernel(CompactArgs a) {
    extern __shared__ __align__(16) uint8_t s_compact[];
    __shared__ uint32_t s_off[kCompactMaxItems];
    __shared__ uint32_t s_src[kCompactMaxItems];  // start of the row in the stage
    __shared__ uint32_t s_cnt[kCompactMaxItems];
    __shared__ uint64_t s_warp_sum[kCompactThreads / 32];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base, s_total, s_r0, s_r1;
    uint8_t *const s_stage = s_compact;
    uint8_t *const s_image = s_compact + a.stage_bytes;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BatchView &b = a.b;
    for (;;) {
        __syncthreads();  // the previous tile has left the image
        if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= a.num_tiles) break;
        const uint64_t item0 = (uint64_t)tile * a.items_per_tile;
        const uint32_t nitems = (uint32_t)min((uint64_t)a.items_per_tile, b.n - item0);
        // ---- offsets: block scan (2 strings per thread), the tile's count goes out at once ------------------------
        uint32_t c0 = 0, c1 = 0;
        uint64_t p0 = 0, p1 = 0;
        if (2 * tid + 1 < nitems) {
            const uint2 c = *reinterpret_cast<const uint2 *>(a.cnt + item0 + 2 * tid);  // (item0 is even)
            const ulonglong2 p = *reinterpret_cast<const ulonglong2 *>(a.row_pos + item0 + 2 * tid);
            c0 = c.x;
            c1 = c.y;
            p0 = p.x;
            p1 = p.y;
        } else if (2 * tid < nitems) {
            c0 = a.cnt[item0 + 2 * tid];
            p0 = a.row_pos[item0 + 2 * tid];
        }
        if (tid == 0) s_r0 = p0;
        if (2 * tid + 1 == nitems - 1) s_r1 = p1 + ((c1 + 15u) & ~15ull);
        if (2 * tid == nitems - 1) s_r1 = p0 + ((c0 + 15u) & ~15ull);
        const uint64_t mine = (uint64_t)c0 + c1;
        const uint64_t incl = warp_inclusive_scan64(mine);
        if (lane == 31) s_warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = lane < kCompactThreads / 32 ? s_warp_sum[lane] : 0;
            const uint64_t wi = warp_inclusive_scan64(w);
            if (lane < kCompactThreads / 32) s_warp_sum[lane] = wi - w;
            const uint64_t total = __shfl_sync(0xffffffffu, wi, kCompactThreads / 32 - 1);
            if (lane == 0) {
                lookback_publish_aggregate(a.tile_state, tile, total);
                s_total = total;
            }
        }
        // ---- the stretch of the scratch comes in while warp 0 resolves the tile's position ---------------------------
        const uint64_t r0 = s_r0, r1 = s_r1;  // (written before the barrier above)
        const bool staged = r1 - r0 <= a.stage_bytes;
        if (staged) {
            const uint4 *g = reinterpret_cast<const uint4 *>(a.rows + r0);  // rows start on 16-byte boundaries
            uint4 *d = reinterpret_cast<uint4 *>(s_stage);
            const uint32_t nvec = (uint32_t)((r1 - r0) >> 4);
            uint32_t v = tid;
            for (; v + 3 * kCompactThreads < nvec; v += 4 * kCompactThreads) {
                const uint4 x0 = __ldcs(g + v), x1 = __ldcs(g + v + kCompactThreads), x2 = __ldcs(g + v + 2 * kCompactThreads),
                            x3 = __ldcs(g + v + 3 * kCompactThreads);
                d[v] = x0;
                d[v + kCompactThreads] = x1;
                d[v + 2 * kCompactThreads] = x2;
                d[v + 3 * kCompactThreads] = x3;
            }
            for (; v < nvec; v += kCompactThreads) d[v] = __ldcs(g + v);
        }
        if (warp == 0) {
            const uint64_t prefix = lookback_resolve(a.tile_state, tile, s_total);
            if (lane == 0) s_base = prefix;
        }
        __syncthreads();
        const uint64_t tile_base = s_base;
        const uint64_t total = s_total;
        const uint64_t e0 = s_warp_sum[warp] + (incl - mine);  // exclusive, within the tile
        const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(b.out + tile_base) & 15);  // image byte 0 = that 16-byte boundary
        const bool fits = staged && (uint64_t)lead + total + 32 <= a.image_bytes;
        if (2 * tid < nitems) {
            b.out_offsets[item0 + 2 * tid] = tile_base + e0;
            if (b.out_lens) b.out_lens[item0 + 2 * tid] = c0;
        }
        if (2 * tid + 1 < nitems) {
            b.out_offsets[item0 + 2 * tid + 1] = tile_base + e0 + c0;
            if (b.out_lens) b.out_lens[item0 + 2 * tid + 1] = c1;
        }
        if (tid == 0 && item0 + nitems == b.n) b.out_offsets[b.n] = tile_base + total;
        s_off[2 * tid] = (uint32_t)e0;
        s_off[2 * tid + 1] = (uint32_t)(e0 + c0);
        s_src[2 * tid] = (uint32_t)(p0 - r0);
        s_src[2 * tid + 1] = (uint32_t)(p1 - r0);
        s_cnt[2 * tid] = c0;
        s_cnt[2 * tid + 1] = c1;
        __syncthreads();
        if (fits) {
            // rows -> image: thread t moves strings t and t + 256 (neighbouring threads, neighbouring rows)
            for (uint32_t it = tid; it < nitems; it += kCompactThreads)
                smem_copy_row(s_stage + s_src[it], s_image + lead + s_off[it], s_cnt[it]);
            __syncthreads();
            // image -> output, clipped to the capacity
            const uint64_t room = tile_base < b.out_capacity ? b.out_capacity - tile_base : 0;
            const uint32_t nout = (uint32_t)min(total, room);
            uint8_t *const dst = b.out + tile_base;
            const uint32_t head = min(nout, (16u - lead) & 15u);
            const uint32_t nvec = (nout - head) >> 4;
            const uint4 *sv = reinterpret_cast<const uint4 *>(s_image + lead + head);  // 16-byte aligned
            uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
            for (uint32_t v = tid; v < nvec; v += kCompactThreads) __stcs(dv + v, sv[v]);
            if (tid < head) dst[tid] = s_image[lead + tid];
            const uint32_t tail0 = head + 16u * nvec;
            if (tid >= 32 && tid - 32 < nout - tail0) dst[tail0 + tid - 32] = s_image[lead + tail0 + tid - 32];
        } else {
            // a tile that does not fit (very long strings): rows go straight to their place
            for (uint32_t it = tid; it < nitems; it += kCompactThreads) {
                const uint64_t off = b.out_offsets[item0 + it];  // (written above, before the barrier)
                const uint64_t n = a.cnt[item0 + it];
                const uint8_t *src = a.rows + a.row_pos[item0 + it];
                for (uint64_t k = 0; k < n && off + k < b.out_capacity; ++k) b.out[off + k] = src[k];
            }
        }
    }
}

}  // namespace hb
