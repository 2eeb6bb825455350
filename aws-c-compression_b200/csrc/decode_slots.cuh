// Batch decoder variant of round 2: decode_batch_kernel (decode_fast.cuh) with the stage and the rows in ONE place.
// OPT-IN (AWS_HUFFMAN_BATCH_SLOTS_DECODE=1 at context creation): correct (the whole GPU suite passes on it) but
// MEASURED SLOWER than decode_batch_kernel — 0.362 against 0.305 ms on the 1M-string batch — and kept as the record.
//
// The idea. What bounds decode_batch_kernel is the shared memory a string in flight takes — its staged bytes plus a
// row at worst-case spacing, 260 B on the benchmark — because that sets the strings per tile (288 = 9 groups of 32 for
// 8 warps), and a tile's decode phase lasts as long as its LONGEST group: with 9 groups of sorted strings of 8..256 B
// the eight warps are busy 63 % of that phase (ncu: 20 % of all stall samples sit behind the barrier that ends it).
// More groups per tile let the warps that finish early pull more work (longest-first pulling is LPT scheduling:
// 12 groups -> 81 %, 14 -> 93 %).
//
// Here a string owns ONE slot of shared memory, as large as its row can get (8 len / min_len bytes + slack): its
// encoded bytes are staged at the END of the slot and its symbols are written from the START. A code is at least
// min_len bits, so after c encoded bytes the row holds at most 8 c / min_len bytes while the unread input starts at
// (8 / min_len - 1) len + c + 1, and the decoder's cursor holds the next three words in registers: the writer never
// reaches a word the reader still needs. That is 37 % less shared memory per string: 384 strings per tile in the
// space of 288. With no separate stage there is no room for the dense image of the tile's output; every thread
// moves its rows straight to global memory (128-bit stores in the middle, one 1/2/4/8-byte store each at the ends).
//
// What the measurement says (profiles/README.md). The kernel executes FEWER instructions than decode_batch_kernel
// (122 M warp instructions against 145 M: no row -> image pass, no image -> global pass) and balances better, but
// issues them at 31 % instead of 42 %: (1) staging is per string — the slots are apart, so it cannot be the flat
// copy of the tile's bytes with every load of the block in flight at once; a thread has two to four loads in
// flight and 11 % of the stall samples wait for them; (2) the scout warp's look-back, which decode_batch_kernel
// hides behind its row -> image pass, is exposed now (6 %): the rows cannot leave before the tile's position is
// known, and there is nothing else to do meanwhile; (3) the copy-out to global memory sits on the tile's critical
// path. The first version with byte-wise edges (15 dependent byte copies per end) took 0.471 ms.
//
// Everything else — string table, counting sort, the lean decode step, group pulling, block scan, look-back by the
// scout warp, the two-pass route for tiles that do not fit — is decode_batch_kernel's, which also stays the framed
// (HPACK literal) decoder.
#pragma once

#include "decode_fast.cuh"

namespace hb {

constexpr int kSlotItemsPerTile = 512;    // strings per tile at most (two per thread in the block scan)
constexpr uint32_t kSlotSlack = 32;       // bytes per slot beyond the worst-case row (16-byte phase of the staged bytes + 1)

struct DecSlotsArgs {
    BatchView b;
    const uint2 *lut2;
    uint32_t lut2_count, lut2_trap, root_bits;
    uint32_t min_len;        // min(shortest code, 8): a string of L bytes decodes to at most 8 L / min_len symbols (>= L)
    uint32_t region_bytes;   // shared memory of one team's slots
    uint64_t *tile_state;
    uint32_t *ticket;
    uint32_t num_tiles;
    uint32_t items_per_tile;  // <= kSlotItemsPerTile, a multiple of 32
};

struct DecSlotsTeam {
    uint32_t start[kSlotItemsPerTile];  // first bit of the staged string in the team's region
    uint32_t bytes[kSlotItemsPerTile];  // encoded length
    uint32_t cnt[kSlotItemsPerTile];    // symbols per string
    uint32_t off[kSlotItemsPerTile];    // exclusive offsets within the tile
    uint32_t row[kSlotItemsPerTile];    // start of the string's slot (= of its row)
    uint16_t perm[kSlotItemsPerTile];   // strings in order of decreasing length
    uint32_t hist[256];
    uint64_t warp_sum[kDecWarps];
    uint64_t total;
    uint32_t tile, next, fits;
    uint32_t hand_tile[2];     // workers -> scout, by hand-off parity
    uint64_t hand_prefix[2];   // scout -> workers
};

__device__ __forceinline__ void sts_u32_generic(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// start of the slot of the string that begins `rel` bytes into the tile's input and is its idx-th string
__device__ __forceinline__ uint32_t slot_start(uint32_t rel, uint32_t idx, uint32_t min_len) {
    return 16u * (rel / (2u * min_len)) + kSlotSlack * idx;
}

// 32 bits of a row in shared memory from any byte offset (two aligned words and a funnel shift)
__device__ __forceinline__ uint32_t row_u32(const uint8_t *srow, uint32_t o) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(srow) + (o >> 2);
    return __funnelshift_r(w[0], w[1], 8u * (o & 3u));
}

// Thread-serial: n bytes of a row in shared memory (16-byte aligned) to any address in global memory. Up to the first
// 16-byte boundary of the destination and behind the last one the bytes leave as one 1-, 2-, 4- and 8-byte store each
// (naturally aligned, in that order resp. the reverse: at most four stores per end instead of fifteen dependent
// byte copies); in between, 128-bit stores (four shared-memory words and four byte permutes each).
__device__ __forceinline__ void copy_row_to_global(const uint8_t *srow, uint8_t *dst, uint32_t n) {
    if (n == 0) return;
    const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
    if (n < head + 16u) {  // no whole vector: (up to 30) bytes
        for (uint32_t i = 0; i < n; ++i) dst[i] = srow[i];
        return;
    }
    uint32_t o = 0;
    if (head & 1u) {
        dst[0] = srow[0];
        o = 1;
    }
    if (head & 2u) {
        *reinterpret_cast<uint16_t *>(dst + o) = (uint16_t)row_u32(srow, o);
        o += 2;
    }
    if (head & 4u) {
        *reinterpret_cast<uint32_t *>(dst + o) = row_u32(srow, o);
        o += 4;
    }
    if (head & 8u) {
        *reinterpret_cast<uint2 *>(dst + o) = make_uint2(row_u32(srow, o), row_u32(srow, o + 4));
        o += 8;
    }
    const uint32_t nvec = (n - head) >> 4;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(srow + (head & ~3u));
    const uint32_t sel = 0x3210u + 0x1111u * (head & 3u);  // bytes (head & 3) .. + 3 of a word pair
    uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
    uint32_t lo = sw[0];
    for (uint32_t v = 0; v < nvec; ++v) {
        const uint32_t h0 = sw[4 * v + 1], h1 = sw[4 * v + 2], h2 = sw[4 * v + 3], h3 = sw[4 * v + 4];
        dv[v] = make_uint4(__byte_perm(lo, h0, sel), __byte_perm(h0, h1, sel), __byte_perm(h1, h2, sel), __byte_perm(h2, h3, sel));
        lo = h3;
    }
    o = head + 16u * nvec;
    const uint32_t rem = n - o;  // < 16; dst + o is 16-byte aligned
    if (rem & 8u) {
        *reinterpret_cast<uint2 *>(dst + o) = make_uint2(row_u32(srow, o), row_u32(srow, o + 4));
        o += 8;
    }
    if (rem & 4u) {
        *reinterpret_cast<uint32_t *>(dst + o) = row_u32(srow, o);
        o += 4;
    }
    if (rem & 2u) {
        *reinterpret_cast<uint16_t *>(dst + o) = (uint16_t)row_u32(srow, o);
        o += 2;
    }
    if (rem & 1u) dst[o] = srow[o];
}

__global__ void __launch_bounds__(kDecTeams * kDecBlock, 1) decode_slots_kernel(DecSlotsArgs a) {
    extern __shared__ __align__(128) uint32_t s_lut[];  // [LUT2][team 0: slots][team 1: slots]
    __shared__ DecSlotsTeam s_teams[kDecTeams];
    const uint32_t team = threadIdx.x / kDecBlock, tid = threadIdx.x - team * kDecBlock;
    DecSlotsTeam &sh = s_teams[team];
    const uint32_t lut_pad = 2u * a.lut2_count;  // words
    uint32_t *s_in = s_lut + lut_pad + team * (a.region_bytes / 4);  // the team's region: big-endian words of staged input
    uint8_t *const s_region = reinterpret_cast<uint8_t *>(s_in);    // ... and little-endian rows, slot by slot
    const Lut2 lut2 = lut2_load(reinterpret_cast<uint2 *>(s_lut), a.lut2, a.lut2_count, a.root_bits, a.lut2_trap);
    const uint32_t lane = lane_id(), warp = tid >> 5;
    const BatchView &b = a.b;
    __syncthreads();  // the LUT is in place
    // ================================ scout (see decode_batch_kernel) ==========================================
    if (warp == kDecWarps) {
        for (uint32_t h = 0;; ++h) {
            dec_bar_sync(5 * team + 2 + (h & 1));  // the workers took a tile
            const uint32_t t = sh.hand_tile[h & 1];
            if (t == kDecDone) return;
            const uint64_t prefix = lookback_exclusive(a.tile_state, t);
            if (lane == 0) sh.hand_prefix[h & 1] = prefix;
            dec_bar_arrive(5 * team + 4 + (h & 1));  // result ready
        }
    }
    // ================================ workers ===================================================================
    uint32_t hand = 0;
    const uint32_t region_addr = (uint32_t)__cvta_generic_to_shared(s_region);
    for (;;) {
        dec_worker_sync(team);  // previous tile fully done
        if (tid == 0) {
            sh.tile = atomicAdd(a.ticket, 1u);
            sh.next = kDecPullWarps;
            sh.fits = 1;
        }
        for (uint32_t i = tid; i < 256; i += kDecThreads) sh.hist[i] = 0;
        dec_worker_sync(team);
        const uint32_t tile = sh.tile;
        if (tile >= a.num_tiles) break;
        if (tid == 0) sh.hand_tile[hand & 1] = tile;
        dec_bar_arrive(5 * team + 2 + (hand & 1));  // the scout starts on the tile's position
        ++hand;
        const uint64_t item0 = (uint64_t)tile * a.items_per_tile;
        const uint32_t nitems = (uint32_t)min((uint64_t)a.items_per_tile, b.n - item0);
        const uint32_t ngroups = (nitems + 31) / 32;

        // ---- string table: the slot of every string, where its bytes are staged in it; length histogram ----------
        const uint64_t byte0 = b.in_offsets[item0], byte1 = b.in_offsets[item0 + nitems];
        const bool small = byte1 - byte0 < (1ull << 27);
        const bool fits = small && (uint64_t)slot_start((uint32_t)(byte1 - byte0), nitems, a.min_len) + 16 <= a.region_bytes;
        for (uint32_t it = tid; it < nitems; it += kDecThreads) {
            const uint64_t in0 = b.in_offsets[item0 + it];
            const uint64_t len = b.in_offsets[item0 + it + 1] - in0;
            sh.bytes[it] = (uint32_t)min(len, (uint64_t)0xffffffffu);
            if (fits) {
                const uint32_t rel = (uint32_t)(in0 - byte0);
                const uint32_t s0 = slot_start(rel, it, a.min_len), s1 = slot_start(rel + (uint32_t)len, it + 1, a.min_len);
                // the staged bytes end as late as the slot allows with the 16-byte phase they have in global memory
                const uint32_t g = (uint32_t)(reinterpret_cast<uintptr_t>(b.in + in0) & 15);
                const uint32_t d = s1 - (uint32_t)len - ((s1 - (uint32_t)len - g) & 15u);
                sh.row[it] = s0;
                sh.start[it] = 8u * d;
                if (8 * len / a.min_len + kDecRowSlack > kDecMaxRow) sh.fits = 0;
            }
            atomicAdd(&sh.hist[255 - (uint32_t)min(len, (uint64_t)255)], 1u);
        }
        dec_worker_sync(team);
        const bool staged = fits && sh.fits != 0;
        if (warp == 0) {
            // exclusive scan of the 256 bins, 8 per lane
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] = sh.hist[lane * 8 + i];
                sum += v[i];
            }
            uint32_t run = warp_inclusive_scan(sum) - sum;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sh.hist[lane * 8 + i] = run;
                run += v[i];
            }
        }
        dec_worker_sync(team);
        for (uint32_t it = tid; it < nitems; it += kDecThreads) {
            const uint32_t rank = atomicAdd(&sh.hist[255 - min(sh.bytes[it], 255u)], 1u);
            sh.perm[rank] = (uint16_t)it;
        }
        dec_worker_sync(team);
        if (staged) {
            // ---- stage: every thread copies strings of its rank (similar lengths in a warp) into their slots, as
            // big-endian words: stream byte k of the region lives at byte k ^ 3 -----------------------------------
            const uint8_t *const in_end = b.in + b.in_offsets[b.n];
            for (uint32_t slot = tid; slot < nitems; slot += kDecThreads) {
                const uint32_t it = sh.perm[slot];
                const uint32_t len = sh.bytes[it];
                if (len == 0) continue;
                const uint8_t *g = b.in + b.in_offsets[item0 + it];
                const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15);
                const uint8_t *const gal = g - lead;                    // the string's first vector in global memory
                const uint32_t d0 = (sh.start[it] >> 3) - lead;         // ... and where it goes in the region
                const uint32_t span = lead + len, nvec = (span + 15u) >> 4;
                // vector j as four big-endian words; the two vectors at the ends of the INPUT BUFFER are read byte by byte
                auto load = [&](uint32_t j) -> uint4 {
                    const uint8_t *p = gal + 16u * j;
                    uint4 x;
                    if (p >= b.in && p + 16 <= in_end) {
                        x = __ldg(reinterpret_cast<const uint4 *>(p));
                    } else {
                        uint32_t w[4] = {0, 0, 0, 0};
                        for (int k = 0; k < 16; ++k)
                            if (p + k >= b.in && p + k < in_end) w[k >> 2] |= (uint32_t)p[k] << (8 * (k & 3));
                        x = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    x.x = __byte_perm(x.x, 0, 0x0123);
                    x.y = __byte_perm(x.y, 0, 0x0123);
                    x.z = __byte_perm(x.z, 0, 0x0123);
                    x.w = __byte_perm(x.w, 0, 0x0123);
                    return x;
                };
                // bytes [lo, hi) of vector j: whole words as words, the others byte by byte (a neighbour's bytes may
                // share the word)
                auto store_part = [&](uint32_t j, const uint4 &x, uint32_t lo, uint32_t hi) {
                    const uint32_t base = region_addr + d0 + 16u * j;
                    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (lo <= 4u * q && 4u * q + 4u <= hi) {
                            sts_u32_generic(base + 4u * q, w[q]);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (lo <= 4u * q + k && 4u * q + k < hi) sts_u8(base + 4u * q + (3u - k), w[q] >> (24 - 8 * k));
                        }
                    }
                };
                const uint4 first = load(0);
                const uint4 last = load(nvec - 1);
                store_part(0, first, lead, min(16u, span));
                uint4 *sv = reinterpret_cast<uint4 *>(s_region + d0);
                uint32_t j = 1;
                for (; j + 2 < nvec; j += 2) {
                    const uint4 x = load(j), y = load(j + 1);
                    sv[j] = x;
                    sv[j + 1] = y;
                }
                if (j + 1 < nvec) sv[j] = load(j);
                if (nvec > 1) store_part(nvec - 1, last, 0u, span - 16u * (nvec - 1));
                // (the decoder's cursor reads up to two words past the string: whatever lies there is never used)
            }
            dec_worker_sync(team);
        }

        // ---- decode (staged: once, in place; otherwise: count) ------------------------------------------------------
        for (uint32_t g = warp < (uint32_t)kDecPullWarps ? warp : ngroups; g < ngroups;) {
            const uint32_t slot = g * 32 + lane;
            if (slot < nitems) {
                const uint32_t it = sh.perm[slot];
                const uint64_t item = item0 + it;
                const uint64_t in0 = b.in_offsets[item];
                const uint64_t len = b.in_offsets[item + 1] - in0;
                const uint8_t *payload = b.in + in0;
                uint64_t cbits = 0;
                uint32_t nsym = 0, term = kTermEnd;
                bool redo = !staged;
                if (staged) {
                    const uint32_t ib = sh.start[it], ie = ib + (uint32_t)len * 8;
                    const SpanS r = decode_span_lean<true, false, false, true>(s_in, lut2, a.root_bits, ib, ie, ie, region_addr + sh.row[it]);
                    cbits = r.pos - ib;
                    nsym = r.nsym;
                    term = r.term;
                    redo = term == kTermTrapped;
                }
                if (redo) {
                    // not staged: count now, write after the scan. Staged and a window matched no code: the slot's
                    // input is partly overwritten, so the string is decoded again from global memory into its row
                    // (it ends at the unknown symbol: the row cannot overflow)
                    ByteWriter wr;
                    if (staged) wr.init(s_region + sh.row[it], ~0ull);
                    const DecodeSpan r = staged ? decode_span_lut2<true>(lut2, payload, 0, len * 8, len, &wr)
                                                : decode_span_lut2<false>(lut2, payload, 0, len * 8, len, nullptr);
                    if (staged) wr.finish();
                    cbits = r.pos;
                    nsym = (uint32_t)r.nsym;
                    term = r.term;
                }
                sh.cnt[it] = nsym;
                if (b.status) b.status[item] = term == kTermUnknown ? kStatusUnknownSymbol : kStatusOk;
                if (b.consumed || b.leftover_working_bits || b.leftover_num_bits)
                    leftover_state(
                        payload, len, cbits, term == kTermUnknown, b.consumed ? b.consumed + item : nullptr,
                        b.leftover_working_bits ? b.leftover_working_bits + item : nullptr,
                        b.leftover_num_bits ? b.leftover_num_bits + item : nullptr);
            }
            uint32_t next = 0;
            if (lane == 0) next = atomicAdd(&sh.next, 1u);
            g = __shfl_sync(0xffffffffu, next, 0);
        }
        dec_worker_sync(team);

        // ---- offsets: block scan (2 strings per thread); the tile's count goes out at once ------------------------
        const uint32_t c0 = 2 * tid < nitems ? sh.cnt[2 * tid] : 0u;
        const uint32_t c1 = 2 * tid + 1 < nitems ? sh.cnt[2 * tid + 1] : 0u;
        const uint64_t mine = (uint64_t)c0 + c1;
        const uint64_t incl = warp_inclusive_scan64(mine);
        if (lane == 31) sh.warp_sum[warp] = incl;
        dec_worker_sync(team);
        if (warp == 0) {
            uint64_t w = lane < kDecWarps ? sh.warp_sum[lane] : 0;
            const uint64_t wi = warp_inclusive_scan64(w);
            if (lane < kDecWarps) sh.warp_sum[lane] = wi - w;
            const uint64_t total = __shfl_sync(0xffffffffu, wi, kDecWarps - 1);
            if (lane == 0) {
                lookback_publish_aggregate(a.tile_state, tile, total);
                sh.total = total;
            }
        }
        dec_worker_sync(team);
        {
            const uint64_t e0 = sh.warp_sum[warp] + (incl - mine);  // exclusive, within the tile
            if (2 * tid < nitems) {
                sh.off[2 * tid] = (uint32_t)e0;
                if (b.out_lens) b.out_lens[item0 + 2 * tid] = c0;
            }
            if (2 * tid + 1 < nitems) {
                sh.off[2 * tid + 1] = (uint32_t)(e0 + c0);
                if (b.out_lens) b.out_lens[item0 + 2 * tid + 1] = c1;
            }
            if (tid == 0) sh.next = kDecWarps;
        }
        // ---- where the tile goes: the scout has been summing its predecessors since the tile was taken --------------
        dec_bar_sync(5 * team + 4 + ((hand - 1) & 1));  // (also publishes sh.off to the team)
        const uint64_t total = sh.total;
        const uint64_t tile_base = sh.hand_prefix[(hand - 1) & 1];
        if (tid == 0 && tile > 0)
            st_relaxed_u64(&a.tile_state[tile], (kLbPrefix << kLbFlagShift) | ((tile_base + total) & kLbValueMask));
        for (uint32_t it = tid; it < nitems; it += kDecThreads) b.out_offsets[item0 + it] = tile_base + sh.off[it];
        if (tid == 0 && item0 + nitems == b.n) b.out_offsets[b.n] = tile_base + total;
        if (staged) {
            // ---- rows -> global memory, strings in order of length ----------------------------------------------------
            for (uint32_t slot = tid; slot < nitems; slot += kDecThreads) {
                const uint32_t it = sh.perm[slot];
                const uint64_t off = tile_base + sh.off[it];
                const uint64_t room = off < b.out_capacity ? b.out_capacity - off : 0;
                copy_row_to_global(s_region + sh.row[it], b.out + off, (uint32_t)min((uint64_t)sh.cnt[it], room));
            }
        } else {
            // ---- write: decode again from global memory, now storing --------------------------------------------------
            for (uint32_t g = warp; g < ngroups;) {
                const uint32_t gslot = g * 32 + lane;
                if (gslot < nitems) {
                    const uint32_t it = sh.perm[gslot];
                    const uint64_t off = tile_base + sh.off[it];
                    const uint64_t room = off < b.out_capacity ? b.out_capacity - off : 0;
                    const uint64_t in0 = b.in_offsets[item0 + it];
                    const uint64_t len = b.in_offsets[item0 + it + 1] - in0;
                    ByteWriter wr;
                    wr.init(b.out + off, room);
                    decode_span_lut2<true>(lut2, b.in + in0, 0, len * 8, len, &wr);
                    wr.finish();
                }
                uint32_t next = 0;
                if (lane == 0) next = atomicAdd(&sh.next, 1u);
                g = __shfl_sync(0xffffffffu, next, 0);
            }
        }
    }
    if (tid == 0) sh.hand_tile[hand & 1] = kDecDone;
    dec_bar_arrive(5 * team + 2 + (hand & 1));
}

}  // namespace hb
