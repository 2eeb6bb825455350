// Batch encoder with ITEM-ALIGNED thread ranges ("slots"), for batches of many strings.
//
// encode_tiled_kernel<true> cuts the batch into fixed 32-byte thread ranges, so item starts fall
// inside them and every symbol of the packing loop pays for the test "does an item start here?" (2.3x
// the instructions of the single-stream kernel). Here a thread range never contains an item start:
// item i owns ceil(len_i / 32) SLOTS of up to 32 symbols, a tile is 256 consecutive slots, thread t of a
// tile encodes slot t. What changes with respect to encode_tiled.cuh (everything else — look-back by
// the scout warp, direct-to-stage packing, carry-flag flush, two-piece stage, copy-out — is shared):
//   * a pre-pass turns item lengths into slot counts and scans them (slot_base); tile_first[j] = first
//     item whose first slot is >= 256 j, stored with the tile's byte range in one 32-byte record per tile
//   * per tile, the threads that load the items' offsets scatter {first byte, length} to the item's first
//     slot and set a bit per slot where an item starts; thread t finds its item by looking for the last
//     such bit at or before t
//   * the tile's bytes are contiguous in the input: they are fetched with cp.async as 16-byte chunks and
//     each thread extracts its 32 symbols from three aligned 128-bit reads
//   * a thread's segment function is "p -> ceil8(p) + bits" (first slot of an item) or "p -> p + bits";
//     the last slot of an item appends the EOS padding itself; symbols past the end of a partial slot
//     have length 0, and an append of length 0 is a no-op
// Items of zero length own no slot; their output offset is that of the slot where the next item starts.
#pragma once

#include "encode_tiled.cuh"

namespace hb {

constexpr int kSlotsPerTile = kEncThreads;                            // 256
constexpr int kSlotSymBytes = kSlotsPerTile * kEncSymsPerThread + 128; // staged symbols: 15 B of lead-in, 48-byte reads

struct EncSlotTile;
struct EncSlotArgs {
    const uint8_t *in;           // 16-byte aligned
    const uint64_t *in_offsets;  // n + 1
    uint64_t n;
    uint64_t total_in;
    uint8_t *out;
    uint64_t out_capacity;
    uint64_t *out_offsets;
    const uint64_t *slot_base;   // n + 1: first slot of item i; slot_base[n] = number of slots
    const EncSlotTile *tiles;    // num_tiles_ub + 1 records
    uint64_t *tile_state;
    uint32_t *ticket;
    uint32_t num_tiles_ub;
    uint32_t eos_padding;
};

// Everything the main kernel needs to start on tile j, in one 32-byte record (one load instead of three
// dependent round trips): the items whose first slot lies in the tile, the tile's first and last input
// byte, and what is left of the item that continues from the previous tile.
struct EncSlotTile {
    uint32_t tf0, tf1;    // items [tf0, tf1) start in the tile (tf0 = first item whose first slot is >= 256 j)
    uint64_t first_byte;  // input byte of the tile's first symbol
    uint64_t end_byte;    // input byte after its last symbol
    uint32_t lead_len;    // bytes left of the item that continues from the previous tile (0: none)
    uint32_t nvalid;      // slots in the tile (0: past the end)
};
static_assert(sizeof(EncSlotTile) == 32, "one 32-byte record per tile");

__device__ __forceinline__ uint64_t slot_first_item(const uint64_t *slot_base, uint64_t n, uint64_t target) {
    uint64_t lo = 0, hi = n;  // first item i in [0, n) with slot_base[i] >= target, else n
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (slot_base[mid] < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void slot_tile_index_kernel(
    const uint64_t *in_offsets, const uint64_t *slot_base, uint64_t n, uint64_t total_in, uint64_t count, EncSlotTile *tiles) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const uint64_t num_slots = slot_base[n];
    const uint64_t s0 = j * (uint64_t)kSlotsPerTile, s1 = min(s0 + kSlotsPerTile, num_slots);
    EncSlotTile t;
    t.tf0 = (uint32_t)slot_first_item(slot_base, n, s0);
    t.tf1 = (uint32_t)slot_first_item(slot_base, n, s0 + kSlotsPerTile);
    t.first_byte = t.end_byte = total_in;
    t.lead_len = 0;
    t.nvalid = s0 < num_slots ? (uint32_t)(s1 - s0) : 0u;
    if (t.nvalid) {
        // byte of slot s: inside the item that owns it; an item of zero length has the offset of its successor
        const uint64_t sbA = slot_base[t.tf0];
        if (t.tf0 > 0 && sbA > s0) {
            t.first_byte = in_offsets[t.tf0 - 1] + (s0 - slot_base[t.tf0 - 1]) * kEncSymsPerThread;
            t.lead_len = (uint32_t)(in_offsets[t.tf0] - t.first_byte);
        } else {
            t.first_byte = in_offsets[t.tf0];
        }
        if (s1 == num_slots) t.end_byte = total_in;
        else if (slot_base[t.tf1] == s1) t.end_byte = in_offsets[t.tf1];
        else t.end_byte = in_offsets[t.tf1 - 1] + (s1 - slot_base[t.tf1 - 1]) * kEncSymsPerThread;
    }
    tiles[j] = t;
}

struct EncSlotHandoff {  // workers -> scout, double buffered by iteration parity
    uint32_t tile;       // kEncDone: no more tiles
    Seg total;
    uint64_t next_byte;  // input byte after the tile's last symbol
    uint64_t open_end;   // where the item that is open at the end of the tile ends (== next_byte: it ends there)
};

// Everything a tile needs to know about its slots; built one tile ahead (after the packing loop).
struct EncSlotMap {
    uint32_t startmask[kSlotsPerTile / 32];  // bit t: an item starts at slot t
    uint32_t first_off[kSlotsPerTile];       // start slots: first byte of the item, relative to `org`
    uint32_t first_len[kSlotsPerTile];       // start slots: length of the item
    uint32_t lead_off, lead_len;             // the item that continues from the previous tile: its next byte, bytes left
    uint32_t nvalid;                         // slots in the tile
    uint64_t org;                            // input byte staged at offset 0 (16-byte aligned)
};

constexpr size_t kEncSlotSmemBytes =
    kEncStageWords * 4 + (kSlotsPerTile + 4) * 2 /* item start positions */ + 2048 * kEncTabCopies /* code table */ + kSlotSymBytes;

__global__ void __launch_bounds__(kEncBlock, 2) encode_slots_kernel(const uint2 *__restrict__ enc_table, EncSlotArgs a) {
    __shared__ Seg s_wseg[kEncWarps];
    __shared__ uint32_t s_tail[kEncWarps], s_brk[kEncWarps], s_wpos[kEncWarps + 1];
    __shared__ uint32_t s_next;
    __shared__ EncSlotHandoff s_hand[2];
    __shared__ EncResult s_res[2];
    __shared__ EncSlotMap s_map;
    __shared__ struct {
        uint32_t tile, tf0, tf1, nvalid;
        Seg total;
    } s_prev;
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint32_t *const stage = reinterpret_cast<uint32_t *>(s_dyn);
    uint16_t *const obpos = reinterpret_cast<uint16_t *>(s_dyn + kEncStageWords * 4);  // [257]
    const uint32_t tab0 = smem_addr(s_dyn + kEncStageWords * 4 + (kSlotsPerTile + 4) * 2 + 8) & ~15u;
    const uint32_t stage_addr = smem_addr(stage);
    const uint32_t sym_addr = tab0 + 2048 * kEncTabCopies;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t tab = tab0 + (lane & (kEncTabCopies - 1)) * 8;  // my copy of the table
    const uint64_t num_slots = a.slot_base[a.n];

    if (tid == 0) s_next = atomicAdd(a.ticket, 1u);
    if (tid < 256) {
        const uint2 e = enc_table[tid];
#pragma unroll
        for (int j = 0; j < kEncTabCopies; ++j)
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(tab0 + (tid * kEncTabCopies + j) * 8), "r"(e.x),
                         "r"(enc_len_fields(e.y))
                         : "memory");
    }
    if (tid < kSlotsPerTile / 32) s_map.startmask[tid] = 0;
    __syncthreads();

    // ================================ scout =========================================================================
    if (warp == kEncWarps) {
        for (uint32_t it = 0;; ++it) {
            enc_bar_sync(2 + (it & 1));  // the workers handed over a tile
            const uint32_t tile = s_hand[it & 1].tile;
            if (tile == kEncDone) return;
            const Seg total = s_hand[it & 1].total;
            const uint64_t next_byte = s_hand[it & 1].next_byte, open_end = s_hand[it & 1].open_end;
            // what may complete the tile's last byte: the next symbols of the item that is open at its end
            uint32_t fb_lo = 0, fb_hi = 0;
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (next_byte + q < open_end) {
                        const uint32_t byte = a.in[next_byte + q];
                        if (q < 4) fb_lo |= byte << (8 * q); else fb_hi |= byte << (8 * (q - 4));
                    }
                }
            }
            const uint64_t G0 = seg_resolve<true>(a.tile_state, tile, total);
            if (lane == 0) {
                const uint64_t Gend = seg_apply(total, G0);
                const uint32_t need = (8u - (uint32_t)(Gend & 7u)) & 7u;
                uint32_t bits = 0;
                if (need) {
                    uint32_t have = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (have < need && next_byte + q < open_end) {
                            const uint2 e = enc_lookup(tab, q < 4 ? fb_lo : fb_hi, q);
                            const uint32_t len = e.y & 63u;
                            const uint32_t take = min(len, need - have);
                            bits = (bits << take) | (e.x >> (len - take));
                            have += take;
                        }
                    }
                    if (have < need) {
                        const uint32_t rem = need - have;
                        bits = (bits << rem) | (a.eos_padding & ((1u << rem) - 1u));
                    }
                }
                s_res[it & 1].G = G0;
                s_res[it & 1].fill = bits;
            }
            enc_bar_arrive(4 + (it & 1));  // result ready
        }
    }

    // ================================ workers =======================================================================
    // Preparing a tile = its record (one uniform 32-byte load), then — in flight together — the cp.async of
    // its bytes and the offsets of "my" item, then the scatter of the items to their first slots. The three
    // steps sit at different places of the previous tile's packing loop: a warp issues in order, so a load
    // and its first use must be far apart or the warp stalls for the whole latency.
    EncSlotTile p_t = {};                  // record of the tile being prepared
    uint64_t p_sb = 0, p_o0 = 0, p_o1 = 0; // "my" item of that tile
    uint64_t p_org = 0;

    auto prepare_record = [&](uint32_t t) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.tiles + t);
        const uint4 lo = __ldg(src), hi = __ldg(src + 1);
        p_t.tf0 = lo.x;
        p_t.tf1 = lo.y;
        p_t.first_byte = (uint64_t)lo.z | ((uint64_t)lo.w << 32);
        p_t.end_byte = (uint64_t)hi.x | ((uint64_t)hi.y << 32);
        p_t.lead_len = hi.z;
        p_t.nvalid = hi.w;
    };
    auto prepare_fetch = [&](uint32_t t) {
        (void)t;
        p_org = p_t.first_byte & ~15ull;
        const uint32_t nchunks = (uint32_t)((p_t.end_byte - p_org + 15) >> 4);
        for (uint32_t c = tid; c < nchunks; c += kEncThreads) {
            const uint64_t src = p_org + 16ull * c;
            if (src + 16 <= a.total_in) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sym_addr + 16 * c), "l"(a.in + src) : "memory");
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t v = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (src + 4 * q + b < a.total_in) v |= (uint32_t)a.in[src + 4 * q + b] << (8 * b);
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sym_addr + 16 * c + 4 * q), "r"(v) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const uint64_t i = (uint64_t)p_t.tf0 + tid;
        p_sb = p_o0 = p_o1 = 0;
        if (i < p_t.tf1) {
            p_sb = a.slot_base[i];
            p_o0 = a.in_offsets[i];
            p_o1 = a.in_offsets[i + 1];
        }
    };
    // scatter the items to their first slots (s_map.startmask is zero at this point)
    auto prepare_map = [&](uint32_t t) {
        const uint64_t s0 = (uint64_t)t * kSlotsPerTile;
        if ((uint64_t)p_t.tf0 + tid < p_t.tf1 && p_o1 > p_o0) {
            const uint32_t sp = (uint32_t)(p_sb - s0);
            atomicOr(&s_map.startmask[sp >> 5], 1u << (sp & 31));
            s_map.first_off[sp] = (uint32_t)(p_o0 - p_org);
            s_map.first_len[sp] = (uint32_t)(p_o1 - p_o0);
        }
        for (uint64_t i = (uint64_t)p_t.tf0 + tid + kEncThreads; i < p_t.tf1; i += kEncThreads) {  // (runs of empty items)
            const uint64_t o0 = a.in_offsets[i], o1 = a.in_offsets[i + 1];
            if (o1 > o0) {
                const uint32_t sp = (uint32_t)(a.slot_base[i] - s0);
                atomicOr(&s_map.startmask[sp >> 5], 1u << (sp & 31));
                s_map.first_off[sp] = (uint32_t)(o0 - p_org);
                s_map.first_len[sp] = (uint32_t)(o1 - o0);
            }
        }
        if (tid == 0) {
            s_map.lead_off = (uint32_t)(p_t.first_byte - p_org);
            s_map.lead_len = p_t.lead_len;
            s_map.nvalid = p_t.nvalid;
            s_map.org = p_org;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    };

    uint32_t tile = s_next;
    uint32_t tf0 = 0, tf1 = 0, nvalid = 0;
    if ((uint64_t)tile * kSlotsPerTile < num_slots) {
        prepare_record(tile);
        prepare_fetch(tile);
        prepare_map(tile);
        tf0 = p_t.tf0;
        tf1 = p_t.tf1;
        nvalid = p_t.nvalid;
    }
    enc_worker_sync();

    // the tile that is packed in the stage and waits for its position (its uniform values live in shared
    // memory: registers are what limits this kernel)
    bool prev_valid = false;

    for (uint32_t it = 0;; ++it) {
        const bool cur_valid = (uint64_t)tile * kSlotsPerTile < num_slots;
        uint32_t c[kEncSymsPerThread], l[kEncSymsPerThread];
        Seg mine = {0, 0, 0}, excl = {0, 0, 0}, total = {0, 0, 0};
        bool last_slot = false;
        if (cur_valid) {
            // ---- 0. my slot: item, offset, symbols ---------------------------------------------------------
            uint32_t off = 0, rem = 0;
            bool first_slot = false;
            if (tid < nvalid) {
                const uint32_t w0 = tid >> 5;
                uint32_t m = s_map.startmask[w0] & (0xffffffffu >> (31 - (tid & 31)));
                int sp0 = -1;
                if (m) {
                    sp0 = (int)(32 * w0 + 31 - __clz(m));
                } else {
                    for (int ww = (int)w0 - 1; ww >= 0; --ww) {
                        m = s_map.startmask[ww];
                        if (m) {
                            sp0 = 32 * ww + 31 - __clz(m);
                            break;
                        }
                    }
                }
                if (sp0 >= 0) {
                    const uint32_t j = tid - (uint32_t)sp0;
                    off = s_map.first_off[sp0] + kEncSymsPerThread * j;
                    rem = s_map.first_len[sp0] - kEncSymsPerThread * j;
                    first_slot = j == 0;
                } else {
                    off = s_map.lead_off + kEncSymsPerThread * tid;
                    rem = s_map.lead_len - kEncSymsPerThread * tid;
                }
            }
            const uint32_t nsym = min(rem, (uint32_t)kEncSymsPerThread);
            last_slot = rem <= (uint32_t)kEncSymsPerThread;  // (also true for the idle threads past nvalid)
            if (tid + 1 == nvalid) {
                s_hand[it & 1].next_byte = s_map.org + off + nsym;
                s_hand[it & 1].open_end = s_map.org + off + rem;
            }
            // my 32 symbols: three aligned 128-bit reads, shifted by off % 16 bytes
            uint32_t w[kEncSymsPerThread / 4];
            {
                uint32_t r[12];
                const uint32_t ra = sym_addr + (off & ~15u);
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(r[4 * q]), "=r"(r[4 * q + 1]), "=r"(r[4 * q + 2]), "=r"(r[4 * q + 3])
                                 : "r"(ra + 16 * q)
                                 : "memory");
                const uint32_t ws = (off >> 2) & 3u, bs = 8 * (off & 3u);
                if (ws & 2u) {
#pragma unroll
                    for (int q = 0; q < 10; ++q) r[q] = r[q + 2];
                }
                if (ws & 1u) {
#pragma unroll
                    for (int q = 0; q < 9; ++q) r[q] = r[q + 1];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) w[q] = __funnelshift_r(r[q], r[q + 1], bs);
            }
            // ---- 1. code and length of my symbols (length 0 past the end of a partial slot) -------------------
            uint32_t s = 0;
#pragma unroll
            for (int k = 0; k < kEncSymsPerThread; ++k) {
                const uint2 e = enc_lookup(tab, w[k >> 2], k);
                const bool valid = (uint32_t)k < nsym;
                c[k] = valid ? e.x : 0u;  // (code 0 of length 0: appending it changes nothing)
                l[k] = valid ? e.y : 0u;
                s += l[k];
            }
            s &= kEncLenMask;
            // ---- 2. my segment function, block scan -----------------------------------------------------------
            if (first_slot) mine = Seg{0, s, 1}; else mine = Seg{s, 0, 0};
            Seg incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Seg up = seg_shfl_up(incl, d);
                if (lane >= (uint32_t)d) incl = seg_combine(up, incl);
            }
            excl = seg_shfl_up(incl, 1);
            if (lane == 0) excl = Seg{0, 0, 0};
            if (lane == 31) s_wseg[warp] = incl;
            enc_worker_sync();  // A
            Seg wi = lane < kEncWarps ? s_wseg[lane] : Seg{0, 0, 0};
#pragma unroll
            for (int d = 1; d < kEncWarps; d <<= 1) {
                const Seg up = seg_shfl_up(wi, d);
                if (lane >= (uint32_t)d) wi = seg_combine(up, wi);
            }
            Seg we = seg_shfl_up(wi, 1);
            if (lane == 0) we = Seg{0, 0, 0};
            total = seg_shfl(wi, kEncWarps - 1);
            excl = seg_combine(seg_shfl(we, warp), excl);
            if (tid == 0) {
                seg_publish_aggregate(a.tile_state, tile, total);
                s_hand[it & 1].tile = tile;
                s_hand[it & 1].total = total;
            }
            if (tid < kSlotsPerTile / 32) s_map.startmask[tid] = 0;  // everybody has found its item
        } else {
            if (tid == 0) s_hand[it & 1].tile = kEncDone;
            if (prev_valid) enc_worker_sync();  // the previous tile is completely packed
        }
        if (prev_valid) enc_fix_warp_boundaries(stage, s_tail, s_brk, s_wpos, warp, lane);
        enc_bar_arrive(2 + (it & 1));                       // the scout may take tile `tile`
        if (prev_valid) enc_bar_sync(4 + ((it - 1) & 1));  // G of the previous tile is in s_res[(it - 1) & 1]
        if (cur_valid && tid == 0) s_next = atomicAdd(a.ticket, 1u);  // nothing this block waits for lies ahead

        // ---- copy out the previous tile ------------------------------------------------------------------------
        if (prev_valid) {
            const uint32_t prev_tile = s_prev.tile, prev_tf0 = s_prev.tf0, prev_tf1 = s_prev.tf1, prev_nvalid = s_prev.nvalid;
            const Seg prev_total = s_prev.total;
            const uint64_t G = s_res[(it - 1) & 1].G;
            const uint32_t fill = s_res[(it - 1) & 1].fill;
            const uint64_t Gend = seg_apply(prev_total, G);
            const uint64_t end_byte = (Gend + 7) >> 3;
            const uint32_t H = prev_total.head;
            uint64_t tail_byte = 0;
            if (!prev_total.hb) {
                enc_copy_piece(stage, 0, H, G, fill, a.out, a.out_capacity, tid);
            } else {
                const uint32_t Q = ((H >> 5) + 2u) << 5;
                const uint32_t pad = (uint32_t)(0 - (G + H)) & 7u;
                enc_copy_piece(stage, 0, H, G, a.eos_padding & ((1u << pad) - 1u), a.out, a.out_capacity, tid);
                enc_copy_piece(stage, Q, prev_total.tail, G + H + pad, fill, a.out, a.out_capacity, tid);
                tail_byte = (G + H + pad) >> 3;
            }
            const uint64_t s0 = (uint64_t)prev_tile * kSlotsPerTile;
            for (uint64_t i = (uint64_t)prev_tf0 + tid; i < prev_tf1; i += kEncThreads) {
                const uint32_t sp = (uint32_t)(a.slot_base[i] - s0);
                a.out_offsets[i] = sp < prev_nvalid ? tail_byte + obpos[sp] : end_byte;
            }
            if (s0 + kSlotsPerTile >= num_slots) {
                // the last tile: trailing empty items and the total
                for (uint64_t i = (uint64_t)prev_tf1 + tid; i <= a.n; i += kEncThreads) a.out_offsets[i] = end_byte;
            }
        }
        if (!cur_valid) return;
        enc_worker_sync();  // B: the stage is free again, the next ticket is visible

        const uint32_t next = s_next;
        const bool next_valid = (uint64_t)next * kSlotsPerTile < num_slots;
        if (next_valid) prepare_record(next);

        // ---- 3. pack straight into the stage --------------------------------------------------------------------
        const uint32_t H = total.head;
        const uint32_t Q = ((H >> 5) + 2u) << 5;
        const bool tstar = mine.hb && !excl.hb;  // the tile's first item start: piece 1 begins with me
        const bool in_tail = mine.hb || excl.hb;
        uint32_t pos0;
        if (mine.hb) pos0 = Q + (excl.hb ? (excl.tail + 7u) & ~7u : 0u);
        else pos0 = excl.hb ? Q + excl.tail : excl.head;
        if (mine.hb) obpos[tid] = (uint16_t)((pos0 - Q) >> 3);
        const uint32_t sp_first = stage_addr + (pos0 >> 5) * 4;
        uint32_t sp = sp_first;
        uint32_t acc = 0;
        uint32_t nb = enc_len_fields(pos0 & 31);
#pragma unroll
        for (int k = 0; k < kEncSymsPerThread; ++k) {
            if (k == 20 && next_valid) prepare_fetch(next);
            enc_append(sp, acc, nb, c[k], l[k]);
        }
        {
            // the last slot of an item pads it with the LOW bits of eos_padding (huffman.c:178-184) — unless the
            // item lies in piece 0, whose padding depends on G and is added by the copy, or it is the tile's
            // last slot (the tile's function ends before that padding: it is the scout's `fill`)
            const uint32_t pad = (last_slot && in_tail && tid + 1 < nvalid) ? (0u - nb) & 7u : 0u;
            enc_append(sp, acc, nb, a.eos_padding & ((1u << pad) - 1u), enc_len_fields(pad));
        }
        // The word I share with my predecessor(s); the thread that opens piece 1 puts what came before into
        // the last word of piece 0 instead.
        {
            const uint32_t rem = nb >> 27;
            uint32_t v = rem ? acc << (32 - rem) : 0u;
            uint32_t f = (sp != sp_first) || tstar;
            const uint32_t brk = f;
            const uint32_t any = __ballot_sync(0xffffffffu, brk);
            if (any != 0xffffffffu) enc_seg_or_scan(v, f, lane);
            uint32_t cin = __shfl_up_sync(0xffffffffu, v, 1);
            if (lane == 0) cin = 0;
            if (tstar) {
                if ((H & 31u) && lane > 0) stage[H >> 5] = cin;  // (a zero leftover must be stored too: nothing else writes that word)
            } else if (brk && cin) {
                stage[pos0 >> 5] |= cin;
            }
            if (lane == 31) {
                s_tail[warp] = v;
                s_brk[warp] = any != 0;
            }
            if (lane == 0) s_wpos[warp] = tstar ? (H | 0x80000000u) : pos0;
            if (tid == 0) s_wpos[kEncWarps] = total.hb ? Q + total.tail : total.head;
        }
        if (next_valid) prepare_map(next);
        enc_worker_sync();  // C: the map and the symbols of the next tile are complete

        prev_valid = true;
        if (tid == 0) {  // (read after the next worker barrier; the copy-out that read the old values is over)
            s_prev.tile = tile;
            s_prev.tf0 = tf0;
            s_prev.tf1 = tf1;
            s_prev.nvalid = nvalid;
            s_prev.total = total;
        }
        tile = next;
        tf0 = p_t.tf0;
        tf1 = p_t.tf1;
        nvalid = p_t.nvalid;
    }
}

}  // namespace hb
