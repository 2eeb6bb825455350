// Tiled, symbol-parallel Huffman encode for the packed layout (BASELINE configs 2, 3, 5).
//
// The whole batch is treated as ONE run of input bytes cut into fixed tiles of kEncTile symbols; a
// block (8 worker warps + 1 scout warp) owns a tile regardless of where item boundaries fall, so loads are
// 128-bit and coalesced and every lane has work. Per tile:
//
//   1. each worker reads its 32 symbols (fetched with cp.async into shared memory while the previous
//      tile was packed) and looks their {code, length} up in shared memory ONCE — in a table replicated
//      so that every lane of a half-warp owns a pair of banks — and keeps the 32 pairs in registers
//   2. the bit counts are reduced to a "segment function"
//          p -> p + head                      (no item starts in the range)
//          p -> ceil8(p + head) + tail        (items start in the range; an item start byte-aligns the
//                                              output because the previous item is padded, huffman.c:178-184)
//      and scanned over the block (warp shuffles + one barrier). The tile's own function goes to its
//      look-back descriptor immediately, so successors rarely wait for it.
//   3. every worker now knows the tile-relative bit position of its first code and packs its codes
//      MSB-first STRAIGHT INTO THE STAGE at that alignment (enc_append: one multiply-add, one add whose
//      carry says "a word is complete", one predicated store): whole words are plain stores, the word a
//      thread shares with its predecessor travels by one warp shuffle (a segmented OR-scan only when some
//      thread is too short to complete a word) and is merged by a read-modify-write of the owning thread —
//      no shared-memory atomics, no zeroing.
//      The stage is tile-relative, in two pieces: the bits before the tile's first item start (whose
//      output position depends on the tile's absolute bit position G, modulo 8) and everything from
//      that item start on (byte aligned in the output whatever G is).
//   4. the SCOUT warp resolves G (single-pass decoupled look-back, seg_resolve) while the workers copy
//      out the previous tile, pack this one and measure the next; it also computes the bits that complete
//      the tile's last byte
//   5. one tile later the workers copy both pieces to global memory through one funnel shift per 32-bit
//      word (shift G mod 32 for the first piece, whole bytes for the second), coalesced. A tile owns the
//      bytes whose first bit it holds; the bits that complete its last byte belong to the next tile's
//      first symbols (or to the EOS padding) and are recomputed from the input, so tiles never write the
//      same byte and no global atomics or pre-zeroed output are needed.
//
// Only used when every symbol has a code of at most 31 bits (no UNKNOWN_SYMBOL possible); otherwise the
// generic kernel runs. Results are bit-identical to the generic kernel and therefore to the reference.
#pragma once

#include "device_common.cuh"

namespace hb {

constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;
constexpr int kEncSymsPerThread = 32;
constexpr int kEncTile = kEncThreads * kEncSymsPerThread;  // 8192 input bytes
// worst case 32 bits/symbol (padding included: ceil8(sum) <= 32 * symbols), the gap between the two
// pieces, and slack for the copy's look-ahead word
constexpr int kEncStageWords = kEncTile + 16;

// p -> hb ? ceil8(p + head) + tail : p + head
struct Seg {
    uint32_t head, tail, hb;
};
struct Seg64 {
    uint64_t head, tail;
    uint32_t hb;
};

__device__ __forceinline__ Seg seg_combine(const Seg &l, const Seg &r) {  // l first, then r
    const uint32_t x = l.tail + r.head;
    const uint32_t xr = r.hb ? ((x + 7u) & ~7u) : x;
    Seg o;
    o.head = l.hb ? l.head : l.head + r.head;
    o.tail = l.hb ? xr + r.tail : r.tail;
    o.hb = l.hb | r.hb;
    return o;
}
__device__ __forceinline__ Seg64 seg_combine64(const Seg64 &l, const Seg64 &r) {
    const uint64_t x = l.tail + r.head;
    const uint64_t xr = r.hb ? ((x + 7ull) & ~7ull) : x;
    Seg64 o;
    o.head = l.hb ? l.head : l.head + r.head;
    o.tail = l.hb ? xr + r.tail : r.tail;
    o.hb = l.hb | r.hb;
    return o;
}
__device__ __forceinline__ uint64_t seg_apply(const Seg &f, uint64_t p) {
    return f.hb ? ((p + f.head + 7ull) & ~7ull) + f.tail : p + f.head;
}
__device__ __forceinline__ uint64_t seg_apply64(const Seg64 &f, uint64_t p) {
    return f.hb ? ((p + f.head + 7ull) & ~7ull) + f.tail : p + f.head;
}
__device__ __forceinline__ Seg seg_shfl_up(const Seg &v, int d) {
    Seg o;
    o.head = __shfl_up_sync(0xffffffffu, v.head, d);
    o.tail = __shfl_up_sync(0xffffffffu, v.tail, d);
    o.hb = __shfl_up_sync(0xffffffffu, v.hb, d);
    return o;
}
__device__ __forceinline__ Seg seg_shfl(const Seg &v, int src) {
    Seg o;
    o.head = __shfl_sync(0xffffffffu, v.head, src);
    o.tail = __shfl_sync(0xffffffffu, v.tail, src);
    o.hb = __shfl_sync(0xffffffffu, v.hb, src);
    return o;
}

// Tile descriptor word: [63:62] status. Aggregate: [61] hb, [60:31] head, [30:0] tail (a tile's own
// function). Prefix: [61:0] absolute output bit position at the END of the tile.
__device__ __forceinline__ uint64_t seg_pack_aggregate(const Seg &f) {
    return (kLbAggregate << kLbFlagShift) | ((uint64_t)f.hb << 61) | ((uint64_t)f.head << 31) | (uint64_t)f.tail;
}

// One thread: make this tile's own function visible as early as possible.
__device__ __forceinline__ void seg_publish_aggregate(uint64_t *tile_state, uint32_t tile, const Seg &agg) {
    if (tile == 0)
        st_relaxed_u64(&tile_state[0], (kLbPrefix << kLbFlagShift) | (seg_apply(agg, 0) & kLbValueMask));
    else
        st_relaxed_u64(&tile_state[tile], seg_pack_aggregate(agg));
}

// One full warp (the scout), some time after seg_publish_aggregate: resolves the absolute start position G
// of the tile by walking back to the nearest tile whose end position is known, then publishes this
// tile's end position.
//
// Strong (gpu-scope) loads of one warp complete one after the other (measured: 16 loads per round cost
// 16 L2 latencies), so a round is ONE 256-bit load instruction: four consecutive descriptors per lane,
// 128 per round; lane 0 holds the closest group, higher lanes earlier tiles. A round whose closest
// descriptors are not published yet (the grid runs in near lockstep, so same-generation predecessors
// publish at about the same time) is simply repeated after a short sleep.
//   kSeg = false  plain sums: nearest end position by a warp min-reduction, sum by a warp add-reduction
//   kSeg = true   segment functions: composed in order inside the lane, then by a shuffle tree
// tile_state must be 32-byte aligned.
template <bool kSeg>
__device__ __forceinline__ uint64_t seg_resolve(uint64_t *tile_state, uint32_t tile, const Seg &agg) {
    const uint32_t lane = lane_id();
    if (tile == 0) return 0;
    uint64_t G = 0;
    Seg64 acc = {0, 0, 0};           // function of the tiles between the current window and `tile`
    int64_t gtop = ((int64_t)tile - 1) >> 2;  // closest group of four descriptors
    uint32_t rtop = (tile - 1) & 3;           // last element of that group that is a predecessor
    while (true) {
        const int64_t g = gtop - (int64_t)lane;
        uint64_t word[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) word[e] = kLbPrefix << kLbFlagShift;  // "tile -1" ends at bit 0
        if (g >= 0)
            asm volatile("ld.relaxed.gpu.global.v4.b64 {%0, %1, %2, %3}, [%4];"
                         : "=l"(word[0]), "=l"(word[1]), "=l"(word[2]), "=l"(word[3])
                         : "l"(tile_state + 4 * g)
                         : "memory");
        const uint32_t last = lane == 0 ? rtop : 3u;  // elements 0..last of my group are predecessors
        // my closest end position (element pe) and whether something closer than it is still missing
        int pe = -1;
        bool missing = false;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t status = (uint32_t)(word[e] >> kLbFlagShift);
            if ((uint32_t)e <= last) {
                if (status == kLbPrefix) {
                    pe = e;
                    missing = false;
                } else if (status == kLbInvalid) {
                    missing = true;
                }
            }
        }
        const uint32_t pmask = __ballot_sync(0xffffffffu, pe >= 0);
        const uint32_t first = pmask ? (uint32_t)(__ffs(pmask) - 1) : 32u;
        if (__any_sync(0xffffffffu, missing && lane <= first)) {
            __nanosleep(100);
            continue;
        }
        // my function: the aggregates after my end position, composed earliest -> latest
        Seg f = {0, 0, 0};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (e > pe && (uint32_t)e <= last) {
                if (kSeg) {
                    Seg d;
                    d.hb = (uint32_t)(word[e] >> 61) & 1u;
                    d.head = (uint32_t)(word[e] >> 31) & 0x3FFFFFFFu;
                    d.tail = (uint32_t)word[e] & 0x7FFFFFFFu;
                    f = seg_combine(f, d);
                } else {
                    f.head += (uint32_t)(word[e] >> 31) & 0x3FFFFFFFu;
                }
            }
        }
        if (lane > first) f = Seg{0, 0, 0};
        Seg64 window = {0, 0, 0};
        if (kSeg) {
            // ordered reduction: higher lanes hold EARLIER tiles
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                Seg o;
                o.head = __shfl_down_sync(0xffffffffu, f.head, d);
                o.tail = __shfl_down_sync(0xffffffffu, f.tail, d);
                o.hb = __shfl_down_sync(0xffffffffu, f.hb, d);
                if (lane + d < 32) f = seg_combine(o, f);
            }
            window.head = __shfl_sync(0xffffffffu, f.head, 0);
            window.tail = __shfl_sync(0xffffffffu, f.tail, 0);
            window.hb = __shfl_sync(0xffffffffu, f.hb, 0);
        } else {
            window.head = __reduce_add_sync(0xffffffffu, f.head);
        }
        acc = seg_combine64(window, acc);
        if (pmask) {
            uint64_t mine_end = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (e == pe) mine_end = word[e] & kLbValueMask;
            const uint64_t end_of_known = __shfl_sync(0xffffffffu, mine_end, first);
            G = seg_apply64(acc, end_of_known);
            break;
        }
        gtop -= 32;
        rtop = 3;
    }
    if (lane == 0) st_relaxed_u64(&tile_state[tile], (kLbPrefix << kLbFlagShift) | (seg_apply(agg, G) & kLbValueMask));
    return G;
}

// tile_first[j] = first item whose start offset is >= min(j * kEncTile, total_in), for j in [0, num_tiles]
// (so tile_first[num_tiles] is the first trailing empty item, or n)
__global__ void tile_index_kernel(
    const uint64_t *in_offsets, uint64_t n, uint64_t total_in, uint64_t num_tiles, uint32_t *tile_first,
    const uint32_t *gate = nullptr) {
    if (gate != nullptr && *gate == 0) return;  // (fallback of str_pack_kernel: nothing to redo)
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > num_tiles) return;
    const uint64_t target = min(j * (uint64_t)kEncTile, total_in);
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (in_offsets[mid] < target) lo = mid + 1; else hi = mid;
    }
    tile_first[j] = (uint32_t)lo;
}

struct EncTiledArgs {
    const uint8_t *in;
    const uint64_t *in_offsets;
    uint64_t n;
    uint64_t total_in;
    uint8_t *out;
    uint64_t out_capacity;
    uint64_t *out_offsets;
    const uint32_t *tile_first;  // kSeg only
    uint64_t *tile_state;
    uint32_t *ticket;
    uint32_t num_tiles;
    uint32_t eos_padding;
    const uint32_t *gate;  // not null: the launch is a fallback that only runs when *gate != 0
};

// Shared-memory table entry of the tiled encoder: x = code, y = len | len << 27 (len <= 31).
// The bit counter `nb` of a thread has the same two fields: the top five bits count the pending
// (not yet stored) bits modulo 32, so "a word is complete" is simply the CARRY of nb += y; the low field
// counts all bits since the thread's first word boundary, and a wrapping funnel shift by nb extracts the
// completed word (it only looks at nb mod 32). One add does the work of add + compare + wrap.
constexpr uint32_t kEncLenMask = 0x07ffffffu;
__device__ __forceinline__ uint32_t enc_len_fields(uint32_t len) { return len | (len << 27); }

// Appends one code to the accumulator. Only the low word survives between appends: it holds the < 32 bits
// not stored yet (above them: stale bits that every later extraction ignores), so "shift left and OR" is
// a multiply-add by 1 << len on the FMA pipe (the integer ALU pipe is the one this kernel saturates).
// The store and the pointer bump are predicated, no branches. `sp` is a shared-window address.
// Written in PTX so that the multiply-add stays one (the compiler turns it back into ALU shifts and adds).
__device__ __forceinline__ void enc_append(uint32_t &sp, uint32_t &acc, uint32_t &nb, uint32_t code, uint32_t ey) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 pw, lo, hi;\n\t"
        "shf.l.wrap.b32 pw, 0, 1, %3;\n\t"
        "mul.hi.u32 hi, %1, pw;\n\t"
        "mad.lo.u32 lo, %1, pw, %4;\n\t"
        "add.u32 %2, %2, %3;\n\t"
        "setp.lt.u32 p, %2, %3;\n\t"
        "shf.r.wrap.b32 hi, lo, hi, %2;\n\t"
        "@p st.shared.u32 [%0], hi;\n\t"
        "@p add.u32 %0, %0, 4;\n\t"
        "mov.b32 %1, lo;\n\t"
        "}"
        : "+r"(sp), "+r"(acc), "+r"(nb)
        : "r"(ey), "r"(code)
        : "memory");
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// Table entry of byte `k` of `word`. The table is replicated 16 times, entry s of copy j at
// (16 s + j) * 8: every lane of a half-warp reads from its own pair of banks (copy lane % 16), so a
// lookup is conflict-free whatever the symbols are. `tab` = shared-window address of MY copy's entry 0.
constexpr int kEncTabCopies = 16;
__device__ __forceinline__ uint2 enc_lookup(uint32_t tab, uint32_t word, int k) {
    const uint32_t byte = __byte_perm(word, 0, 0x4440 | (k & 3));
    uint32_t addr;
    asm("mad.lo.u32 %0, %1, %3, %2;" : "=r"(addr) : "r"(byte), "r"(tab), "n"(8 * kEncTabCopies));
    uint2 e;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(addr));
    return e;
}

// Starts fetching the kEncSymsPerThread symbols of thread `tid` of `tile` into its slots of the symbol
// buffer (cp.async: no registers are held while the previous tile is packed). The buffer keeps the j-th
// 16-byte piece of every thread together ([j][tid]) so that reading it back is bank-conflict free.
// Bytes past the end of the input read as 0.
__device__ __forceinline__ void enc_fetch_symbols(const EncTiledArgs &a, uint32_t tile, uint32_t tid, uint32_t sym_addr) {
    const uint64_t p0 = (uint64_t)tile * kEncTile + (uint64_t)tid * kEncSymsPerThread;
    if (p0 + kEncSymsPerThread <= a.total_in && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < kEncSymsPerThread / 16; ++j)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sym_addr + (j * kEncThreads + tid) * 16),
                         "l"(a.in + p0 + 16 * j)
                         : "memory");
    } else {
#pragma unroll
        for (int j = 0; j < kEncSymsPerThread / 4; ++j) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (p0 + 4 * j + k < a.total_in) v |= (uint32_t)a.in[p0 + 4 * j + k] << (8 * k);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(sym_addr + ((j >> 2) * kEncThreads + tid) * 16 + (j & 3) * 4), "r"(v)
                         : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
// Waits for this thread's fetch and reads its symbols, little-endian in w[].
__device__ __forceinline__ void enc_read_symbols(uint32_t tid, uint32_t sym_addr, uint32_t (&w)[kEncSymsPerThread / 4]) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int j = 0; j < kEncSymsPerThread / 16; ++j)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[4 * j]), "=r"(w[4 * j + 1]), "=r"(w[4 * j + 2]), "=r"(w[4 * j + 3])
                     : "r"(sym_addr + (j * kEncThreads + tid) * 16)
                     : "memory");
}

// Inclusive segmented OR-scan over the warp: a lane with f set starts a new segment.
__device__ __forceinline__ void enc_seg_or_scan(uint32_t &v, uint32_t &f, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t uv = __shfl_up_sync(0xffffffffu, v, d);
        const uint32_t uf = __shfl_up_sync(0xffffffffu, f, d);
        if (lane >= (uint32_t)d && !f) {
            v |= uv;
            f = uf;
        }
    }
}

// Copies stage bits [s0, s0 + nbits) to output bits [o0, o0 + nbits). The piece owns the output bytes
// whose FIRST bit it holds. `closing` (already in the low bits of its byte) completes the last byte when
// the piece ends inside one. All positions are taken relative to the 4-byte aligned address at or below
// the piece's first byte, so the arithmetic is 32-bit.
__device__ __forceinline__ void enc_copy_piece(
    const uint32_t *st, uint32_t s0, uint32_t nbits, uint64_t o0, uint32_t closing, uint8_t *out, uint64_t cap,
    uint32_t tid) {
    if (nbits == 0) return;
    const uint64_t byte0 = o0 >> 3;
    if (byte0 >= cap) return;
    uint8_t *const first = out + byte0;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 3);
    uint8_t *const base = first - mis;                      // 4-byte aligned
    const uint32_t ob = 8 * mis + (uint32_t)(o0 & 7);       // where the piece starts, in bits from `base`
    const uint32_t oe = ob + nbits;                         // where it ends
    const uint64_t room = cap - byte0 + mis;                // bytes addressable from `base`
    const uint32_t b_lo = (ob + 7) >> 3;                    // owned bytes [b_lo, b_hi)
    const uint32_t b_hi = (uint32_t)min((uint64_t)((oe + 7) >> 3), room);
    if (b_lo >= b_hi) return;
    // whole words inside the piece: [j_lo, j_hi)
    const uint32_t j_lo = (ob + 31) >> 5;
    const uint32_t j_hi = min(oe >> 5, b_hi >> 2);
    const int d = (int)s0 - (int)ob;  // stage bit = bit from `base` + d
    if (j_lo < j_hi) {
        const uint32_t sh = (uint32_t)d & 31;
        const uint32_t *src = st + ((d + 32 * (int)j_lo) >> 5);
        uint32_t *dst = reinterpret_cast<uint32_t *>(base) + j_lo;
        const uint32_t nwords = j_hi - j_lo;
#pragma unroll 2
        for (uint32_t i = tid; i < nwords; i += kEncThreads)
            dst[i] = __byte_perm(__funnelshift_l(src[i + 1], src[i], sh), 0, 0x0123);
    }
    // edge bytes: before the first whole word and after the last one (at most 3 + 4), or all of a short piece
    const uint32_t e_lo = j_lo < j_hi ? 4 * j_lo : b_hi, e_hi = j_lo < j_hi ? 4 * j_hi : b_hi;
    const uint32_t nhead = e_lo - b_lo;
    const uint32_t nedge = nhead + (b_hi - e_hi);
    if (tid < nedge) {
        const uint32_t B = tid < nhead ? b_lo + tid : e_hi + (tid - nhead);
        const uint32_t sb = (uint32_t)((int)(8 * B) + d);
        uint32_t v = __funnelshift_l(st[(sb >> 5) + 1], st[sb >> 5], sb & 31) >> 24;
        const uint32_t valid = oe - 8 * B;  // >= 1
        if (valid < 8) v = (v & (0xff00u >> valid) & 0xffu) | closing;
        base[B] = (uint8_t)v;
    }
}

// The words shared by neighbouring warps and the tile's last, partial word: warp w's first completed word
// still lacks what the warps before it left over (s_tail). Called by every worker warp after a barrier
// that follows the packing. Normally every warp completed a word, so lane 0 of warp w simply ORs warp
// w-1's leftover into its first word; otherwise (tiny tiles) warp 1 runs a segmented OR-scan over the warps.
__device__ __forceinline__ void enc_fix_warp_boundaries(
    uint32_t *stage, const uint32_t *s_tail, const uint32_t *s_brk, const uint32_t *s_wpos, uint32_t warp, uint32_t lane) {
    // s_wpos[w] = stage bit where warp w starts; bit 31 set: the leftover is STORED there instead of ORed in
    // (encode_slots_kernel: the warp opens piece 1 and the leftover is the last word of piece 0)
    const uint32_t brk = lane < kEncWarps ? s_brk[lane] : 1u;
    const uint32_t endpos = s_wpos[kEncWarps];
    if (__all_sync(0xffffffffu, brk)) {
        if (lane == 0 && warp > 0) {
            const uint32_t cin = s_tail[warp - 1];
            const uint32_t wp = s_wpos[warp];
            if (wp & 0x80000000u) {
                if (wp & 31u) stage[(wp & 0x7fffffffu) >> 5] = cin;  // (stored even when zero: nothing else writes that word)
            } else if (cin) {
                stage[wp >> 5] |= cin;
            }
        }
        if (lane == 1 && warp == kEncWarps - 1 && (endpos & 31u)) stage[endpos >> 5] = s_tail[kEncWarps - 1];
    } else if (warp == 1) {
        uint32_t v = lane < kEncWarps ? s_tail[lane] : 0u;
        uint32_t f = lane < kEncWarps ? s_brk[lane] : 0u;
        enc_seg_or_scan(v, f, lane);
        uint32_t cin = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) cin = 0;
        if (lane < kEncWarps && brk) {
            const uint32_t wp = s_wpos[lane];
            if (wp & 0x80000000u) {
                if (wp & 31u) stage[(wp & 0x7fffffffu) >> 5] = cin;
            } else if (cin) {
                stage[wp >> 5] |= cin;
            }
        }
        if (lane == kEncWarps - 1 && (endpos & 31u)) stage[endpos >> 5] = v;
    }
}

constexpr size_t kEncSmemBytes =
    kEncStageWords * 4 + kEncTile * 2 /* item start positions */ + 2048 * kEncTabCopies /* code table */ + kEncTile /* symbols */;

// Persistent blocks of 8 WORKER warps + 1 SCOUT warp; tiles come from an atomic ticket.
//
// Look-back rule: between taking a ticket and publishing that tile's function a block must never wait
// for another tile (otherwise waits chain from block to block and the grid degenerates into a convoy).
// The workers therefore never wait for a look-back before they publish: per tile t they run
//
//     measure(t)  lookups, scan, publish the tile's function          (worker barrier A)
//     hand tile t to the scout (arrive, no wait); wait for G(t-1) (normally there since long)
//     ticket(t+1) + copy-out(t-1)                                      (worker barrier B)
//     fetch(t+1)  symbols, item range (loads in flight during the packing)
//     pack(t)     into the stage                                       (batches: worker barrier C)
//
// and the scout resolves G(t) — the decoupled look-back, a chain of dependent L2 round trips — while
// the workers copy out t-1, pack t and measure t+1: a whole tile of slack. It also computes the bits
// that complete the tile's last byte.
constexpr int kEncBlock = kEncThreads + 32;
constexpr uint32_t kEncDone = 0xffffffffu;

// Named barriers. Workers among themselves: 1. Hand-off of a tile to the scout: 2 + parity (workers
// arrive without waiting, the scout waits). Result of a look-back: 4 + parity (the scout arrives, the
// workers wait — normally not at all, the result has been there for most of a tile).
__device__ __forceinline__ void enc_worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEncThreads) : "memory"); }
__device__ __forceinline__ void enc_bar_arrive(uint32_t id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kEncBlock) : "memory"); }
__device__ __forceinline__ void enc_bar_sync(uint32_t id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kEncBlock) : "memory"); }

struct EncHandoff {   // worker thread 0 -> scout, double buffered by iteration parity
    uint32_t tile;    // kEncDone: no more tiles
    uint32_t tf1;     // first item that starts after the tile (kSeg)
    Seg total;
};
struct EncResult {    // scout -> workers, double buffered by tile parity
    uint64_t G;
    uint32_t fill;
};

template <bool kSeg>
__global__ void __launch_bounds__(kEncBlock, kEncSymsPerThread == 32 ? 2 : 3) encode_tiled_kernel(const uint2 *__restrict__ enc_table, EncTiledArgs a) {
    __shared__ uint32_t s_mask[kEncTile / 32];
    __shared__ Seg s_wseg[kEncWarps];
    __shared__ uint32_t s_tail[kEncWarps], s_brk[kEncWarps], s_wpos[kEncWarps + 1];
    __shared__ uint32_t s_next;
    __shared__ EncHandoff s_hand[2];
    __shared__ EncResult s_res[2];
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint32_t *const stage = reinterpret_cast<uint32_t *>(s_dyn);
    uint16_t *const obpos = reinterpret_cast<uint16_t *>(s_dyn + kEncStageWords * 4);
    const uint32_t tab0 = smem_addr(s_dyn + kEncStageWords * 4 + kEncTile * 2);
    const uint32_t stage_addr = smem_addr(stage);
    const uint32_t sym_addr = tab0 + 2048 * kEncTabCopies;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t tab = tab0 + (lane & (kEncTabCopies - 1)) * 8;  // my copy of the table

    if (a.gate != nullptr && *a.gate == 0) return;
    if (tid == 0) s_next = atomicAdd(a.ticket, 1u);
    if (tid < 256) {
        const uint2 e = enc_table[tid];
#pragma unroll
        for (int j = 0; j < kEncTabCopies; ++j)
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(tab0 + (tid * kEncTabCopies + j) * 8), "r"(e.x),
                         "r"(enc_len_fields(e.y))
                         : "memory");
    }
    if (kSeg && tid < kEncTile / 32) s_mask[tid] = 0;
    __syncthreads();

    // ================================ scout =========================================================================
    if (warp == kEncWarps) {
        for (uint32_t it = 0;; ++it) {
            enc_bar_sync(2 + (it & 1));  // the workers handed over a tile
            const uint32_t tile = s_hand[it & 1].tile;
            if (tile == kEncDone) return;
            const Seg total = s_hand[it & 1].total;
            const uint32_t tf1 = s_hand[it & 1].tf1;
            const uint64_t t1 = min(((uint64_t)tile + 1) * kEncTile, a.total_in);
            // what may complete the tile's last byte: the next symbols of the item that is open at its end
            uint32_t fb_lo = 0, fb_hi = 0;
            uint64_t stop = a.total_in;
            if (lane == 0) {
                if (kSeg && tf1 < a.n) stop = a.in_offsets[tf1];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (t1 + q < a.total_in) {
                        const uint32_t byte = a.in[t1 + q];
                        if (q < 4) fb_lo |= byte << (8 * q); else fb_hi |= byte << (8 * (q - 4));
                    }
                }
            }
            const uint64_t G0 = seg_resolve<kSeg>(a.tile_state, tile, total);
            if (lane == 0) {
                const uint64_t Gend = seg_apply(total, G0);
                const uint32_t need = (8u - (uint32_t)(Gend & 7u)) & 7u;
                uint32_t bits = 0;
                if (need) {
                    uint32_t have = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (have < need && t1 + q < stop) {
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
    uint32_t tile = s_next;
    uint32_t tf0 = 0, tf1 = 0;  // items that start inside `tile`: [tf0, tf1)
    if (tile < a.num_tiles) {
        enc_fetch_symbols(a, tile, tid, sym_addr);
        if (kSeg) {
            tf0 = a.tile_first[tile];
            tf1 = a.tile_first[tile + 1];
            const uint64_t t0 = (uint64_t)tile * kEncTile;
            for (uint32_t i = tf0 + tid; i < tf1; i += kEncThreads) {
                const uint32_t p = (uint32_t)(a.in_offsets[i] - t0);
                atomicOr(&s_mask[p >> 5], 1u << (p & 31));
            }
        }
    }
    enc_worker_sync();

    // the tile that is packed in the stage and waits for its position
    bool prev_valid = false;
    uint32_t prev_tile = 0, prev_tf0 = 0, prev_tf1 = 0;
    Seg prev_total = {0, 0, 0};

    for (uint32_t it = 0;; ++it) {
        const bool cur_valid = tile < a.num_tiles;
        uint32_t c[kEncSymsPerThread], l[kEncSymsPerThread];
        uint32_t m = 0;
        Seg excl = {0, 0, 0}, total = {0, 0, 0}, we = {0, 0, 0};
        if (cur_valid) {
            if (kSeg) {
                m = kEncSymsPerThread == 32 ? s_mask[tid] : (s_mask[tid >> 1] >> ((tid & 1) * 16)) & 0xffffu;
            }
            uint32_t w[kEncSymsPerThread / 4];
            enc_read_symbols(tid, sym_addr, w);
            // ---- 1. code and length of my 16 symbols -------------------------------------------------------
            const uint64_t p0 = (uint64_t)tile * kEncTile + (uint64_t)tid * kEncSymsPerThread;
            if (p0 + kEncSymsPerThread <= a.total_in) {
#pragma unroll
                for (int k = 0; k < kEncSymsPerThread; ++k) {
                    const uint2 e = enc_lookup(tab, w[k >> 2], k);
                    c[k] = e.x;
                    l[k] = e.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kEncSymsPerThread; ++k) {
                    const uint2 e = enc_lookup(tab, w[k >> 2], k);
                    const bool valid = p0 + k < a.total_in;
                    c[k] = valid ? e.x : 0u;
                    l[k] = valid ? e.y : 0u;
                }
            }

            // ---- 2. my segment function, block scan ----------------------------------------------------------
            // (l[] holds the two-field form; sums are masked once)
            Seg mine = {0, 0, 0};
            if (kSeg && m != 0) {
                uint32_t run = 0;
#pragma unroll
                for (int k = 0; k < kEncSymsPerThread; ++k) {
                    if ((m >> k) & 1u) {
                        run &= kEncLenMask;
                        if (mine.hb) mine.tail += (run + 7u) & ~7u; else mine.head = run;
                        mine.hb = 1;
                        run = 0;
                    }
                    run += l[k];
                }
                mine.tail += run & kEncLenMask;
            } else {
                uint32_t s = 0;
#pragma unroll
                for (int k = 0; k < kEncSymsPerThread; ++k) s += l[k];
                mine.head = s & kEncLenMask;
            }
            if (!kSeg) {
                const uint32_t incl = warp_inclusive_scan(mine.head);
                if (lane == 31) s_wseg[warp].head = incl;
                enc_worker_sync();  // A
                const uint32_t wsum = lane < kEncWarps ? s_wseg[lane].head : 0u;
                uint32_t wincl = wsum;
#pragma unroll
                for (int d = 1; d < kEncWarps; d <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, wincl, d);
                    if (lane >= (uint32_t)d) wincl += up;
                }
                we = Seg{wincl - wsum, 0, 0};
                total = Seg{__shfl_sync(0xffffffffu, wincl, kEncWarps - 1), 0, 0};
                excl = Seg{__shfl_sync(0xffffffffu, wincl - wsum, warp) + incl - mine.head, 0, 0};
            } else {
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
                // every warp redoes the 8-entry scan (cheaper than another barrier)
                Seg wi = lane < kEncWarps ? s_wseg[lane] : Seg{0, 0, 0};
#pragma unroll
                for (int d = 1; d < kEncWarps; d <<= 1) {
                    const Seg up = seg_shfl_up(wi, d);
                    if (lane >= (uint32_t)d) wi = seg_combine(up, wi);
                }
                we = seg_shfl_up(wi, 1);
                if (lane == 0) we = Seg{0, 0, 0};
                total = seg_shfl(wi, kEncWarps - 1);
                excl = seg_combine(seg_shfl(we, warp), excl);
            }
            if (tid == 0) {
                seg_publish_aggregate(a.tile_state, tile, total);
                s_hand[it & 1].tile = tile;
                s_hand[it & 1].tf1 = tf1;
                s_hand[it & 1].total = total;
            }
            if (kSeg && tid < kEncTile / 32) s_mask[tid] = 0;  // everybody has read its bits of this tile
        } else {
            if (tid == 0) s_hand[it & 1].tile = kEncDone;
            if (prev_valid) enc_worker_sync();  // the previous tile is completely packed
        }
        // (barrier A, or the one just above, separates this from the packing of the previous tile)
        if (prev_valid) enc_fix_warp_boundaries(stage, s_tail, s_brk, s_wpos, warp, lane);
        enc_bar_arrive(2 + (it & 1));                       // the scout may take tile `tile`
        if (prev_valid) enc_bar_sync(4 + ((it - 1) & 1));  // G of the previous tile is in s_res[(it - 1) & 1]
        if (cur_valid && tid == 0) s_next = atomicAdd(a.ticket, 1u);  // nothing this block waits for lies ahead

        // ---- copy out the previous tile ------------------------------------------------------------------------
        if (prev_valid) {
            const uint64_t G = s_res[(it - 1) & 1].G;
            const uint32_t fill = s_res[(it - 1) & 1].fill;
            const uint64_t Gend = seg_apply(prev_total, G);
            const uint32_t H = prev_total.head;
            if (!kSeg || !prev_total.hb) {
                enc_copy_piece(stage, 0, H, G, fill, a.out, a.out_capacity, tid);
            } else {
                const uint32_t Q = ((H >> 5) + 2u) << 5;
                const uint32_t pad = (uint32_t)(0 - (G + H)) & 7u;
                enc_copy_piece(stage, 0, H, G, a.eos_padding & ((1u << pad) - 1u), a.out, a.out_capacity, tid);
                enc_copy_piece(stage, Q, prev_total.tail, G + H + pad, fill, a.out, a.out_capacity, tid);
                const uint64_t tail_byte = (G + H + pad) >> 3;
                const uint64_t t0 = (uint64_t)prev_tile * kEncTile;
                for (uint32_t i = prev_tf0 + tid; i < prev_tf1; i += kEncThreads) {
                    const uint32_t p = (uint32_t)(a.in_offsets[i] - t0);
                    a.out_offsets[i] = tail_byte + obpos[p];
                }
            }
            if (prev_tile + 1 == a.num_tiles) {
                const uint64_t total_out = (Gend + 7) >> 3;
                if (kSeg) {
                    // trailing empty items start at total_in
                    for (uint64_t i = (uint64_t)prev_tf1 + tid; i <= a.n; i += kEncThreads) a.out_offsets[i] = total_out;
                } else if (tid == 0) {
                    a.out_offsets[0] = 0;
                    a.out_offsets[a.n] = total_out;
                }
            }
        }
        if (!cur_valid) return;
        enc_worker_sync();  // B: the stage is free again, the next ticket is visible

        // ---- fetch the next tile (in flight during the packing) ---------------------------------------------------
        const uint32_t next = s_next;
        uint32_t ntf0 = 0, ntf1 = 0;
        if (next < a.num_tiles) {
            enc_fetch_symbols(a, next, tid, sym_addr);
            if (kSeg) {
                ntf0 = a.tile_first[next];
                ntf1 = a.tile_first[next + 1];
            }
        }

        // ---- 3. pack straight into the stage --------------------------------------------------------------------
        // piece 0 = stage bits [0, H): everything before the tile's first item start; piece 1 starts at stage
        // bit Q (word aligned, one spare word after piece 0)
        const uint32_t H = total.head;
        const uint32_t Q = ((H >> 5) + 2u) << 5;
        const uint32_t tail_base = stage_addr + (Q >> 5) * 4;
        bool in_head = !excl.hb;
        const uint32_t pos0 = in_head ? excl.head : Q + excl.tail;
        const uint32_t sp_first = stage_addr + (pos0 >> 5) * 4;
        uint32_t sp = sp_first;
        uint32_t acc = 0;
        uint32_t nb = enc_len_fields(pos0 & 31);
        uint64_t next_item_off = 0;  // start offset of "my" item of the next tile (kSeg)
#pragma unroll
        for (int k = 0; k < kEncSymsPerThread; ++k) {
            if (kSeg && k == kEncSymsPerThread / 2) {
                // the item range of the next tile has arrived by now; its offsets arrive during the second half
                if (ntf0 + tid < ntf1) next_item_off = a.in_offsets[ntf0 + tid];
            }
            if (kSeg && ((m >> k) & 1u)) {
                if (in_head) {
                    // the tile's first item start: close piece 0 (its EOS padding depends on G and is added by
                    // the copy) and continue at the start of piece 1
                    const uint32_t rem = nb >> 27;
                    stage[(sp - stage_addr) >> 2] = rem ? acc << (32 - rem) : 0u;
                    sp = tail_base;
                    acc = 0;
                    nb = 0;
                    in_head = false;
                } else {
                    // byte-align: pad the item that just ended with the LOW bits of eos_padding (huffman.c:178-184)
                    const uint32_t pad = (0u - nb) & 7u;
                    enc_append(sp, acc, nb, a.eos_padding & ((1u << pad) - 1u), enc_len_fields(pad));
                }
                // byte offset of the new item from the start of piece 1
                obpos[tid * kEncSymsPerThread + k] = (uint16_t)((sp - tail_base) + (nb >> 30));
            }
            enc_append(sp, acc, nb, c[k], l[k]);
        }
        // The word I share with my predecessor(s): a thread that completed at least one word (or moved to
        // piece 1) ORs what came before into its first word; the others pass their bits on.
        {
            const uint32_t rem = nb >> 27;
            uint32_t v = rem ? acc << (32 - rem) : 0u;
            uint32_t f = sp != sp_first;
            const uint32_t brk = f;
            const uint32_t any = __ballot_sync(0xffffffffu, brk);
            if (any != 0xffffffffu) enc_seg_or_scan(v, f, lane);
            uint32_t cin = __shfl_up_sync(0xffffffffu, v, 1);
            if (lane == 0) cin = 0;
            if (brk && cin) stage[pos0 >> 5] |= cin;
            if (lane == 31) {
                s_tail[warp] = v;
                s_brk[warp] = any != 0;
            }
            if (lane == 0) s_wpos[warp] = pos0;
            if (tid == 0) s_wpos[kEncWarps] = total.hb ? Q + total.tail : total.head;
        }
        if (kSeg) {
            if (next < a.num_tiles) {
                // item starts of the next tile -> one bit per symbol
                const uint64_t nt0 = (uint64_t)next * kEncTile;
                if (ntf0 + tid < ntf1) {
                    const uint32_t p = (uint32_t)(next_item_off - nt0);
                    atomicOr(&s_mask[p >> 5], 1u << (p & 31));
                }
                for (uint32_t i = ntf0 + tid + kEncThreads; i < ntf1; i += kEncThreads) {
                    const uint32_t p = (uint32_t)(a.in_offsets[i] - nt0);
                    atomicOr(&s_mask[p >> 5], 1u << (p & 31));
                }
            }
            enc_worker_sync();  // C: the mask of the next tile is complete
        }

        prev_valid = true;
        prev_tile = tile;
        prev_tf0 = tf0;
        prev_tf1 = tf1;
        prev_total = total;
        tile = next;
        tf0 = ntf0;
        tf1 = ntf1;
    }
}

// Optional per-item arrays in the packed layout when no symbol can be unknown: every item succeeds.
__global__ void fill_packed_meta_kernel(
    uint64_t n, const uint64_t *in_offsets, const uint64_t *out_offsets, uint64_t *out_lens, int32_t *status,
    uint64_t *consumed, uint32_t *overflow_pattern, uint8_t *overflow_num_bits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (out_lens) out_lens[i] = out_offsets[i + 1] - out_offsets[i];
    if (status) status[i] = kStatusOk;
    if (consumed) consumed[i] = in_offsets[i + 1] - in_offsets[i];
    if (overflow_pattern) overflow_pattern[i] = 0;
    if (overflow_num_bits) overflow_num_bits[i] = 0;
}

}  // namespace hb
