// Tiled, symbol-parallel Huffman encode for the packed layout (BASELINE configs 2, 3, 5).
//
// The whole batch is treated as ONE run of input bytes cut into fixed tiles of kEncTile symbols; a
// block owns a tile regardless of where item boundaries fall, so loads are 128-bit and coalesced
// and every lane has work. Per tile:
//
//   A. each thread loads 16 symbols (one uint4), looks the code lengths up in shared memory and
//      reduces them to a "segment function"  p -> p + head            (no item starts in its range)
//                                            p -> ceil8(p + head) + tail   (items start in its range;
//      an item start byte-aligns the output because the previous item is padded, huffman.c:178-184)
//   B. warp-shuffle scan + block scan of those functions. The tile's own function is published to its
//      look-back descriptor immediately (before the expensive packing) so that successors rarely wait;
//      after packing one warp resolves the absolute output bit position G of the tile (single-pass
//      decoupled look-back)
//      (A is done while packing: each thread packs its 16 codes MSB-first, left-aligned, into a private
//      bank-conflict-free shared-memory slot with a 64-bit funnel accumulator and a branch-free,
//      predicated word flush; the bit count of that packing is the function's operand)
//   C. once positions are known each thread shift-copies its packed words to their exact bit position
//      in the tile's staging buffer: interior words are plain stores, the first and last word of a run
//      (shared with a neighbour) are merged with shared-memory atomicOr
//   D. the staged bytes this tile owns (those whose first bit lies in the tile) are copied to global
//      memory with 128-bit stores. The bits that complete the tile's last byte belong to the next
//      tile's first symbols (or to the EOS padding); one thread recomputes them from the input, so
//      tiles never write the same byte and no global atomics or pre-zeroed output are needed.
//
// Only used when every symbol has a code (no UNKNOWN_SYMBOL possible); otherwise the generic kernel
// runs. Results are bit-identical to the generic kernel and therefore to the reference.
#pragma once

#include "device_common.cuh"

namespace hb {

constexpr int kEncThreads = 256;
constexpr int kEncSymsPerThread = 16;
constexpr int kEncTile = kEncThreads * kEncSymsPerThread;  // 4096 input bytes
constexpr int kEncStageBytes = kEncTile * 4 + 64;          // worst case 32 bits/symbol (+ alignment slack)

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

// Tile descriptor word: [63:62] status. Aggregate: [61] hb, [60:31] head, [30:0] tail (a tile's own
// function). Prefix: [61:0] absolute output bit position at the END of the tile.
__device__ __forceinline__ uint64_t seg_pack_aggregate(const Seg &f) {
    return (kLbAggregate << kLbFlagShift) | ((uint64_t)f.hb << 61) | ((uint64_t)f.head << 31) | (uint64_t)f.tail;
}

// Lane 0 of any warp: make this tile's own function visible as early as possible.
__device__ __forceinline__ void seg_publish_aggregate(uint64_t *tile_state, uint32_t tile, const Seg &agg) {
    if (tile == 0)
        st_relaxed_u64(&tile_state[0], (kLbPrefix << kLbFlagShift) | (seg_apply(agg, 0) & kLbValueMask));
    else
        st_relaxed_u64(&tile_state[tile], seg_pack_aggregate(agg));
}

// One full warp, some time after seg_publish_aggregate. Resolves the absolute start position G of the
// tile by walking back to the nearest tile whose end position is known, then publishes this tile's
// end position.
__device__ __forceinline__ uint64_t seg_resolve(uint64_t *tile_state, uint32_t tile, const Seg &agg) {
    // Every lane inspects kLbPerLane consecutive descriptors per round trip, so one window covers
    // 32 * kLbPerLane tiles: with hundreds of tiles in flight the distance to the nearest resolved tile is
    // a few hundred descriptors, and the number of dependent L2 round trips is what a tile waits for.
    constexpr int kLbPerLane = 4;
    const uint32_t lane = lane_id();
    if (tile == 0) return 0;
    uint64_t G = 0;
    Seg64 acc = {0, 0, 0};  // function of tiles (look+1 .. tile-1), identity so far
    int64_t look = (int64_t)tile - 1;
    while (true) {
        const int64_t base = look - (int64_t)lane * kLbPerLane;  // my closest descriptor; q steps further back
        uint64_t word[kLbPerLane];
#pragma unroll
        for (int q = 0; q < kLbPerLane; ++q) {
            const int64_t idx = base - q;
            word[q] = kLbPrefix << kLbFlagShift;  // "tile -1" ends at bit 0
            if (idx >= 0) {
                while (true) {
                    word[q] = ld_relaxed_u64(&tile_state[idx]);
                    if ((word[q] >> kLbFlagShift) != kLbInvalid) break;
                    __nanosleep(20);  // a spinning warp must not steal issue slots from the packing warps
                }
            }
        }
        // my function: descriptors closer than my first end position, composed earliest -> latest
        int qp = kLbPerLane;
#pragma unroll
        for (int q = kLbPerLane - 1; q >= 0; --q)
            if ((word[q] >> kLbFlagShift) == kLbPrefix) qp = q;
        Seg64 f = {0, 0, 0};
#pragma unroll
        for (int q = kLbPerLane - 1; q >= 0; --q) {
            if (q < qp) {
                Seg64 d;
                d.hb = (uint32_t)(word[q] >> 61) & 1u;
                d.head = (word[q] >> 31) & 0x3FFFFFFFull;
                d.tail = word[q] & 0x7FFFFFFFull;
                f = seg_combine64(f, d);
            }
        }
        const uint32_t pmask = __ballot_sync(0xffffffffu, qp < kLbPerLane);
        const uint32_t first = pmask ? (uint32_t)(__ffs(pmask) - 1) : 32u;
        if (lane > first) f = Seg64{0, 0, 0};
        // ordered reduction: higher lanes hold EARLIER tiles
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Seg64 o;
            o.head = __shfl_down_sync(0xffffffffu, f.head, d);
            o.tail = __shfl_down_sync(0xffffffffu, f.tail, d);
            o.hb = __shfl_down_sync(0xffffffffu, f.hb, d);
            if (lane + d < 32) f = seg_combine64(o, f);
        }
        Seg64 window;
        window.head = __shfl_sync(0xffffffffu, f.head, 0);
        window.tail = __shfl_sync(0xffffffffu, f.tail, 0);
        window.hb = __shfl_sync(0xffffffffu, f.hb, 0);
        acc = seg_combine64(window, acc);
        if (pmask) {
            uint64_t mine_end = 0;
#pragma unroll
            for (int q = 0; q < kLbPerLane; ++q)
                if (q == qp) mine_end = word[q] & kLbValueMask;
            const uint64_t end_of_known = __shfl_sync(0xffffffffu, mine_end, first);
            G = seg_apply64(acc, end_of_known);
            break;
        }
        look -= 32 * kLbPerLane;
    }
    if (lane == 0) st_relaxed_u64(&tile_state[tile], (kLbPrefix << kLbFlagShift) | (seg_apply(agg, G) & kLbValueMask));
    return G;
}

// tile_first[j] = first item whose start offset is >= min(j * kEncTile, total_in), for j in [0, num_tiles]
// (so tile_first[num_tiles] is the first trailing empty item, or n)
__global__ void tile_index_kernel(
    const uint64_t *in_offsets, uint64_t n, uint64_t total_in, uint64_t num_tiles, uint32_t *tile_first) {
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
    uint32_t debug;  // timing experiments only (AWS_HUFFMAN_BATCH_EXPERIMENT); results are wrong when set
};

// Appends one code to a right-aligned 64-bit accumulator; whenever 32 bits are complete they go to the
// thread's private slot (row-interleaved: word j of thread t at s_slot[j * kEncThreads + t], so lanes
// never collide on a bank). `nbm` = bits in acc minus 32, so "a word is ready" is nbm >= 0 and the
// wrap-around after a flush is a single OR. No branches: the store and pointer bump are predicated.
#define HB_ENC_APPEND(code_, len_)                                                                                     \
    do {                                                                                                               \
        acc = (acc << (len_)) | (uint64_t)(code_);                                                                     \
        nbm += (int)(len_);                                                                                            \
        const bool full_ = nbm >= 0;                                                                                   \
        const uint32_t word_ = __funnelshift_r((uint32_t)acc, (uint32_t)(acc >> 32), (uint32_t)nbm);                   \
        if (full_) *sp = word_;                                                                                        \
        sp += full_ ? kEncThreads : 0;                                                                                 \
        nbm |= ~31;                                                                                                    \
    } while (0)

// One look-back descriptor covers a MACRO tile of kEncSub consecutive tiles: the chain of dependent
// L2 round trips a block waits for is paid once per 16 KiB of input instead of once per 4 KiB.
constexpr int kEncSub = 4;
constexpr int kEncMaskWords = kEncSub * kEncTile / 32;

struct EncShared {
    const uint2 *tab;      // [256] code table
    uint32_t *mask;        // [kEncMaskWords] item-start bits of the macro tile
    Seg *warp;             // [kEncThreads / 32]
    Seg *totals;           // [kEncSub] function of each tile of the macro tile
    uint64_t *pos;         // [1] absolute bit position of the macro tile
    uint32_t *stage;       // staging buffer of one tile's output
    uint32_t *slot;        // private packed words
    uint16_t *runbits;     // bit counts of finished runs
    uint16_t *obpos;       // stage byte where the item starting at a given symbol begins
};

template <bool kFull>
__device__ __forceinline__ void enc_load_symbols(const EncTiledArgs &a, uint64_t p0, uint32_t nsym, uint32_t (&w)[4]) {
    w[0] = w[1] = w[2] = w[3] = 0;
    if (kFull && ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0)) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(a.in + p0));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < kEncSymsPerThread; ++k)
            if ((uint32_t)k < nsym) w[k >> 2] |= (uint32_t)a.in[p0 + k] << (8 * (k & 3));
    }
}

// Phases A + B of one tile: code lengths -> my segment function -> block scan. Returns my exclusive
// function from the start of the tile; the tile's own function lands in sh.totals[sub].
template <bool kSeg, bool kFull>
__device__ __forceinline__ Seg enc_tile_measure(const EncTiledArgs &a, const EncShared &sh, uint32_t tile, uint32_t sub, uint32_t m) {
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint64_t t0 = (uint64_t)tile * kEncTile;
    const uint64_t t1 = kFull ? t0 + kEncTile : a.total_in;
    const uint64_t p0 = t0 + (uint64_t)tid * kEncSymsPerThread;
    const uint32_t nsym = kFull ? (uint32_t)kEncSymsPerThread
                                : (p0 >= t1 ? 0u : (uint32_t)min((uint64_t)kEncSymsPerThread, t1 - p0));
    uint32_t w[4];
    enc_load_symbols<kFull>(a, p0, nsym, w);

    Seg mine = {0, 0, 0};
    {
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < kEncSymsPerThread; ++k) {
            if (!kFull && (uint32_t)k >= nsym) break;
            if (kSeg && ((m >> k) & 1u)) {
                if (mine.hb) mine.tail += (run + 7u) & ~7u; else mine.head = run;
                mine.hb = 1;
                run = 0;
            }
            const uint32_t sym = (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
            run += sh.tab[sym].y;
        }
        if (mine.hb) mine.tail += run; else mine.head = run;
    }

    Seg excl;
    if (!kSeg) {
        // a single stream has no item starts: plain prefix sums of bit counts
        const uint32_t bits = mine.head;
        const uint32_t incl = warp_inclusive_scan(bits);
        uint32_t *warp_sums = reinterpret_cast<uint32_t *>(sh.warp);
        __syncthreads();  // sh.warp free again
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        uint32_t wsum = lane < kEncThreads / 32 ? warp_sums[lane] : 0u;
        uint32_t wincl = wsum;
#pragma unroll
        for (int d = 1; d < kEncThreads / 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, wincl, d);
            if (lane >= (uint32_t)d) wincl += up;
        }
        const uint32_t total_bits = __shfl_sync(0xffffffffu, wincl, kEncThreads / 32 - 1);
        const uint32_t before_my_warp = __shfl_sync(0xffffffffu, wincl - wsum, warp);
        excl = Seg{before_my_warp + incl - bits, 0, 0};
        if (tid == 0) sh.totals[sub] = Seg{total_bits, 0, 0};
        return excl;
    }
    Seg incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Seg up = seg_shfl_up(incl, d);
        if (lane >= (uint32_t)d) incl = seg_combine(up, incl);
    }
    excl = seg_shfl_up(incl, 1);
    if (lane == 0) excl = Seg{0, 0, 0};
    __syncthreads();  // sh.warp free again
    if (lane == 31) sh.warp[warp] = incl;
    __syncthreads();
    {
        // every warp redoes the 8-entry scan (cheaper than another barrier)
        Seg wv = lane < kEncThreads / 32 ? sh.warp[lane] : Seg{0, 0, 0};
        Seg wi = wv;
#pragma unroll
        for (int d = 1; d < kEncThreads / 32; d <<= 1) {
            const Seg up = seg_shfl_up(wi, d);
            if (lane >= (uint32_t)d) wi = seg_combine(up, wi);
        }
        Seg we = seg_shfl_up(wi, 1);
        if (lane == 0) we = Seg{0, 0, 0};
        Seg total;
        total.head = __shfl_sync(0xffffffffu, wi.head, kEncThreads / 32 - 1);
        total.tail = __shfl_sync(0xffffffffu, wi.tail, kEncThreads / 32 - 1);
        total.hb = __shfl_sync(0xffffffffu, wi.hb, kEncThreads / 32 - 1);
        Seg mywarp;
        mywarp.head = __shfl_sync(0xffffffffu, we.head, warp);
        mywarp.tail = __shfl_sync(0xffffffffu, we.tail, warp);
        mywarp.hb = __shfl_sync(0xffffffffu, we.hb, warp);
        excl = seg_combine(mywarp, excl);  // my exclusive function from the start of the tile
        if (tid == 0) sh.totals[sub] = total;
    }
    return excl;
}

// Phase C of one tile: pack my symbols into my private slot, run by run (position independent).
// A "run" is a maximal stretch of my symbols inside one item. Runs are packed left-aligned, each
// starting on a fresh slot word; finished runs leave their bit count in runbits[run][tid].
template <bool kSeg, bool kFull>
__device__ __forceinline__ void enc_tile_pack(
    const EncTiledArgs &a, const EncShared &sh, uint32_t tile, uint32_t m, uint32_t &runs_done, uint32_t &last_bits) {
    const uint32_t tid = threadIdx.x;
    const uint64_t t0 = (uint64_t)tile * kEncTile;
    const uint64_t t1 = kFull ? t0 + kEncTile : a.total_in;
    const uint64_t p0 = t0 + (uint64_t)tid * kEncSymsPerThread;
    const uint32_t nsym = kFull ? (uint32_t)kEncSymsPerThread
                                : (p0 >= t1 ? 0u : (uint32_t)min((uint64_t)kEncSymsPerThread, t1 - p0));
    uint32_t w[4];
    enc_load_symbols<kFull>(a, p0, nsym, w);

    uint64_t acc = 0;
    int nbm = -32;
    uint32_t *sp = sh.slot + tid;
    uint32_t *run_begin = sp;
    runs_done = 0;
#pragma unroll
    for (int k = 0; k < kEncSymsPerThread; ++k) {
        if (!kFull && (uint32_t)k >= nsym) break;
        if (kSeg && ((m >> k) & 1u)) {
            // an item starts at symbol k: close the current run
            const uint32_t rem = (uint32_t)(nbm + 32);
            const uint32_t bits = (uint32_t)(sp - run_begin) / kEncThreads * 32u + rem;
            if (rem) {
                *sp = (uint32_t)(acc << (32 - rem));
                sp += kEncThreads;
            }
            sh.runbits[runs_done * kEncThreads + tid] = (uint16_t)bits;
            ++runs_done;
            run_begin = sp;
            acc = 0;
            nbm = -32;
        }
        const uint32_t sym = (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
        const uint2 e = sh.tab[sym];
        HB_ENC_APPEND(e.x, e.y);
    }
    const uint32_t rem = (uint32_t)(nbm + 32);
    last_bits = (uint32_t)(sp - run_begin) / kEncThreads * 32u + rem;
    if (rem) *sp = (uint32_t)(acc << (32 - rem));
}

// Phases E.. of one tile whose absolute bit range [G, Gend) is known: shift-copy my packed runs into
// the stage, complete the last byte, copy the owned bytes out, record item start offsets, re-zero.
// Requires a block barrier between enc_tile_pack and this call.
template <bool kSeg, bool kFull>
__device__ __forceinline__ void enc_tile_emit(
    const EncTiledArgs &a, const EncShared &sh, uint32_t tile, uint32_t m, const Seg &excl, uint64_t G, uint64_t Gend,
    uint32_t runs_done, uint32_t last_bits) {
    const uint32_t tid = threadIdx.x;
    const uint64_t t0 = (uint64_t)tile * kEncTile;
    const uint64_t t1 = kFull ? t0 + kEncTile : a.total_in;
    uint32_t *s_stage = sh.stage;
    const uint64_t P = seg_apply(excl, G);  // absolute bit position of my first code

    // Staging origin: stage byte 0 <-> global address (out + G/8) rounded down to 16 bytes.
    const uint64_t g_byte0 = G >> 3;
    const uint32_t misalign = (uint32_t)((reinterpret_cast<uintptr_t>(a.out) + g_byte0) & 15);
    const int64_t origin_byte = (int64_t)g_byte0 - (int64_t)misalign;  // absolute output byte of stage byte 0

    {
        uint32_t D = (uint32_t)((int64_t)P - origin_byte * 8);  // stage bit index of the current run
        const uint32_t *src = sh.slot + tid;
        uint32_t boundary_bits = m;
        const uint32_t num_runs = runs_done + 1;
        for (uint32_t r = 0; r < num_runs; ++r) {
            const uint32_t bits = (r + 1 == num_runs) ? last_bits : (uint32_t)sh.runbits[r * kEncThreads + tid];
            if (kSeg && r > 0) {
                // byte-align: pad the item that just ended with the LOW bits of eos_padding
                // (reference huffman.c:178-184), then note where the new item starts
                const uint32_t pad = (8u - (D & 7u)) & 7u;
                if (pad) {
                    const uint32_t v = a.eos_padding & ((1u << pad) - 1u);
                    atomicOr(&s_stage[D >> 5], __byte_perm(v << (32 - (D & 31) - pad), 0, 0x0123));
                    D += pad;
                }
                const uint32_t k = (uint32_t)__ffs(boundary_bits) - 1;
                boundary_bits &= boundary_bits - 1;
                sh.obpos[tid * kEncSymsPerThread + k] = (uint16_t)(D >> 3);  // stage byte where the item starts
            }
            if (bits) {
                const uint32_t shift = D & 31;
                uint32_t *dst = s_stage + (D >> 5);
                const uint32_t nsrc = (bits + 31) >> 5;
                const uint32_t ndst = (shift + bits + 31) >> 5;
                uint32_t prev = 0;
                for (uint32_t j = 0; j < ndst; ++j) {
                    const uint32_t cur = j < nsrc ? src[j * kEncThreads] : 0u;
                    const uint32_t word = __byte_perm(__funnelshift_r(cur, prev, shift), 0, 0x0123);
                    if (j == 0 || j + 1 == ndst) atomicOr(dst + j, word); else dst[j] = word;
                    prev = cur;
                }
                src += nsrc * kEncThreads;
                D += bits;
            }
        }
    }

    // ---- the bits that complete this tile's last byte ------------------------------------------------
    uint32_t first_in_tile = 0, end_in_tile = 0;
    if (kSeg) {
        first_in_tile = a.tile_first[tile];
        end_in_tile = a.tile_first[tile + 1];
    }
    const bool last_tile = tile + 1 == a.num_tiles;
    if (tid == 0) {
        const uint32_t need = (8u - (uint32_t)(Gend & 7u)) & 7u;
        if (need) {
            uint64_t stop = a.total_in;  // where the open item ends
            if (kSeg && end_in_tile < a.n) stop = a.in_offsets[end_in_tile];
            uint32_t bits = 0, have = 0;
            for (uint64_t p = t1; have < need && p < stop; ++p) {
                const uint2 e = sh.tab[a.in[p]];
                const uint32_t take = min(e.y, need - have);
                bits = (bits << take) | (e.x >> (e.y - take));
                have += take;
            }
            if (have < need) {
                const uint32_t rem = need - have;
                bits = (bits << rem) | (a.eos_padding & ((1u << rem) - 1u));
            }
            const uint32_t sbyte = (uint32_t)((int64_t)(Gend >> 3) - origin_byte);
            atomicOr(&s_stage[sbyte >> 2], bits << (8 * (sbyte & 3)));
        }
        if (!kSeg && tile == 0) a.out_offsets[0] = 0;
        if (last_tile) {
            const uint64_t total_out = (Gend + 7) >> 3;
            a.out_offsets[a.n] = total_out;
            if (kSeg) {
                // trailing empty items start at total_in
                for (uint64_t i = end_in_tile; i < a.n; ++i) a.out_offsets[i] = total_out;
            }
        }
    }
    __syncthreads();

    // ---- copy the bytes this tile owns; item start offsets (coalesced) -----------------------------------
    {
        const uint64_t own_lo = (G + 7) >> 3, own_hi = min((Gend + 7) >> 3, a.out_capacity);
        if (own_hi > own_lo) {
            const uint32_t s_lo = (uint32_t)((int64_t)own_lo - origin_byte);
            const uint32_t s_hi = (uint32_t)((int64_t)own_hi - origin_byte);
            uint8_t *gbase = a.out + origin_byte;  // 16-byte aligned address of stage byte 0
            const uint8_t *sb = reinterpret_cast<const uint8_t *>(s_stage);
            const uint32_t v_lo = (s_lo + 15) & ~15u, v_hi = s_hi & ~15u;
            if (v_lo < v_hi) {
                for (uint32_t i = s_lo + tid; i < v_lo; i += kEncThreads) gbase[i] = sb[i];
                const uint4 *s4 = reinterpret_cast<const uint4 *>(s_stage);
                uint4 *g4 = reinterpret_cast<uint4 *>(gbase);
                for (uint32_t i = (v_lo >> 4) + tid; i < (v_hi >> 4); i += kEncThreads) g4[i] = s4[i];
                for (uint32_t i = v_hi + tid; i < s_hi; i += kEncThreads) gbase[i] = sb[i];
            } else {
                for (uint32_t i = s_lo + tid; i < s_hi; i += kEncThreads) gbase[i] = sb[i];
            }
        }
        if (kSeg) {
            for (uint32_t i = first_in_tile + tid; i < end_in_tile; i += kEncThreads) {
                const uint32_t p = (uint32_t)(a.in_offsets[i] - t0);
                a.out_offsets[i] = (uint64_t)(origin_byte + (int64_t)sh.obpos[p]);
            }
        }
    }
    // the stage must be all zero again for the next tile: clear exactly what this tile touched
    __syncthreads();
    {
        const uint32_t z_lo = (uint32_t)((int64_t)(G >> 3) - origin_byte) >> 4;
        const uint32_t z_hi = ((uint32_t)((int64_t)((Gend + 7) >> 3) - origin_byte) + 15) >> 4;
        uint4 *z = reinterpret_cast<uint4 *>(s_stage);
        for (uint32_t i = z_lo + tid; i < z_hi; i += kEncThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
}

constexpr int kEncSlotWords = kEncTile;  // 16 words per thread: every symbol fits one word
constexpr size_t kEncSmemBytes =
    kEncSlotWords * 4 + kEncStageBytes + kEncTile * 2 /* runbits */ + kEncTile * 2 /* item start positions */;

// Persistent blocks: the code table is loaded and the stage zeroed once; macro tiles come from an
// atomic ticket taken only when the block is ready to start (a held-but-idle ticket would stall every
// successor's look-back).
template <bool kSeg>
__global__ void __launch_bounds__(kEncThreads, 4) encode_tiled_kernel(const uint2 *__restrict__ enc_table, EncTiledArgs a) {
    __shared__ uint2 s_tab[256];
    __shared__ uint32_t s_mask[kEncMaskWords];
    __shared__ Seg s_warp[kEncThreads / 32];
    __shared__ Seg s_totals[kEncSub];
    __shared__ uint32_t s_ticket;
    __shared__ uint64_t s_pos[1];
    extern __shared__ __align__(16) uint8_t s_dyn[];
    EncShared sh;
    sh.tab = s_tab;
    sh.mask = s_mask;
    sh.warp = s_warp;
    sh.totals = s_totals;
    sh.pos = s_pos;
    sh.stage = reinterpret_cast<uint32_t *>(s_dyn);
    sh.slot = reinterpret_cast<uint32_t *>(s_dyn + kEncStageBytes);
    sh.runbits = reinterpret_cast<uint16_t *>(s_dyn + kEncStageBytes + kEncSlotWords * 4);
    sh.obpos = sh.runbits + kEncTile;

    const uint32_t tid = threadIdx.x;
    const uint32_t num_macro = (a.num_tiles + kEncSub - 1) / kEncSub;
    if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u);
    s_tab[tid] = enc_table[tid];
    {
        uint4 *z = reinterpret_cast<uint4 *>(sh.stage);
        for (uint32_t i = tid; i < kEncStageBytes / 16; i += kEncThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    uint32_t macro = s_ticket;
    while (macro < num_macro) {
        const uint32_t tile0 = macro * kEncSub;
        const uint32_t nsub = min((uint32_t)kEncSub, a.num_tiles - tile0);

        // ---- item starts inside the macro tile -> one bit per symbol ----------------------------------------
        uint32_t m[kEncSub];
#pragma unroll
        for (int s = 0; s < kEncSub; ++s) m[s] = 0;
        if (kSeg) {
            for (uint32_t i = tid; i < kEncMaskWords; i += kEncThreads) s_mask[i] = 0;
            __syncthreads();
            const uint64_t t0 = (uint64_t)tile0 * kEncTile;
            const uint32_t first = a.tile_first[tile0], end = a.tile_first[tile0 + nsub];
            for (uint32_t i = first + tid; i < end; i += kEncThreads) {
                const uint32_t p = (uint32_t)(a.in_offsets[i] - t0);
                atomicOr(&s_mask[p >> 5], 1u << (p & 31));
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < kEncSub; ++s)
                m[s] = (s_mask[s * (kEncTile / 32) + (tid >> 1)] >> ((tid & 1) * 16)) & 0xffffu;
        }

        // ---- A + B for every tile of the macro tile, then publish the macro tile's function ------------------
        Seg excl[kEncSub];
#pragma unroll
        for (int s = 0; s < kEncSub; ++s) {
            excl[s] = Seg{0, 0, 0};
            if ((uint32_t)s < nsub) {
                const uint32_t tile = tile0 + s;
                const bool full = (uint64_t)(tile + 1) * kEncTile <= a.total_in;
                excl[s] = full ? enc_tile_measure<kSeg, true>(a, sh, tile, s, m[s])
                               : enc_tile_measure<kSeg, false>(a, sh, tile, s, m[s]);
            }
        }
        __syncthreads();  // s_totals complete
        Seg macro_total = s_totals[0];
        for (uint32_t s = 1; s < nsub; ++s) macro_total = seg_combine(macro_total, s_totals[s]);
        if (tid == 0) seg_publish_aggregate(a.tile_state, macro, macro_total);

        // ---- C.. per tile; the macro tile's position is resolved after the first packing --------------------
        uint64_t G = 0;
#pragma unroll
        for (int s = 0; s < kEncSub; ++s) {
            if ((uint32_t)s < nsub) {
                const uint32_t tile = tile0 + s;
                const bool full = (uint64_t)(tile + 1) * kEncTile <= a.total_in;
                uint32_t runs_done, last_bits;
                if (full) enc_tile_pack<kSeg, true>(a, sh, tile, m[s], runs_done, last_bits);
                else enc_tile_pack<kSeg, false>(a, sh, tile, m[s], runs_done, last_bits);
                if (s == 0) {
                    if (tid < 32) {
                        const uint64_t G0 = (a.debug & 1u) ? (uint64_t)macro * kEncSub * kEncTile * 6
                                                           : seg_resolve(a.tile_state, macro, macro_total);
                        if (tid == 0) s_pos[0] = G0;
                    }
                    __syncthreads();
                    G = s_pos[0];
                } else {
                    __syncthreads();  // the previous tile's stage clearing is complete
                }
                const uint64_t Gend = seg_apply(s_totals[s], G);
                if (full) enc_tile_emit<kSeg, true>(a, sh, tile, m[s], excl[s], G, Gend, runs_done, last_bits);
                else enc_tile_emit<kSeg, false>(a, sh, tile, m[s], excl[s], G, Gend, runs_done, last_bits);
                G = Gend;
            }
        }
        if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u);
        __syncthreads();
        macro = s_ticket;
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
