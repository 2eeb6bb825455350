// Batch encoder for MANY SHORT STRINGS (BASELINE configs 2 and 5): one thread per string, two kernels.
//
// Every item of a packed batch starts on a byte boundary of the output (the previous item is EOS-padded,
// huffman.c:178-184), so the only thing an item needs from the others is a BYTE offset. That splits the
// work into
//
//   str_measure_kernel   encoded length of every string (huffman.c:107-129: sum of code lengths, rounded up
//                        to bytes), block scan + single-pass decoupled look-back -> out_offsets[0..n];
//                        also cuts the OUTPUT into tiles of kStrTileBytes (tile k = the strings that start
//                        in output bytes [kT, (k+1)T)) and records each tile's first string
//   str_pack_kernel      per output tile: every thread packs whole strings with the carry-flag append of
//                        encode_tiled.cuh (one multiply-add, one add, one predicated store per symbol)
//                        straight into a shared-memory image of the tile at the string's own byte offset;
//                        the image leaves with coalesced 128-bit stores
//
// and removes what the slot/tile encoders pay for generality: no segment functions, no slot maps, no
// two-piece stage whose position depends on a look-back, no scout warp, no partially filled slots; the
// symbols are read straight from global memory (aligned 128-bit loads per thread: 16 symbols per load).
//
// Balance: strings of a tile are counting-sorted by length and dealt out in zigzag order (thread t takes the
// t-th longest and the t-th shortest of every 512), so lanes of a warp run strings of almost the same
// length and every thread gets about the same number of symbols.
//
// Words shared by two strings: a thread stores the words it COMPLETES with plain stores (the leading bytes
// of its first word are preloaded from the image: zero, or what an earlier sub-batch left there) and keeps
// its last, partial word back; after a barrier the partial words are ORed in (one shared-memory atomic per
// string). No other atomics, no read-modify-write in the packing loop.
//
// Strings whose encoding exceeds kStrSlackBytes (or tiles whose input spans >= 4 GiB) raise a flag in the
// control block: str_pack_kernel then does nothing and the gated launches of encode_tiled_kernel<true>
// behind it redo the batch (no host round trip). Results are bit-identical either way.
#pragma once

#include "encode_tiled.cuh"

namespace hb {

#ifndef HB_STR_TILE_BYTES
#define HB_STR_TILE_BYTES (40 * 1024)
#endif
#ifndef HB_STR_TAB_COPIES
#define HB_STR_TAB_COPIES 8
#endif
#ifndef HB_STR_PACK_BLOCKS
#define HB_STR_PACK_BLOCKS 3
#endif

constexpr int kStrThreads = 256;
constexpr int kStrWarps = kStrThreads / 32;
constexpr int kStrBatch = 2 * kStrThreads;                   // strings sorted and dealt out together
constexpr uint32_t kStrTileBytes = HB_STR_TILE_BYTES;        // T: output bytes per tile (multiple of 16)
constexpr uint32_t kStrSlackBytes = 8 * 1024;                // longest encoded string this path takes
constexpr uint32_t kStrMaxInput = 8 * kStrSlackBytes;       // (>= 1 bit per symbol) longer strings cannot fit
constexpr uint32_t kStrStageBytes = kStrTileBytes + kStrSlackBytes + 32;
constexpr int kStrTabCopies = HB_STR_TAB_COPIES;             // copies of the code table (bank spreading)
constexpr uint32_t kStrTabStride = 8 * kStrTabCopies;        // bytes between entries of one copy
constexpr uint32_t kStrTabBytes = 257 * kStrTabStride;       // entry 256 = {0, 0}: "no symbol"
static_assert(kStrTileBytes % 16 == 0, "tiles start on 16-byte boundaries of the output");

// control block (device memory, zeroed before str_measure_kernel)
enum : int { kStrCtlTicket = 0, kStrCtlFallback = 1, kStrCtlNumTiles = 2, kStrCtlWords = 8 };

struct StrArgs {
    const uint8_t *in;           // 16-byte aligned
    const uint64_t *in_offsets;  // n + 1
    uint64_t n;
    uint64_t total_in;
    uint8_t *out;
    uint64_t out_capacity;
    uint64_t *out_offsets;       // n + 1, written by str_measure_kernel
    uint64_t *tile_state;        // look-back descriptors of str_measure_kernel (32-byte aligned, zeroed)
    uint32_t *control;           // kStrCtlWords words, zeroed
    uint32_t *tile_first;        // output tiles: first string of tile k; [0] zeroed
    uint32_t num_measure_tiles;
    uint32_t out_phase;          // out & 15: tiles are cut in the address space of `out` rounded down to 16
    uint32_t eos_padding;
};

__device__ __forceinline__ uint32_t str_bucket(uint32_t len) { return 255u - min((len + 3u) >> 2, 255u); }

// Block-wide counting sort of cnt <= kStrBatch strings by decreasing length (4-byte buckets).
// s_perm[p] = index of the p-th longest string. Ends with a barrier.
__device__ __forceinline__ void str_sort(
    const uint32_t *s_len, uint32_t cnt, uint16_t *s_perm, uint32_t *s_hist, uint32_t *s_wsum) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_hist[tid] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < cnt; i += kStrThreads) atomicAdd(&s_hist[str_bucket(s_len[i])], 1u);
    __syncthreads();
    const uint32_t v = s_hist[tid];
    const uint32_t incl = warp_inclusive_scan(v);
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < kStrWarps; ++w)
        if ((uint32_t)w < warp) before += s_wsum[w];
    s_hist[tid] = before + incl - v;
    __syncthreads();
    for (uint32_t i = tid; i < cnt; i += kStrThreads) s_perm[atomicAdd(&s_hist[str_bucket(s_len[i])], 1u)] = (uint16_t)i;
    __syncthreads();
}

// Walks one string in aligned 16-byte vectors: f.masked(v, lo, hi) for the first and the last vector (bytes
// [lo, hi) of it belong to the string), f.full(v) for the ones in between. `nsafe` = vectors from the
// aligned start of the string that lie completely inside the input buffer (the others are read byte-wise).
template <class F>
__device__ __forceinline__ void str_walk(const uint8_t *p, uint32_t len, const uint8_t *in_end, F &f) {
    if (len == 0) return;
    const uint32_t r = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15);
    const uint8_t *v0 = p - r;
    const uint4 *vp = reinterpret_cast<const uint4 *>(v0);
    const uint32_t span = r + len;
    const uint32_t nvec = (span + 15u) >> 4;
    const uint64_t nsafe = (uint64_t)(in_end - v0) >> 4;
    auto load = [&](uint32_t j) -> uint4 {
        if (j < nsafe) return __ldg(vp + j);
        uint32_t w[4] = {0, 0, 0, 0};
        const uint8_t *a = v0 + 16ull * j;
        for (int b = 0; b < 16; ++b)
            if (a + b < in_end) w[b >> 2] |= (uint32_t)a[b] << (8 * (b & 3));
        return make_uint4(w[0], w[1], w[2], w[3]);
    };
    uint4 cur = load(0);
    uint4 nxt = cur;
    if (nvec > 1) nxt = load(1);
    f.masked(cur, r, min(16u, span));
    for (uint32_t j = 2; j < nvec; ++j) {
        cur = nxt;
        nxt = load(j);
        f.full(cur);
    }
    if (nvec > 1) f.masked(nxt, 0u, span - 16u * (nvec - 1));
}

__device__ __forceinline__ uint32_t str_word(const uint4 &v, int k) {
    return (k >> 2) == 0 ? v.x : (k >> 2) == 1 ? v.y : (k >> 2) == 2 ? v.z : v.w;
}

// ---- measure ------------------------------------------------------------------------------------------------
struct StrMeasure {
    uint32_t bits;
    uint32_t tab;  // shared-window address of the 256-byte code length table
    __device__ __forceinline__ uint32_t len_of(uint32_t word, int k) const {
        const uint32_t byte = __byte_perm(word, 0, 0x4440 | (k & 3));
        uint32_t l;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(l) : "r"(tab + byte));
        return l;
    }
    __device__ __forceinline__ void full(const uint4 &v) {
        uint32_t s0 = 0, s1 = 0;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            s0 += len_of(str_word(v, k), k);
            s1 += len_of(str_word(v, k + 1), k + 1);
        }
        bits += s0 + s1;
    }
    __device__ __forceinline__ void masked(const uint4 &v, uint32_t lo, uint32_t hi) {
        const uint32_t width = hi - lo;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t l = len_of(str_word(v, k), k);
            bits += ((uint32_t)k - lo < width) ? l : 0u;
        }
    }
};

__global__ void __launch_bounds__(kStrThreads, 4) str_measure_kernel(const uint2 *__restrict__ enc_table, StrArgs a) {
    __shared__ __align__(16) uint8_t s_lentab[256];
    __shared__ uint32_t s_inrel[kStrBatch];
    __shared__ uint32_t s_len[kStrBatch];
    __shared__ uint32_t s_bytes[kStrBatch];
    __shared__ uint16_t s_perm[kStrBatch];
    __shared__ uint32_t s_hist[kStrThreads];
    __shared__ uint32_t s_wsum[kStrWarps];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_prefix;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_lentab[tid] = (uint8_t)enc_table[tid].y;
    const uint8_t *const in_end = a.in + a.total_in;
    StrMeasure m;
    m.tab = smem_addr(s_lentab);

    for (;;) {
        __syncthreads();  // previous tile done with the shared arrays (first trip: the table is in place)
        if (tid == 0) s_tile = atomicAdd(&a.control[kStrCtlTicket], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= a.num_measure_tiles) break;
        const uint64_t item0 = (uint64_t)tile * kStrBatch;
        const uint32_t cnt = (uint32_t)min((uint64_t)kStrBatch, a.n - item0);
        const uint64_t base0 = a.in_offsets[item0];
        bool wide = false;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t idx = tid + q * kStrThreads;
            if (idx < cnt) {
                const uint64_t o0 = a.in_offsets[item0 + idx], o1 = a.in_offsets[item0 + idx + 1];
                // (a string this path does not take is measured as empty: the flag sends the batch elsewhere)
                const bool big = (o1 - base0) >= (1ull << 32) || (o1 - o0) > kStrMaxInput;
                wide |= big;
                s_inrel[idx] = big ? 0u : (uint32_t)(o0 - base0);
                s_len[idx] = big ? 0u : (uint32_t)(o1 - o0);
            }
        }
        if (wide) atomicOr(&a.control[kStrCtlFallback], 1u);
        str_sort(s_len, cnt, s_perm, s_hist, s_wsum);
        // zigzag: the tid-th longest, then the tid-th shortest of the 512
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const uint32_t pos = q == 0 ? tid : (uint32_t)kStrBatch - 1u - tid;
            if (pos < cnt) {
                const uint32_t idx = s_perm[pos];
                m.bits = 0;
                str_walk(a.in + base0 + s_inrel[idx], s_len[idx], in_end, m);
                s_bytes[idx] = (m.bits + 7u) >> 3;
            }
        }
        __syncthreads();
        // block scan in item order, two items per thread
        const uint32_t b0 = 2 * tid < cnt ? s_bytes[2 * tid] : 0u;
        const uint32_t b1 = 2 * tid + 1 < cnt ? s_bytes[2 * tid + 1] : 0u;
        const uint32_t incl = warp_inclusive_scan(b0 + b1);
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kStrWarps; ++w) {
            const uint32_t ws = s_wsum[w];
            if ((uint32_t)w < warp) before += ws;
            total += ws;
        }
        if (tid == 0) lookback_publish_aggregate(a.tile_state, tile, total);
        if (warp == 0) {
            const uint64_t prefix = lookback_resolve(a.tile_state, tile, total);
            if (lane == 0) s_prefix = prefix;
        }
        __syncthreads();
        const uint64_t prefix = s_prefix;
        uint64_t o = prefix + before + incl - (b0 + b1);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t idx = 2 * tid + q;
            const uint32_t bytes = q == 0 ? b0 : b1;
            if (idx < cnt) {
                const uint64_t item = item0 + idx;
                a.out_offsets[item] = o;
                if (bytes > kStrSlackBytes) atomicOr(&a.control[kStrCtlFallback], 1u);
                // output tiles: the next string opens tile k1 when this one crosses into it
                const uint64_t k0 = (o + a.out_phase) / kStrTileBytes, k1 = (o + bytes + a.out_phase) / kStrTileBytes;
                if (a.tile_first && bytes <= kStrSlackBytes) {
                    if (k1 != k0) a.tile_first[k1] = (uint32_t)(item + 1);
                    if (item + 1 == a.n) {
                        a.tile_first[k1 + 1] = (uint32_t)a.n;
                        a.control[kStrCtlNumTiles] = (uint32_t)(k1 + 1);
                    }
                }
                if (item + 1 == a.n) a.out_offsets[a.n] = o + bytes;
            }
            o += bytes;
        }
    }
}

// ---- pack ---------------------------------------------------------------------------------------------------
struct StrPack {
    uint32_t sp, acc, nb;
    uint32_t tab;   // shared-window address of MY copy's entry 0
    uint32_t zero;  // ... of my copy's entry 256 ({0, 0})
    __device__ __forceinline__ static uint2 fetch(uint32_t addr) {
        uint2 e;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(addr));
        return e;
    }
    __device__ __forceinline__ uint32_t entry(uint32_t word, int k) const {
        const uint32_t byte = __byte_perm(word, 0, 0x4440 | (k & 3));
        uint32_t addr;
        asm("mad.lo.u32 %0, %1, %3, %2;" : "=r"(addr) : "r"(byte), "r"(tab), "n"(kStrTabStride));
        return addr;
    }
    // (all sixteen lookups first: the loads are in flight together, the appends are one dependent chain)
    __device__ __forceinline__ void full(const uint4 &v) {
        uint2 e[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = fetch(entry(str_word(v, k), k));
#pragma unroll
        for (int k = 0; k < 16; ++k) enc_append(sp, acc, nb, e[k].x, e[k].y);
    }
    __device__ __forceinline__ void masked(const uint4 &v, uint32_t lo, uint32_t hi) {
        const uint32_t width = hi - lo;
        uint2 e[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t addr = entry(str_word(v, k), k);
            e[k] = fetch(((uint32_t)k - lo < width) ? addr : zero);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) enc_append(sp, acc, nb, e[k].x, e[k].y);
    }
};

constexpr size_t kStrPackSmemBytes = kStrStageBytes + kStrTabBytes;

__global__ void __launch_bounds__(kStrThreads, HB_STR_PACK_BLOCKS) str_pack_kernel(const uint2 *__restrict__ enc_table, StrArgs a) {
    __shared__ uint32_t s_inrel[kStrBatch];
    __shared__ uint32_t s_len[kStrBatch];
    __shared__ uint32_t s_orel[kStrBatch];
    __shared__ uint16_t s_perm[kStrBatch];
    __shared__ uint32_t s_hist[kStrThreads];
    __shared__ uint32_t s_wsum[kStrWarps];
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint32_t *const stage = reinterpret_cast<uint32_t *>(s_dyn);  // big-endian words of the tile's output
    const uint32_t stage_addr = smem_addr(stage);
    const uint32_t tab0 = stage_addr + kStrStageBytes;

    if (a.control[kStrCtlFallback] != 0) return;
    const uint32_t num_tiles = a.control[kStrCtlNumTiles];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    {
        const uint2 e = enc_table[tid];
#pragma unroll
        for (int j = 0; j < kStrTabCopies; ++j)
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(tab0 + tid * kStrTabStride + j * 8), "r"(e.x),
                         "r"(enc_len_fields(e.y))
                         : "memory");
        if (tid < kStrTabCopies)
            asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(tab0 + 256 * kStrTabStride + tid * 8), "r"(0u) : "memory");
    }
    StrPack pk;
    pk.tab = tab0 + (lane & (kStrTabCopies - 1)) * 8;
    pk.zero = pk.tab + 256 * kStrTabStride;
    const uint8_t *const in_end = a.in + a.total_in;
    uint8_t *const out_al = a.out - a.out_phase;  // 16-byte aligned
    const uint64_t cap_al = a.out_capacity + a.out_phase;

    for (uint32_t k = blockIdx.x; k < num_tiles; k += gridDim.x) {
        const uint32_t f0 = a.tile_first[k], f1 = a.tile_first[k + 1];
        if (f0 >= f1) continue;
        const uint64_t tile0 = (uint64_t)k * kStrTileBytes;  // position of stage byte 0 in the aligned output space
        const uint32_t lo = (uint32_t)(a.out_offsets[f0] + a.out_phase - tile0);
        const uint32_t hi = (uint32_t)(a.out_offsets[f1] + a.out_phase - tile0);
        __syncthreads();  // the previous tile has left the stage (first trip: the table is in place)
        {
            uint4 *z = reinterpret_cast<uint4 *>(s_dyn);
            const uint32_t z0 = lo >> 4, z1 = (hi + 15u) >> 4;
            for (uint32_t i = z0 + tid; i < z1; i += kStrThreads) z[i] = make_uint4(0, 0, 0, 0);
        }
        for (uint32_t s0 = f0; s0 < f1; s0 += kStrBatch) {
            const uint32_t cnt = min((uint32_t)kStrBatch, f1 - s0);
            const uint64_t base0 = a.in_offsets[s0];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t idx = tid + q * kStrThreads;
                if (idx < cnt) {
                    const uint64_t o0 = a.in_offsets[s0 + idx], o1 = a.in_offsets[s0 + idx + 1];
                    s_inrel[idx] = (uint32_t)(o0 - base0);
                    s_len[idx] = (uint32_t)(o1 - o0);
                    s_orel[idx] = (uint32_t)(a.out_offsets[s0 + idx] + a.out_phase - tile0);
                }
            }
            str_sort(s_len, cnt, s_perm, s_hist, s_wsum);  // (its barriers also cover the zeroing above)
            uint32_t tail_addr[2], tail_word[2];
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {
                tail_addr[q] = 0;
                tail_word[q] = 0;
                const uint32_t pos = q == 0 ? tid : (uint32_t)kStrBatch - 1u - tid;
                if (pos < cnt) {
                    const uint32_t idx = s_perm[pos];
                    const uint32_t orel = s_orel[idx];
                    const uint32_t lead = 8u * (orel & 3u);  // bits of my first word that belong to earlier strings
                    pk.sp = stage_addr + (orel & ~3u);
                    uint32_t first;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(first) : "r"(pk.sp));
                    pk.acc = lead ? first >> (32u - lead) : 0u;
                    pk.nb = enc_len_fields(lead);
                    str_walk(a.in + base0 + s_inrel[idx], s_len[idx], in_end, pk);
                    // pad the last byte with the LOW bits of eos_padding (huffman.c:178-184)
                    const uint32_t pad = (0u - pk.nb) & 7u;
                    enc_append(pk.sp, pk.acc, pk.nb, a.eos_padding & ((1u << pad) - 1u), enc_len_fields(pad));
                    const uint32_t rem = pk.nb >> 27;
                    if (rem) {
                        tail_addr[q] = pk.sp;
                        tail_word[q] = pk.acc << (32u - rem);
                    }
                }
            }
            __syncthreads();  // every completed word is stored
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (tail_addr[q]) atomicOr(stage + ((tail_addr[q] - stage_addr) >> 2), tail_word[q]);
            __syncthreads();
        }
        // ---- the image leaves: bytes [lo, hi) of the stage, clipped to the capacity -------------------------
        const uint64_t room = cap_al > tile0 ? cap_al - tile0 : 0;
        const uint32_t hi_c = (uint32_t)min((uint64_t)hi, room);
        if (lo < hi_c) {
            uint8_t *const dst = out_al + tile0;
            const uint32_t v0 = (lo + 15u) >> 4, v1 = hi_c >> 4;
            const uint4 *sv = reinterpret_cast<const uint4 *>(s_dyn);
            uint4 *dv = reinterpret_cast<uint4 *>(dst);
            for (uint32_t v = v0 + tid; v < v1; v += kStrThreads) {
                uint4 w = sv[v];
                w.x = __byte_perm(w.x, 0, 0x0123);
                w.y = __byte_perm(w.y, 0, 0x0123);
                w.z = __byte_perm(w.z, 0, 0x0123);
                w.w = __byte_perm(w.w, 0, 0x0123);
                dv[v] = w;
            }
            // edge bytes: before the first whole vector and after the last one (or all of a short range)
            const uint32_t e_lo = v0 < v1 ? 16u * v0 : hi_c, e_hi = v0 < v1 ? 16u * v1 : hi_c;
            const uint32_t nhead = e_lo - lo, nedge = nhead + (hi_c - e_hi);
            if (tid < nedge) {
                const uint32_t B = tid < nhead ? lo + tid : e_hi + (tid - nhead);
                dst[B] = (uint8_t)(stage[B >> 2] >> (24u - 8u * (B & 3u)));
            }
        }
    }
}

}  // namespace hb
