// Batch encoder for MANY SHORT STRINGS (BASELINE configs 2 and 5): one thread per string, two kernels.
//
// Every item of a packed batch starts on a byte boundary of the output (the previous item is EOS-padded,
// huffman.c:178-184), so the only thing an item needs from the others is a BYTE offset. That splits the
// work into
//
//   str_bits_kernel      code-length sums (huffman.c:107-129) over FLAT 16 KiB tiles of the input, prefix sums
//                        left in shared memory; a string's bits are two look-ups (str_prep_kernel zeroes
//                        the scratch and finds each flat tile's first string)
//   str_scan_kernel      bits -> bytes, block scan + single-pass decoupled look-back -> out_offsets[0..n];
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
#define HB_STR_TAB_COPIES 16
#endif
#ifndef HB_STR_PREFETCH2
#define HB_STR_PREFETCH2 0
#endif
#ifndef HB_STR_PACK_BLOCKS
#define HB_STR_PACK_BLOCKS 2
#endif

constexpr int kStrThreads = 256;  // (the sort's 256 buckets and the scan's tiles are laid out for this)
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

// control block (device memory, zeroed by str_prep_kernel)
enum : int { kStrCtlTicket = 0, kStrCtlFallback = 1, kStrCtlNumTiles = 2, kStrCtlWords = 8 };

struct StrArgs {
    const uint8_t *in;           // 16-byte aligned
    const uint64_t *in_offsets;  // n + 1
    uint64_t n;
    uint64_t total_in;
    uint8_t *out;
    uint64_t out_capacity;
    uint64_t *out_offsets;       // n + 1, written by str_scan_kernel
    uint64_t *tile_state;        // look-back descriptors of str_scan_kernel (32-byte aligned, zeroed by str_prep_kernel)
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
#if HB_STR_PREFETCH2
    // two vectors in flight ahead of the one being packed (the packing loop's top stall was the wait for its loads)
    uint4 cur = load(0);
    uint4 n1 = cur, n2 = cur;
    if (nvec > 1) n1 = load(1);
    if (nvec > 2) n2 = load(2);
    f.masked(cur, r, min(16u, span));
    for (uint32_t j = 1; j + 1 < nvec; ++j) {
        cur = n1;
        n1 = n2;
        if (j + 2 < nvec) n2 = load(j + 2);
        f.full(cur);
    }
    if (nvec > 1) f.masked(n1, 0u, span - 16u * (nvec - 1));
#else
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
#endif
}

__device__ __forceinline__ uint32_t str_word(const uint4 &v, int k) {
    return (k >> 2) == 0 ? v.x : (k >> 2) == 1 ? v.y : (k >> 2) == 2 ? v.z : v.w;
}

// ---- measure ------------------------------------------------------------------------------------------------
// The encoded length of a string is a plain SUM over its bytes, so the measuring does not have to follow the
// strings: the input is cut into flat tiles of kBitsTileBytes, every thread sums whole aligned 16-byte
// vectors (coalesced 128-bit loads, no sorting, no partial vectors, every lane busy) and leaves the running
// sums behind in shared memory — per vector the inclusive prefix after each of its 16 bytes (u16), per tile the
// exclusive prefix over its vectors. The bits of a string are then cum(end) - cum(start), two look-ups; a
// string that crosses tiles collects its pieces with atomicAdd. (First version: one thread per string with
// the sorted walk of the pack kernel: 89 us for 1M strings; this one: see profiles/README.md.)
#ifndef HB_BITS_THREADS
#define HB_BITS_THREADS 256
#endif
constexpr int kBitsThreads = HB_BITS_THREADS;
constexpr int kBitsVecsPerThread = 4;
constexpr int kBitsTileVecs = kBitsThreads * kBitsVecsPerThread;  // 1024
constexpr uint32_t kBitsTileBytes = 16u * kBitsTileVecs;          // 16 KiB of input per tile

struct StrPrepArgs {
    const uint64_t *in_offsets;
    uint64_t n, total_in;
    uint32_t *bits;             // n (rounded up to 4) words, zeroed here
    uint32_t *bits_tile_first;  // num_bits_tiles + 1: first string that starts at or after the tile's first byte
    uint64_t *tile_state;       // num_scan_tiles descriptors, zeroed here
    uint32_t *control;          // zeroed here
    uint32_t *tile_first;       // [0] zeroed here
    uint32_t num_bits_tiles, num_scan_tiles;
};

// One launch instead of three memsets and an index kernel. The flat tiles' first strings are SCATTERED by the
// strings themselves (string i opens every tile whose first byte lies in (in_offsets[i-1], in_offsets[i]]): one
// coalesced pass over the offsets instead of a 20-step binary search per tile (12 us of dependent loads).
__global__ void str_prep_kernel(StrPrepArgs a) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = g; i < a.num_scan_tiles; i += stride) a.tile_state[i] = 0;
    if (g < kStrCtlWords) a.control[g] = 0;
    if (g == 0) a.tile_first[0] = 0;
    for (uint64_t i = g; i <= a.n; i += stride) {
        if (i < a.n) a.bits[i] = 0;
        const uint64_t cur = a.in_offsets[i];
        // tiles j with prev < j * T <= cur (prev = "-1" for the first string)
        uint64_t j = i ? a.in_offsets[i - 1] / kBitsTileBytes + 1 : 0;
        const bool last = (i == 0 || a.in_offsets[i - 1] < a.total_in) && a.total_in <= cur;
        for (; j * kBitsTileBytes <= cur && j < a.num_bits_tiles; ++j) a.bits_tile_first[j] = (uint32_t)i;
        if (last) a.bits_tile_first[a.num_bits_tiles] = (uint32_t)i;
    }
}

struct StrBitsArgs {
    const uint8_t *in;  // 16-byte aligned
    const uint64_t *in_offsets;
    uint64_t n, total_in;
    const uint32_t *bits_tile_first;
    uint32_t *bits;
    uint32_t *control;
    uint32_t num_bits_tiles;
};

__global__ void __launch_bounds__(kBitsThreads, 1024 / kBitsThreads) str_bits_kernel(const uint2 *__restrict__ enc_table, StrBitsArgs a) {
    __shared__ __align__(16) uint8_t s_lentab[256];
    __shared__ __align__(16) uint32_t s_pre[kBitsTileVecs][8];  // per vector: inclusive prefix after byte 2j | after byte 2j+1 << 16
    __shared__ uint32_t s_vex[kBitsTileVecs + 1];               // exclusive prefix over the vectors; [nvec] = the tile's total
    // (16-byte aligned: the compiler reads the warp sums as vectors, and a vector that started in s_vex's last word
    // made compute-sanitizer's racecheck flag that word — written in the same phase, never used by the reader)
    __shared__ __align__(16) uint32_t s_wsum[kBitsThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < 256; i += kBitsThreads) s_lentab[i] = (uint8_t)enc_table[i].y;

    // bits of the tile's bytes [0, x)
    auto cum = [&](uint32_t x) -> uint32_t {
        const uint32_t v = x >> 4, k = x & 15u;
        uint32_t c = s_vex[v];
        if (k) {
            const uint32_t w = s_pre[v][(k - 1) >> 1];
            c += ((k - 1) & 1u) ? w >> 16 : w & 0xffffu;
        }
        return c;
    };

    // the tile's vectors of this thread (zeros past the input's end)
    uint4 v[kBitsVecsPerThread];
    auto fetch = [&](uint32_t tile) {
        const uint64_t t0 = (uint64_t)tile * kBitsTileBytes;
        const uint32_t tile_len = (uint32_t)min((uint64_t)kBitsTileBytes, a.total_in - t0);
        const uint32_t nvec = (tile_len + 15u) >> 4;
        const uint4 *vp = reinterpret_cast<const uint4 *>(a.in + t0);
#pragma unroll
        for (int i = 0; i < kBitsVecsPerThread; ++i) {
            const uint32_t vi = i * kBitsThreads + tid;
            v[i] = make_uint4(0, 0, 0, 0);
            if (16u * vi + 16u <= tile_len) {
                v[i] = __ldg(vp + vi);
            } else if (vi < nvec) {  // the input's last, partial vector
                uint32_t w[4] = {0, 0, 0, 0};
                for (uint32_t b = 0; 16u * vi + b < tile_len; ++b) w[b >> 2] |= (uint32_t)a.in[t0 + 16u * vi + b] << (8 * (b & 3));
                v[i] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    };
    if (blockIdx.x < a.num_bits_tiles) fetch(blockIdx.x);
    for (uint32_t tile = blockIdx.x; tile < a.num_bits_tiles; tile += gridDim.x) {
        const uint64_t t0 = (uint64_t)tile * kBitsTileBytes;
        const uint32_t tile_len = (uint32_t)min((uint64_t)kBitsTileBytes, a.total_in - t0);
        __syncthreads();  // the previous tile's look-ups are over (first trip: the table is in place)
#pragma unroll
        for (int i = 0; i < kBitsVecsPerThread; ++i) {
            const uint32_t vi = i * kBitsThreads + tid;
            uint32_t run = 0, pk[8];
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const uint32_t w = str_word(v[i], k);
                const uint32_t l0 = s_lentab[__byte_perm(w, 0, 0x4440 | (k & 3))];
                const uint32_t l1 = s_lentab[__byte_perm(w, 0, 0x4440 | ((k + 1) & 3))];
                const uint32_t p0 = run + l0;
                run = p0 + l1;
                pk[k >> 1] = p0 | (run << 16);
            }
            *reinterpret_cast<uint4 *>(&s_pre[vi][0]) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4 *>(&s_pre[vi][4]) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            s_vex[vi] = run;  // (raw totals first; scanned below)
        }
        // the next tile's vectors are requested now (the registers are free again): their latency passes behind
        // the scan and the strings' look-ups instead of in front of the next tile's work
        if (tile + gridDim.x < a.num_bits_tiles) fetch(tile + gridDim.x);
        __syncthreads();
        // exclusive scan over the vectors in order: four consecutive vectors per thread
        uint32_t q[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            q[j] = s_vex[4 * tid + j];
            sum += q[j];
        }
        const uint32_t incl = warp_inclusive_scan(sum);
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
#pragma unroll
        for (int w = 0; w < kBitsThreads / 32; ++w)
            if ((uint32_t)w < warp) run += s_wsum[w];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s_vex[4 * tid + j] = run;
            run += q[j];
        }
        if (tid == kBitsThreads - 1) s_vex[kBitsTileVecs] = run;
        __syncthreads();
        // the strings that start in this tile, and the one that continues into it
        const uint32_t f0 = a.bits_tile_first[tile], f1 = a.bits_tile_first[tile + 1];
        if (tid == 0 && f0 > 0) {
            const uint64_t endc = a.in_offsets[f0];
            if (endc > t0) atomicAdd(&a.bits[f0 - 1], cum((uint32_t)min(endc - t0, (uint64_t)tile_len)));
        }
        for (uint32_t i = f0 + tid; i < f1; i += kBitsThreads) {
            const uint64_t s = a.in_offsets[i] - t0, e = a.in_offsets[i + 1] - t0;
            if (e - s > kStrMaxInput) atomicOr(&a.control[kStrCtlFallback], 1u);  // (its bit count may not fit 32 bits)
            if (e <= tile_len) a.bits[i] = cum((uint32_t)e) - cum((uint32_t)s);
            else atomicAdd(&a.bits[i], cum(tile_len) - cum((uint32_t)s));
        }
    }
}

// bits -> bytes -> out_offsets (block scan + single-pass decoupled look-back), the output tiles of
// str_pack_kernel, and the flag that sends batches with strings it does not take to the tiled kernel.
constexpr int kStrScanPerThread = 8;
constexpr int kStrScanTile = kStrThreads * kStrScanPerThread;  // 2048 strings per block

__global__ void __launch_bounds__(kStrThreads) str_scan_kernel(const uint32_t *__restrict__ bits, StrArgs a) {
    __shared__ uint32_t s_wsum[kStrWarps];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_prefix;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&a.control[kStrCtlTicket], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t item0 = (uint64_t)tile * kStrScanTile + (uint64_t)kStrScanPerThread * tid;
    uint32_t by[kStrScanPerThread];
    uint32_t sum = 0;
    bool big = false;
    if (item0 + kStrScanPerThread <= a.n) {  // (the scratch is 16-byte aligned and a multiple of 8 strings from it)
        const uint4 lo = *reinterpret_cast<const uint4 *>(bits + item0), hi = *reinterpret_cast<const uint4 *>(bits + item0 + 4);
        by[0] = lo.x; by[1] = lo.y; by[2] = lo.z; by[3] = lo.w; by[4] = hi.x; by[5] = hi.y; by[6] = hi.z; by[7] = hi.w;
    } else {
#pragma unroll
        for (int q = 0; q < kStrScanPerThread; ++q) by[q] = item0 + q < a.n ? bits[item0 + q] : 0u;
    }
#pragma unroll
    for (int q = 0; q < kStrScanPerThread; ++q) {
        by[q] = (by[q] + 7u) >> 3;
        big |= by[q] > kStrSlackBytes;
        sum += by[q];
    }
    if (big) atomicOr(&a.control[kStrCtlFallback], 1u);
    const uint32_t incl = warp_inclusive_scan(sum);
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
    uint64_t o = s_prefix + before + incl - sum;
#pragma unroll
    for (int q = 0; q < kStrScanPerThread; ++q) {
        const uint64_t item = item0 + q;
        const uint32_t bytes = by[q];
        if (item < a.n) {
            a.out_offsets[item] = o;
            // output tiles: the next string opens tile k1 when this one crosses into it
            const uint64_t k0 = (o + a.out_phase) / kStrTileBytes, k1 = (o + bytes + a.out_phase) / kStrTileBytes;
            if (bytes <= kStrSlackBytes) {
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

// ---- pack ---------------------------------------------------------------------------------------------------
// enc_append of encode_tiled.cuh without the store pointer: the low field of the bit counter holds the ABSOLUTE
// bit address in the shared window (8 * byte address + bit), so the word that a carry completes lies at
// ((nb >> 3) & ~3) - 4. A pointer that is bumped right behind the predicated store it addresses makes the warp
// wait until the store has read its operands (measured: a third of the stall samples of the packing loop sat
// on that add); a fresh temporary per symbol has no such dependency.
#ifndef HB_STR_PTR_APPEND
#define HB_STR_PTR_APPEND 0
#endif
__device__ __forceinline__ void str_append(uint32_t &acc, uint32_t &nb, uint32_t code, uint32_t ey) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 pw, hi, t;\n\t"
        "shf.l.wrap.b32 pw, 0, 1, %2;\n\t"
        "mul.hi.u32 hi, %0, pw;\n\t"
        "mad.lo.u32 %0, %0, pw, %3;\n\t"
        "add.u32 %1, %1, %2;\n\t"
        "setp.lt.u32 p, %1, %2;\n\t"
        "shf.r.wrap.b32 hi, %0, hi, %1;\n\t"
        "shr.u32 t, %1, 3;\n\t"
        "and.b32 t, t, 0x00fffffc;\n\t"
        "@p st.shared.u32 [t+-4], hi;\n\t"
        "}"
        : "+r"(acc), "+r"(nb)
        : "r"(ey), "r"(code)
        : "memory");
}

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
        for (int k = 0; k < 16; ++k) put(e[k].x, e[k].y);
    }
    __device__ __forceinline__ void put(uint32_t code, uint32_t ey) {
#if HB_STR_PTR_APPEND
        enc_append(sp, acc, nb, code, ey);
#else
        str_append(acc, nb, code, ey);
#endif
    }
    // where the next completed word goes / how many bits of it are there
    __device__ __forceinline__ void start(uint32_t word_addr, uint32_t lead_bits, uint32_t lead_value) {
        acc = lead_value;
#if HB_STR_PTR_APPEND
        sp = word_addr;
        nb = enc_len_fields(lead_bits);
#else
        nb = (lead_bits << 27) | (8u * word_addr + lead_bits);
#endif
    }
    __device__ __forceinline__ uint32_t word_addr() const {
#if HB_STR_PTR_APPEND
        return sp;
#else
        return (nb >> 3) & 0x00fffffcu;
#endif
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
        for (int k = 0; k < 16; ++k) put(e[k].x, e[k].y);
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
                    const uint32_t w0 = stage_addr + (orel & ~3u);
                    uint32_t first;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(first) : "r"(w0));
                    pk.start(w0, lead, lead ? first >> (32u - lead) : 0u);
                    str_walk(a.in + base0 + s_inrel[idx], s_len[idx], in_end, pk);
                    // pad the last byte with the LOW bits of eos_padding (huffman.c:178-184)
                    const uint32_t pad = (0u - pk.nb) & 7u;
                    pk.put(a.eos_padding & ((1u << pad) - 1u), enc_len_fields(pad));
                    const uint32_t rem = pk.nb >> 27;
                    if (rem) {
                        tail_addr[q] = pk.word_addr();
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
