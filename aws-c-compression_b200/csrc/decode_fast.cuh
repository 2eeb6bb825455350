// Fast decode paths for the packed layout.
//
//   decode_batch_kernel        batches of independent short strings (BASELINE configs 2, 5): one thread
//                              per string, persistent blocks with the whole decode LUT in shared memory.
//                              Each tile of strings is decoded twice from L1-resident input — a counting
//                              pass, then (after a block scan and a single-pass decoupled look-back give
//                              every string its output offset) a writing pass — so the output is packed
//                              back to back without a separate size query.
//
//   stream_*                   one long stream (BASELINE config 4): chunked speculative decode that
//                              exploits Huffman self-synchronisation. The stream is cut into fixed
//                              chunks of kChunkBits; each thread starts kPrerollBits before its chunk at
//                              an arbitrary bit, and by the time it reaches its chunk it is (almost
//                              always) on a true code boundary; it then decodes its chunk and records
//                              where it entered, where it left and how many symbols it saw. A chunk is
//                              confirmed when it entered exactly where its predecessor left; the few that
//                              are not are re-decoded from the right spot until the chain is a fixed
//                              point (correct for ANY prefix code, self-synchronising or not). A scan of
//                              the symbol counts gives the output offsets and a final pass writes.
//
// Termination rules, status, cursor position and leftover register follow the reference loop
// (source/huffman.c:230-281, refill :196-211); see SURVEY.md App. B.6/B.7 for the closed forms.
#pragma once

#include "device_common.cuh"

namespace hb {

constexpr uint32_t kDecLutMaxSmem = 8192;  // entries (32 KiB); larger tables use the generic kernels

// Device LUT entry (converted from host/huffman_lut.h at context creation):
//   leaf: len << 8 | symbol   (len 1..32)      link: 0x80000000 | width << 24 | base      hole: 0
__device__ __forceinline__ uint32_t dec_lookup(const uint32_t *s_lut, uint32_t root_bits, uint32_t window) {
    uint32_t e = s_lut[window >> (32 - root_bits)];
    if ((int32_t)e < 0) {
        uint32_t used = root_bits;
        do {
            const uint32_t width = (e >> 24) & 0x7fu;
            e = s_lut[(e & 0xFFFFFFu) + ((window << used) >> (32 - width))];
            used += width;
        } while ((int32_t)e < 0);
    }
    return e;
}

enum : uint32_t {
    kTermStop = 0,     // reached the caller's stop position (more stream follows)
    kTermEnd = 1,      // the stream ended (all bits used, padding, or a code cut short): SUCCESS
    kTermUnknown = 2,  // a window with >= 32 bits left matched no code: UNKNOWN_SYMBOL
};

struct DecodeSpan {
    uint64_t pos;   // bit position after the last decoded symbol
    uint64_t nsym;  // symbols decoded
    uint32_t term;
};

// Big-endian bit reader over global memory using aligned 32-bit loads; zero beyond `end_byte`.
struct BitReader {
    const uint32_t *wp;     // next aligned word to load
    const uint32_t *wlast;  // aligned word holding the last valid byte
    uint32_t last_mask;     // keeps the valid leading bytes of *wlast (big-endian value)
    uint64_t buf;           // unread bits, left-aligned
    int nb;                 // how many bits of buf came from loaded words

    __device__ __forceinline__ uint32_t load_word() {
        uint32_t w = 0;
        if (wp <= wlast) {
            w = __byte_perm(__ldg(wp), 0, 0x0123);
            if (wp == wlast) w &= last_mask;
        }
        ++wp;
        return w;
    }
    __device__ __forceinline__ void init(const uint8_t *base, uint64_t bit_pos, uint64_t end_byte) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(base) + (bit_pos >> 3);
        const uintptr_t e = reinterpret_cast<uintptr_t>(base) + end_byte;  // one past the last valid byte
        wp = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
        if (end_byte == 0 || e <= (a & ~uintptr_t(3))) {
            wlast = wp - 1;  // nothing valid from here on
            last_mask = 0;
        } else {
            wlast = reinterpret_cast<const uint32_t *>((e - 1) & ~uintptr_t(3));
            const uint32_t valid = (uint32_t)(e - reinterpret_cast<uintptr_t>(wlast));  // 1..4
            last_mask = 0xffffffffu << (8 * (4 - valid));
        }
        const uint32_t lead = (uint32_t)(8 * (a & 3) + (bit_pos & 7));
        const uint32_t w = load_word();
        buf = (uint64_t)w << (32 + lead);
        nb = 32 - (int)lead;
    }
    __device__ __forceinline__ void refill() {
        if (nb <= 32) {
            buf |= (uint64_t)load_word() << (32 - nb);
            nb += 32;
        }
    }
    __device__ __forceinline__ uint32_t window() const { return (uint32_t)(buf >> 32); }
    __device__ __forceinline__ void consume(uint32_t n) {
        buf <<= n;
        nb -= (int)n;
    }
};

// Packs symbols into aligned 32-bit global stores (byte stores only for a ragged head or tail).
struct ByteWriter {
    uint8_t *wordp;   // aligned address of the word being assembled
    uint32_t pack;
    uint32_t count;   // bytes already in `pack` (including the skipped head bytes)
    uint32_t head;    // bytes at the front of the first word that are not ours
    uint64_t room;    // bytes we may still write

    __device__ __forceinline__ void init(uint8_t *out, uint64_t out_room) {
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 3);
        wordp = out - mis;
        pack = 0;
        count = mis;
        head = mis;
        room = out_room;
    }
    __device__ __forceinline__ void flush_partial(uint32_t upto) {
        for (uint32_t i = head; i < upto; ++i) {
            if (room) {
                wordp[i] = (uint8_t)(pack >> (8 * i));
                --room;
            }
        }
    }
    __device__ __forceinline__ void put(uint32_t sym) {
        pack |= (sym & 0xffu) << (8 * count);
        if (++count == 4) {
            if (head == 0 && room >= 4) {
                *reinterpret_cast<uint32_t *>(wordp) = pack;
                room -= 4;
            } else {
                flush_partial(4);
            }
            head = 0;
            wordp += 4;
            pack = 0;
            count = 0;
        }
    }
    __device__ __forceinline__ void finish() {
        if (count > head) flush_partial(count);
    }
};

// Decodes from bit `start` while the position is below `stop` (<= end_bit = 8 * end_byte).
// kSkipHoles: pre-roll policy — a hole just advances one bit (we are only hunting for a boundary).
template <bool kWrite, bool kSkipHoles>
__device__ __forceinline__ DecodeSpan decode_span(
    const uint32_t *s_lut, uint32_t root_bits, const uint8_t *base, uint64_t start, uint64_t stop, uint64_t end_byte,
    ByteWriter *writer) {
    const uint64_t end_bit = end_byte * 8;
    DecodeSpan r;
    r.pos = start;
    r.nsym = 0;
    r.term = kTermStop;
    if (start >= stop) {
        if (start >= end_bit) r.term = kTermEnd;
        return r;
    }
    BitReader br;
    br.init(base, start, end_byte);
    uint64_t pos = start;
    uint64_t nsym = 0;
    while (true) {
        br.refill();
        const uint32_t e = dec_lookup(s_lut, root_bits, br.window());
        const uint64_t bits_left = end_bit - pos;
        if (e == 0) {
            if (kSkipHoles) {
                if (bits_left <= 1) { r.term = kTermEnd; break; }
                br.consume(1);
                pos += 1;
                if (pos >= stop) break;
                continue;
            }
            r.term = bits_left < 32 ? kTermEnd : kTermUnknown;
            break;
        }
        const uint32_t used = e >> 8;
        if (used > bits_left) { r.term = kTermEnd; break; }
        br.consume(used);
        pos += used;
        ++nsym;
        if (kWrite) writer->put(e);
        if (pos >= stop) {
            if (pos >= end_bit) r.term = kTermEnd;
            break;
        }
    }
    r.pos = pos;
    r.nsym = nsym;
    return r;
}

// decoder->working_bits / num_bits and the cursor position after a call that emitted `cbits` bits of
// symbols from an item of `len` bytes (fresh decoder). SURVEY.md App. B.7.
__device__ __forceinline__ void leftover_state(
    const uint8_t *item, uint64_t len, uint64_t cbits, bool stopped_early, uint64_t *consumed, uint64_t *working,
    uint8_t *num_bits) {
    uint64_t pulled = len;
    if (stopped_early) {
        const uint64_t need = (cbits + 32 + 7) >> 3;
        pulled = need < len ? need : len;
    }
    const uint64_t first = cbits >> 3;
    uint64_t reg = 0;
    for (uint64_t i = first; i < pulled && i < first + 8; ++i) reg |= (uint64_t)item[i] << (56 - 8 * (i - first));
    reg <<= (cbits & 7);
    if (consumed) *consumed = pulled;
    if (working) *working = reg;
    if (num_bits) *num_bits = (uint8_t)(pulled * 8 - cbits);
}

// ---------------------------------------------------------------------------------------------
// Batch of independent strings
// ---------------------------------------------------------------------------------------------
constexpr int kDecThreads = 128;

struct DecBatchArgs {
    BatchView b;
    const uint32_t *lut;
    uint32_t lut_count;
    uint32_t root_bits;
    uint64_t *tile_state;
    uint32_t *ticket;
    uint32_t num_tiles;
};

__global__ void __launch_bounds__(kDecThreads) decode_batch_kernel(DecBatchArgs a) {
    extern __shared__ uint32_t s_lut[];
    __shared__ uint64_t s_warp_sum[kDecThreads / 32];
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_tile;

    for (uint32_t i = threadIdx.x; i < a.lut_count; i += kDecThreads) s_lut[i] = a.lut[i];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const BatchView &b = a.b;

    while (true) {
        __syncthreads();  // s_tile / s_prefix / s_warp_sum reuse, and the LUT on the first trip
        if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= a.num_tiles) break;
        const uint64_t item = (uint64_t)tile * kDecThreads + threadIdx.x;
        const bool live = item < b.n;

        uint64_t in0 = 0, len = 0;
        if (live) {
            in0 = b.in_offsets[item];
            len = b.in_offsets[item + 1] - in0;
        }
        const uint8_t *src = b.in + in0;

        // ---- count ---------------------------------------------------------------------------------------
        DecodeSpan c = {0, 0, kTermEnd};
        if (live) c = decode_span<false, false>(s_lut, a.root_bits, src, 0, len * 8, len, nullptr);

        // ---- offsets: block scan + look-back ----------------------------------------------------------------
        const uint64_t incl = warp_inclusive_scan64(c.nsym);
        if (lane == 31) s_warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = lane < kDecThreads / 32 ? s_warp_sum[lane] : 0;
            const uint64_t wi = warp_inclusive_scan64(w);
            if (lane < kDecThreads / 32) s_warp_sum[lane] = wi - w;
            const uint64_t total = __shfl_sync(0xffffffffu, wi, kDecThreads / 32 - 1);
            const uint64_t prefix = lookback_exclusive_prefix(a.tile_state, tile, total);
            if (lane == 0) s_prefix = prefix;
        }
        __syncthreads();
        const uint64_t off = s_prefix + s_warp_sum[warp] + (incl - c.nsym);

        if (live) {
            b.out_offsets[item] = off;
            if (item + 1 == b.n) b.out_offsets[b.n] = off + c.nsym;
            if (b.out_lens) b.out_lens[item] = c.nsym;
            if (b.status) b.status[item] = c.term == kTermUnknown ? kStatusUnknownSymbol : kStatusOk;
            if (b.consumed || b.leftover_working_bits || b.leftover_num_bits)
                leftover_state(
                    src, len, c.pos, c.term == kTermUnknown, b.consumed ? b.consumed + item : nullptr,
                    b.leftover_working_bits ? b.leftover_working_bits + item : nullptr,
                    b.leftover_num_bits ? b.leftover_num_bits + item : nullptr);

            // ---- write -----------------------------------------------------------------------------------------
            if (c.nsym) {
                ByteWriter wr;
                wr.init(b.out + off, off < b.out_capacity ? b.out_capacity - off : 0);
                decode_span<true, false>(s_lut, a.root_bits, src, 0, len * 8, len, &wr);
                wr.finish();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One long stream
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kChunkBits = 1024;   // 128 encoded bytes per thread
constexpr uint32_t kPrerollBits = 256;  // HPACK on Zipf data: 99.4 % of starts are in sync by then
constexpr int kStreamThreads = 128;

// Per-chunk record, one 64-bit word so it is always read and written whole:
//   [15:0]  entry offset  (first code boundary at or after the chunk start, relative to it)
//   [31:16] exit offset   (first code boundary at or after the chunk end, relative to it)
//   [47:32] symbols decoded in the chunk
//   [49:48] termination (kTerm*)
__device__ __forceinline__ uint64_t chunk_pack(uint32_t entry, uint32_t exit, uint32_t nsym, uint32_t term) {
    return (uint64_t)(entry & 0xffffu) | ((uint64_t)(exit & 0xffffu) << 16) | ((uint64_t)(nsym & 0xffffu) << 32) |
           ((uint64_t)(term & 3u) << 48);
}
__device__ __forceinline__ uint32_t chunk_entry(uint64_t w) { return (uint32_t)(w & 0xffffu); }
__device__ __forceinline__ uint32_t chunk_exit(uint64_t w) { return (uint32_t)((w >> 16) & 0xffffu); }
__device__ __forceinline__ uint32_t chunk_nsym(uint64_t w) { return (uint32_t)((w >> 32) & 0xffffu); }
__device__ __forceinline__ uint32_t chunk_term(uint64_t w) { return (uint32_t)((w >> 48) & 3u); }

struct StreamArgs {
    const uint8_t *in;       // the item's first byte
    uint64_t len;            // encoded bytes
    uint64_t num_chunks;
    uint64_t *chunks;        // num_chunks records
    uint64_t *chunk_offsets; // num_chunks + 1 output offsets (after the scan)
    uint64_t *control;       // [0] = first inconsistent chunk (or ~0), [1] = first terminated chunk (or ~0)
    const uint32_t *lut;
    uint32_t lut_count;
    uint32_t root_bits;
};

__device__ __forceinline__ uint64_t decode_chunk_record(
    const uint32_t *s_lut, const StreamArgs &a, uint64_t k, uint32_t entry) {
    const uint64_t begin = k * kChunkBits;
    const uint64_t stop = min(begin + kChunkBits, a.len * 8);
    const DecodeSpan r = decode_span<false, false>(s_lut, a.root_bits, a.in, begin + entry, stop, a.len, nullptr);
    const uint32_t exit = r.term == kTermStop ? (uint32_t)(r.pos - stop) : 0u;
    return chunk_pack(entry, exit, (uint32_t)r.nsym, r.term);
}

// Speculative pass: find an entry point by pre-rolling, then decode the chunk once.
__global__ void __launch_bounds__(kStreamThreads) stream_sync_kernel(StreamArgs a) {
    extern __shared__ uint32_t s_lut[];
    for (uint32_t i = threadIdx.x; i < a.lut_count; i += kStreamThreads) s_lut[i] = a.lut[i];
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * kStreamThreads + threadIdx.x; k < a.num_chunks;
         k += (uint64_t)gridDim.x * kStreamThreads) {
        const uint64_t begin = k * kChunkBits;
        uint32_t entry = 0;
        if (k != 0) {
            const uint64_t from = begin > kPrerollBits ? begin - kPrerollBits : 0;
            const DecodeSpan pre = decode_span<false, true>(s_lut, a.root_bits, a.in, from, begin, a.len, nullptr);
            entry = pre.pos >= begin ? (uint32_t)(pre.pos - begin) : 0u;
        }
        a.chunks[k] = decode_chunk_record(s_lut, a, k, entry);
    }
}

// One relaxation round: a chunk whose entry differs from its predecessor's exit is re-decoded from
// there. Records are single words, so updating in place is safe; the fixed point is the true chain.
__global__ void __launch_bounds__(kStreamThreads) stream_fix_kernel(StreamArgs a) {
    extern __shared__ uint32_t s_lut[];
    for (uint32_t i = threadIdx.x; i < a.lut_count; i += kStreamThreads) s_lut[i] = a.lut[i];
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * kStreamThreads + threadIdx.x + 1; k < a.num_chunks;
         k += (uint64_t)gridDim.x * kStreamThreads) {
        const uint64_t prev = ld_relaxed_u64(&a.chunks[k - 1]);
        const uint64_t mine = ld_relaxed_u64(&a.chunks[k]);
        if (chunk_term(prev) != kTermStop) continue;  // the stream ended before this chunk
        if (chunk_entry(mine) != chunk_exit(prev))
            st_relaxed_u64(&a.chunks[k], decode_chunk_record(s_lut, a, k, chunk_exit(prev)));
    }
}

// Finds the first chunk that still disagrees with its predecessor and the first terminated chunk.
__global__ void __launch_bounds__(256) stream_verify_kernel(StreamArgs a) {
    uint64_t bad = ~0ull, term = ~0ull;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.num_chunks;
         k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t mine = a.chunks[k];
        if (chunk_term(mine) != kTermStop && k < term) term = k;
        if (k > 0) {
            const uint64_t prev = a.chunks[k - 1];
            if (chunk_term(prev) == kTermStop && chunk_entry(mine) != chunk_exit(prev) && k < bad) bad = k;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        bad = min(bad, __shfl_xor_sync(0xffffffffu, bad, d));
        term = min(term, __shfl_xor_sync(0xffffffffu, term, d));
    }
    if (lane_id() == 0) {
        if (bad != ~0ull) atomicMin(reinterpret_cast<unsigned long long *>(&a.control[0]), (unsigned long long)bad);
        if (term != ~0ull) atomicMin(reinterpret_cast<unsigned long long *>(&a.control[1]), (unsigned long long)term);
    }
}

// Last resort for inputs that do not self-synchronise within a few rounds: one thread walks the chain
// from the first inconsistent chunk. Does nothing when the verify pass found no inconsistency.
__global__ void __launch_bounds__(32) stream_repair_kernel(StreamArgs a) {
    extern __shared__ uint32_t s_lut[];
    if (a.control[0] == ~0ull) return;
    for (uint32_t i = threadIdx.x; i < a.lut_count; i += 32) s_lut[i] = a.lut[i];
    __syncwarp();
    if (threadIdx.x != 0) return;
    uint64_t first_term = a.control[1];
    for (uint64_t k = a.control[0]; k < a.num_chunks; ++k) {
        const uint64_t prev = a.chunks[k - 1];
        if (chunk_term(prev) != kTermStop) break;
        uint64_t mine = a.chunks[k];
        if (chunk_entry(mine) != chunk_exit(prev)) {
            mine = decode_chunk_record(s_lut, a, k, chunk_exit(prev));
            a.chunks[k] = mine;
        }
    }
    // termination may have moved: recompute the first terminated chunk
    first_term = ~0ull;
    for (uint64_t k = 0; k < a.num_chunks; ++k) {
        if (chunk_term(a.chunks[k]) != kTermStop) { first_term = k; break; }
    }
    a.control[1] = first_term;
    a.control[0] = ~0ull;
}

// chunk symbol counts (zero after the first terminated chunk) -> lens array for scan_lens_kernel
__global__ void __launch_bounds__(256) stream_counts_kernel(StreamArgs a, uint64_t *lens) {
    const uint64_t first_term = a.control[1];
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.num_chunks;
         k += (uint64_t)gridDim.x * blockDim.x)
        lens[k] = k <= first_term ? chunk_nsym(a.chunks[k]) : 0;
}

// Final pass: every chunk up to the terminating one decodes again, now writing.
__global__ void __launch_bounds__(kStreamThreads) stream_write_kernel(StreamArgs a, BatchView b) {
    extern __shared__ uint32_t s_lut[];
    for (uint32_t i = threadIdx.x; i < a.lut_count; i += kStreamThreads) s_lut[i] = a.lut[i];
    __syncthreads();
    const uint64_t first_term = a.control[1];
    const uint64_t last = first_term < a.num_chunks ? first_term : a.num_chunks - 1;
    for (uint64_t k = (uint64_t)blockIdx.x * kStreamThreads + threadIdx.x; k <= last;
         k += (uint64_t)gridDim.x * kStreamThreads) {
        const uint64_t rec = a.chunks[k];
        const uint64_t begin = k * kChunkBits;
        const uint64_t stop = min(begin + kChunkBits, a.len * 8);
        const uint64_t off = a.chunk_offsets[k];
        ByteWriter wr;
        wr.init(b.out + off, off < b.out_capacity ? b.out_capacity - off : 0);
        const DecodeSpan r =
            decode_span<true, false>(s_lut, a.root_bits, a.in, begin + chunk_entry(rec), stop, a.len, &wr);
        wr.finish();
        if (k == last) {
            // item-level results (n == 1)
            const uint64_t total = off + r.nsym;
            b.out_offsets[0] = 0;
            b.out_offsets[1] = total;
            if (b.out_lens) b.out_lens[0] = total;
            if (b.status) b.status[0] = r.term == kTermUnknown ? kStatusUnknownSymbol : kStatusOk;
            if (b.consumed || b.leftover_working_bits || b.leftover_num_bits)
                leftover_state(a.in, a.len, r.pos, r.term == kTermUnknown, b.consumed, b.leftover_working_bits,
                               b.leftover_num_bits);
        }
    }
}

}  // namespace hb
