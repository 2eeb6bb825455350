// Generic batched kernels: every per-item rule of the reference codec (capacity limits,
// SHORT_BUFFER with overflow bits, UNKNOWN_SYMBOL, cursor positions, leftover decoder state) in
// closed form on the device. These serve the slotted layout and any item shape; the tiled kernels
// in encode_tiled.cuh / decode_tiled.cuh take over the throughput cases.
//
// Per-item contract: one aws_huffman_encode / aws_huffman_decode call on a fresh encoder/decoder
// (reference source/huffman.c:131-187 and :213-286; closed forms in SURVEY.md App. B.4-B.7).
#pragma once

#include "device_common.cuh"

namespace hb {

constexpr int kWarpsPerBlock = 8;
constexpr int kStageWords = 36;  // 7 carry bits + 32 codes x 32 bits = 1031 bits -> 33 words (+ spill)

// ---------------------------------------------------------------------------------------------
// Encode: one warp per item, 32 symbols per step.
//
// With B_k the inclusive bit prefix over symbols, C the item's capacity in bytes and u the first
// symbol without a code (reference huffman.c:161-173 with :59-105):
//   * C == 0 and the item is not empty                      -> SHORT_BUFFER, nothing consumed
//   * j = first k < u with B_k >= 8C:
//       B_j == 8C and j is the last symbol                  -> SUCCESS (exact fit, no padding)
//       otherwise                                           -> SHORT_BUFFER, consumed j+1, len C,
//                                                              overflow = low (B_j - 8C) bits of code j
//   * else u exists                                         -> UNKNOWN_SYMBOL, consumed u+1,
//                                                              len floor(B_{u-1}/8) (pending bits dropped)
//   * else                                                  -> SUCCESS, len ceil(B/8), last byte padded
//                                                              with the LOW bits of eos_padding (:178-184)
// kWrite=false only measures (lengths / status); kWrite=true also produces the bytes.
// ---------------------------------------------------------------------------------------------
template <bool kWrite>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) encode_items_warp_kernel(DeviceTables t, BatchView b) {
    __shared__ uint2 s_enc[256];
    __shared__ uint32_t s_stage[kWarpsPerBlock][kStageWords];

    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_enc[i] = t.enc[i];
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const uint64_t item = (uint64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (item >= b.n) return;

    const uint64_t in0 = b.in_offsets[item];
    const uint64_t L = b.in_offsets[item + 1] - in0;
    const uint8_t *src = b.in + in0;
    const uint64_t C = b.out_caps ? b.out_caps[item] : kNoCap;
    const uint64_t cap_bits = (C >= (1ull << 60)) ? ~0ull : C * 8;
    uint64_t out0 = 0;
    if (kWrite) out0 = b.out_offsets[item];
    // writes never leave [out, out + out_capacity)
    const uint64_t phys_room = (kWrite && out0 < b.out_capacity) ? b.out_capacity - out0 : 0;
    uint8_t *dst = b.out + out0;
    uint32_t *stage = s_stage[warp];

    int32_t status = kStatusOk;
    uint64_t consumed = L;
    uint64_t out_len = 0;
    uint32_t ovf_pattern = 0, ovf_bits = 0;

    uint64_t bits_base = 0;  // bits of all earlier steps
    uint64_t flushed = 0;    // bytes already copied out of the stage
    uint32_t carry = 0;      // bits (<8) sitting at the top of stage[0]
    bool stopped = false;

    // Continuing a stream (huffman.c:150-160): the bits the previous call could not place go first, as if
    // they were the code of a symbol at index -1 (consuming it advances the cursor by nothing).
    uint32_t in_pattern = 0, in_bits = 0;
    if (b.resume) {
        in_bits = b.overflow_num_bits[item];
        in_pattern = in_bits ? b.overflow_pattern[item] & (0xffffffffu >> (32 - in_bits)) : 0u;
    }

    if (kWrite) {
        for (int w = lane; w < kStageWords; w += 32) stage[w] = 0;
        __syncwarp();
    }

    if (C == 0 && (L > 0 || in_bits > 0)) {
        status = kStatusShortBuffer;
        consumed = 0;
        stopped = true;
        ovf_pattern = in_pattern;  // nothing was written: the pending bits stay pending (:151-153)
        ovf_bits = in_bits;
    }

    // (the step that carries the pending bits runs with base = -32: lane 31 is "symbol -1")
    for (int64_t base = in_bits ? -32 : 0; base < (int64_t)L && !stopped; base += 32) {
        const int64_t k = base + (int64_t)lane;
        const bool pending = k == -1;
        const bool valid = pending || (k >= 0 && k < (int64_t)L);
        uint32_t code = 0, len = 0;
        if (pending) {
            code = in_pattern;
            len = in_bits;
        } else if (valid) {
            const uint2 e = s_enc[src[k]];
            code = e.x;
            len = e.y;
        }
        const uint32_t incl = warp_inclusive_scan(len);
        const uint64_t Bk = bits_base + incl;
        const uint32_t unknown_mask = __ballot_sync(0xffffffffu, valid && len == 0);
        const uint32_t full_mask = __ballot_sync(0xffffffffu, valid && len != 0 && Bk >= cap_bits);
        const uint32_t uu = unknown_mask ? (uint32_t)(__ffs(unknown_mask) - 1) : 32u;
        uint32_t jj = full_mask ? (uint32_t)(__ffs(full_mask) - 1) : 32u;

        uint32_t active_lanes = 32;  // lanes [0, active_lanes) contribute bits
        uint64_t clip_bytes = ~0ull; // bytes of this item that may be written
        if (jj < uu) {
            const uint64_t Bj = __shfl_sync(0xffffffffu, Bk, jj);
            const uint32_t code_j = __shfl_sync(0xffffffffu, code, jj);
            const bool exact_fit_at_end = (Bj == cap_bits) && (base + jj + 1 == (int64_t)L);
            if (!exact_fit_at_end) {
                status = kStatusShortBuffer;
                consumed = (uint64_t)(base + jj + 1);
                out_len = C;
                ovf_bits = (uint32_t)(Bj - cap_bits);
                ovf_pattern = ovf_bits ? (code_j & (0xffffffffu >> (32 - ovf_bits))) : 0;
                active_lanes = jj + 1;
                clip_bytes = C;
                stopped = true;
            }
        } else if (uu < 32) {
            const uint64_t Bprev = __shfl_sync(0xffffffffu, Bk - len, uu);
            status = kStatusUnknownSymbol;
            consumed = (uint64_t)(base + uu + 1);
            out_len = Bprev >> 3;
            active_lanes = uu;
            clip_bytes = out_len;
            stopped = true;
        }

        // bits contributed by this step
        const uint32_t last_active = active_lanes ? active_lanes - 1 : 0;
        const uint32_t step_bits = active_lanes ? __shfl_sync(0xffffffffu, incl, last_active) : 0;

        if (kWrite) {
            if (lane < active_lanes && len != 0) {
                const uint32_t q = carry + (incl - len);
                const uint32_t w = q >> 5, sh = q & 31;
                const uint64_t v = (uint64_t)code << (64 - len - sh);
                atomicOr(&stage[w], (uint32_t)(v >> 32));
                if ((uint32_t)v) atomicOr(&stage[w + 1], (uint32_t)v);
            }
            __syncwarp();
            const uint32_t have_bits = carry + step_bits;
            uint64_t nbytes = have_bits >> 3;
            if (flushed + nbytes > clip_bytes) nbytes = clip_bytes > flushed ? clip_bytes - flushed : 0;
            for (uint32_t i = lane; i < nbytes; i += 32) {
                const uint8_t byte = (uint8_t)(stage[i >> 2] >> (24 - 8 * (i & 3)));
                if (kWrite && flushed + i < phys_room) dst[flushed + i] = byte;
            }
            const uint32_t whole = have_bits >> 3;
            const uint32_t carry_byte = (stage[whole >> 2] >> (24 - 8 * (whole & 3))) & 0xffu;
            __syncwarp();
            for (int w = lane; w < kStageWords; w += 32) stage[w] = 0;
            __syncwarp();
            carry = have_bits & 7;
            if (lane == 0) stage[0] = carry ? (carry_byte << 24) : 0;
            __syncwarp();
            flushed += nbytes;
        } else {
            carry = (carry + step_bits) & 7;
        }
        bits_base += step_bits;
    }

    if (!stopped) {
        // every symbol placed
        out_len = (bits_base + 7) >> 3;
        if (kWrite && carry) {
            const uint32_t pad = 8 - carry;
            const uint8_t last = (uint8_t)((stage[0] >> 24) | (t.eos_padding & ((1u << pad) - 1u)));
            if (kWrite && lane == 0 && flushed < phys_room) dst[flushed] = last;
        }
    }

    if (lane == 0) {
        b.out_lens[item] = out_len;
        // (packed layout: a measuring launch precedes the writing one; the state arrays are inputs of both)
        if (kWrite || !b.resume) {
            if (b.status) b.status[item] = status;
            if (b.consumed) b.consumed[item] = consumed;
            if (b.overflow_pattern) b.overflow_pattern[item] = ovf_pattern;
            if (b.overflow_num_bits) b.overflow_num_bits[item] = (uint8_t)ovf_bits;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Decode: one thread per item, walking the reference loop (huffman.c:230-281) with the register
// refill of :196-211, so the cursor position and leftover register come out identical.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lut_lookup(
    const uint32_t *s_lut, uint32_t s_count, const uint32_t *g_lut, uint32_t root_bits, uint32_t window) {
    // returns a leaf entry (only its first symbol is used here) or 0 for a hole
    uint32_t e = s_lut[window >> (32 - root_bits)];
    uint32_t used = root_bits;
    while (e != 0 && !dlut_is_leaf(e)) {
        const uint32_t width = dlut_link_width(e);
        const uint32_t idx = dlut_link_base(e) + ((window << used) >> (32 - width));
        e = idx < s_count ? s_lut[idx] : __ldg(&g_lut[idx]);
        used += width;
    }
    return e;
}

template <bool kWrite>
__global__ void __launch_bounds__(256) decode_items_thread_kernel(DeviceTables t, BatchView b, uint32_t smem_entries) {
    extern __shared__ uint32_t s_lut[];
    for (uint32_t i = threadIdx.x; i < smem_entries; i += blockDim.x) s_lut[i] = t.lut[i];
    __syncthreads();

    const uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= b.n) return;

    const uint64_t in0 = b.in_offsets[item];
    const uint64_t len = b.in_offsets[item + 1] - in0;
    const uint8_t *src = b.in + in0;
    const uint64_t C = b.out_caps ? b.out_caps[item] : kNoCap;
    uint64_t out0 = 0;
    if (kWrite) out0 = b.out_offsets[item];
    const uint64_t phys_room = (kWrite && out0 < b.out_capacity) ? b.out_capacity - out0 : 0;
    uint8_t *dst = b.out + out0;

    uint64_t reg = 0;
    uint32_t have = 0;
    uint64_t pos = 0;
    if (b.resume) {  // decoder->working_bits / num_bits of the previous call (huffman.c:222)
        have = b.leftover_num_bits[item];
        reg = have ? b.leftover_working_bits[item] : 0;
    }
    uint64_t bits_left = len * 8 + have;
    uint64_t out_len = 0;
    int32_t status = kStatusOk;

    while (true) {
        if (have < 32 && pos < len) {
            uint64_t want = (32 - have + 7) >> 3;
            if (want > len - pos) want = len - pos;
            for (uint64_t i = 0; i < want; ++i) {
                reg |= (uint64_t)src[pos + i] << (56 - have);
                have += 8;
            }
            pos += want;
        }
        const uint32_t e = lut_lookup(s_lut, smem_entries, t.lut, t.lut_root_bits, (uint32_t)(reg >> 32));
        if (e == 0) {
            if (bits_left >= 32) status = kStatusUnknownSymbol;
            break;
        }
        const uint32_t used = dlut_len1(e);
        if (used > bits_left) break;
        if (out_len == C) {
            status = kStatusShortBuffer;
            break;
        }
        bits_left -= used;
        reg <<= used;
        have -= used;
        if (kWrite && out_len < phys_room) dst[out_len] = (uint8_t)dlut_sym1(e);
        ++out_len;
        if (bits_left == 0) break;
    }

    b.out_lens[item] = out_len;
    if (kWrite || !b.resume) {
        if (b.status) b.status[item] = status;
        if (b.consumed) b.consumed[item] = pos;
        if (b.leftover_working_bits) b.leftover_working_bits[item] = reg;
        if (b.leftover_num_bits) b.leftover_num_bits[item] = (uint8_t)have;
    }
}

// ---------------------------------------------------------------------------------------------
// lens[0..n) -> offsets[0..n] (exclusive prefix sum, offsets[n] = total). Single pass with
// decoupled look-back; tiles are handed out by an atomic ticket so a tile's predecessors are
// always already running.
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItemsPerThread = 8;
constexpr int kScanTile = kScanThreads * kScanItemsPerThread;

__global__ void __launch_bounds__(kScanThreads)
    scan_lens_kernel(const uint64_t *lens, uint64_t *offsets, uint64_t n, uint64_t *tile_state, uint32_t *ticket,
                     const uint32_t *gate = nullptr, uint32_t slots_of = 0) {
    // slots_of != 0: `lens` is a CSR offsets array (n + 1 entries) and the values scanned are
    // ceil((lens[i + 1] - lens[i]) / slots_of), the slot counts of encode_slots_kernel
    if (gate != nullptr && *gate == 0) return;  // (fallback of the fused stream decoder: nothing to redo)
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_warp_sums[kScanThreads / 32];
    __shared__ uint64_t s_tile_prefix;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t first = (uint64_t)tile * kScanTile + (uint64_t)threadIdx.x * kScanItemsPerThread;

    uint64_t v[kScanItemsPerThread];
    uint64_t sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItemsPerThread; ++i) {
        v[i] = 0;
        if (first + i < n) v[i] = slots_of ? (lens[first + i + 1] - lens[first + i] + slots_of - 1) / slots_of : lens[first + i];
        sum += v[i];
    }
    const uint64_t incl = warp_inclusive_scan64(sum);
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    if (lane == 31) s_warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < kScanThreads / 32 ? s_warp_sums[lane] : 0;
        const uint64_t wi = warp_inclusive_scan64(w);
        if (lane < kScanThreads / 32) s_warp_sums[lane] = wi - w;  // exclusive warp offsets
        const uint64_t tile_total = __shfl_sync(0xffffffffu, wi, kScanThreads / 32 - 1);
        const uint64_t prefix = lookback_exclusive_prefix(tile_state, tile, tile_total);
        if (lane == 0) s_tile_prefix = prefix;
    }
    __syncthreads();
    uint64_t run = s_tile_prefix + s_warp_sums[warp] + (incl - sum);
#pragma unroll
    for (int i = 0; i < kScanItemsPerThread; ++i) {
        if (first + i < n) offsets[first + i] = run;
        run += v[i];
        if (first + i + 1 == n) offsets[n] = run;
    }
    if (n == 0 && tile == 0 && threadIdx.x == 0) offsets[0] = 0;
}

}  // namespace hb
