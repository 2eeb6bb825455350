// HPACK string literals (RFC 7541 section 5.2) around the Huffman codec: SURVEY.md 8f.1.
//
//     +---+---+---+---+---+---+---+---+
//     | H |    String Length (7+)     |     H = 1: the payload is Huffman coded (RFC 7541 Appendix B table)
//     +---+---------------------------+     length: prefix integer of section 5.1 (N = 7): < 127 in the first
//     |  String Data (Length octets)  |             byte, else 127 + little-endian groups of 7 bits with a
//     +-------------------------------+             continuation flag in bit 7
//
// The payload kernels are the packed-layout codecs of this library; the kernels here are the thin per-item
// layer around them: plan the frames, write the prefixes, move payloads between the framed and the packed
// layout, parse incoming literals and apply the padding rule of section 5.2 (padding is the most significant
// bits of EOS, i.e. all ones, and strictly shorter than 8 bits; a payload that contains the EOS symbol cannot
// satisfy it) from the decoder's leftover register, which is what the reference's README.md:176-183
// describes the register for.
#pragma once

#include "device_common.cuh"

namespace hb {

constexpr int32_t kStatusInvalidArgument = 34;    // AWS_ERROR_INVALID_ARGUMENT (aws-c-common)
constexpr int32_t kStatusInvalidPadding = 3075;   // AWS_ERROR_COMPRESSION_INVALID_PADDING

enum : uint32_t { kHpackSmallest = 0, kHpackNever = 1, kHpackAlways = 2 };  // aws-c-http's aws_hpack_huffman_mode

__device__ __forceinline__ uint32_t hpack_prefix_bytes(uint64_t len) {
    if (len < 127) return 1;
    uint32_t nb = 2;
    for (uint64_t rem = len - 127; rem >= 128; rem >>= 7) ++nb;
    return nb;
}

// One warp copies n bytes, any alignment on either side: destination words are whole 32-bit stores, each made of
// the two aligned source words that cover it (one funnel shift); only the ragged ends go byte by byte. Reads
// stay inside the aligned words that hold the first and the last source byte.
__device__ __forceinline__ void warp_copy_bytes(uint8_t *dst, const uint8_t *src, uint64_t n, uint32_t lane) {
    if (n < 16) {
        if (lane < n) dst[lane] = src[lane];
        return;
    }
    const uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3)) & 3u;
    if (lane < head) dst[lane] = src[lane];
    const uint64_t words = (n - head) >> 2;
    const uintptr_t s0 = reinterpret_cast<uintptr_t>(src) + head;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(s0 & ~uintptr_t(3));
    const uint32_t r8 = (uint32_t)(s0 & 3) * 8u;
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
    for (uint64_t j = lane; j < words; j += 32) {
        const uint32_t lo = sw[j];
        const uint32_t hi = r8 ? sw[j + 1] : 0u;  // (not read when it could lie past the last source byte)
        dw[j] = __funnelshift_r(lo, hi, r8);
    }
    const uint64_t tail0 = head + 4 * words;
    if (lane < n - tail0) dst[tail0 + lane] = src[tail0 + lane];
}

// ---- encode ---------------------------------------------------------------------------------------------------
// Per item: Huffman or not (mode), prefix size, frame size.
__global__ void hpack_plan_kernel(
    uint64_t n, const uint64_t *raw_offsets, const uint64_t *enc_offsets, uint32_t mode, uint8_t *huff, uint64_t *frame_lens) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t raw = raw_offsets[i + 1] - raw_offsets[i];
    const uint64_t enc = enc_offsets ? enc_offsets[i + 1] - enc_offsets[i] : raw;
    const bool h = mode == kHpackAlways || (mode == kHpackSmallest && enc < raw);
    const uint64_t payload = h ? enc : raw;
    huff[i] = h ? 1 : 0;
    frame_lens[i] = hpack_prefix_bytes(payload) + payload;
}

// A warp takes 32 literals: every lane looks up one of them (so the dependent loads of offsets and flags are paid
// once per 32 items, not once per item) and writes its prefix; then the warp moves the 32 payloads one after the
// other, all lanes on each. Never writes at or after out + out_capacity.
struct MoveJob {
    const uint8_t *src;
    uint8_t *dst;
    uint64_t len;
};

__device__ __forceinline__ void warp_run_jobs(MoveJob (*jobs)[32], const MoveJob &mine, uint32_t warp, uint32_t lane) {
    jobs[warp][lane] = mine;
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
        const MoveJob j = jobs[warp][k];
        if (j.len) warp_copy_bytes(j.dst, j.src, j.len, lane);
    }
}

__global__ void __launch_bounds__(256) hpack_frame_kernel(
    uint64_t n, const uint8_t *raw, const uint64_t *raw_offsets, const uint8_t *enc, const uint64_t *enc_offsets,
    const uint8_t *huff, uint8_t *out, uint64_t out_capacity, const uint64_t *out_offsets) {
    __shared__ MoveJob s_jobs[8][32];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t i = ((uint64_t)blockIdx.x * 8 + warp) * 32 + lane;
    MoveJob job = {nullptr, nullptr, 0};
    if (i < n) {
        const bool h = huff[i] != 0;
        const uint8_t *src = h ? enc + enc_offsets[i] : raw + raw_offsets[i];
        const uint64_t len = h ? enc_offsets[i + 1] - enc_offsets[i] : raw_offsets[i + 1] - raw_offsets[i];
        const uint64_t o0 = out_offsets[i];
        const uint32_t np = hpack_prefix_bytes(len);
        const uint8_t hbit = h ? 0x80 : 0x00;
        if (len < 127) {
            if (o0 < out_capacity) out[o0] = hbit | (uint8_t)len;
        } else {
            if (o0 < out_capacity) out[o0] = hbit | 127;
            uint64_t rem = len - 127;
            uint64_t p = o0 + 1;
            while (rem >= 128) {
                if (p < out_capacity) out[p] = (uint8_t)(rem & 127) | 0x80;
                rem >>= 7;
                ++p;
            }
            if (p < out_capacity) out[p] = (uint8_t)rem;
        }
        const uint64_t room = o0 + np < out_capacity ? out_capacity - (o0 + np) : 0;
        job.src = src;
        job.dst = out + o0 + np;
        job.len = min(len, room);
    }
    warp_run_jobs(s_jobs, job, warp, lane);
}

// ---- decode ---------------------------------------------------------------------------------------------------
// Item i must be exactly one literal. Thread per item: H bit, prefix integer, consistency with the item's size.
//   status 0                        well formed
//   AWS_ERROR_SHORT_BUFFER          the literal is cut short (no prefix byte, prefix not finished, or fewer payload
//                                   bytes than it announces)
//   AWS_ERROR_INVALID_ARGUMENT      bytes left over after the payload, or a length that does not fit 62 bits
// gather_lens[i] = bytes of Huffman payload to decode (0 for raw and malformed items).
// H bit, prefix integer (RFC 7541 5.1) and consistency of one literal of `size` bytes at p.
__device__ __forceinline__ int32_t hpack_parse_literal(const uint8_t *p, uint64_t size, uint32_t &h, uint32_t &np, uint64_t &len) {
    h = 0;
    np = 0;
    len = 0;
    if (size == 0) return kStatusShortBuffer;
    const uint8_t b0 = p[0];
    h = b0 >> 7;
    len = b0 & 127;
    np = 1;
    if (len == 127) {
        uint32_t shift = 0;
        bool more = true;
        while (more) {
            if (np >= size) return kStatusShortBuffer;
            const uint8_t b = p[np];
            ++np;
            if (shift > 56) return kStatusInvalidArgument;
            len += (uint64_t)(b & 127) << shift;
            shift += 7;
            more = (b & 128) != 0;
        }
    }
    if (len > size - np) return kStatusShortBuffer;
    if (len < size - np) return kStatusInvalidArgument;
    return kStatusOk;
}

// The padding rule of section 5.2 on what a decoder left over: `nb` bits, left-aligned in `bits`.
__device__ __forceinline__ bool hpack_padding_ok(uint64_t bits, uint32_t nb) {
    return nb < 8 && (nb == 0 || (bits >> (64 - nb)) == ((1ull << nb) - 1));
}

__global__ void hpack_parse_kernel(
    uint64_t n, const uint8_t *in, const uint64_t *in_offsets, uint8_t *huff, uint8_t *prefix_len, uint64_t *pay_lens,
    uint64_t *gather_lens, int32_t *status) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t a = in_offsets[i], size = in_offsets[i + 1] - a;
    uint32_t h, np;
    uint64_t len;
    const int32_t st = hpack_parse_literal(in + a, size, h, np, len);
    huff[i] = (uint8_t)h;
    prefix_len[i] = (uint8_t)np;
    pay_lens[i] = st == kStatusOk ? len : 0;
    gather_lens[i] = (st == kStatusOk && h) ? len : 0;
    status[i] = st;
}

// dst[dst_offsets[i] ..) = the item's payload, taken from `alt` (packed, alt_offsets) when pick[i] != 0, else from
// the framed input behind its prefix. skip_unpicked: items with pick[i] == 0 are left out (the gather of Huffman
// payloads); otherwise they are copied as they are (raw literals in the final pass). 32 items per warp, as above.
__global__ void __launch_bounds__(256) hpack_move_kernel(
    uint64_t n, const uint8_t *framed, const uint64_t *framed_offsets, const uint8_t *prefix_len, const uint64_t *pay_lens,
    const uint8_t *alt, const uint64_t *alt_offsets, const uint8_t *pick, bool pick_from_alt, bool skip_unpicked,
    const int32_t *status, uint8_t *dst, uint64_t dst_capacity, const uint64_t *dst_offsets) {
    __shared__ MoveJob s_jobs[8][32];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t i = ((uint64_t)blockIdx.x * 8 + warp) * 32 + lane;
    MoveJob job = {nullptr, nullptr, 0};
    if (i < n && status[i] == kStatusOk) {
        const bool picked = pick[i] != 0;
        if (picked || !skip_unpicked) {
            const uint8_t *src;
            uint64_t len;
            if (picked && pick_from_alt) {
                src = alt + alt_offsets[i];
                len = alt_offsets[i + 1] - alt_offsets[i];
            } else {
                src = framed + framed_offsets[i] + prefix_len[i];
                len = pay_lens[i];
            }
            const uint64_t o0 = dst_offsets[i];
            const uint64_t room = o0 < dst_capacity ? dst_capacity - o0 : 0;
            job.src = src;
            job.dst = dst + o0;
            job.len = min(len, room);
        }
    }
    warp_run_jobs(s_jobs, job, warp, lane);
}

// After the Huffman payloads were decoded (packed, dec_offsets; leftover register per item): the padding rule,
// the final status and the final length of every item.
__global__ void hpack_finish_kernel(
    uint64_t n, const uint8_t *huff, const uint64_t *pay_lens, const uint64_t *dec_offsets, const int32_t *dec_status,
    const uint64_t *left_bits, const uint8_t *left_num, int32_t *status, uint64_t *final_lens) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t st = status[i];
    uint64_t len = 0;
    if (st == kStatusOk) {
        if (huff[i]) {
            const uint32_t nb = left_num[i];
            if (dec_status[i] != kStatusOk) st = dec_status[i];
            else if (!hpack_padding_ok(left_bits[i], nb)) st = kStatusInvalidPadding;  // a byte or more, or not all ones
            if (st == kStatusOk) len = dec_offsets[i + 1] - dec_offsets[i];
        } else {
            len = pay_lens[i];
        }
    }
    status[i] = st;
    final_lens[i] = len;
}

}  // namespace hb
