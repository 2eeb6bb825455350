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

// One warp per item: lane 0 writes the prefix, the warp copies the payload behind it. Never writes at or after
// out + out_capacity.
__global__ void __launch_bounds__(256) hpack_frame_kernel(
    uint64_t n, const uint8_t *raw, const uint64_t *raw_offsets, const uint8_t *enc, const uint64_t *enc_offsets,
    const uint8_t *huff, uint8_t *out, uint64_t out_capacity, const uint64_t *out_offsets) {
    const uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint32_t lane = lane_id();
    const bool h = huff[i] != 0;
    const uint8_t *src = h ? enc + enc_offsets[i] : raw + raw_offsets[i];
    const uint64_t len = h ? enc_offsets[i + 1] - enc_offsets[i] : raw_offsets[i + 1] - raw_offsets[i];
    const uint64_t o0 = out_offsets[i];
    const uint32_t np = hpack_prefix_bytes(len);
    if (lane == 0) {
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
    }
    uint8_t *dst = out + o0 + np;
    const uint64_t room = o0 + np < out_capacity ? out_capacity - (o0 + np) : 0;
    const uint64_t ncopy = min(len, room);
    for (uint64_t k = lane; k < ncopy; k += 32) dst[k] = src[k];
}

// ---- decode ---------------------------------------------------------------------------------------------------
// Item i must be exactly one literal. Thread per item: H bit, prefix integer, consistency with the item's size.
//   status 0                        well formed
//   AWS_ERROR_SHORT_BUFFER          the literal is cut short (no prefix byte, prefix not finished, or fewer payload
//                                   bytes than it announces)
//   AWS_ERROR_INVALID_ARGUMENT      bytes left over after the payload, or a length that does not fit 62 bits
// gather_lens[i] = bytes of Huffman payload to decode (0 for raw and malformed items).
__global__ void hpack_parse_kernel(
    uint64_t n, const uint8_t *in, const uint64_t *in_offsets, uint8_t *huff, uint8_t *prefix_len, uint64_t *pay_lens,
    uint64_t *gather_lens, int32_t *status) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t a = in_offsets[i], size = in_offsets[i + 1] - a;
    int32_t st = kStatusOk;
    uint32_t h = 0, np = 0;
    uint64_t len = 0;
    if (size == 0) {
        st = kStatusShortBuffer;
    } else {
        const uint8_t b0 = in[a];
        h = b0 >> 7;
        len = b0 & 127;
        np = 1;
        if (len == 127) {
            uint32_t shift = 0;
            bool more = true;
            while (more) {
                if (np >= size) { st = kStatusShortBuffer; break; }
                const uint8_t b = in[a + np];
                ++np;
                if (shift > 56) { st = kStatusInvalidArgument; break; }
                len += (uint64_t)(b & 127) << shift;
                shift += 7;
                more = (b & 128) != 0;
            }
        }
        if (st == kStatusOk) {
            if (len > size - np) st = kStatusShortBuffer;
            else if (len < size - np) st = kStatusInvalidArgument;
        }
    }
    huff[i] = (uint8_t)h;
    prefix_len[i] = (uint8_t)np;
    pay_lens[i] = st == kStatusOk ? len : 0;
    gather_lens[i] = (st == kStatusOk && h) ? len : 0;
    status[i] = st;
}

// One warp per item: dst[dst_offsets[i] ..) = the item's payload, taken from `alt` (packed, alt_offsets) when
// pick[i] != 0, else from the framed input behind its prefix. skip_unpicked: items with pick[i] == 0 are left
// out (the gather of Huffman payloads); otherwise they are copied as they are (raw literals in the final pass).
__global__ void __launch_bounds__(256) hpack_move_kernel(
    uint64_t n, const uint8_t *framed, const uint64_t *framed_offsets, const uint8_t *prefix_len, const uint64_t *pay_lens,
    const uint8_t *alt, const uint64_t *alt_offsets, const uint8_t *pick, bool pick_from_alt, bool skip_unpicked,
    const int32_t *status, uint8_t *dst, uint64_t dst_capacity, const uint64_t *dst_offsets) {
    const uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    if (status[i] != kStatusOk) return;
    const bool picked = pick[i] != 0;
    if (!picked && skip_unpicked) return;
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
    const uint64_t ncopy = min(len, room);
    const uint32_t lane = lane_id();
    for (uint64_t k = lane; k < ncopy; k += 32) dst[o0 + k] = src[k];
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
            else if (nb >= 8) st = kStatusInvalidPadding;                                 // a byte or more of padding (or EOS itself)
            else if (nb && (left_bits[i] >> (64 - nb)) != ((1ull << nb) - 1)) st = kStatusInvalidPadding;  // not a prefix of EOS
            if (st == kStatusOk) len = dec_offsets[i + 1] - dec_offsets[i];
        } else {
            len = pay_lens[i];
        }
    }
    status[i] = st;
    final_lens[i] = len;
}

}  // namespace hb
