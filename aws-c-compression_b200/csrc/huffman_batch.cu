// C-ABI of the batched Huffman codec (include/aws/compression/huffman_batch.h) and the host-side
// orchestration around the CUDA kernels: context (device tables, stream, scratch), host<->device
// staging for the host-pointer entry points, kernel selection and launch.
//
// No CPU fallback lives here: every failure to reach the device surfaces as
// AWS_ERROR_COMPRESSION_DEVICE_FAILURE.
#include <aws/compression/huffman_batch.h>
#include <aws/compression/hpack_string_batch.h>
#include <aws/compression/huffman_table_builder.h>

#include "../host/huffman_lut.h"
#include "device_common.cuh"
#include "generic_kernels.cuh"
#include "encode_tiled.cuh"
#include "encode_slots.cuh"
#include "encode_strings.cuh"
#include "hpack_literals.cuh"
#include "decode_fast.cuh"
#include "decode_rows.cuh"
#include "decode_slots.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

namespace {

using namespace hb;

#define HB_CUDA_TRY(expr)                                                                                              \
    do {                                                                                                               \
        cudaError_t hb_err_ = (expr);                                                                                  \
        if (hb_err_ != cudaSuccess) {                                                                                  \
            hb_note_cuda_error(hb_err_, #expr, __LINE__);                                                              \
            return aws_raise_error(AWS_ERROR_COMPRESSION_DEVICE_FAILURE);                                              \
        }                                                                                                              \
    } while (0)

void hb_note_cuda_error(cudaError_t err, const char *what, int line) {
    if (getenv("AWS_HUFFMAN_BATCH_DEBUG")) {
        fprintf(stderr, "aws-c-compression(b200): %s failed at line %d: %s\n", what, line, cudaGetErrorString(err));
    }
    (void)cudaGetLastError(); /* clear the sticky-less error slot */
}

}  // namespace

namespace hb_host {
// Device buffer that only ever grows.
struct GrowBuf {
    void *ptr = nullptr;
    size_t bytes = 0;

    cudaError_t reserve(size_t want) {
        if (want <= bytes) return cudaSuccess;
        size_t grown = std::max(want, bytes + bytes / 2);
        grown = (grown + 255) & ~size_t(255);
        void *fresh = nullptr;
        cudaError_t err = cudaMalloc(&fresh, grown);
        if (err != cudaSuccess) return err;
        if (ptr) cudaFree(ptr);
        ptr = fresh;
        bytes = grown;
        return cudaSuccess;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    template <typename T>
    T *as() const {
        return static_cast<T *>(ptr);
    }
};
// Per-launch kernel scratch (tile descriptors, chunk records, ...). One per concurrent stream.
struct Scratch {
    GrowBuf lens, tile_state, tile_first, chunks, chunk_lens, chunk_offsets, fused, slot_base, deferred, str_ctl, str_state,
        str_tiles, str_bits, str_bits_first, rows, row_pos, row_cnt, rows_ctl;
    void release() {
        GrowBuf *all[] = {&lens,          &tile_state, &tile_first,  &chunks,   &chunk_lens, &chunk_offsets,
                          &fused,         &slot_base,  &deferred,    &str_ctl,  &str_state,  &str_tiles,
                          &str_bits,      &str_bits_first, &rows,       &row_pos,     &row_cnt,  &rows_ctl};
        for (GrowBuf *g : all) g->release();
    }
};

constexpr int kLanes = 8;
struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t kernels_done = nullptr, retired = nullptr;
    bool in_flight = false;
    uint64_t *h_total = nullptr;  // pinned: the sub-batch's output size
    Scratch scratch;
    GrowBuf in, in_off, out, out_off, lens, status, consumed, aux32, aux8, aux64;
    void release() {
        GrowBuf *all[] = {&in, &in_off, &out, &out_off, &lens, &status, &consumed, &aux32, &aux8, &aux64};
        for (GrowBuf *g : all) g->release();
        scratch.release();
        if (h_total) cudaFreeHost(h_total);
        if (kernels_done) cudaEventDestroy(kernels_done);
        if (retired) cudaEventDestroy(retired);
        if (stream) cudaStreamDestroy(stream);
        h_total = nullptr;
        kernels_done = retired = nullptr;
        stream = nullptr;
    }
};
}  // namespace hb_host
using hb_host::GrowBuf;
using hb_host::Lane;
using hb_host::Scratch;

struct aws_huffman_batch_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    hb::DeviceTables tables{};
    uint2 *d_enc = nullptr;
    uint32_t *d_lut = nullptr;
    uint2 *d_lut2 = nullptr;  // 64-bit entries of the lean decode step (decode_fast.cuh)
    uint32_t lut2_count = 0, lut2_trap = 0;
    uint32_t lut_smem_entries = 0;
    uint64_t launches = 0;

    // kernel scratch for the device entry points (and the simple host path)
    hb_host::Scratch scratch;
    // pipelined host path: independent lanes (stream + scratch + staging) so that the H2D copy of one
    // sub-batch, the kernels of the next and the D2H copy of the previous one overlap
    hb_host::Lane lanes[hb_host::kLanes];
    int sm_count = 148;
    int enc_blocks_per_sm[2] = {0, 0};  // resident blocks per SM of encode_tiled_kernel<seg>
    int enc_slots_blocks_per_sm = 0;
    int str_blocks_per_sm[3] = {0, 0, 0};  // str_bits_kernel, str_pack_kernel, str_scan_kernel
    // staging for the host entry points
    GrowBuf s_in, s_in_off, s_out, s_out_off, s_caps, s_status, s_consumed, s_ovf_pattern, s_ovf_bits, s_left_bits,
        s_left_num, s_chain;
    // HPACK string literals (hpack_literals.cuh): packed payloads between the framing and the codec, per-item plans
    GrowBuf hp_pay, hp_pay_off, hp_dec, hp_dec_off, hp_huff, hp_prefix, hp_lens, hp_pay_lens, hp_dec_status, hp_left_bits,
        hp_left_num, hp_status;
    uint64_t *h_scalar = nullptr;  // pinned
    // The *_device entry points enqueue on the caller's stream but share ctx->scratch (tile descriptors, ...).
    // `scratch_free` is recorded after every such call; a call on ANOTHER stream waits for it first, so two calls
    // in flight on different streams never touch the scratch at the same time.
    cudaEvent_t scratch_free = nullptr;
    cudaStream_t scratch_stream = nullptr;
    bool scratch_busy = false;
    // switches, read once at context creation
    bool force_generic = false, no_slots = false, no_strings = false, no_fused_stream = false, no_rows = false;
    int rows_blocks_per_sm = 0, compact_blocks_per_sm = 0;
    bool no_slots_decode = false;
    size_t slots_static_smem = 0, dec_static_smem[2] = {0, 0};
};

namespace {

// Stream of a *_device call: the caller's, or the context's own; ordered behind the previous device call when
// that one ran on a different stream (they share ctx->scratch).
int device_call_begin(aws_huffman_batch_ctx *ctx, void *cuda_stream, cudaStream_t *st) {
    HB_CUDA_TRY(cudaSetDevice(ctx->device));
    *st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    if (ctx->scratch_busy && ctx->scratch_stream != *st) HB_CUDA_TRY(cudaStreamWaitEvent(*st, ctx->scratch_free, 0));
    return AWS_OP_SUCCESS;
}
int device_call_end(aws_huffman_batch_ctx *ctx, cudaStream_t st, int rc) {
    if (ctx->scratch_free && cudaEventRecord(ctx->scratch_free, st) == cudaSuccess) {
        ctx->scratch_stream = st;
        ctx->scratch_busy = true;
    } else {
        (void)cudaGetLastError();
    }
    return rc;
}

// The dynamic shared-memory limit belongs to a kernel ON A DEVICE, not to a context; tables differ in size. It is
// raised when a launch needs more than was ever asked for on that device (not on every call: the host path issues
// one call per 8 MiB sub-batch).
template <typename K>
int ensure_dynamic_smem(K kernel, int device, size_t bytes) {
    struct Seen {
        const void *kernel;
        size_t bytes[16];
    };
    static Seen seen[16] = {};  // (a handful of kernels ask; keyed by the kernel's address, then by device)
    const void *key = reinterpret_cast<const void *>(kernel);
    for (Seen &e : seen) {
        if (e.kernel != key && e.kernel != nullptr) continue;
        e.kernel = key;
        size_t &cur = e.bytes[device & 15];
        if (bytes > cur) {
            HB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            cur = bytes;
        }
        return AWS_OP_SUCCESS;
    }
    HB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return AWS_OP_SUCCESS;
}

#ifndef HB_LUT_ROOT_BITS
#define HB_LUT_ROOT_BITS 12
#endif
constexpr uint32_t kLutRootBits = HB_LUT_ROOT_BITS;
constexpr uint32_t kLutSubBits = 8;
constexpr uint32_t kLutMaxSmemEntries = 8192;  // 32 KiB

int check_batch(const aws_huffman_batch *b) {
    if (!b) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (b->n == 0) return AWS_OP_SUCCESS;
    if (!b->in_offsets || !b->out_offsets) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    return AWS_OP_SUCCESS;
}

hb::BatchView make_view(const aws_huffman_batch *b) {
    hb::BatchView v{};
    v.n = b->n;
    v.in = b->in;
    v.in_offsets = b->in_offsets;
    v.out = b->out;
    v.out_capacity = b->out_capacity;
    v.out_offsets = b->out_offsets;
    v.out_caps = b->out_caps;
    v.out_lens = b->out_lens;
    v.status = b->status;
    v.consumed = b->consumed;
    v.overflow_pattern = b->overflow_pattern;
    v.overflow_num_bits = b->overflow_num_bits;
    v.leftover_working_bits = b->leftover_working_bits;
    v.leftover_num_bits = b->leftover_num_bits;
    return v;
}

// lens -> offsets on the device
int launch_scan(
    aws_huffman_batch_ctx *ctx, Scratch &sc, const uint64_t *lens, uint64_t *offsets, uint64_t n, cudaStream_t stream,
    const uint32_t *gate = nullptr, uint32_t slots_of = 0) {
    const uint64_t tiles = std::max<uint64_t>(1, (n + kScanTile - 1) / kScanTile);
    const size_t state_bytes = tiles * sizeof(uint64_t) + 256;
    HB_CUDA_TRY(sc.tile_state.reserve(state_bytes));
    HB_CUDA_TRY(cudaMemsetAsync(sc.tile_state.ptr, 0, state_bytes, stream));
    uint64_t *state = sc.tile_state.as<uint64_t>();
    uint32_t *ticket = reinterpret_cast<uint32_t *>(state + tiles);
    scan_lens_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(lens, offsets, n, state, ticket, gate, slots_of);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

// Packed layout, many items, every symbol has a code: the tiled kernel with item-aligned thread ranges.
int encode_slots_on_device(
    aws_huffman_batch_ctx *ctx, Scratch &sc, const hb::BatchView &v, uint64_t total_in, cudaStream_t stream) {
    const uint64_t num_tiles_ub = (total_in / kEncSymsPerThread + v.n + kSlotsPerTile - 1) / kSlotsPerTile + 1;
    HB_CUDA_TRY(sc.slot_base.reserve((v.n + 1) * sizeof(uint64_t)));
    HB_CUDA_TRY(sc.tile_first.reserve((num_tiles_ub + 1) * sizeof(EncSlotTile)));
    // slot_base = exclusive scan of the items' slot counts, straight from the offsets
    if (launch_scan(ctx, sc, v.in_offsets, sc.slot_base.as<uint64_t>(), v.n, stream, nullptr, kEncSymsPerThread)) return AWS_OP_ERR;
    slot_tile_index_kernel<<<(unsigned)((num_tiles_ub + 1 + 255) / 256), 256, 0, stream>>>(
        v.in_offsets, sc.slot_base.as<uint64_t>(), v.n, total_in, num_tiles_ub + 1, sc.tile_first.as<EncSlotTile>());
    ++ctx->launches;
    const size_t state_bytes = num_tiles_ub * sizeof(uint64_t) + 256;
    HB_CUDA_TRY(sc.tile_state.reserve(state_bytes));
    HB_CUDA_TRY(cudaMemsetAsync(sc.tile_state.ptr, 0, state_bytes, stream));
    EncSlotArgs a{};
    a.in = v.in;
    a.in_offsets = v.in_offsets;
    a.n = v.n;
    a.total_in = total_in;
    a.out = v.out;
    a.out_capacity = v.out_capacity;
    a.out_offsets = v.out_offsets;
    a.slot_base = sc.slot_base.as<uint64_t>();
    a.tiles = sc.tile_first.as<EncSlotTile>();
    a.tile_state = sc.tile_state.as<uint64_t>();
    a.ticket = reinterpret_cast<uint32_t *>(a.tile_state + num_tiles_ub);
    a.num_tiles_ub = (uint32_t)num_tiles_ub;
    a.eos_padding = ctx->tables.eos_padding;
    if (!ctx->enc_slots_blocks_per_sm) {
        int per_sm = 0;
        HB_CUDA_TRY(cudaFuncSetAttribute(
            encode_slots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEncSlotSmemBytes));
        HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_slots_kernel, kEncBlock, kEncSlotSmemBytes));
        ctx->enc_slots_blocks_per_sm = std::max(1, per_sm);
    }
    const unsigned grid = (unsigned)std::min<uint64_t>(num_tiles_ub, (uint64_t)ctx->sm_count * ctx->enc_slots_blocks_per_sm);
    encode_slots_kernel<<<grid, kEncBlock, kEncSlotSmemBytes, stream>>>(ctx->tables.enc, a);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    if (v.out_lens || v.status || v.consumed || v.overflow_pattern || v.overflow_num_bits) {
        fill_packed_meta_kernel<<<(unsigned)((v.n + 255) / 256), 256, 0, stream>>>(
            v.n, v.in_offsets, v.out_offsets, v.out_lens, v.status, v.consumed, v.overflow_pattern,
            v.overflow_num_bits);
        ++ctx->launches;
        HB_CUDA_TRY(cudaGetLastError());
    }
    return AWS_OP_SUCCESS;
}

int encode_tiled_on_device(
    aws_huffman_batch_ctx *ctx, Scratch &sc, const hb::BatchView &v, uint64_t total_in, cudaStream_t stream,
    const uint32_t *gate);

// Packed layout, many short strings, every symbol has a code: one thread per string (encode_strings.cuh).
int encode_strings_on_device(
    aws_huffman_batch_ctx *ctx, Scratch &sc, const hb::BatchView &v, uint64_t total_in, cudaStream_t stream) {
    const uint64_t scan_tiles = (v.n + kStrScanTile - 1) / kStrScanTile;
    const uint64_t bits_tiles = (total_in + kBitsTileBytes - 1) / kBitsTileBytes;
    const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(v.out) & 15);
    // a symbol encodes to at most 4 bytes
    const uint64_t cap_tiles = (4 * total_in + v.n + phase) / kStrTileBytes + 4;
    HB_CUDA_TRY(sc.str_state.reserve(scan_tiles * sizeof(uint64_t) + 64));
    HB_CUDA_TRY(sc.str_ctl.reserve(kStrCtlWords * sizeof(uint32_t)));
    HB_CUDA_TRY(sc.str_tiles.reserve(cap_tiles * sizeof(uint32_t)));
    HB_CUDA_TRY(sc.str_bits.reserve((v.n + 8) * sizeof(uint32_t)));
    HB_CUDA_TRY(sc.str_bits_first.reserve((bits_tiles + 1) * sizeof(uint32_t)));
    StrArgs a{};
    a.in = v.in;
    a.in_offsets = v.in_offsets;
    a.n = v.n;
    a.total_in = total_in;
    a.out = v.out;
    a.out_capacity = v.out_capacity;
    a.out_offsets = v.out_offsets;
    a.tile_state = sc.str_state.as<uint64_t>();
    a.control = sc.str_ctl.as<uint32_t>();
    a.tile_first = sc.str_tiles.as<uint32_t>();
    a.num_measure_tiles = (uint32_t)scan_tiles;
    a.out_phase = phase;
    a.eos_padding = ctx->tables.eos_padding;
    StrPrepArgs p{};
    p.in_offsets = v.in_offsets;
    p.n = v.n;
    p.total_in = total_in;
    p.bits = sc.str_bits.as<uint32_t>();
    p.bits_tile_first = sc.str_bits_first.as<uint32_t>();
    p.tile_state = a.tile_state;
    p.control = a.control;
    p.tile_first = a.tile_first;
    p.num_bits_tiles = (uint32_t)bits_tiles;
    p.num_scan_tiles = (uint32_t)scan_tiles;
    StrBitsArgs m{};
    m.in = v.in;
    m.in_offsets = v.in_offsets;
    m.n = v.n;
    m.total_in = total_in;
    m.bits_tile_first = p.bits_tile_first;
    m.bits = p.bits;
    m.control = a.control;
    m.num_bits_tiles = (uint32_t)bits_tiles;
    if (!ctx->str_blocks_per_sm[0]) {
        int per_sm = 0;
        HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, str_bits_kernel, kBitsThreads, 0));
        ctx->str_blocks_per_sm[0] = std::max(1, per_sm);
        HB_CUDA_TRY(cudaFuncSetAttribute(str_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStrPackSmemBytes));
        HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, str_pack_kernel, kStrThreads, kStrPackSmemBytes));
        ctx->str_blocks_per_sm[1] = std::max(1, per_sm);
    }
    const unsigned grid_prep = (unsigned)std::min<uint64_t>((v.n + 256) / 256, (uint64_t)ctx->sm_count * 16);
    str_prep_kernel<<<grid_prep, 256, 0, stream>>>(p);
    const unsigned grid_b = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, bits_tiles), (uint64_t)ctx->sm_count * ctx->str_blocks_per_sm[0]);
    str_bits_kernel<<<grid_b, kBitsThreads, 0, stream>>>(ctx->tables.enc, m);
    const unsigned grid_s = (unsigned)scan_tiles;
    str_scan_kernel<<<grid_s, kStrThreads, 0, stream>>>(p.bits, a);
    const unsigned grid_p = (unsigned)std::min<uint64_t>(cap_tiles, (uint64_t)ctx->sm_count * ctx->str_blocks_per_sm[1]);
    str_pack_kernel<<<grid_p, kStrThreads, kStrPackSmemBytes, stream>>>(ctx->tables.enc, a);
    ctx->launches += 4;
    HB_CUDA_TRY(cudaGetLastError());
    // strings this path does not take (very long ones): the tiled kernel redoes the batch, on the device's own flag
    if (encode_tiled_on_device(ctx, sc, v, total_in, stream, a.control + kStrCtlFallback)) return AWS_OP_ERR;
    if (v.out_lens || v.status || v.consumed || v.overflow_pattern || v.overflow_num_bits) {
        fill_packed_meta_kernel<<<(unsigned)((v.n + 255) / 256), 256, 0, stream>>>(
            v.n, v.in_offsets, v.out_offsets, v.out_lens, v.status, v.consumed, v.overflow_pattern,
            v.overflow_num_bits);
        ++ctx->launches;
        HB_CUDA_TRY(cudaGetLastError());
    }
    return AWS_OP_SUCCESS;
}

// Packed layout, every symbol has a code: the tiled symbol-parallel kernel.
int encode_tiled_on_device(
    aws_huffman_batch_ctx *ctx, Scratch &sc, const hb::BatchView &v, uint64_t total_in, cudaStream_t stream,
    const uint32_t *gate = nullptr) {
    const uint64_t num_tiles = (total_in + kEncTile - 1) / kEncTile;
    const bool seg = v.n > 1;
    const size_t state_bytes = num_tiles * sizeof(uint64_t) + 256;
    HB_CUDA_TRY(sc.tile_state.reserve(state_bytes));
    HB_CUDA_TRY(cudaMemsetAsync(sc.tile_state.ptr, 0, state_bytes, stream));
    EncTiledArgs a{};
    a.in = v.in;
    a.in_offsets = v.in_offsets;
    a.n = v.n;
    a.total_in = total_in;
    a.out = v.out;
    a.out_capacity = v.out_capacity;
    a.out_offsets = v.out_offsets;
    a.tile_state = sc.tile_state.as<uint64_t>();
    a.ticket = reinterpret_cast<uint32_t *>(a.tile_state + num_tiles);
    a.num_tiles = (uint32_t)num_tiles;
    a.eos_padding = ctx->tables.eos_padding;
    a.gate = gate;
    if (seg) {
        HB_CUDA_TRY(sc.tile_first.reserve((num_tiles + 1) * sizeof(uint32_t)));
        a.tile_first = sc.tile_first.as<uint32_t>();
        tile_index_kernel<<<(unsigned)((num_tiles + 1 + 255) / 256), 256, 0, stream>>>(
            v.in_offsets, v.n, total_in, num_tiles, sc.tile_first.as<uint32_t>(), gate);
        ++ctx->launches;
    }
    if (!ctx->enc_blocks_per_sm[seg]) {
        int per_sm = 0;
        if (seg) {
            HB_CUDA_TRY(cudaFuncSetAttribute(
                encode_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEncSmemBytes));
            HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &per_sm, encode_tiled_kernel<true>, kEncBlock, kEncSmemBytes));
        } else {
            HB_CUDA_TRY(cudaFuncSetAttribute(
                encode_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEncSmemBytes));
            HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &per_sm, encode_tiled_kernel<false>, kEncBlock, kEncSmemBytes));
        }
        ctx->enc_blocks_per_sm[seg] = std::max(1, per_sm);
    }
    const unsigned grid =
        (unsigned)std::min<uint64_t>(num_tiles, (uint64_t)ctx->sm_count * ctx->enc_blocks_per_sm[seg]);
    if (seg) encode_tiled_kernel<true><<<grid, kEncBlock, kEncSmemBytes, stream>>>(ctx->tables.enc, a);
    else encode_tiled_kernel<false><<<grid, kEncBlock, kEncSmemBytes, stream>>>(ctx->tables.enc, a);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    if (gate) return AWS_OP_SUCCESS;  // (the caller fills the per-item arrays once)
    if (v.out_lens || v.status || v.consumed || v.overflow_pattern || v.overflow_num_bits) {
        fill_packed_meta_kernel<<<(unsigned)((v.n + 255) / 256), 256, 0, stream>>>(
            v.n, v.in_offsets, v.out_offsets, v.out_lens, v.status, v.consumed, v.overflow_pattern,
            v.overflow_num_bits);
        ++ctx->launches;
        HB_CUDA_TRY(cudaGetLastError());
    }
    return AWS_OP_SUCCESS;
}

int encode_on_device(
    aws_huffman_batch_ctx *ctx, Scratch &sc, hb::BatchView v, uint64_t total_in, cudaStream_t stream) {
    if (v.n == 0) return AWS_OP_SUCCESS;
    const bool force_generic = ctx->force_generic;
    // (the tiled kernel's multiply-add accumulator needs 1 << len to fit a word: codes of up to 31 bits)
    if (!v.resume && !v.out_caps && !ctx->tables.has_unknown && ctx->tables.max_len <= 31 && total_in > 0 && v.n < 0xffffffffull &&
        (total_in + kEncTile - 1) / kEncTile < 0xffffffffull && !force_generic) {
        // many short strings: one thread per string
        if (v.n > 1 && total_in / v.n <= 2048 && (reinterpret_cast<uintptr_t>(v.in) & 15) == 0 &&
            (4 * total_in + v.n) / kStrTileBytes < 0xfffffff0ull && !ctx->no_strings)
            return encode_strings_on_device(ctx, sc, v, total_in, stream);
        // many items of some length: thread ranges aligned to the items
        if (v.n > 1 && total_in / v.n >= 24 && total_in < (1ull << 31) && (reinterpret_cast<uintptr_t>(v.in) & 15) == 0 &&
            !ctx->no_slots)
            return encode_slots_on_device(ctx, sc, v, total_in, stream);
        return encode_tiled_on_device(ctx, sc, v, total_in, stream);
    }
    if (!v.out_lens) {
        HB_CUDA_TRY(sc.lens.reserve(v.n * sizeof(uint64_t)));
        v.out_lens = sc.lens.as<uint64_t>();
    }
    const unsigned blocks = (unsigned)((v.n + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (!v.out_caps) {
        encode_items_warp_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(ctx->tables, v);
        ++ctx->launches;
        HB_CUDA_TRY(cudaGetLastError());
        if (launch_scan(ctx, sc, v.out_lens, v.out_offsets, v.n, stream)) return AWS_OP_ERR;
    }
    encode_items_warp_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(ctx->tables, v);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

// Packed layout, many strings: fused count -> look-back -> write kernel (one thread per string).
int decode_batch_fast(
    aws_huffman_batch_ctx *ctx, Scratch &sc, const hb::BatchView &v, uint64_t total_in, cudaStream_t stream,
    bool framed = false) {
    DecBatchArgs a{};
    a.b = v;
    a.lut = ctx->tables.lut;
    a.lut2 = ctx->d_lut2;
    a.lut2_count = ctx->lut2_count;
    a.lut2_trap = ctx->lut2_trap;
    a.root_bits = ctx->tables.lut_root_bits;
    a.min_len = std::max<uint32_t>(1, ctx->tables.min_len);
    // Shared memory of one block (two blocks per SM): [LUT][stage][rows]. A string of L bytes decodes to at
    // most 8 L / min_len symbols, so the row area is that much larger than the stage; and the dense output
    // image (which reuses the stage and the front of the rows) must end before row offset `front`
    // (decode_batch_kernel step 4): rows <= stage + front - 32.
    const size_t lut_bytes = (size_t)a.lut2_count * 8;
    // one block per SM: [LUT2][team 0: stage, rows][team 1: stage, rows] + the teams' static arrays (what the kernel
    // declares is asked from the runtime: every KB of the SM's 227 is a string more per tile)
    if (!ctx->dec_static_smem[framed]) {
        cudaFuncAttributes attr{};
        HB_CUDA_TRY(framed ? cudaFuncGetAttributes(&attr, decode_batch_kernel<true>) : cudaFuncGetAttributes(&attr, decode_batch_kernel<false>));
        ctx->dec_static_smem[framed] = attr.sharedSizeBytes;
    }
    const size_t budget = (((size_t)227 * 1024 - ctx->dec_static_smem[framed] - lut_bytes) / kDecTeams - 128) & ~size_t(15);
    const double expand = 8.0 / a.min_len;
    size_t stage_bytes = (size_t)((double)budget / (1.0 + expand)) & ~size_t(15);
    stage_bytes = std::max<size_t>(stage_bytes, 2 * kDecMaxRow);
    size_t rows_bytes = budget - stage_bytes;
    const size_t front = (stage_bytes - kDecMaxRow - 32) & ~size_t(3);
    rows_bytes = std::min(rows_bytes, stage_bytes + front - 64) & ~size_t(15);
    a.stage_words = (uint32_t)(stage_bytes / 4 - 2);
    a.rows_bytes = (uint32_t)rows_bytes;
    // strings per tile: as many as fit the stage and the row area at the batch's average length, with room to spare
    // for tiles above the average (a tile that does not fit takes the two-pass global route and every tile behind it
    // waits for its count)
    {
        const double avg = std::max(1.0, (double)total_in / (double)v.n);
        // (17 % to spare: on the benchmark's uniform 8..256 B strings a tile's bytes vary by 3.2 % (one sigma); with 12 %
        // two tiles in 4,500 overflowed, 20 % cost the ninth group of the tile)
        const double by_stage = (double)(stage_bytes - 64) / (1.17 * avg);
        const double by_rows = (double)rows_bytes / (1.17 * avg * expand + kDecRowSlack);
        const uint64_t fit = (uint64_t)std::max(32.0, std::min(by_stage, by_rows));
        // (in steps of 8 strings: the last group of a tile may be partly empty — it holds the tile's shortest strings —
        // and every string more per tile is decode-phase time shared by one more: 272 instead of 256 on the benchmark)
        a.items_per_tile = (uint32_t)std::min<uint64_t>(kDecItemsPerTile, std::max<uint64_t>(32, fit & ~uint64_t(7)));
    }
    const uint64_t num_tiles = (v.n + a.items_per_tile - 1) / a.items_per_tile;
    const size_t state_bytes = num_tiles * sizeof(uint64_t) + 256;
    HB_CUDA_TRY(sc.tile_state.reserve(state_bytes));
    HB_CUDA_TRY(cudaMemsetAsync(sc.tile_state.ptr, 0, state_bytes, stream));
    a.tile_state = sc.tile_state.as<uint64_t>();
    a.ticket = reinterpret_cast<uint32_t *>(a.tile_state + num_tiles);
    a.num_tiles = (uint32_t)num_tiles;
    const size_t team_bytes = ((stage_bytes + 15) & ~size_t(15)) + rows_bytes + 64;
    const size_t smem = lut_bytes + kDecTeams * team_bytes;
    if (framed ? ensure_dynamic_smem(decode_batch_kernel<true>, ctx->device, smem) : ensure_dynamic_smem(decode_batch_kernel<false>, ctx->device, smem))
        return AWS_OP_ERR;
    const unsigned blocks = (unsigned)std::min<uint64_t>((num_tiles + kDecTeams - 1) / kDecTeams, (uint64_t)ctx->sm_count);
    if (framed) decode_batch_kernel<true><<<blocks, kDecTeams * kDecBlock, smem, stream>>>(a);
    else decode_batch_kernel<false><<<blocks, kDecTeams * kDecBlock, smem, stream>>>(a);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

// Packed layout, many strings: decode_slots_kernel (decode_slots.cuh) — decode_batch_kernel with stage and rows in
// one place, larger tiles for the same shared memory, rows straight to global memory.
int decode_batch_slots(
    aws_huffman_batch_ctx *ctx, Scratch &sc, const hb::BatchView &v, uint64_t total_in, cudaStream_t stream) {
    if (!ctx->slots_static_smem) {
        cudaFuncAttributes attr{};
        HB_CUDA_TRY(cudaFuncGetAttributes(&attr, decode_slots_kernel));
        ctx->slots_static_smem = attr.sharedSizeBytes;
        // (the limit belongs to the kernel, not to the context: set to what the device allows, launches ask for less)
        HB_CUDA_TRY(cudaFuncSetAttribute(
            decode_slots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024 - 1024 - attr.sharedSizeBytes)));
    }
    DecSlotsArgs a{};
    a.b = v;
    a.lut2 = ctx->d_lut2;
    a.lut2_count = ctx->lut2_count;
    a.lut2_trap = ctx->lut2_trap;
    a.root_bits = ctx->tables.lut_root_bits;
    a.min_len = std::min<uint32_t>(8, std::max<uint32_t>(1, ctx->tables.min_len));
    const size_t lut_bytes = (size_t)a.lut2_count * 8;
    const size_t avail = (size_t)227 * 1024 - 1024 - ctx->slots_static_smem - lut_bytes;
    const size_t region = (avail / kDecTeams) & ~size_t(15);
    a.region_bytes = (uint32_t)region;
    // strings per tile: as many slots as fit the region at the batch's average length with 15 % to spare, in whole
    // warps (a tile that does not fit takes the two-pass global route and every tile behind it waits for its count)
    {
        const double avg = std::max(1.0, (double)total_in / (double)v.n);
        const uint64_t fit = (uint64_t)std::max(32.0, (double)region / (1.15 * (avg * 8.0 / a.min_len + kSlotSlack)));
        a.items_per_tile = (uint32_t)std::min<uint64_t>(kSlotItemsPerTile, fit & ~uint64_t(31));
    }
    const uint64_t num_tiles = (v.n + a.items_per_tile - 1) / a.items_per_tile;
    const size_t state_bytes = num_tiles * sizeof(uint64_t) + 256;
    HB_CUDA_TRY(sc.tile_state.reserve(state_bytes));
    HB_CUDA_TRY(cudaMemsetAsync(sc.tile_state.ptr, 0, state_bytes, stream));
    a.tile_state = sc.tile_state.as<uint64_t>();
    a.ticket = reinterpret_cast<uint32_t *>(a.tile_state + num_tiles);
    a.num_tiles = (uint32_t)num_tiles;
    const size_t smem = lut_bytes + kDecTeams * region;
    const unsigned blocks = (unsigned)std::min<uint64_t>((num_tiles + kDecTeams - 1) / kDecTeams, (uint64_t)ctx->sm_count);
    decode_slots_kernel<<<blocks, kDecTeams * kDecBlock, smem, stream>>>(a);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

// Packed layout, many strings: decode into worst-case spaced rows of a global scratch (decode_rows_kernel), then
// counts -> offsets and rows -> dense output (compact_rows_kernel). decode_rows.cuh.
constexpr size_t kCompactStageBytes = 56 * 1024, kCompactImageBytes = 42 * 1024;  // two blocks per SM
int decode_batch_rows(
    aws_huffman_batch_ctx *ctx, Scratch &sc, const hb::BatchView &v, uint64_t total_in, cudaStream_t stream) {
    const uint32_t min_len = std::max<uint32_t>(1, ctx->tables.min_len);
    const uint64_t rows_bytes = 16ull * (total_in / (2ull * min_len) + 1) + (uint64_t)kRowsSlack * v.n + 256;
    const uint64_t num_pools = (v.n + kRowsPool - 1) / kRowsPool;
    // strings per compaction tile: as many as fit the stage (rows at their worst-case spacing) with 15 % to spare, and
    // the image if strings of average length decode to 85 % of their worst case (8 / min_len symbols per byte; the
    // benchmark's data: 1.36 of 1.6); even. A tile that does not fit takes the slow route.
    const double avg_row = std::max(1.0, (double)total_in * 8.0 / min_len / (double)v.n);
    uint64_t per_tile = (uint64_t)std::min((double)kCompactStageBytes / (1.15 * (avg_row + kRowsSlack)),
                                           (double)(kCompactImageBytes - 64) / (0.85 * 1.15 * avg_row));
    per_tile = std::max<uint64_t>(2, std::min<uint64_t>(kCompactMaxItems, per_tile & ~uint64_t(1)));
    const uint64_t num_tiles = (v.n + per_tile - 1) / per_tile;
    if (num_pools >= 0xffffffffull || num_tiles >= 0xffffffffull) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    HB_CUDA_TRY(sc.rows.reserve(rows_bytes));
    HB_CUDA_TRY(sc.row_pos.reserve(v.n * sizeof(uint64_t)));
    HB_CUDA_TRY(sc.row_cnt.reserve((v.n + 2) * sizeof(uint32_t)));
    // [look-back descriptors of the compaction: num_tiles][tickets: 2 words (+ pad)]
    const size_t ctl_bytes = num_tiles * sizeof(uint64_t) + 64;
    HB_CUDA_TRY(sc.rows_ctl.reserve(ctl_bytes));
    HB_CUDA_TRY(cudaMemsetAsync(sc.rows_ctl.ptr, 0, ctl_bytes, stream));
    uint32_t *tickets = reinterpret_cast<uint32_t *>(sc.rows_ctl.as<uint64_t>() + num_tiles);
    DecRowsArgs a{};
    a.b = v;
    a.total_in = total_in;
    a.lut2 = ctx->d_lut2;
    a.lut2_count = ctx->lut2_count;
    a.lut2_trap = ctx->lut2_trap;
    a.root_bits = ctx->tables.lut_root_bits;
    a.min_len = min_len;
    a.rows = sc.rows.as<uint8_t>();
    a.row_pos = sc.row_pos.as<uint64_t>();
    a.cnt = sc.row_cnt.as<uint32_t>();
    a.ticket = tickets;
    a.num_pools = (uint32_t)num_pools;
    const size_t rows_smem = (size_t)ctx->lut2_count * 8 + 1024 + (size_t)kRowsWarps * kRowsRingBytes;
    if (!ctx->rows_blocks_per_sm) {
        // (the attribute belongs to the kernel, not to the context, and tables differ in size: the limit is set to
        // what the device allows once, the launches ask for what they need)
        int per_sm = 0;
        HB_CUDA_TRY(cudaFuncSetAttribute(decode_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
        HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_rows_kernel, kRowsThreads, rows_smem));
        ctx->rows_blocks_per_sm = std::max(1, per_sm);
        HB_CUDA_TRY(cudaFuncSetAttribute(compact_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kCompactStageBytes + kCompactImageBytes)));
        HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, compact_rows_kernel, kCompactThreads, kCompactStageBytes + kCompactImageBytes));
        ctx->compact_blocks_per_sm = std::max(1, per_sm);
    }
    const unsigned grid_r = (unsigned)std::min<uint64_t>(num_pools, (uint64_t)ctx->sm_count * ctx->rows_blocks_per_sm);
    decode_rows_kernel<<<grid_r, kRowsThreads, rows_smem, stream>>>(a);
    CompactArgs c{};
    c.b = v;
    c.rows = a.rows;
    c.row_pos = a.row_pos;
    c.cnt = a.cnt;
    c.tile_state = sc.rows_ctl.as<uint64_t>();
    c.ticket = tickets + 1;
    c.num_tiles = (uint32_t)num_tiles;
    c.items_per_tile = (uint32_t)per_tile;
    c.stage_bytes = (uint32_t)kCompactStageBytes;
    c.image_bytes = (uint32_t)kCompactImageBytes;
    const unsigned grid_c = (unsigned)std::min<uint64_t>(num_tiles, (uint64_t)ctx->sm_count * ctx->compact_blocks_per_sm);
    compact_rows_kernel<<<grid_c, kCompactThreads, kCompactStageBytes + kCompactImageBytes, stream>>>(c);
    ctx->launches += 2;
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

// Packed layout, one long stream: chunked speculative decode. The fused single-pass kernel does the
// work; the multi-kernel path behind it only runs (on the device's own decision, no host round trip)
// when the fused kernel raised its `fail` flag: a stream that does not self-synchronise, or that stops
// at an unknown symbol before its end.
int decode_stream_fast(
    aws_huffman_batch_ctx *ctx, Scratch &sc, const hb::BatchView &v, uint64_t len, cudaStream_t stream) {
    const uint64_t lead = reinterpret_cast<uintptr_t>(v.in) & 15;  // "aligned space" starts on a 16-byte boundary
    const uint64_t end_bit = (lead + len) * 8;
    const uint64_t num_chunks = (end_bit + kChunkBits - 1) / kChunkBits;
    HB_CUDA_TRY(sc.chunks.reserve(num_chunks * sizeof(uint64_t) + 64));
    HB_CUDA_TRY(sc.chunk_lens.reserve(num_chunks * sizeof(uint64_t)));
    HB_CUDA_TRY(sc.chunk_offsets.reserve((num_chunks + 1) * sizeof(uint64_t)));
    StreamArgs a{};
    a.in_aligned = v.in - lead;
    a.begin_bit = lead * 8;
    a.end_bit = end_bit;
    a.num_chunks = num_chunks;
    a.chunks = sc.chunks.as<uint64_t>();
    a.chunk_offsets = sc.chunk_offsets.as<uint64_t>();
    a.control = a.chunks + num_chunks;  // two words after the records
    a.lut = ctx->tables.lut;
    a.lut_count = ctx->tables.lut_count;
    a.root_bits = ctx->tables.lut_root_bits;
    const size_t smem = ctx->tables.lut_count * sizeof(uint32_t);
    const size_t smem_staged = smem + kStreamStageWords * sizeof(uint32_t);

    // ---- fused single pass ------------------------------------------------------------------------------------
    const uint32_t min_len = std::max<uint32_t>(1, ctx->tables.min_len);
    const uint32_t row_words = (((kChunkBits + 31) / min_len + 4 + 3) / 4) | 1u;
    const size_t lut_bytes = (size_t)ctx->lut2_count * 8;
    const size_t stage_bytes = (kStreamStageWords * 4 + 15) & ~size_t(15);
    const size_t team_bytes = (stage_bytes + (size_t)kStreamThreads * row_words * 4 + 16 + 15) & ~size_t(15);
    const size_t fused_smem = lut_bytes + kStreamTeams * team_bytes;  // one block per SM, two teams, one table
    // (the dense output image reuses the stage and the front of the rows: stream_fused_kernel step 5)
    const bool fused = (size_t)kStreamThreads * row_words * 4 + 16 <= 2 * stage_bytes - row_words * 4 - 32 &&
                       fused_smem <= 220 * 1024 && !ctx->no_fused_stream;
    if (fused) {
        StreamFusedArgs f{};
        f.s = a;
        f.s.num_chunks = std::max<uint64_t>(1, (end_bit > 31 ? end_bit - 31 + kChunkBits - 1 : 0) / kChunkBits);
        f.b = v;
        f.num_tiles = (uint32_t)((f.s.num_chunks + kStreamThreads - 1) / kStreamThreads);
        f.row_words = row_words;
        f.lut2 = ctx->d_lut2;
        f.lut2_count = ctx->lut2_count;
        f.lut2_trap = ctx->lut2_trap;
        // [tile_state: num_tiles][tile_rec: num_tiles][ticket][fail]
        const size_t words = 2 * (size_t)f.num_tiles + 2;
        HB_CUDA_TRY(sc.fused.reserve(words * sizeof(uint64_t)));
        HB_CUDA_TRY(cudaMemsetAsync(sc.fused.ptr, 0, words * sizeof(uint64_t), stream));
        f.tile_state = sc.fused.as<uint64_t>();
        f.tile_rec = f.tile_state + f.num_tiles;
        f.ticket = reinterpret_cast<uint32_t *>(f.tile_rec + f.num_tiles);
        f.fail = f.ticket + 2;
        a.gate = f.fail;
        if (ensure_dynamic_smem(stream_fused_kernel, ctx->device, fused_smem)) return AWS_OP_ERR;
        const unsigned blocks = (unsigned)std::min<uint64_t>((f.num_tiles + kStreamTeams - 1) / kStreamTeams, (uint64_t)ctx->sm_count);
        stream_fused_kernel<<<blocks, kStreamTeams * kStreamTeamThreads, fused_smem, stream>>>(f);
        stream_fused_verify_kernel<<<(unsigned)std::min<uint64_t>((f.num_tiles + 255) / 256, 1024), 256, 0, stream>>>(f);
        ctx->launches += 2;
        HB_CUDA_TRY(cudaGetLastError());
    }

    // ---- multi-kernel path (alone, or gated behind the fused kernel's fail flag) ----------------------------------
    HB_CUDA_TRY(cudaMemsetAsync(a.control, 0xff, 2 * sizeof(uint64_t), stream));
    if (ensure_dynamic_smem(stream_sync_kernel, ctx->device, smem_staged) || ensure_dynamic_smem(stream_write_kernel, ctx->device, smem_staged))
        return AWS_OP_ERR;
    // a gated launch that finds the flag clear costs a few microseconds: keep those grids small
    const uint64_t cap_wide = fused ? (uint64_t)ctx->sm_count * 4 : (uint64_t)ctx->sm_count * 12;
    const uint64_t cap_flat = fused ? (uint64_t)ctx->sm_count * 4 : (uint64_t)ctx->sm_count * 8;
    const unsigned wide = (unsigned)std::min<uint64_t>((num_chunks + kStreamThreads - 1) / kStreamThreads, cap_wide);
    const unsigned flat = (unsigned)std::min<uint64_t>((num_chunks + 255) / 256, cap_flat);
    stream_sync_kernel<<<wide, kStreamThreads, smem_staged, stream>>>(a);
    for (int round = 0; round < 2; ++round) stream_fix_kernel<<<wide, kStreamThreads, smem, stream>>>(a);
    stream_verify_kernel<<<flat, 256, 0, stream>>>(a);
    stream_repair_kernel<<<1, 32, smem, stream>>>(a);
    stream_counts_kernel<<<flat, 256, 0, stream>>>(a, sc.chunk_lens.as<uint64_t>());
    ctx->launches += 6;
    HB_CUDA_TRY(cudaGetLastError());
    if (launch_scan(ctx, sc, sc.chunk_lens.as<uint64_t>(), a.chunk_offsets, num_chunks, stream, a.gate)) return AWS_OP_ERR;
    stream_write_kernel<<<wide, kStreamThreads, smem_staged, stream>>>(a, v);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

inline bool num_tiles_ok(uint64_t n) { return n / 32 < 0xfffffff0ull; }
constexpr uint64_t kStreamMinBytes = 64 * 1024;  // shorter single items go through the batch kernel

int decode_on_device(
    aws_huffman_batch_ctx *ctx, Scratch &sc, hb::BatchView v, uint64_t total_in, cudaStream_t stream) {
    if (v.n == 0) return AWS_OP_SUCCESS;
    const bool force_generic = ctx->force_generic;
    if (!v.resume && !v.out_caps && ctx->tables.lut_count <= kDecLutMaxSmem && !force_generic) {
        if (v.n == 1 && total_in >= kStreamMinBytes) return decode_stream_fast(ctx, sc, v, total_in, stream);
        if (!ctx->no_rows) return decode_batch_rows(ctx, sc, v, total_in, stream);
        if (!ctx->no_slots_decode && num_tiles_ok(v.n)) return decode_batch_slots(ctx, sc, v, total_in, stream);
        return decode_batch_fast(ctx, sc, v, total_in, stream);
    }
    if (!v.out_lens) {
        HB_CUDA_TRY(sc.lens.reserve(v.n * sizeof(uint64_t)));
        v.out_lens = sc.lens.as<uint64_t>();
    }
    const unsigned threads = 256;
    const unsigned blocks = (unsigned)((v.n + threads - 1) / threads);
    const size_t smem = ctx->lut_smem_entries * sizeof(uint32_t);
    if (!v.out_caps) {
        decode_items_thread_kernel<false><<<blocks, threads, smem, stream>>>(ctx->tables, v, ctx->lut_smem_entries);
        ++ctx->launches;
        HB_CUDA_TRY(cudaGetLastError());
        if (launch_scan(ctx, sc, v.out_lens, v.out_offsets, v.n, stream)) return AWS_OP_ERR;
    }
    decode_items_thread_kernel<true><<<blocks, threads, smem, stream>>>(ctx->tables, v, ctx->lut_smem_entries);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

__global__ void rebase_offsets_kernel(const uint64_t *src, uint64_t count, uint64_t base, uint64_t *dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[i] - base;
}

__global__ void add_base_kernel(uint64_t *offsets, uint64_t count, uint64_t base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) offsets[i] += base;
}

// Sub-batch j of the pipelined host path leaves the device WITHOUT the host knowing its size: this kernel reads the
// sub-batch's size and the running output position (written by sub-batch j - 1's instance; the streams are ordered
// by an event), publishes the next position, rebases the sub-batch's packed offsets and copies its payload straight
// into the caller's pinned buffer (`dst`: the device's alias of it) with 128-bit stores.
constexpr int kChainBlocks = 96, kChainThreads = 256;
constexpr uint32_t kChainSegment = 16 * 1024;  // bytes a block moves per turn (multiple of 16)
__global__ void __launch_bounds__(kChainThreads) d2h_chain_kernel(
    const uint8_t *src, uint64_t *out_off, uint64_t nj, const uint64_t *base_in, uint64_t *base_out, uint8_t *dst,
    uint64_t capacity, uint32_t *overflow) {
    const uint64_t base = base_in ? *base_in : 0;
    const uint64_t total = out_off[nj];
    const bool fits = base + total <= capacity;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *base_out = base + total;
        if (!fits) *overflow = 1u;
    }
    if (fits) {
        for (uint64_t seg = (uint64_t)blockIdx.x * kChainSegment; seg < total; seg += (uint64_t)gridDim.x * kChainSegment)
            hb::block_copy_realign(src + seg, dst + base + seg, (uint32_t)min((uint64_t)kChainSegment, total - seg), threadIdx.x, kChainThreads);
    }
    if (base)
        for (uint64_t i = (uint64_t)blockIdx.x * kChainThreads + threadIdx.x; i < nj; i += (uint64_t)gridDim.x * kChainThreads)
            out_off[i] += base;
}

template <typename T>
T *mapped_alias(T *p, uint64_t size);

int lane_prepare(Lane &lane) {
    if (lane.stream) return AWS_OP_SUCCESS;
    HB_CUDA_TRY(cudaStreamCreateWithFlags(&lane.stream, cudaStreamNonBlocking));
    HB_CUDA_TRY(cudaEventCreateWithFlags(&lane.kernels_done, cudaEventDisableTiming));
    HB_CUDA_TRY(cudaEventCreateWithFlags(&lane.retired, cudaEventDisableTiming));
    HB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&lane.h_total), sizeof(uint64_t), cudaHostAllocDefault));
    return AWS_OP_SUCCESS;
}

constexpr uint64_t kZeroCopyMinBytes = 4096;          // pinned payload buffers: kernels work on them in place
constexpr uint64_t kPipelineMinBytes = 8ull << 20;   // below this one shot is as good
constexpr uint64_t kPipelineShardBytes = 12ull << 20;  // measured on PCIe Gen5 x16: 4 / 6 / 8 / 12 / 16 MiB: 53.2 / 57.0 / 58.3 / 60.4 / 59.6 GB/s
                                                       // end to end (round 1, slower kernels: 8 MiB)

// Packed layout with many items: the batch is cut into sub-batches (contiguous item ranges balanced by
// bytes) that flow through kLanes lanes, so the host->device copy of one sub-batch, the kernels of the
// next and the device->host copy of the previous one run at the same time. Sub-batches are ordinary
// independent batches on the device; their packed offsets are rebased on the host at the end
// (the same concatenation step as multi-GPU sharding).
int run_host_batch_pipelined(aws_huffman_batch_ctx *ctx, const aws_huffman_batch *b, bool encode) {
    const size_t n = b->n;
    const uint64_t total_in = b->in_offsets[n];
    uint64_t shard_bytes = kPipelineShardBytes;
    if (const char *mb = getenv("AWS_HUFFMAN_BATCH_SHARD_MB")) shard_bytes = std::max<uint64_t>(1, atoi(mb)) << 20;
    size_t shards = (size_t)std::min<uint64_t>(256, std::max<uint64_t>(2, total_in / shard_bytes));
    shards = std::min(shards, n);
    std::vector<size_t> begin(shards + 1);
    if (aws_huffman_batch_plan_shards(b->in_offsets, n, shards, begin.data())) return AWS_OP_ERR;
    // One item larger than a sub-batch's share makes several byte targets fall inside it: plan_shards then
    // returns empty ranges. An empty sub-batch has nothing to issue or retire (its out_offsets[0] would never
    // be written on the device), so the ranges are compacted first.
    begin.erase(std::unique(begin.begin(), begin.end()), begin.end());
    shards = begin.size() - 1;
    // (cutting the first and last sub-batches smaller, to shorten pipeline fill and drain, was measured: no
    // effect — 59.5 vs 60.2 GB/s end to end)
    std::vector<uint64_t> shard_base(shards + 1, 0);

    for (Lane &lane : ctx->lanes)
        if (lane_prepare(lane)) return AWS_OP_ERR;

    const uint32_t max_len = std::max<uint32_t>(1, ctx->tables.max_len);
    const uint32_t min_len = std::max<uint32_t>(1, ctx->tables.min_len);
    // a sub-batch is retired `depth` issues after it was issued; more lanes than that keep re-use from waiting
    size_t depth = 1, lanes_used = 4;  // (round 2, 12 MiB sub-batches: depth 1 / 4 lanes 60.9, depth 2 / 5 lanes 59.7, depth 1 / 3 lanes 59.6 GB/s)
    if (const char *d = getenv("AWS_HUFFMAN_BATCH_PIPE_DEPTH")) depth = std::min<size_t>(6, std::max(1, atoi(d)));
    if (const char *l = getenv("AWS_HUFFMAN_BATCH_PIPE_LANES")) lanes_used = std::min<size_t>(hb_host::kLanes, std::max(2, atoi(l)));
    lanes_used = std::max(lanes_used, depth + 2);
    lanes_used = std::min<size_t>(lanes_used, hb_host::kLanes);
    uint64_t base_out = 0;
    bool overflow = false;
    // CHAINED mode (opt-in: AWS_HUFFMAN_BATCH_CHAIN=1 and a pinned output buffer): the host never learns a sub-batch's size
    // while the call runs. A small kernel behind each sub-batch's codec kernels (d2h_chain_kernel) reads the size on the
    // device, takes the running output position from its predecessor (streams ordered by an event), and writes the
    // payload straight into the caller's buffer; the fixed-size arrays follow by ordinary copies. No
    // cudaEventSynchronize per sub-batch, no copy that waits for the host to have seen a size, no drain of late
    // sub-batches at the end. Correct (tests/test_gpu_multi.py) and MEASURED SLOWER than the copy engines: 53.6 GB/s end
    // to end against 60.4 on the 1M-string batch (4 / 6 / 8 lanes: 56.5 / 53.6 / 48 GB/s: the more sub-batches in
    // flight, the worse — kernel-issued PCIe writes hold their SMs while the copy engines work for free). Round 1
    // measured the same for whole kernels working on pinned buffers; it holds for a dedicated copy kernel too.
    uint8_t *const zout = getenv("AWS_HUFFMAN_BATCH_CHAIN") ? mapped_alias(b->out, b->out_capacity) : nullptr;
    const bool chained = zout != nullptr;
    uint64_t *chain = nullptr;      // [0 .. shards]: output position before sub-batch j; [shards + 1]: overflow flag
    cudaEvent_t prev_copied = nullptr;
    if (chained) {
        HB_CUDA_TRY(ctx->s_chain.reserve((shards + 2) * sizeof(uint64_t)));
        chain = ctx->s_chain.as<uint64_t>();
        HB_CUDA_TRY(cudaMemsetAsync(chain + shards + 1, 0, sizeof(uint64_t), ctx->lanes[0].stream));
    }

    auto issue = [&](size_t j) -> int {
        Lane &lane = ctx->lanes[j % lanes_used];
        if (lane.in_flight) {
            HB_CUDA_TRY(cudaEventSynchronize(lane.retired));
            lane.in_flight = false;
        }
        const size_t a = begin[j], nj = begin[j + 1] - a;
        const uint64_t in0 = b->in_offsets[a], bytes_in = b->in_offsets[a + nj] - in0;
        const uint64_t room = encode ? bytes_in * ((max_len + 7) / 8) + 64 : (bytes_in * 8) / min_len + 64;
        HB_CUDA_TRY(lane.in.reserve(bytes_in + 64));
        HB_CUDA_TRY(lane.in_off.reserve((nj + 1) * sizeof(uint64_t)));
        HB_CUDA_TRY(lane.out.reserve(room + 64));
        HB_CUDA_TRY(lane.out_off.reserve((nj + 1) * sizeof(uint64_t)));
        if (b->out_lens) HB_CUDA_TRY(lane.lens.reserve(nj * sizeof(uint64_t)));
        if (b->status) HB_CUDA_TRY(lane.status.reserve(nj * sizeof(int32_t)));
        if (b->consumed) HB_CUDA_TRY(lane.consumed.reserve(nj * sizeof(uint64_t)));
        if (encode ? b->overflow_pattern != nullptr : false) HB_CUDA_TRY(lane.aux32.reserve(nj * sizeof(uint32_t)));
        if (encode ? b->overflow_num_bits != nullptr : b->leftover_num_bits != nullptr) HB_CUDA_TRY(lane.aux8.reserve(nj));
        if (!encode && b->leftover_working_bits) HB_CUDA_TRY(lane.aux64.reserve(nj * sizeof(uint64_t)));

        // the sub-batch's slice of the offsets goes up with it and is rebased in place on the device
        HB_CUDA_TRY(cudaMemcpyAsync(
            lane.in_off.ptr, b->in_offsets + a, (nj + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, lane.stream));
        if (bytes_in)
            HB_CUDA_TRY(cudaMemcpyAsync(lane.in.ptr, b->in + in0, bytes_in, cudaMemcpyHostToDevice, lane.stream));
        rebase_offsets_kernel<<<(unsigned)((nj + 1 + 255) / 256), 256, 0, lane.stream>>>(
            lane.in_off.as<uint64_t>(), nj + 1, in0, lane.in_off.as<uint64_t>());
        ++ctx->launches;
        hb::BatchView v{};
        v.n = nj;
        v.in = lane.in.as<uint8_t>();
        v.in_offsets = lane.in_off.as<uint64_t>();
        v.out = lane.out.as<uint8_t>();
        v.out_capacity = room;
        v.out_offsets = lane.out_off.as<uint64_t>();
        v.out_lens = b->out_lens ? lane.lens.as<uint64_t>() : nullptr;
        v.status = b->status ? lane.status.as<int32_t>() : nullptr;
        v.consumed = b->consumed ? lane.consumed.as<uint64_t>() : nullptr;
        if (encode) {
            v.overflow_pattern = b->overflow_pattern ? lane.aux32.as<uint32_t>() : nullptr;
            v.overflow_num_bits = b->overflow_num_bits ? lane.aux8.as<uint8_t>() : nullptr;
        } else {
            v.leftover_working_bits = b->leftover_working_bits ? lane.aux64.as<uint64_t>() : nullptr;
            v.leftover_num_bits = b->leftover_num_bits ? lane.aux8.as<uint8_t>() : nullptr;
        }
        if ((encode ? encode_on_device(ctx, lane.scratch, v, bytes_in, lane.stream)
                    : decode_on_device(ctx, lane.scratch, v, bytes_in, lane.stream)) != AWS_OP_SUCCESS)
            return AWS_OP_ERR;
        if (chained) {
            cudaStream_t st = lane.stream;
            if (prev_copied) HB_CUDA_TRY(cudaStreamWaitEvent(st, prev_copied, 0));
            d2h_chain_kernel<<<kChainBlocks, kChainThreads, 0, st>>>(
                lane.out.as<uint8_t>(), lane.out_off.as<uint64_t>(), nj, j ? chain + j : nullptr, chain + j + 1, zout,
                b->out_capacity, reinterpret_cast<uint32_t *>(chain + shards + 1));
            ++ctx->launches;
            HB_CUDA_TRY(cudaGetLastError());
            HB_CUDA_TRY(cudaEventRecord(lane.kernels_done, st));
            prev_copied = lane.kernels_done;
            HB_CUDA_TRY(cudaMemcpyAsync(b->out_offsets + a, lane.out_off.ptr, nj * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            if (b->out_lens)
                HB_CUDA_TRY(cudaMemcpyAsync(b->out_lens + a, lane.lens.ptr, nj * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            if (b->status)
                HB_CUDA_TRY(cudaMemcpyAsync(b->status + a, lane.status.ptr, nj * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            if (b->consumed)
                HB_CUDA_TRY(cudaMemcpyAsync(b->consumed + a, lane.consumed.ptr, nj * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            if (encode) {
                if (b->overflow_pattern)
                    HB_CUDA_TRY(cudaMemcpyAsync(b->overflow_pattern + a, lane.aux32.ptr, nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                if (b->overflow_num_bits)
                    HB_CUDA_TRY(cudaMemcpyAsync(b->overflow_num_bits + a, lane.aux8.ptr, nj, cudaMemcpyDeviceToHost, st));
            } else {
                if (b->leftover_working_bits)
                    HB_CUDA_TRY(cudaMemcpyAsync(b->leftover_working_bits + a, lane.aux64.ptr, nj * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
                if (b->leftover_num_bits)
                    HB_CUDA_TRY(cudaMemcpyAsync(b->leftover_num_bits + a, lane.aux8.ptr, nj, cudaMemcpyDeviceToHost, st));
            }
            HB_CUDA_TRY(cudaEventRecord(lane.retired, st));
            lane.in_flight = true;
            return AWS_OP_SUCCESS;
        }
        HB_CUDA_TRY(cudaMemcpyAsync(
            lane.h_total, lane.out_off.as<uint64_t>() + nj, sizeof(uint64_t), cudaMemcpyDeviceToHost, lane.stream));
        HB_CUDA_TRY(cudaEventRecord(lane.kernels_done, lane.stream));
        lane.in_flight = true;
        return AWS_OP_SUCCESS;
    };

    auto retire = [&](size_t j) -> int {
        Lane &lane = ctx->lanes[j % lanes_used];
        const size_t a = begin[j], nj = begin[j + 1] - a;
        HB_CUDA_TRY(cudaEventSynchronize(lane.kernels_done));
        const uint64_t total = *lane.h_total;
        shard_base[j] = base_out;
        cudaStream_t st = lane.stream;
        if (base_out + total > b->out_capacity) overflow = true;
        if (!overflow && total)
            HB_CUDA_TRY(cudaMemcpyAsync(b->out + base_out, lane.out.ptr, total, cudaMemcpyDeviceToHost, st));
        if (base_out) {
            // the concatenation step: shard-local packed offsets -> global, before they leave the device
            add_base_kernel<<<(unsigned)((nj + 255) / 256), 256, 0, st>>>(lane.out_off.as<uint64_t>(), nj, base_out);
            ++ctx->launches;
        }
        HB_CUDA_TRY(cudaMemcpyAsync(b->out_offsets + a, lane.out_off.ptr, nj * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        if (b->out_lens)
            HB_CUDA_TRY(cudaMemcpyAsync(b->out_lens + a, lane.lens.ptr, nj * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        if (b->status)
            HB_CUDA_TRY(cudaMemcpyAsync(b->status + a, lane.status.ptr, nj * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (b->consumed)
            HB_CUDA_TRY(cudaMemcpyAsync(b->consumed + a, lane.consumed.ptr, nj * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        if (encode) {
            if (b->overflow_pattern)
                HB_CUDA_TRY(cudaMemcpyAsync(b->overflow_pattern + a, lane.aux32.ptr, nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            if (b->overflow_num_bits)
                HB_CUDA_TRY(cudaMemcpyAsync(b->overflow_num_bits + a, lane.aux8.ptr, nj, cudaMemcpyDeviceToHost, st));
        } else {
            if (b->leftover_working_bits)
                HB_CUDA_TRY(cudaMemcpyAsync(b->leftover_working_bits + a, lane.aux64.ptr, nj * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            if (b->leftover_num_bits)
                HB_CUDA_TRY(cudaMemcpyAsync(b->leftover_num_bits + a, lane.aux8.ptr, nj, cudaMemcpyDeviceToHost, st));
        }
        HB_CUDA_TRY(cudaEventRecord(lane.retired, st));
        base_out += total;
        return AWS_OP_SUCCESS;
    };

    // A sub-batch is retired (its D2H copies queued) two issues after it was issued; with more lanes than
    // that, re-using a lane never has to wait for a D2H copy that is still running, so uploads of later
    // sub-batches and downloads of earlier ones stay concurrent.
    static_assert(hb_host::kLanes > 3, "lanes must outnumber the issue-to-retire distance");
    // On an error no copy into the caller's buffers may still be running when the call returns.
    auto drain = [&]() {
        for (Lane &lane : ctx->lanes) {
            if (lane.stream) (void)cudaStreamSynchronize(lane.stream);
            lane.in_flight = false;
        }
        (void)cudaGetLastError();
    };
    if (chained) depth = 0;  // (nothing to retire: a sub-batch is complete once issued)
    if (chained && !getenv("AWS_HUFFMAN_BATCH_PIPE_LANES")) lanes_used = 6;  // (a lane is re-used only when everything of its last sub-batch has left)
    for (size_t j = 0; j < shards + depth; ++j) {
        if ((j < shards && issue(j)) || (!chained && j >= depth && j - depth < shards && retire(j - depth))) {
            const int err = aws_last_error();
            drain();
            return aws_raise_error(err ? err : AWS_ERROR_COMPRESSION_DEVICE_FAILURE);
        }
    }
    for (Lane &lane : ctx->lanes) {
        if (lane.in_flight) HB_CUDA_TRY(cudaEventSynchronize(lane.retired));
        lane.in_flight = false;
    }
    if (chained) {
        uint64_t tail[2] = {0, 0};  // the output position behind the last sub-batch, the overflow flag
        HB_CUDA_TRY(cudaMemcpy(tail, chain + shards, sizeof(tail), cudaMemcpyDeviceToHost));
        base_out = tail[0];
        overflow = tail[1] != 0;
    }
    b->out_offsets[n] = base_out;
    if (overflow) return aws_raise_error(AWS_ERROR_SHORT_BUFFER);
    return AWS_OP_SUCCESS;
}

// Host-pointer entry point shared by encode and decode: stage in, run, stage out.
int check_resume(const aws_huffman_batch *b, bool encode) {
    // the state arrays are inputs as well as outputs
    if (b && b->n &&
        (encode ? (!b->overflow_pattern || !b->overflow_num_bits) : (!b->leftover_working_bits || !b->leftover_num_bits)))
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    return AWS_OP_SUCCESS;
}

// The device's alias of [p, p + size) when the range lies in pinned (page-locked, mapped) host memory, else
// nullptr. Kernels can then read and write the caller's buffers in place over PCIe (zero copy).
template <typename T>
T *mapped_alias(T *p, uint64_t size) {
    if (!p || !size) return nullptr;
    cudaPointerAttributes first{}, last{};
    if (cudaPointerGetAttributes(&first, p) != cudaSuccess ||
        cudaPointerGetAttributes(&last, reinterpret_cast<const uint8_t *>(p) + size - 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (first.type != cudaMemoryTypeHost || last.type != cudaMemoryTypeHost || !first.devicePointer || !last.devicePointer)
        return nullptr;
    // one allocation: the aliases are as far apart as the host addresses
    if (reinterpret_cast<const uint8_t *>(last.devicePointer) - reinterpret_cast<const uint8_t *>(first.devicePointer) !=
        (ptrdiff_t)(size - 1))
        return nullptr;
    return reinterpret_cast<T *>(first.devicePointer);
}

int run_host_batch(aws_huffman_batch_ctx *ctx, const aws_huffman_batch *b, bool encode, bool resume = false) {
    if (!ctx) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (check_batch(b)) return AWS_OP_ERR;
    if (resume && check_resume(b, encode)) return AWS_OP_ERR;
    const size_t n = b->n;
    if (n == 0) {
        if (!b->out_caps && b->out_offsets) b->out_offsets[0] = 0;
        return AWS_OP_SUCCESS;
    }
    HB_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t total_in = b->in_offsets[n];
    const bool slotted = b->out_caps != nullptr;
    if ((total_in && !b->in) || (b->out_capacity && !b->out)) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    // ZERO COPY (opt-in: AWS_HUFFMAN_BATCH_ZEROCOPY=1, AWS_HUFFMAN_BATCH_ZC_MODE = 1 input, 2 output, 3 both):
    // payload buffers in pinned host memory are read and written by the kernels themselves, in place, over
    // PCIe; only the small per-item arrays are staged. Correct (tests/test_gpu_batch.py::test_zero_copy_host_path)
    // but MEASURED SLOWER than the copy engines on B200 / PCIe Gen5: kernel-issued reads of host memory run at
    // 18 GB/s, kernel-issued writes at 48 GB/s, against 55 / 57 GB/s for cudaMemcpyAsync; end to end
    // 41.8 vs 58.5 GB/s on the 1M-string batch and 26.6 vs 52.1 GB/s on the 1 GiB stream. Kept for callers
    // with small latency-bound batches and as the measured record of why the host path stages its copies.
    const uint8_t *zin = nullptr;
    uint8_t *zout = nullptr;
    if (!resume && !slotted && total_in >= kZeroCopyMinBytes && getenv("AWS_HUFFMAN_BATCH_ZEROCOPY")) {
        zin = mapped_alias(b->in, total_in);
        zout = zin ? mapped_alias(b->out, b->out_capacity) : nullptr;
    }
    int zc_mode = 3;
    if (const char *m = getenv("AWS_HUFFMAN_BATCH_ZC_MODE")) zc_mode = atoi(m);
    const bool zc_possible = zin && zout;
    const bool zero_copy_in = zc_possible && (zc_mode & 1), zero_copy_out = zc_possible && (zc_mode & 2);
    const bool zero_copy = zero_copy_in || zero_copy_out;
    if (!zero_copy && !resume && !slotted && n >= 2 && total_in >= kPipelineMinBytes && !getenv("AWS_HUFFMAN_BATCH_NO_PIPELINE"))
        return run_host_batch_pipelined(ctx, b, encode);

    if (!zero_copy_in) HB_CUDA_TRY(ctx->s_in.reserve(total_in + 16));
    HB_CUDA_TRY(ctx->s_in_off.reserve((n + 1) * sizeof(uint64_t)));
    if (!zero_copy_out) HB_CUDA_TRY(ctx->s_out.reserve(b->out_capacity + 16));
    HB_CUDA_TRY(ctx->s_out_off.reserve((n + 1) * sizeof(uint64_t)));
    HB_CUDA_TRY(ctx->scratch.lens.reserve(n * sizeof(uint64_t)));
    if (slotted) HB_CUDA_TRY(ctx->s_caps.reserve(n * sizeof(uint64_t)));
    if (b->status) HB_CUDA_TRY(ctx->s_status.reserve(n * sizeof(int32_t)));
    if (b->consumed) HB_CUDA_TRY(ctx->s_consumed.reserve(n * sizeof(uint64_t)));
    if (encode && b->overflow_pattern) HB_CUDA_TRY(ctx->s_ovf_pattern.reserve(n * sizeof(uint32_t)));
    if (encode && b->overflow_num_bits) HB_CUDA_TRY(ctx->s_ovf_bits.reserve(n));
    if (!encode && b->leftover_working_bits) HB_CUDA_TRY(ctx->s_left_bits.reserve(n * sizeof(uint64_t)));
    if (!encode && b->leftover_num_bits) HB_CUDA_TRY(ctx->s_left_num.reserve(n));

    if (total_in && !zero_copy_in) HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_in.ptr, b->in, total_in, cudaMemcpyHostToDevice, st));
    HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_in_off.ptr, b->in_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    if (slotted) {
        HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_out_off.ptr, b->out_offsets, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_caps.ptr, b->out_caps, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        // bytes the codec does not touch must come back unchanged
        if (b->out_capacity)
            HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_out.ptr, b->out, b->out_capacity, cudaMemcpyHostToDevice, st));
    }

    if (resume) {
        if (encode) {
            HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_ovf_pattern.ptr, b->overflow_pattern, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_ovf_bits.ptr, b->overflow_num_bits, n, cudaMemcpyHostToDevice, st));
        } else {
            HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_left_bits.ptr, b->leftover_working_bits, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
            HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_left_num.ptr, b->leftover_num_bits, n, cudaMemcpyHostToDevice, st));
        }
    }

    hb::BatchView v{};
    v.resume = resume;
    v.n = n;
    v.in = zero_copy_in ? zin : ctx->s_in.as<uint8_t>();
    v.in_offsets = ctx->s_in_off.as<uint64_t>();
    v.out = zero_copy_out ? zout : ctx->s_out.as<uint8_t>();
    v.out_capacity = b->out_capacity;
    v.out_offsets = ctx->s_out_off.as<uint64_t>();
    v.out_caps = slotted ? ctx->s_caps.as<uint64_t>() : nullptr;
    v.out_lens = ctx->scratch.lens.as<uint64_t>();
    v.status = b->status ? ctx->s_status.as<int32_t>() : nullptr;
    v.consumed = b->consumed ? ctx->s_consumed.as<uint64_t>() : nullptr;
    if (encode) {
        v.overflow_pattern = b->overflow_pattern ? ctx->s_ovf_pattern.as<uint32_t>() : nullptr;
        v.overflow_num_bits = b->overflow_num_bits ? ctx->s_ovf_bits.as<uint8_t>() : nullptr;
    } else {
        v.leftover_working_bits = b->leftover_working_bits ? ctx->s_left_bits.as<uint64_t>() : nullptr;
        v.leftover_num_bits = b->leftover_num_bits ? ctx->s_left_num.as<uint8_t>() : nullptr;
    }

    v.out_lens = b->out_lens ? ctx->scratch.lens.as<uint64_t>() : nullptr;
    if ((encode ? encode_on_device(ctx, ctx->scratch, v, total_in, st)
                : decode_on_device(ctx, ctx->scratch, v, total_in, st)) != AWS_OP_SUCCESS)
        return AWS_OP_ERR;

    if (b->out_lens)
        HB_CUDA_TRY(cudaMemcpyAsync(b->out_lens, ctx->scratch.lens.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (b->status) HB_CUDA_TRY(cudaMemcpyAsync(b->status, v.status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (b->consumed)
        HB_CUDA_TRY(cudaMemcpyAsync(b->consumed, v.consumed, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (encode) {
        if (b->overflow_pattern)
            HB_CUDA_TRY(cudaMemcpyAsync(
                b->overflow_pattern, v.overflow_pattern, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (b->overflow_num_bits)
            HB_CUDA_TRY(cudaMemcpyAsync(b->overflow_num_bits, v.overflow_num_bits, n, cudaMemcpyDeviceToHost, st));
    } else {
        if (b->leftover_working_bits)
            HB_CUDA_TRY(cudaMemcpyAsync(
                b->leftover_working_bits, v.leftover_working_bits, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        if (b->leftover_num_bits)
            HB_CUDA_TRY(cudaMemcpyAsync(b->leftover_num_bits, v.leftover_num_bits, n, cudaMemcpyDeviceToHost, st));
    }

    if (slotted) {
        if (b->out_capacity)
            HB_CUDA_TRY(cudaMemcpyAsync(b->out, v.out, b->out_capacity, cudaMemcpyDeviceToHost, st));
        HB_CUDA_TRY(cudaStreamSynchronize(st));
        return AWS_OP_SUCCESS;
    }

    HB_CUDA_TRY(
        cudaMemcpyAsync(b->out_offsets, v.out_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    const uint64_t total_out = b->out_offsets[n];
    if (total_out > b->out_capacity) return aws_raise_error(AWS_ERROR_SHORT_BUFFER);
    if (total_out && !zero_copy_out) {
        HB_CUDA_TRY(cudaMemcpyAsync(b->out, v.out, total_out, cudaMemcpyDeviceToHost, st));
        HB_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return AWS_OP_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// HPACK string literals. Device pointers in, device pointers out.
// ---------------------------------------------------------------------------------------------
int hpack_encode_on_device(
    aws_huffman_batch_ctx *ctx, uint64_t n, const uint8_t *raw, const uint64_t *raw_off, uint64_t total_in, uint32_t mode,
    uint8_t *out, uint64_t out_capacity, uint64_t *out_off, cudaStream_t st) {
    if (n == 0) return AWS_OP_SUCCESS;
    const unsigned flat = (unsigned)((n + 255) / 256), warps = (unsigned)((n + 255) / 256);  // (move kernels: 32 items per warp)
    HB_CUDA_TRY(ctx->hp_huff.reserve(n));
    HB_CUDA_TRY(ctx->hp_lens.reserve(n * sizeof(uint64_t)));
    const uint8_t *pay = nullptr;
    const uint64_t *pay_off = nullptr;
    if (mode != hb::kHpackNever) {
        const uint64_t cap = total_in * ((std::max<uint32_t>(1, ctx->tables.max_len) + 7) / 8) + 64;
        HB_CUDA_TRY(ctx->hp_pay.reserve(cap + 64));
        HB_CUDA_TRY(ctx->hp_pay_off.reserve((n + 1) * sizeof(uint64_t)));
        hb::BatchView v{};
        v.n = n;
        v.in = raw;
        v.in_offsets = raw_off;
        v.out = ctx->hp_pay.as<uint8_t>();
        v.out_capacity = cap;
        v.out_offsets = ctx->hp_pay_off.as<uint64_t>();
        if (encode_on_device(ctx, ctx->scratch, v, total_in, st)) return AWS_OP_ERR;
        pay = ctx->hp_pay.as<uint8_t>();
        pay_off = ctx->hp_pay_off.as<uint64_t>();
    }
    hb::hpack_plan_kernel<<<flat, 256, 0, st>>>(n, raw_off, pay_off, mode, ctx->hp_huff.as<uint8_t>(), ctx->hp_lens.as<uint64_t>());
    ++ctx->launches;
    if (launch_scan(ctx, ctx->scratch, ctx->hp_lens.as<uint64_t>(), out_off, n, st)) return AWS_OP_ERR;
    hb::hpack_frame_kernel<<<warps, 256, 0, st>>>(n, raw, raw_off, pay, pay_off, ctx->hp_huff.as<uint8_t>(), out, out_capacity, out_off);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

int hpack_decode_on_device(
    aws_huffman_batch_ctx *ctx, uint64_t n, const uint8_t *in, const uint64_t *in_off, uint64_t total_in, uint8_t *out,
    uint64_t out_capacity, uint64_t *out_off, int32_t *status, cudaStream_t st) {
    if (n == 0) return AWS_OP_SUCCESS;
    // (the one-kernel route copies raw literals into rows sized 8 len / min_len: a table whose shortest code is longer
    // than a byte — not HPACK's — takes the multi-pass route)
    if (n > 1 && ctx->tables.lut_count <= kDecLutMaxSmem && ctx->tables.min_len <= 8 && !getenv("AWS_HUFFMAN_BATCH_FORCE_GENERIC") &&
        !getenv("AWS_HUFFMAN_HPACK_PASSES")) {
        // one kernel: the batch decoder parses the literals in its string table, copies raw payloads, applies
        // the padding rule to what it leaves over and writes strings, offsets and status itself
        hb::BatchView v{};
        v.n = n;
        v.in = in;
        v.in_offsets = in_off;
        v.out = out;
        v.out_capacity = out_capacity;
        v.out_offsets = out_off;
        if (!status) {
            HB_CUDA_TRY(ctx->hp_status.reserve(n * sizeof(int32_t)));
            status = ctx->hp_status.as<int32_t>();
        }
        v.status = status;
        return decode_batch_fast(ctx, ctx->scratch, v, total_in, st, true);
    }
    // one long literal (chunked stream kernels) or a table too large for shared memory: framing passes around the codec
    const unsigned flat = (unsigned)((n + 255) / 256), warps = (unsigned)((n + 255) / 256);  // (move kernels: 32 items per warp)
    const uint32_t min_len = std::max<uint32_t>(1, ctx->tables.min_len);
    const uint64_t dec_cap = total_in * 8 / min_len + 64;
    HB_CUDA_TRY(ctx->hp_huff.reserve(n));
    HB_CUDA_TRY(ctx->hp_prefix.reserve(n));
    HB_CUDA_TRY(ctx->hp_lens.reserve(n * sizeof(uint64_t)));
    HB_CUDA_TRY(ctx->hp_pay_lens.reserve(n * sizeof(uint64_t)));
    HB_CUDA_TRY(ctx->hp_pay.reserve(total_in + 64));
    HB_CUDA_TRY(ctx->hp_pay_off.reserve((n + 1) * sizeof(uint64_t)));
    HB_CUDA_TRY(ctx->hp_dec.reserve(dec_cap + 64));
    HB_CUDA_TRY(ctx->hp_dec_off.reserve((n + 1) * sizeof(uint64_t)));
    HB_CUDA_TRY(ctx->hp_dec_status.reserve(n * sizeof(int32_t)));
    HB_CUDA_TRY(ctx->hp_left_bits.reserve(n * sizeof(uint64_t)));
    HB_CUDA_TRY(ctx->hp_left_num.reserve(n));
    if (!status) {
        HB_CUDA_TRY(ctx->hp_status.reserve(n * sizeof(int32_t)));
        status = ctx->hp_status.as<int32_t>();
    }
    uint8_t *huff = ctx->hp_huff.as<uint8_t>(), *prefix = ctx->hp_prefix.as<uint8_t>();
    uint64_t *pay_lens = ctx->hp_pay_lens.as<uint64_t>(), *lens = ctx->hp_lens.as<uint64_t>();
    uint64_t *pay_off = ctx->hp_pay_off.as<uint64_t>(), *dec_off = ctx->hp_dec_off.as<uint64_t>();

    // 1. parse the literals; 2. the Huffman payloads, packed
    hb::hpack_parse_kernel<<<flat, 256, 0, st>>>(n, in, in_off, huff, prefix, pay_lens, lens, status);
    ++ctx->launches;
    if (launch_scan(ctx, ctx->scratch, lens, pay_off, n, st)) return AWS_OP_ERR;
    hb::hpack_move_kernel<<<warps, 256, 0, st>>>(
        n, in, in_off, prefix, pay_lens, nullptr, nullptr, huff, false, true, status, ctx->hp_pay.as<uint8_t>(), total_in, pay_off);
    ++ctx->launches;
    // 3. decode them (the single-stream kernels size themselves from the payload length: needed on the host for n == 1)
    uint64_t pay_total = total_in;
    if (n == 1) {
        if (!ctx->h_scalar) HB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_scalar), sizeof(uint64_t), cudaHostAllocDefault));
        HB_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalar, pay_off + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        HB_CUDA_TRY(cudaStreamSynchronize(st));
        pay_total = *ctx->h_scalar;
    }
    hb::BatchView v{};
    v.n = n;
    v.in = ctx->hp_pay.as<uint8_t>();
    v.in_offsets = pay_off;
    v.out = ctx->hp_dec.as<uint8_t>();
    v.out_capacity = dec_cap;
    v.out_offsets = dec_off;
    v.status = ctx->hp_dec_status.as<int32_t>();
    v.leftover_working_bits = ctx->hp_left_bits.as<uint64_t>();
    v.leftover_num_bits = ctx->hp_left_num.as<uint8_t>();
    if (decode_on_device(ctx, ctx->scratch, v, pay_total, st)) return AWS_OP_ERR;
    // 4. padding rule, final lengths; 5. the strings, packed
    hb::hpack_finish_kernel<<<flat, 256, 0, st>>>(
        n, huff, pay_lens, dec_off, v.status, v.leftover_working_bits, v.leftover_num_bits, status, lens);
    ++ctx->launches;
    if (launch_scan(ctx, ctx->scratch, lens, out_off, n, st)) return AWS_OP_ERR;
    hb::hpack_move_kernel<<<warps, 256, 0, st>>>(
        n, in, in_off, prefix, pay_lens, ctx->hp_dec.as<uint8_t>(), dec_off, huff, true, false, status, out, out_capacity, out_off);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

// Host pointers: stage in, run, stage out.
int hpack_run_host(
    aws_huffman_batch_ctx *ctx, bool encode, size_t n, const uint8_t *in, const uint64_t *in_off, uint32_t mode, uint8_t *out,
    uint64_t out_capacity, uint64_t *out_off, int32_t *status) {
    if (!ctx || (n && (!in_off || !out_off))) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (n == 0) {
        if (out_off) out_off[0] = 0;
        return AWS_OP_SUCCESS;
    }
    const uint64_t total_in = in_off[n];
    if ((total_in && !in) || (out_capacity && !out)) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    HB_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    HB_CUDA_TRY(ctx->s_in.reserve(total_in + 16));
    HB_CUDA_TRY(ctx->s_in_off.reserve((n + 1) * sizeof(uint64_t)));
    HB_CUDA_TRY(ctx->s_out.reserve(out_capacity + 16));
    HB_CUDA_TRY(ctx->s_out_off.reserve((n + 1) * sizeof(uint64_t)));
    HB_CUDA_TRY(ctx->s_status.reserve(n * sizeof(int32_t)));
    if (total_in) HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_in.ptr, in, total_in, cudaMemcpyHostToDevice, st));
    HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_in_off.ptr, in_off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const int rc = encode ? hpack_encode_on_device(ctx, n, ctx->s_in.as<uint8_t>(), ctx->s_in_off.as<uint64_t>(), total_in, mode,
                                                   ctx->s_out.as<uint8_t>(), out_capacity, ctx->s_out_off.as<uint64_t>(), st)
                          : hpack_decode_on_device(ctx, n, ctx->s_in.as<uint8_t>(), ctx->s_in_off.as<uint64_t>(), total_in,
                                                   ctx->s_out.as<uint8_t>(), out_capacity, ctx->s_out_off.as<uint64_t>(),
                                                   ctx->s_status.as<int32_t>(), st);
    if (rc) return AWS_OP_ERR;
    HB_CUDA_TRY(cudaMemcpyAsync(out_off, ctx->s_out_off.ptr, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (!encode && status) HB_CUDA_TRY(cudaMemcpyAsync(status, ctx->s_status.ptr, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    const uint64_t total_out = out_off[n];
    if (total_out > out_capacity) return aws_raise_error(AWS_ERROR_SHORT_BUFFER);
    if (total_out) {
        HB_CUDA_TRY(cudaMemcpyAsync(out, ctx->s_out.ptr, total_out, cudaMemcpyDeviceToHost, st));
        HB_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return AWS_OP_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// Byte histogram (the counts a code table is built from). HBM-bound by construction: 16 bytes per
// lane and load, and a block-wide histogram with ONE COLUMN PER LANE (hist[value][lane]): the 32 lanes of a
// shared-memory atomic never meet in an address and every lane stays in its own bank, however skewed the
// data is (the benchmark's top symbol is 40 % of the bytes: per-warp histograms would serialise 13-fold).
// ---------------------------------------------------------------------------------------------
constexpr int kHistThreads = 512;

__global__ void __launch_bounds__(kHistThreads) histogram_kernel(const uint8_t *in, uint64_t size, unsigned long long *counts) {
    __shared__ uint32_t s_hist[256 * 32];
    for (uint32_t i = threadIdx.x; i < 256 * 32; i += kHistThreads) s_hist[i] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t *const mine = s_hist + lane;
    const uintptr_t base = reinterpret_cast<uintptr_t>(in);
    const uint64_t head = min((unsigned long long)size, (unsigned long long)((16 - (base & 15)) & 15));  // bytes before the first aligned 16
    const uint64_t nvec = (size - head) >> 4;
    const uint4 *v = reinterpret_cast<const uint4 *>(in + head);
    auto add_word = [&](uint32_t w) {
        atomicAdd(mine + ((w & 0xffu) << 5), 1u);
        atomicAdd(mine + (((w >> 8) & 0xffu) << 5), 1u);
        atomicAdd(mine + (((w >> 16) & 0xffu) << 5), 1u);
        atomicAdd(mine + ((w >> 24) << 5), 1u);
    };
    for (uint64_t i = (uint64_t)blockIdx.x * kHistThreads + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * kHistThreads) {
        const uint4 q = __ldg(v + i);
        add_word(q.x);
        add_word(q.y);
        add_word(q.z);
        add_word(q.w);
    }
    if (blockIdx.x == 0) {  // the ragged ends
        const uint64_t tail0 = head + 16 * nvec;
        for (uint64_t i = threadIdx.x; i < head; i += kHistThreads) atomicAdd(mine + ((uint32_t)in[i] << 5), 1u);
        for (uint64_t i = tail0 + threadIdx.x; i < size; i += kHistThreads) atomicAdd(mine + ((uint32_t)in[i] << 5), 1u);
    }
    __syncthreads();
    // 256 values x 32 columns -> one count per value (a warp per value, 16 values per warp)
    for (uint32_t value = threadIdx.x >> 5; value < 256; value += kHistThreads / 32) {
        uint32_t c = s_hist[value * 32 + lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0 && c) atomicAdd(counts + value, (unsigned long long)c);
    }
}

int histogram_on_device(const uint8_t *in, uint64_t size, uint64_t *counts, cudaStream_t st) {
    HB_CUDA_TRY(cudaMemsetAsync(counts, 0, 256 * sizeof(uint64_t), st));
    if (size == 0) return AWS_OP_SUCCESS;
    int device = 0, sms = 148;
    HB_CUDA_TRY(cudaGetDevice(&device));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    // (a block's 32-bit columns hold its share of the bytes: size / blocks / 32 per column at most)
    const uint64_t want = (size / 16 + kHistThreads - 1) / kHistThreads;
    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)sms * 4));
    histogram_kernel<<<blocks, kHistThreads, 0, st>>>(in, size, reinterpret_cast<unsigned long long *>(counts));
    HB_CUDA_TRY(cudaGetLastError());
    return AWS_OP_SUCCESS;
}

__global__ void encoded_length_kernel(hb::DeviceTables t, const uint8_t *in, const uint64_t *in_offsets, uint64_t n, uint64_t *lens) {
    __shared__ uint32_t s_len[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_len[i] = t.enc[i].y;
    __syncthreads();
    const uint64_t item = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= n) return;
    const uint64_t a = in_offsets[item], e = in_offsets[item + 1];
    uint64_t bits = 0;
    for (uint64_t k = a + hb::lane_id(); k < e; k += 32) bits += s_len[in[k]];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, d);
    if (hb::lane_id() == 0) lens[item] = (bits + 7) >> 3;
}

}  // namespace

extern "C" {

static int hb_ctx_from_codes(
    struct aws_huffman_batch_ctx **out_ctx,
    const struct aws_huffman_code *codes,
    struct aws_huffman_symbol_coder *coder, /* optional: its decode callback is cross-checked */
    uint8_t eos_padding,
    int device_id);

int aws_huffman_batch_ctx_new(
    struct aws_huffman_batch_ctx **out_ctx,
    struct aws_huffman_symbol_coder *coder,
    uint8_t eos_padding,
    int device_id) {

    if (!out_ctx || !coder || !coder->encode) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    *out_ctx = nullptr;
    // Materialise the coder once over all 256 symbols (host callbacks never run again).
    struct aws_huffman_code codes[256];
    for (int s = 0; s < 256; ++s) codes[s] = coder->encode((uint8_t)s, coder->userdata);
    return hb_ctx_from_codes(out_ctx, codes, coder, eos_padding, device_id);
}

int aws_huffman_batch_ctx_new_from_code_table(
    struct aws_huffman_batch_ctx **out_ctx,
    const struct aws_huffman_code *code_table,
    uint8_t eos_padding,
    int device_id) {

    if (!out_ctx || !code_table) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    *out_ctx = nullptr;
    return hb_ctx_from_codes(out_ctx, code_table, nullptr, eos_padding, device_id);
}

static int hb_ctx_from_codes(
    struct aws_huffman_batch_ctx **out_ctx,
    const struct aws_huffman_code *codes,
    struct aws_huffman_symbol_coder *coder,
    uint8_t eos_padding,
    int device_id) {

    uint32_t patterns[256];
    uint8_t num_bits[256];
    uint2 enc[256];
    for (int s = 0; s < 256; ++s) {
        const struct aws_huffman_code c = codes[s];
        if (c.num_bits > 32) return aws_raise_error(AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE);
        const uint32_t mask = c.num_bits >= 32 ? 0xffffffffu : ((1u << c.num_bits) - 1u);
        patterns[s] = c.pattern & mask;
        num_bits[s] = c.num_bits;
        enc[s] = make_uint2(patterns[s], c.num_bits);
    }
    struct huffman_lut lut;
    const int lut_rc = huffman_lut_build(&lut, patterns, num_bits, kLutRootBits, kLutSubBits);
    if (lut_rc == HUFFMAN_LUT_ERR_OOM) return aws_raise_error(AWS_ERROR_OOM);
    if (lut_rc != HUFFMAN_LUT_OK) return aws_raise_error(AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE);

    // Optional cross-check of the coder's own decode callback on every code (what the reference's
    // huffman_symbol_decoder test does, tests/huffman_test.c:199-220).
    if (coder && coder->decode) {
        for (int s = 0; s < 256; ++s) {
            if (!num_bits[s]) continue;
            uint8_t sym = 0;
            const uint8_t used = coder->decode(patterns[s] << (32 - num_bits[s]), &sym, coder->userdata);
            if (used != num_bits[s] || sym != s) {
                huffman_lut_clean_up(&lut);
                return aws_raise_error(AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE);
            }
        }
    }

    aws_huffman_batch_ctx *ctx = new (std::nothrow) aws_huffman_batch_ctx();
    if (!ctx) {
        huffman_lut_clean_up(&lut);
        return aws_raise_error(AWS_ERROR_OOM);
    }
    ctx->device = device_id;

    auto fail = [&](cudaError_t err, const char *what, int line) {
        hb_note_cuda_error(err, what, line);
        huffman_lut_clean_up(&lut);
        aws_huffman_batch_ctx_destroy(ctx);
        return aws_raise_error(AWS_ERROR_COMPRESSION_DEVICE_FAILURE);
    };
#define HB_CTX_TRY(expr)                                                                                               \
    do {                                                                                                               \
        cudaError_t e_ = (expr);                                                                                       \
        if (e_ != cudaSuccess) return fail(e_, #expr, __LINE__);                                                       \
    } while (0)

    int count = 0;
    HB_CTX_TRY(cudaGetDeviceCount(&count));
    if (device_id < 0 || device_id >= count) return fail(cudaErrorInvalidDevice, "device_id", __LINE__);
    HB_CTX_TRY(cudaSetDevice(device_id));
    HB_CTX_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    HB_CTX_TRY(cudaEventCreateWithFlags(&ctx->scratch_free, cudaEventDisableTiming));
    ctx->force_generic = getenv("AWS_HUFFMAN_BATCH_FORCE_GENERIC") != nullptr;
    ctx->no_slots = getenv("AWS_HUFFMAN_BATCH_NO_SLOTS") != nullptr;
    ctx->no_strings = getenv("AWS_HUFFMAN_BATCH_NO_STRINGS") != nullptr;
    ctx->no_fused_stream = getenv("AWS_HUFFMAN_BATCH_NO_FUSED_STREAM") != nullptr;
    ctx->no_slots_decode = getenv("AWS_HUFFMAN_BATCH_SLOTS_DECODE") == nullptr;  // (opt-in: measured slower, decode_slots.cuh)
    ctx->no_rows = getenv("AWS_HUFFMAN_BATCH_ROWS") == nullptr;  // (the two-kernel decoder is opt-in: measured slower, decode_rows.cuh)
    HB_CTX_TRY(cudaMalloc(&ctx->d_enc, sizeof(enc)));
    HB_CTX_TRY(cudaMalloc(&ctx->d_lut, (size_t)lut.count * sizeof(uint32_t)));
    HB_CTX_TRY(cudaMemcpy(ctx->d_enc, enc, sizeof(enc), cudaMemcpyHostToDevice));
    {
        // re-encode for the device (format: device_common.cuh); root entries get a second symbol when
        // two complete codes fit in the root index
        std::vector<uint32_t> dev(lut.count);
        auto single = [](uint32_t len, uint32_t sym) { return (len << 24) | (sym << 8) | (len << 2) | 1u; };
        for (uint32_t i = 0; i < lut.count; ++i) {
            const uint32_t e = lut.entries[i];
            if (e == 0) dev[i] = 0;
            else if (HUFFMAN_LUT_IS_LEAF(e)) dev[i] = single(HUFFMAN_LUT_LEAF_LEN(e), HUFFMAN_LUT_LEAF_SYMBOL(e));
            else dev[i] = (HUFFMAN_LUT_LINK_WIDTH(e) << 20) | HUFFMAN_LUT_LINK_BASE(e);
        }
        if (lut.count >= (1u << 20)) {
            huffman_lut_clean_up(&lut);
            aws_huffman_batch_ctx_destroy(ctx);
            return aws_raise_error(AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE);
        }
        for (uint32_t i = 0; i < (1u << lut.root_bits); ++i) {
            const uint32_t e = lut.entries[i];
            if (!HUFFMAN_LUT_IS_LEAF(e)) continue;
            const uint32_t len1 = HUFFMAN_LUT_LEAF_LEN(e);
            if (len1 >= lut.root_bits) continue;
            const uint32_t rest = lut.root_bits - len1;  // index bits left after the first code
            const uint32_t window2 = (i << (32 - lut.root_bits)) << len1;
            uint8_t sym2 = 0;
            const uint8_t len2 = huffman_lut_decode(&lut, window2, &sym2);
            if (len2 != 0 && len2 <= rest)
                dev[i] = ((len1 + len2) << 24) | ((uint32_t)sym2 << 16) | ((uint32_t)HUFFMAN_LUT_LEAF_SYMBOL(e) << 8) |
                         (len1 << 2) | 2u;
        }
        HB_CTX_TRY(cudaMemcpy(ctx->d_lut, dev.data(), (size_t)lut.count * sizeof(uint32_t), cudaMemcpyHostToDevice));
        // LUT2 (decode_span_lean): the same tables as 64-bit entries {symbols, count | bits consumed, next table};
        // tables are 32-byte aligned (the step takes the table address and the index shift out of one word)
        std::vector<uint2> lut2;
        const uint32_t root_y = 32u - lut.root_bits;  // next step: root table (offset 0)
        lut2.resize((size_t)1 << lut.root_bits);
        while (lut2.size() & 3) lut2.push_back(uint2{0, 0});
        const uint32_t trap_base = (uint32_t)lut2.size();
        const uint32_t trap_y = (trap_base * 8u) | 31u;  // index width 1, nothing consumed: a trapped lane stays
        for (int i = 0; i < 4; ++i) lut2.push_back(uint2{0, trap_y});
        struct Conv {
            std::vector<uint2> &out;
            const std::vector<uint32_t> &dev;
            uint32_t root_y, trap_y;
            void table(uint32_t new_base, uint32_t old_base, uint32_t width, uint32_t used) {
                for (uint32_t idx = 0; idx < (1u << width); ++idx) {
                    const uint32_t e = dev[old_base + idx];
                    uint2 o;
                    if (e == 0) {
                        o = uint2{0, trap_y};
                    } else if (e >= 0x01000000u) {
                        const uint32_t total = e >> 24, len1 = (e >> 2) & 63u, cnt = e & 3u;
                        o.x = ((e >> 8) & 0xffu) | (((e >> 16) & 0xffu) << 8) | (len1 << 16) | (cnt << 30);
                        o.y = root_y | ((total - used) << 24);
                    } else {
                        const uint32_t w2 = (e >> 20) & 0xfu, b2 = e & 0xFFFFFu;
                        size_t child = (out.size() + 3) & ~size_t(3);
                        out.resize(child + std::max<size_t>(4, (size_t)1 << w2), uint2{0, trap_y});
                        table((uint32_t)child, b2, w2, used + width);
                        o.x = 0;
                        o.y = ((uint32_t)child * 8u) | (32u - w2) | (width << 24);
                    }
                    out[new_base + idx] = o;
                }
            }
        } conv{lut2, dev, root_y, trap_y};
        conv.table(0, 0, lut.root_bits, 0);
        while (lut2.size() & 3) lut2.push_back(uint2{0, trap_y});
        ctx->lut2_count = (uint32_t)lut2.size();
        ctx->lut2_trap = trap_base;
        HB_CTX_TRY(cudaMalloc(&ctx->d_lut2, lut2.size() * sizeof(uint2)));
        HB_CTX_TRY(cudaMemcpy(ctx->d_lut2, lut2.data(), lut2.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    }
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_id) == cudaSuccess && sms > 0)
            ctx->sm_count = sms;
    }
#undef HB_CTX_TRY

    ctx->tables.enc = ctx->d_enc;
    ctx->tables.lut = ctx->d_lut;
    ctx->tables.lut_count = lut.count;
    ctx->tables.lut_root_bits = lut.root_bits;
    ctx->tables.min_len = lut.min_len;
    ctx->tables.max_len = lut.max_len;
    ctx->tables.has_unknown = lut.has_unknown_symbols;
    ctx->tables.eos_padding = eos_padding;
    ctx->lut_smem_entries = std::min<uint32_t>(lut.count, kLutMaxSmemEntries);
    if (ctx->lut_smem_entries < (1u << lut.root_bits)) ctx->lut_smem_entries = 1u << lut.root_bits;
    huffman_lut_clean_up(&lut);

    *out_ctx = ctx;
    return AWS_OP_SUCCESS;
}

void aws_huffman_batch_ctx_destroy(struct aws_huffman_batch_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
        if (ctx->scratch_free) cudaEventDestroy(ctx->scratch_free);
    }
    if (ctx->d_enc) cudaFree(ctx->d_enc);
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    if (ctx->d_lut2) cudaFree(ctx->d_lut2);
    GrowBuf *bufs[] = {&ctx->s_in,       &ctx->s_in_off,      &ctx->s_out,      &ctx->s_out_off,
                       &ctx->s_caps,     &ctx->s_status,      &ctx->s_consumed, &ctx->s_ovf_pattern,
                       &ctx->s_ovf_bits, &ctx->s_left_bits,   &ctx->s_left_num, &ctx->s_chain, &ctx->hp_pay,
                       &ctx->hp_pay_off, &ctx->hp_dec,        &ctx->hp_dec_off, &ctx->hp_huff,
                       &ctx->hp_prefix,  &ctx->hp_lens,       &ctx->hp_pay_lens, &ctx->hp_dec_status,
                       &ctx->hp_left_bits, &ctx->hp_left_num, &ctx->hp_status};
    for (GrowBuf *g : bufs) g->release();
    if (ctx->h_scalar) cudaFreeHost(ctx->h_scalar);
    ctx->scratch.release();
    for (Lane &lane : ctx->lanes) lane.release();
    (void)cudaGetLastError();
    delete ctx;
}

int aws_huffman_encode_batch(struct aws_huffman_batch_ctx *ctx, const struct aws_huffman_batch *batch) {
    return run_host_batch(ctx, batch, true);
}

int aws_huffman_decode_batch(struct aws_huffman_batch_ctx *ctx, const struct aws_huffman_batch *batch) {
    return run_host_batch(ctx, batch, false);
}

int aws_huffman_encode_batch_resume(struct aws_huffman_batch_ctx *ctx, const struct aws_huffman_batch *batch) {
    return run_host_batch(ctx, batch, true, true);
}

int aws_huffman_decode_batch_resume(struct aws_huffman_batch_ctx *ctx, const struct aws_huffman_batch *batch) {
    return run_host_batch(ctx, batch, false, true);
}

int aws_huffman_encode_batch_resume_device(
    struct aws_huffman_batch_ctx *ctx,
    const struct aws_huffman_batch *batch,
    void *cuda_stream) {
    if (!ctx) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (check_batch(batch) || check_resume(batch, true)) return AWS_OP_ERR;
    cudaStream_t st;
    if (device_call_begin(ctx, cuda_stream, &st)) return AWS_OP_ERR;
    hb::BatchView v = make_view(batch);
    v.resume = true;
    return device_call_end(ctx, st, encode_on_device(ctx, ctx->scratch, v, batch->in_size, st));
}

int aws_huffman_decode_batch_resume_device(
    struct aws_huffman_batch_ctx *ctx,
    const struct aws_huffman_batch *batch,
    void *cuda_stream) {
    if (!ctx) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (check_batch(batch) || check_resume(batch, false)) return AWS_OP_ERR;
    cudaStream_t st;
    if (device_call_begin(ctx, cuda_stream, &st)) return AWS_OP_ERR;
    hb::BatchView v = make_view(batch);
    v.resume = true;
    return device_call_end(ctx, st, decode_on_device(ctx, ctx->scratch, v, batch->in_size, st));
}

int aws_huffman_encode_batch_device(
    struct aws_huffman_batch_ctx *ctx,
    const struct aws_huffman_batch *batch,
    void *cuda_stream) {
    if (!ctx) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (check_batch(batch)) return AWS_OP_ERR;
    cudaStream_t st;
    if (device_call_begin(ctx, cuda_stream, &st)) return AWS_OP_ERR;
    return device_call_end(ctx, st, encode_on_device(ctx, ctx->scratch, make_view(batch), batch->in_size, st));
}

int aws_huffman_decode_batch_device(
    struct aws_huffman_batch_ctx *ctx,
    const struct aws_huffman_batch *batch,
    void *cuda_stream) {
    if (!ctx) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (check_batch(batch)) return AWS_OP_ERR;
    cudaStream_t st;
    if (device_call_begin(ctx, cuda_stream, &st)) return AWS_OP_ERR;
    return device_call_end(ctx, st, decode_on_device(ctx, ctx->scratch, make_view(batch), batch->in_size, st));
}

int aws_huffman_histogram_device(const uint8_t *in, uint64_t size, uint64_t *counts, void *cuda_stream) {
    if (!counts || (size && !in)) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    return histogram_on_device(in, size, counts, static_cast<cudaStream_t>(cuda_stream));
}

int aws_huffman_histogram(int device_id, const uint8_t *in, uint64_t size, uint64_t counts[256]) {
    if (!counts || (size && !in)) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    HB_CUDA_TRY(cudaSetDevice(device_id));
    uint8_t *d_in = nullptr;
    uint64_t *d_counts = nullptr;
    cudaStream_t st = nullptr;
    int rc = AWS_OP_ERR;
    do {
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) break;
        if (cudaMalloc(&d_counts, 256 * sizeof(uint64_t)) != cudaSuccess) break;
        if (size && cudaMalloc(&d_in, size) != cudaSuccess) break;
        if (size && cudaMemcpyAsync(d_in, in, size, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        if (histogram_on_device(d_in, size, d_counts, st)) break;
        if (cudaMemcpyAsync(counts, d_counts, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        rc = AWS_OP_SUCCESS;
    } while (false);
    if (d_in) cudaFree(d_in);
    if (d_counts) cudaFree(d_counts);
    if (st) cudaStreamDestroy(st);
    if (rc) {
        (void)cudaGetLastError();
        return aws_raise_error(AWS_ERROR_COMPRESSION_DEVICE_FAILURE);
    }
    return AWS_OP_SUCCESS;
}

int aws_hpack_string_encode_batch(
    struct aws_huffman_batch_ctx *ctx, size_t n, const uint8_t *in, const uint64_t *in_offsets,
    enum aws_hpack_huffman_mode mode, uint8_t *out, uint64_t out_capacity, uint64_t *out_offsets) {
    if ((unsigned)mode > (unsigned)AWS_HPACK_HUFFMAN_ALWAYS) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    return hpack_run_host(ctx, true, n, in, in_offsets, (uint32_t)mode, out, out_capacity, out_offsets, nullptr);
}

int aws_hpack_string_decode_batch(
    struct aws_huffman_batch_ctx *ctx, size_t n, const uint8_t *in, const uint64_t *in_offsets, uint8_t *out,
    uint64_t out_capacity, uint64_t *out_offsets, int32_t *status) {
    return hpack_run_host(ctx, false, n, in, in_offsets, 0, out, out_capacity, out_offsets, status);
}

int aws_hpack_string_encode_batch_device(
    struct aws_huffman_batch_ctx *ctx, size_t n, const uint8_t *in, const uint64_t *in_offsets, uint64_t in_size,
    enum aws_hpack_huffman_mode mode, uint8_t *out, uint64_t out_capacity, uint64_t *out_offsets, void *cuda_stream) {
    if (!ctx || (n && (!in_offsets || !out_offsets)) || (unsigned)mode > (unsigned)AWS_HPACK_HUFFMAN_ALWAYS)
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    cudaStream_t st;
    if (device_call_begin(ctx, cuda_stream, &st)) return AWS_OP_ERR;
    return device_call_end(
        ctx, st, hpack_encode_on_device(ctx, n, in, in_offsets, in_size, (uint32_t)mode, out, out_capacity, out_offsets, st));
}

int aws_hpack_string_decode_batch_device(
    struct aws_huffman_batch_ctx *ctx, size_t n, const uint8_t *in, const uint64_t *in_offsets, uint64_t in_size,
    uint8_t *out, uint64_t out_capacity, uint64_t *out_offsets, int32_t *status, void *cuda_stream) {
    if (!ctx || (n && (!in_offsets || !out_offsets))) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    cudaStream_t st;
    if (device_call_begin(ctx, cuda_stream, &st)) return AWS_OP_ERR;
    return device_call_end(
        ctx, st, hpack_decode_on_device(ctx, n, in, in_offsets, in_size, out, out_capacity, out_offsets, status, st));
}

int aws_huffman_get_encoded_length_batch(
    struct aws_huffman_batch_ctx *ctx,
    const uint8_t *in,
    const uint64_t *in_offsets,
    size_t n,
    uint64_t *lens) {
    if (!ctx || (n && (!in_offsets || !lens))) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (n == 0) return AWS_OP_SUCCESS;
    HB_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t total_in = in_offsets[n];
    if (total_in && !in) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    HB_CUDA_TRY(ctx->s_in.reserve(total_in + 16));
    HB_CUDA_TRY(ctx->s_in_off.reserve((n + 1) * sizeof(uint64_t)));
    HB_CUDA_TRY(ctx->scratch.lens.reserve(n * sizeof(uint64_t)));
    if (total_in) HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_in.ptr, in, total_in, cudaMemcpyHostToDevice, st));
    HB_CUDA_TRY(cudaMemcpyAsync(ctx->s_in_off.ptr, in_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)((n + kWarpsPerBlock - 1) / kWarpsPerBlock);
    encoded_length_kernel<<<blocks, kWarpsPerBlock * 32, 0, st>>>(
        ctx->tables, ctx->s_in.as<uint8_t>(), ctx->s_in_off.as<uint64_t>(), n, ctx->scratch.lens.as<uint64_t>());
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    HB_CUDA_TRY(cudaMemcpyAsync(lens, ctx->scratch.lens.ptr, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    return AWS_OP_SUCCESS;
}

int aws_huffman_batch_ctx_synchronize(struct aws_huffman_batch_ctx *ctx) {
    if (!ctx) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    HB_CUDA_TRY(cudaSetDevice(ctx->device));
    HB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return AWS_OP_SUCCESS;
}

void *aws_huffman_batch_ctx_stream(struct aws_huffman_batch_ctx *ctx) {
    return ctx ? ctx->stream : nullptr;
}

int aws_huffman_batch_ctx_device(struct aws_huffman_batch_ctx *ctx) {
    return ctx ? ctx->device : -1;
}

uint64_t aws_huffman_batch_ctx_launch_count(struct aws_huffman_batch_ctx *ctx) {
    return ctx ? ctx->launches : 0;
}

int aws_huffman_batch_plan_shards(const uint64_t *in_offsets, size_t n, size_t num_shards, size_t *shard_begin) {
    if (!shard_begin || num_shards == 0 || (n && !in_offsets)) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    const uint64_t base = n ? in_offsets[0] : 0;
    const uint64_t total = n ? in_offsets[n] - base : 0;
    shard_begin[0] = 0;
    for (size_t s = 1; s < num_shards; ++s) {
        // first item whose start offset reaches the s-th equal share of the bytes
        const uint64_t target = base + (uint64_t)(((unsigned __int128)total * s) / num_shards);
        const uint64_t *it = std::lower_bound(in_offsets, in_offsets + n, target);
        size_t idx = (size_t)(it - in_offsets);
        if (total == 0) idx = (size_t)(((unsigned __int128)n * s) / num_shards);  // all items empty: split by count
        shard_begin[s] = std::max(idx, shard_begin[s - 1]);
    }
    shard_begin[num_shards] = n;
    return AWS_OP_SUCCESS;
}

int aws_huffman_batch_concat_offsets(
    const uint64_t *const *shard_offsets,
    const size_t *shard_items,
    size_t num_shards,
    uint64_t *global_offsets) {
    if (!global_offsets || (num_shards && (!shard_offsets || !shard_items)))
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    uint64_t base = 0;
    size_t at = 0;
    for (size_t s = 0; s < num_shards; ++s) {
        const uint64_t *local = shard_offsets[s];
        for (size_t i = 0; i < shard_items[s]; ++i) global_offsets[at++] = base + local[i];
        if (shard_items[s]) base += local[shard_items[s]];
    }
    global_offsets[at] = base;
    return AWS_OP_SUCCESS;
}

// One batch over several contexts (one per GPU of the box, or several on one GPU): plan by bytes, one host thread
// per context drives its shard through that context's own (pipelined) host path into a private buffer, and the
// host concatenates payloads and offsets. No collective: the shards never talk to each other.
static int hb_run_multi(struct aws_huffman_batch_ctx *const *ctxs, size_t n_ctx, const struct aws_huffman_batch *b, bool encode) {
    if (!ctxs || n_ctx == 0) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    for (size_t d = 0; d < n_ctx; ++d)
        if (!ctxs[d]) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (check_batch(b)) return AWS_OP_ERR;
    if (b->out_caps) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);  // packed layout only
    const size_t n = b->n;
    if (n == 0) {
        if (b->out_offsets) b->out_offsets[0] = 0;
        return AWS_OP_SUCCESS;
    }
    if (n_ctx == 1) return encode ? aws_huffman_encode_batch(ctxs[0], b) : aws_huffman_decode_batch(ctxs[0], b);
    std::vector<size_t> begin(n_ctx + 1);
    if (aws_huffman_batch_plan_shards(b->in_offsets, n, n_ctx, begin.data())) return AWS_OP_ERR;
    struct Shard {
        std::vector<uint64_t> in_off, out_off;
        std::vector<uint8_t> out;
        int rc = AWS_OP_SUCCESS, err = 0;
    };
    std::vector<Shard> shards(n_ctx);
    std::vector<std::thread> threads;
    for (size_t d = 0; d < n_ctx; ++d) {
        threads.emplace_back([&, d]() {
            Shard &sh = shards[d];
            const size_t a = begin[d], nd = begin[d + 1] - a;
            if (nd == 0) return;
            const uint64_t in0 = b->in_offsets[a], bytes_in = b->in_offsets[a + nd] - in0;
            const uint32_t max_len = std::max<uint32_t>(1, ctxs[d]->tables.max_len), min_len = std::max<uint32_t>(1, ctxs[d]->tables.min_len);
            const uint64_t room = (encode ? bytes_in * ((max_len + 7) / 8) : (bytes_in * 8) / min_len) + 64;
            sh.in_off.resize(nd + 1);
            for (size_t i = 0; i <= nd; ++i) sh.in_off[i] = b->in_offsets[a + i] - in0;
            sh.out_off.assign(nd + 1, 0);
            sh.out.resize(room);
            aws_huffman_batch sb = *b;
            sb.n = nd;
            sb.in = b->in + in0;
            sb.in_offsets = sh.in_off.data();
            sb.in_size = bytes_in;
            sb.out = sh.out.data();
            sb.out_capacity = room;
            sb.out_offsets = sh.out_off.data();
            if (b->out_lens) sb.out_lens = b->out_lens + a;
            if (b->status) sb.status = b->status + a;
            if (b->consumed) sb.consumed = b->consumed + a;
            if (b->overflow_pattern) sb.overflow_pattern = b->overflow_pattern + a;
            if (b->overflow_num_bits) sb.overflow_num_bits = b->overflow_num_bits + a;
            if (b->leftover_working_bits) sb.leftover_working_bits = b->leftover_working_bits + a;
            if (b->leftover_num_bits) sb.leftover_num_bits = b->leftover_num_bits + a;
            sh.rc = encode ? aws_huffman_encode_batch(ctxs[d], &sb) : aws_huffman_decode_batch(ctxs[d], &sb);
            if (sh.rc != AWS_OP_SUCCESS) sh.err = aws_last_error();  // (the error slot is per thread)
        });
    }
    for (std::thread &t : threads) t.join();
    for (const Shard &sh : shards)
        if (sh.rc != AWS_OP_SUCCESS) return aws_raise_error(sh.err ? sh.err : AWS_ERROR_COMPRESSION_DEVICE_FAILURE);
    // the concatenation step
    std::vector<const uint64_t *> ptrs(n_ctx);
    std::vector<size_t> items(n_ctx);
    for (size_t d = 0; d < n_ctx; ++d) {
        items[d] = begin[d + 1] - begin[d];
        static const uint64_t zero = 0;
        ptrs[d] = items[d] ? shards[d].out_off.data() : &zero;
    }
    if (aws_huffman_batch_concat_offsets(ptrs.data(), items.data(), n_ctx, b->out_offsets)) return AWS_OP_ERR;
    if (b->out_offsets[n] > b->out_capacity) return aws_raise_error(AWS_ERROR_SHORT_BUFFER);
    for (size_t d = 0; d < n_ctx; ++d)
        if (items[d]) memcpy(b->out + b->out_offsets[begin[d]], shards[d].out.data(), (size_t)shards[d].out_off[items[d]]);
    return AWS_OP_SUCCESS;
}

int aws_huffman_encode_batch_multi(
    struct aws_huffman_batch_ctx *const *ctxs, size_t n_ctx, const struct aws_huffman_batch *batch) {
    return hb_run_multi(ctxs, n_ctx, batch, true);
}

int aws_huffman_decode_batch_multi(
    struct aws_huffman_batch_ctx *const *ctxs, size_t n_ctx, const struct aws_huffman_batch *batch) {
    return hb_run_multi(ctxs, n_ctx, batch, false);
}

int aws_huffman_get_encoded_length_batch_device(
    struct aws_huffman_batch_ctx *ctx,
    const uint8_t *in,
    const uint64_t *in_offsets,
    size_t n,
    uint64_t *lens,
    void *cuda_stream) {
    if (!ctx || (n && (!in_offsets || !lens))) return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    if (n == 0) return AWS_OP_SUCCESS;
    cudaStream_t st;
    if (device_call_begin(ctx, cuda_stream, &st)) return AWS_OP_ERR;
    const unsigned blocks = (unsigned)((n + kWarpsPerBlock - 1) / kWarpsPerBlock);
    encoded_length_kernel<<<blocks, kWarpsPerBlock * 32, 0, st>>>(ctx->tables, in, in_offsets, n, lens);
    ++ctx->launches;
    HB_CUDA_TRY(cudaGetLastError());
    return device_call_end(ctx, st, AWS_OP_SUCCESS);
}

}  // extern "C"

#ifdef HB_PHASE_TIMING
// development builds only (tools/phase_probe.py): cycles per phase of decode_batch_kernel, summed over blocks
extern "C" int aws_huffman_batch_debug_phase_cycles(unsigned long long *out16, int reset) {
    if (cudaMemcpyFromSymbol(out16, hb::hb_phase_cycles, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long zero[16] = {0};
        if (cudaMemcpyToSymbol(hb::hb_phase_cycles, zero, sizeof(zero)) != cudaSuccess) return -1;
    }
    return 0;
}
extern "C" int aws_huffman_batch_debug_tile_times(unsigned long long *out /* 4 x 8192 */) {
    return cudaMemcpyFromSymbol(out, hb::hb_tile_times, sizeof(unsigned long long) * 4 * 8192) == cudaSuccess ? 0 : -1;
}
#endif
