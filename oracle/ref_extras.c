/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE. Linked only into oracle/_ref/libref_huffman.so.
 *
 * Batch drivers around the UNMODIFIED reference functions (compiled in place from
 * /root/reference/source/huffman.c): item i = aws_huffman_{en,de}coder_init + one
 * aws_huffman_{en,de}code call. Same array contract as oracle_encode_batch /
 * oracle_decode_batch (huffman_oracle.h), so tests and bench.py can swap one for the other.
 */
#include <aws/compression/huffman.h>

void ref_encode_batch(
    struct aws_huffman_symbol_coder *coder,
    uint8_t eos_padding,
    const uint8_t *in,
    const uint64_t *in_offsets,
    size_t n,
    uint8_t *out,
    uint64_t out_capacity,
    uint64_t *out_offsets,
    const uint64_t *out_caps,
    uint64_t *out_lens,
    int32_t *status,
    uint64_t *consumed,
    uint32_t *overflow_pattern,
    uint8_t *overflow_num_bits) {

    uint64_t cursor = 0;
    for (size_t i = 0; i < n; ++i) {
        struct aws_huffman_encoder encoder;
        aws_huffman_encoder_init(&encoder, coder);
        encoder.eos_padding = eos_padding;
        uint64_t base, cap;
        if (out_caps) {
            base = out_offsets[i];
            cap = out_caps[i];
        } else {
            base = cursor;
            cap = out_capacity > cursor ? out_capacity - cursor : 0;
            out_offsets[i] = cursor;
        }
        const size_t in_len = (size_t)(in_offsets[i + 1] - in_offsets[i]);
        struct aws_byte_cursor cur = {in_len, (uint8_t *)in + in_offsets[i]};
        struct aws_byte_buf buf = {0, out + base, (size_t)cap, NULL};
        aws_reset_error();
        const int rc = aws_huffman_encode(&encoder, &cur, &buf);
        cursor += buf.len;
        if (out_lens) out_lens[i] = buf.len;
        if (status) status[i] = rc == AWS_OP_SUCCESS ? 0 : aws_last_error();
        if (consumed) consumed[i] = in_len - cur.len;
        if (overflow_pattern) overflow_pattern[i] = encoder.overflow_bits.num_bits ? encoder.overflow_bits.pattern : 0;
        if (overflow_num_bits) overflow_num_bits[i] = encoder.overflow_bits.num_bits;
    }
    if (!out_caps) {
        out_offsets[n] = cursor;
    }
}

void ref_decode_batch(
    struct aws_huffman_symbol_coder *coder,
    const uint8_t *in,
    const uint64_t *in_offsets,
    size_t n,
    uint8_t *out,
    uint64_t out_capacity,
    uint64_t *out_offsets,
    const uint64_t *out_caps,
    uint64_t *out_lens,
    int32_t *status,
    uint64_t *consumed,
    uint64_t *leftover_working_bits,
    uint8_t *leftover_num_bits) {

    uint64_t cursor = 0;
    for (size_t i = 0; i < n; ++i) {
        struct aws_huffman_decoder decoder;
        aws_huffman_decoder_init(&decoder, coder);
        uint64_t base, cap;
        if (out_caps) {
            base = out_offsets[i];
            cap = out_caps[i];
        } else {
            base = cursor;
            cap = out_capacity > cursor ? out_capacity - cursor : 0;
            out_offsets[i] = cursor;
        }
        const size_t in_len = (size_t)(in_offsets[i + 1] - in_offsets[i]);
        struct aws_byte_cursor cur = {in_len, (uint8_t *)in + in_offsets[i]};
        struct aws_byte_buf buf = {0, out + base, (size_t)cap, NULL};
        aws_reset_error();
        const int rc = aws_huffman_decode(&decoder, &cur, &buf);
        cursor += buf.len;
        if (out_lens) out_lens[i] = buf.len;
        if (status) status[i] = rc == AWS_OP_SUCCESS ? 0 : aws_last_error();
        if (consumed) consumed[i] = in_len - cur.len;
        if (leftover_working_bits) leftover_working_bits[i] = decoder.working_bits;
        if (leftover_num_bits) leftover_num_bits[i] = decoder.num_bits;
    }
    if (!out_caps) {
        out_offsets[n] = cursor;
    }
}

/* A coder with holes in its ENCODE table, for the unknown-symbol paths: wraps another coder and
 * reports num_bits == 0 for the symbols flagged in `unknown`. */
struct ref_masked_coder {
    struct aws_huffman_symbol_coder coder;
    struct aws_huffman_symbol_coder *inner;
    uint8_t unknown[256];
};

static struct aws_huffman_code s_masked_encode(uint8_t symbol, void *userdata) {
    struct ref_masked_coder *m = userdata;
    if (m->unknown[symbol]) {
        struct aws_huffman_code none = {0, 0};
        return none;
    }
    return m->inner->encode(symbol, m->inner->userdata);
}

static uint8_t s_masked_decode(uint32_t bits, uint8_t *symbol, void *userdata) {
    struct ref_masked_coder *m = userdata;
    uint8_t found = 0;
    const uint8_t used = m->inner->decode(bits, &found, m->inner->userdata);
    if (used == 0 || m->unknown[found]) {
        return 0;
    }
    *symbol = found;
    return used;
}

struct aws_huffman_symbol_coder *ref_masked_coder_new(struct aws_huffman_symbol_coder *inner, const uint8_t *unknown) {
    struct ref_masked_coder *m = calloc(1, sizeof(*m));
    m->inner = inner;
    memcpy(m->unknown, unknown, 256);
    m->coder.encode = s_masked_encode;
    m->coder.decode = s_masked_decode;
    m->coder.userdata = m;
    return &m->coder;
}

void ref_masked_coder_free(struct aws_huffman_symbol_coder *coder) {
    free(coder);
}
