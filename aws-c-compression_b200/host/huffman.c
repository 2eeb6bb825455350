/*
 * Host streaming Huffman codec — the drop-in for the reference's source/huffman.c, written from
 * scratch. Same exported functions, same struct state, same observable behaviour (bytes, cursor
 * and buffer movement, error codes, overflow_bits / working_bits resume state; the contract is
 * spelled out in SURVEY.md Appendix B and enforced by tests/test_host_codec.py against the
 * unmodified reference).
 *
 * Differences in HOW: the encoder assembles bits in a 64-bit accumulator and stores whole bytes
 * (big-endian words on the roomy path) instead of looping per output byte per code
 * (reference huffman.c:59-105); the decoder refills its register with one multi-byte load instead
 * of a byte loop (reference huffman.c:196-211).
 *
 * This file is the CPU API only. The batched GPU path is csrc/; it never calls into this file.
 */
#include <aws/compression/huffman.h>

enum { WINDOW_BITS = 32 };

void aws_huffman_encoder_init(struct aws_huffman_encoder *encoder, struct aws_huffman_symbol_coder *coder) {
    AWS_ASSERT(encoder);
    AWS_ASSERT(coder);
    AWS_ZERO_STRUCT(*encoder);
    encoder->coder = coder;
    encoder->eos_padding = UINT8_MAX;
}

void aws_huffman_encoder_reset(struct aws_huffman_encoder *encoder) {
    AWS_ASSERT(encoder);
    AWS_ZERO_STRUCT(encoder->overflow_bits);
}

void aws_huffman_decoder_init(struct aws_huffman_decoder *decoder, struct aws_huffman_symbol_coder *coder) {
    AWS_ASSERT(decoder);
    AWS_ASSERT(coder);
    AWS_ZERO_STRUCT(*decoder);
    decoder->coder = coder;
}

void aws_huffman_decoder_reset(struct aws_huffman_decoder *decoder) {
    decoder->working_bits = 0;
    decoder->num_bits = 0;
}

void aws_huffman_decoder_allow_growth(struct aws_huffman_decoder *decoder, bool allow_growth) {
    decoder->allow_growth = allow_growth;
}

size_t aws_huffman_get_encoded_length(struct aws_huffman_encoder *encoder, struct aws_byte_cursor to_encode) {
    AWS_PRECONDITION(encoder);
    AWS_PRECONDITION(aws_byte_cursor_is_valid(&to_encode));

    struct aws_huffman_symbol_coder *coder = encoder->coder;
    size_t total_bits = 0;
    for (size_t i = 0; i < to_encode.len; ++i) {
        total_bits += coder->encode(to_encode.ptr[i], coder->userdata).num_bits;
    }
    return (total_bits + 7) / 8;
}

/* ---- encode ---- */

static inline uint32_t s_low_bits(uint32_t value, unsigned count) {
    return count >= 32 ? value : (value & ((1u << count) - 1u));
}

/*
 * Bit assembler. `acc` holds `pending` not-yet-stored stream bits right-aligned (pending < 8
 * between codes). Returns AWS_OP_SUCCESS, or raises SHORT_BUFFER after parking the unwritten low
 * bits of `code` in encoder->overflow_bits when a stored byte fills the output mid-code.
 */
struct bit_sink {
    struct aws_huffman_encoder *encoder;
    struct aws_byte_buf *out;
    uint64_t acc;
    unsigned pending;
};

static int s_sink_code(struct bit_sink *sink, uint32_t pattern, unsigned num_bits) {
    struct aws_byte_buf *out = sink->out;
    sink->acc = (sink->acc << num_bits) | s_low_bits(pattern, num_bits);
    sink->pending += num_bits;
    while (sink->pending >= 8) {
        sink->pending -= 8;
        out->buffer[out->len++] = (uint8_t)(sink->acc >> sink->pending);
        if (out->len == out->capacity) {
            /* Whatever is still pending belongs to this code: the first stored byte swallowed the
             * (<8) bits older codes had left behind. */
            sink->encoder->overflow_bits.num_bits = (uint8_t)sink->pending;
            if (sink->pending) {
                sink->encoder->overflow_bits.pattern = (uint32_t)(sink->acc & ((1ull << sink->pending) - 1ull));
                return aws_raise_error(AWS_ERROR_SHORT_BUFFER);
            }
        }
    }
    return AWS_OP_SUCCESS;
}

int aws_huffman_encode(
    struct aws_huffman_encoder *encoder,
    struct aws_byte_cursor *to_encode,
    struct aws_byte_buf *output) {

    AWS_ASSERT(encoder);
    AWS_ASSERT(encoder->coder);
    AWS_ASSERT(to_encode);
    AWS_ASSERT(output);

    struct aws_huffman_symbol_coder *coder = encoder->coder;
    struct bit_sink sink = {encoder, output, 0, 0};

    if (encoder->overflow_bits.num_bits) {
        if (output->len == output->capacity) {
            return aws_raise_error(AWS_ERROR_SHORT_BUFFER);
        }
        if (s_sink_code(&sink, encoder->overflow_bits.pattern, encoder->overflow_bits.num_bits)) {
            return AWS_OP_ERR;
        }
        encoder->overflow_bits.num_bits = 0;
    }

    /* Roomy path: while 4 bytes per remaining symbol (+ the pad byte) are guaranteed to fit, no
     * capacity checks are needed and bits leave 32 at a time. */
    while (to_encode->len && (output->capacity - output->len) / 4 > to_encode->len) {
        uint8_t *dst = output->buffer + output->len;
        uint64_t acc = sink.acc;
        unsigned pending = sink.pending;
        const uint8_t *src = to_encode->ptr;
        size_t left = to_encode->len;
        int unknown = 0;
        while (left) {
            const struct aws_huffman_code code = coder->encode(*src++, coder->userdata);
            --left;
            if (code.num_bits == 0) {
                unknown = 1;
                break;
            }
            acc = (acc << code.num_bits) | s_low_bits(code.pattern, code.num_bits);
            pending += code.num_bits;
            if (pending >= 32) {
                pending -= 32;
                const uint32_t word = (uint32_t)(acc >> pending);
                dst[0] = (uint8_t)(word >> 24);
                dst[1] = (uint8_t)(word >> 16);
                dst[2] = (uint8_t)(word >> 8);
                dst[3] = (uint8_t)word;
                dst += 4;
            }
        }
        while (pending >= 8) {
            pending -= 8;
            *dst++ = (uint8_t)(acc >> pending);
        }
        output->len = (size_t)(dst - output->buffer);
        to_encode->ptr = (uint8_t *)src;
        to_encode->len = left;
        sink.acc = acc;
        sink.pending = pending;
        if (unknown) {
            /* bits of earlier symbols that did not complete a byte are dropped, like the reference */
            return aws_raise_error(AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL);
        }
    }

    /* Tight path: one capacity check per symbol and per stored byte. */
    while (to_encode->len) {
        if (output->len == output->capacity) {
            return aws_raise_error(AWS_ERROR_SHORT_BUFFER);
        }
        const struct aws_huffman_code code = coder->encode(*to_encode->ptr, coder->userdata);
        ++to_encode->ptr;
        --to_encode->len;
        if (code.num_bits == 0) {
            return aws_raise_error(AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL);
        }
        if (s_sink_code(&sink, code.pattern, code.num_bits)) {
            return AWS_OP_ERR;
        }
    }

    /* All symbols placed: pad the open byte with the LOW bits of eos_padding. Room for it is
     * certain: a byte store that filled the output with bits pending would have returned above. */
    if (sink.pending) {
        const unsigned pad = 8 - sink.pending;
        const uint8_t last = (uint8_t)((sink.acc << pad) | (encoder->eos_padding & ((1u << pad) - 1u)));
        output->buffer[output->len++] = last;
    }
    return AWS_OP_SUCCESS;
}

/* ---- decode ---- */

int aws_huffman_decode(
    struct aws_huffman_decoder *decoder,
    struct aws_byte_cursor *to_decode,
    struct aws_byte_buf *output) {

    AWS_ASSERT(decoder);
    AWS_ASSERT(decoder->coder);
    AWS_ASSERT(to_decode);
    AWS_ASSERT(output);

    struct aws_huffman_symbol_coder *coder = decoder->coder;
    uint64_t reg = decoder->working_bits;
    unsigned have = decoder->num_bits;
    const uint8_t *src = to_decode->ptr;
    size_t left = to_decode->len;
    /* stream bits not yet turned into symbols: buffered + unread */
    size_t bits_left = (size_t)have + left * 8;
    int result = AWS_OP_SUCCESS;

    for (;;) {
        /* Keep at least 32 bits in the register (or everything, near the end): pull the exact
         * number of bytes the byte-at-a-time rule would (reference huffman.c:196-211). */
        if (have < WINDOW_BITS && left) {
            size_t want = (WINDOW_BITS - have + 7) / 8;
            if (want > left) {
                want = left;
            }
            for (size_t i = 0; i < want; ++i) {
                reg |= (uint64_t)src[i] << (56 - have);
                have += 8;
            }
            src += want;
            left -= want;
        }

        uint8_t symbol = 0;
        const unsigned used = coder->decode((uint32_t)(reg >> 32), &symbol, coder->userdata);

        if (used == 0) {
            if (bits_left >= WINDOW_BITS) {
                result = aws_raise_error(AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL);
            }
            break; /* fewer than 32 bits left: padding or a code cut short, wait for more input */
        }
        if (used > bits_left) {
            break; /* the match leans on the zero fill below the real bits */
        }
        if (output->len == output->capacity) {
            if (!decoder->allow_growth) {
                result = aws_raise_error(AWS_ERROR_SHORT_BUFFER);
                break;
            }
            if (aws_byte_buf_reserve_relative(output, output->capacity)) {
                result = AWS_OP_ERR;
                break;
            }
        }
        bits_left -= used;
        reg <<= used;
        have -= used;
        aws_byte_buf_write_u8(output, symbol);
        if (bits_left == 0) {
            break;
        }
    }

    decoder->working_bits = reg;
    decoder->num_bits = (uint8_t)have;
    to_decode->ptr = (uint8_t *)src;
    to_decode->len = left;
    return result;
}
