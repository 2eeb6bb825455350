/*
 * Round-trip helpers exported from the library (coder authors and fuzzers call them).
 * Replaces the reference's source/huffman_testing.c:15-173: same two entry points, same failure
 * strings, same acceptance rules; scratch space comes from the heap rather than VLAs so large
 * inputs do not depend on stack size.
 */
#include <aws/compression/private/huffman_testing.h>

struct round_trip {
    struct aws_huffman_encoder encoder;
    struct aws_huffman_decoder decoder;
    uint8_t *encoded; /* 2 * size bytes, zeroed */
    uint8_t *decoded; /* size bytes, zeroed */
    size_t encoded_room;
};

static int s_round_trip_begin(struct round_trip *rt, struct aws_huffman_symbol_coder *coder, size_t size) {
    aws_huffman_encoder_init(&rt->encoder, coder);
    aws_huffman_decoder_init(&rt->decoder, coder);
    rt->encoded_room = size * 2;
    rt->encoded = calloc(rt->encoded_room ? rt->encoded_room : 1, 1);
    rt->decoded = calloc(size ? size : 1, 1);
    return (rt->encoded && rt->decoded) ? AWS_OP_SUCCESS : AWS_OP_ERR;
}

static int s_round_trip_end(struct round_trip *rt, const char **error_string, const char *why) {
    free(rt->encoded);
    free(rt->decoded);
    if (why) {
        *error_string = why;
        return AWS_OP_ERR;
    }
    return AWS_OP_SUCCESS;
}

int huffman_test_transitive(
    struct aws_huffman_symbol_coder *coder,
    const char *input,
    size_t size,
    size_t encoded_size,
    const char **error_string) {

    struct round_trip rt;
    if (s_round_trip_begin(&rt, coder, size)) {
        return s_round_trip_end(&rt, error_string, "out of memory");
    }

    struct aws_byte_cursor to_encode = aws_byte_cursor_from_array(input, size);
    struct aws_byte_buf encoded = aws_byte_buf_from_empty_array(rt.encoded, rt.encoded_room);
    struct aws_byte_buf decoded = aws_byte_buf_from_empty_array(rt.decoded, size);

    if (aws_huffman_encode(&rt.encoder, &to_encode, &encoded) != AWS_OP_SUCCESS) {
        return s_round_trip_end(&rt, error_string, "aws_huffman_encode failed");
    }
    if (to_encode.len != 0) {
        return s_round_trip_end(&rt, error_string, "not all data encoded");
    }
    if (encoded_size && encoded.len != encoded_size) {
        return s_round_trip_end(&rt, error_string, "encoded length is incorrect");
    }

    struct aws_byte_cursor to_decode = aws_byte_cursor_from_buf(&encoded);
    if (aws_huffman_decode(&rt.decoder, &to_decode, &decoded) != AWS_OP_SUCCESS) {
        return s_round_trip_end(&rt, error_string, "aws_huffman_decode failed");
    }
    if (to_decode.len != 0) {
        return s_round_trip_end(&rt, error_string, "not all encoded data was decoded");
    }
    if (decoded.len != size) {
        return s_round_trip_end(&rt, error_string, "decode output size incorrect");
    }
    if (size && memcmp(input, rt.decoded, size) != 0) {
        return s_round_trip_end(&rt, error_string, "decoded data does not match input data");
    }
    return s_round_trip_end(&rt, error_string, NULL);
}

int huffman_test_transitive_chunked(
    struct aws_huffman_symbol_coder *coder,
    const char *input,
    size_t size,
    size_t encoded_size,
    size_t output_chunk_size,
    const char **error_string) {

    struct round_trip rt;
    if (s_round_trip_begin(&rt, coder, size)) {
        return s_round_trip_end(&rt, error_string, "out of memory");
    }

    /* Encode with an output window that opens output_chunk_size bytes per call. */
    struct aws_byte_cursor to_encode = aws_byte_cursor_from_array(input, size);
    struct aws_byte_buf encoded = {.len = 0, .buffer = rt.encoded, .capacity = 0, .allocator = NULL};
    int rc;
    do {
        const size_t before = encoded.len;
        encoded.capacity += output_chunk_size;
        rc = aws_huffman_encode(&rt.encoder, &to_encode, &encoded);
        if (encoded.len == before) {
            return s_round_trip_end(&rt, error_string, "encode didn't write any data");
        }
        if (rc != AWS_OP_SUCCESS && aws_last_error() != AWS_ERROR_SHORT_BUFFER) {
            return s_round_trip_end(&rt, error_string, "encode returned wrong error code");
        }
    } while (rc != AWS_OP_SUCCESS);

    if (encoded.len > rt.encoded_room) {
        return s_round_trip_end(&rt, error_string, "too much data encoded");
    }
    if (encoded_size && encoded.len != encoded_size) {
        return s_round_trip_end(&rt, error_string, "encoded length is incorrect");
    }

    /* Decode the same way; the window never opens past `size`. */
    struct aws_byte_cursor to_decode = aws_byte_cursor_from_buf(&encoded);
    struct aws_byte_buf decoded = {.len = 0, .buffer = rt.decoded, .capacity = 0, .allocator = NULL};
    do {
        const size_t before = decoded.len;
        decoded.capacity += output_chunk_size;
        if (decoded.capacity > size) {
            decoded.capacity = size;
        }
        rc = aws_huffman_decode(&rt.decoder, &to_decode, &decoded);
        if (decoded.len == before) {
            return s_round_trip_end(&rt, error_string, "decode didn't write any data");
        }
        if (rc != AWS_OP_SUCCESS && aws_last_error() != AWS_ERROR_SHORT_BUFFER) {
            return s_round_trip_end(&rt, error_string, "decode returned wrong error code");
        }
    } while (rc != AWS_OP_SUCCESS);

    if (decoded.len != size) {
        return s_round_trip_end(&rt, error_string, "decode output size incorrect");
    }
    if (size && memcmp(input, rt.decoded, size) != 0) {
        return s_round_trip_end(&rt, error_string, "decoded data does not match input data");
    }
    return s_round_trip_end(&rt, error_string, NULL);
}
