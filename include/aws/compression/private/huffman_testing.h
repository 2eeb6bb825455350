#ifndef AWS_COMPRESSION_HUFFMAN_TESTING_H
#define AWS_COMPRESSION_HUFFMAN_TESTING_H
/*
 * Round-trip helpers exported from the library for coder authors and fuzzers, plus the
 * HUFFMAN_CODE macro that turns a .def table into an array of code points.
 * Replaces the reference's include/aws/compression/private/huffman_testing.h:35-97.
 */
#include <aws/compression/huffman.h>

struct huffman_test_code_point {
    uint8_t symbol;
    struct aws_huffman_code code;
};

/* Usage:  static struct huffman_test_code_point table[] = {
 *         #include "my_table.def"
 *         };                                                                          */
#define HUFFMAN_CODE(psymbol, pbit_string, pbit_pattern, pnum_bits)                                                    \
    {.symbol = (psymbol), .code = {.pattern = (pbit_pattern), .num_bits = (pnum_bits)}},

AWS_EXTERN_C_BEGIN

/* encode `input` in one call, decode it in one call, compare. encoded_size == 0 skips the length
 * check. On failure returns AWS_OP_ERR and points *error_string at a static description. */
AWS_COMPRESSION_API
int huffman_test_transitive(
    struct aws_huffman_symbol_coder *coder,
    const char *input,
    size_t size,
    size_t encoded_size,
    const char **error_string);

/* Same, but both directions run with an output buffer that grows by output_chunk_size per call;
 * every call must make progress and fail only with AWS_ERROR_SHORT_BUFFER. */
AWS_COMPRESSION_API
int huffman_test_transitive_chunked(
    struct aws_huffman_symbol_coder *coder,
    const char *input,
    size_t size,
    size_t encoded_size,
    size_t output_chunk_size,
    const char **error_string);

AWS_EXTERN_C_END

#endif /* AWS_COMPRESSION_HUFFMAN_TESTING_H */
