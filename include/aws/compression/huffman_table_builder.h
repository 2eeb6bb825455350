#ifndef AWS_COMPRESSION_HUFFMAN_TABLE_BUILDER_H
#define AWS_COMPRESSION_HUFFMAN_TABLE_BUILDER_H
/*
 * From data to a code table: the step before the path. The reference only CONSUMES tables
 * (a .def file through source/huffman_generator/generator.c, tests/test_huffman_static_table.def); nothing in
 * it builds one. Here:
 *     symbol counts      aws_huffman_histogram[_device]       one pass over the bytes on the B200 (HBM-bound)
 *     code lengths       aws_huffman_code_lengths_from_counts package-merge: the optimal prefix code whose
 *                                                             longest code has at most max_bits bits
 *     codes              aws_huffman_canonical_codes          canonical, ordered by (length, symbol) like
 *                                                             RFC 7541 Appendix B; optional EOS = all ones
 *     consumers          aws_huffman_batch_ctx_new_from_code_table (huffman_batch.h), aws_huffman_table_coder
 *                        (a run-time aws_huffman_symbol_coder for the streaming API), aws_huffman_code_table_write_def
 *                        (the generator's input grammar)
 */
#include <aws/compression/huffman.h>

AWS_PUSH_SANE_WARNING_LEVEL

/* A symbol coder over a code table built at run time: `coder` can be handed to aws_huffman_encoder_init /
 * aws_huffman_decoder_init / aws_huffman_batch_ctx_new. Must outlive its users; not copyable (userdata points
 * into the struct). */
struct aws_huffman_table_coder {
    struct aws_huffman_symbol_coder coder;
    struct aws_huffman_code codes[256];
    uint32_t *lut_entries; /* multi-level decode table (owned) */
    uint32_t lut_count;
    uint8_t lut_root_bits;
};

AWS_EXTERN_C_BEGIN

/* counts[v] = how often byte value v occurs in in[0, size). Host pointers; runs on CUDA device `device_id`
 * (AWS_ERROR_COMPRESSION_DEVICE_FAILURE without one: no CPU fallback). */
AWS_COMPRESSION_API
int aws_huffman_histogram(int device_id, const uint8_t *in, uint64_t size, uint64_t counts[256]);

/* Device pointers; enqueued on cuda_stream (a cudaStream_t, may be NULL = the default stream) of the current
 * device. counts (256 x uint64 in device memory) is overwritten. */
AWS_COMPRESSION_API
int aws_huffman_histogram_device(const uint8_t *in, uint64_t size, uint64_t *counts, void *cuda_stream);

/*
 * Optimal code lengths under a length limit (1 <= max_bits <= 32). cover_all_symbols: symbols that never
 * occur still get a (long) code, so that aws_huffman_encode never meets an unknown symbol; otherwise their
 * length is 0 (no code). reserve_eos: a 257th symbol lighter than all others takes part, so that the all-ones
 * code is left to it (HPACK's padding convention); its length is returned through eos_length (may be NULL).
 * AWS_ERROR_INVALID_ARGUMENT when 2^max_bits codes cannot hold the symbols.
 */
AWS_COMPRESSION_API
int aws_huffman_code_lengths_from_counts(
    const uint64_t counts[256],
    unsigned max_bits,
    bool cover_all_symbols,
    bool reserve_eos,
    uint8_t lengths[256],
    uint8_t *eos_length);

/* Canonical codes for given lengths (0 = no code): numerically increasing with (length, symbol), EOS last.
 * AWS_ERROR_INVALID_ARGUMENT when the lengths violate Kraft's inequality. eos_code may be NULL. */
AWS_COMPRESSION_API
int aws_huffman_canonical_codes(
    const uint8_t lengths[256],
    uint8_t eos_length,
    struct aws_huffman_code codes[256],
    struct aws_huffman_code *eos_code);

/* The two steps above in one call. */
AWS_COMPRESSION_API
int aws_huffman_code_table_from_counts(
    const uint64_t counts[256],
    unsigned max_bits,
    bool cover_all_symbols,
    bool reserve_eos,
    struct aws_huffman_code codes[256],
    struct aws_huffman_code *eos_code);

/* Writes the table in the .def grammar huffman_generator reads: HUFFMAN_CODE(symbol, "bits", 0xhex, length). */
AWS_COMPRESSION_API
int aws_huffman_code_table_write_def(const struct aws_huffman_code codes[256], const char *path);

/* AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE when the table is not a prefix code. */
AWS_COMPRESSION_API
int aws_huffman_table_coder_init(struct aws_huffman_table_coder *coder, const struct aws_huffman_code codes[256]);

AWS_COMPRESSION_API
void aws_huffman_table_coder_clean_up(struct aws_huffman_table_coder *coder);

AWS_EXTERN_C_END
AWS_POP_SANE_WARNING_LEVEL

#endif /* AWS_COMPRESSION_HUFFMAN_TABLE_BUILDER_H */
