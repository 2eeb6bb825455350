#ifndef HUFFMAN_ORACLE_H
#define HUFFMAN_ORACLE_H
/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement of the reference codec (awslabs/aws-c-compression source/huffman.c) used only
 * as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs. The product library never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED. tests/test_oracle_pins.py checks this restatement against every golden
 * vector the reference's own tests hold for the path (tests/huffman_test.c:20-37,175-194,408 and
 * all 256 rows of tests/test_huffman_static_table.def, committed as tests/golden/), against the
 * RFC 7541 Appendix C strings for the HPACK table, and — when oracle/_ref was built from
 * /root/reference — differentially against the unmodified reference on random inputs, including
 * the short-buffer and unknown-symbol paths.
 *
 * Plain pointers and sizes only, so it does not need aws-c-common.
 */
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORACLE_OK = 0,
    ORACLE_ERR_SHORT_BUFFER = 4,      /* AWS_ERROR_SHORT_BUFFER */
    ORACLE_ERR_UNKNOWN_SYMBOL = 3072, /* AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL, compression.h:17 */
};

/* huffman.h:18-26 */
struct oracle_code {
    uint32_t pattern;
    uint8_t num_bits;
};

/* One trie node per code prefix; child index 0 = none. Built MSB-first like
 * source/huffman_generator/generator.c:239-278. */
struct oracle_trie_node {
    int32_t child[2]; /* >0: inner node index; <0: -(symbol+1) leaf; 0: hole */
};

struct oracle_table {
    struct oracle_code enc[256];
    struct oracle_trie_node *nodes;
    int32_t num_nodes;
    int32_t cap_nodes;
};

/* huffman.h:63-70 */
struct oracle_encoder {
    const struct oracle_table *table;
    uint8_t eos_padding;
    struct oracle_code overflow_bits;
};

/* huffman.h:76-84 (allow_growth is a byte_buf matter and is left to the caller) */
struct oracle_decoder {
    const struct oracle_table *table;
    uint64_t working_bits;
    uint8_t num_bits;
};

/* Returns 0, or -1 when two codes collide / one is a prefix of another / a length is > 32. */
int oracle_table_init(struct oracle_table *table, const uint32_t *patterns, const uint8_t *num_bits);
void oracle_table_clean_up(struct oracle_table *table);

/* The generated coder's two callbacks (tests/test_huffman_static.c:269-273 and :276-2381). */
struct oracle_code oracle_encode_symbol(const struct oracle_table *table, uint8_t symbol);
uint8_t oracle_decode_symbol(const struct oracle_table *table, uint32_t bits, uint8_t *symbol);

void oracle_encoder_init(struct oracle_encoder *encoder, const struct oracle_table *table);
void oracle_encoder_reset(struct oracle_encoder *encoder);
void oracle_decoder_init(struct oracle_decoder *decoder, const struct oracle_table *table);
void oracle_decoder_reset(struct oracle_decoder *decoder);

size_t oracle_get_encoded_length(const struct oracle_encoder *encoder, const uint8_t *in, size_t in_len);

/* One aws_huffman_encode call. *in_consumed = how far the cursor advanced; out_len is in/out like
 * aws_byte_buf.len. Returns ORACLE_OK or the error the reference would raise. */
int oracle_encode(
    struct oracle_encoder *encoder,
    const uint8_t *in,
    size_t in_len,
    size_t *in_consumed,
    uint8_t *out,
    size_t out_capacity,
    size_t *out_len);

/* One aws_huffman_decode call (no growth). */
int oracle_decode(
    struct oracle_decoder *decoder,
    const uint8_t *in,
    size_t in_len,
    size_t *in_consumed,
    uint8_t *out,
    size_t out_capacity,
    size_t *out_len);

/*
 * Batch drivers: item i is a fresh encoder/decoder and ONE call, i.e. the per-item contract of
 * aws_huffman_encode_batch / aws_huffman_decode_batch (include/aws/compression/huffman_batch.h).
 * out_caps == NULL: packed layout, out_offsets[0..n] is written (item i sees the remaining
 * capacity). out_caps != NULL: slotted layout, item i is written at out + out_offsets[i] with
 * capacity out_caps[i]. Optional arrays may be NULL.
 */
void oracle_encode_batch(
    const struct oracle_table *table,
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
    uint8_t *overflow_num_bits);

void oracle_decode_batch(
    const struct oracle_table *table,
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
    uint8_t *leftover_num_bits);

#ifdef __cplusplus
}
#endif

#endif /* HUFFMAN_ORACLE_H */
