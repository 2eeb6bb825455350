#ifndef AWS_COMPRESSION_B200_HUFFMAN_LUT_H
#define AWS_COMPRESSION_B200_HUFFMAN_LUT_H
/*
 * Multi-level decode lookup table built from the 256 (pattern, num_bits) pairs of a symbol coder.
 * Shared by the .def generator (which prints it into the emitted C file) and by the batched CUDA
 * context (which uploads it to the device). Internal header, not installed.
 *
 * Entry layout (uint32):
 *     leaf    : bit 31 set | len << 8 | symbol        (len = full code length, 1..32)
 *     link    : bit 31 clear | width << 24 | base     (width 1..8 = index bits of the sub-table
 *                                                      starting at entries[base]; base < 2^24)
 *     invalid : 0                                     (no code starts with these bits: a "hole")
 *
 * Lookup of a left-aligned 32-bit window `w`:
 *     e = entries[w >> (32 - root_bits)]; used = root_bits;
 *     while e is a link: e = entries[base + ((w << used) >> (32 - width))]; used += width;
 *     e == 0 -> no match (the generated decode_symbol returns 0); else (symbol, len).
 * This is the table form of the bit-at-a-time tree walk the reference generator emits
 * (reference source/huffman_generator/generator.c:175-214): the unique code that prefixes the
 * window, else 0.
 */
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HUFFMAN_LUT_LEAF_FLAG 0x80000000u
#define HUFFMAN_LUT_IS_LEAF(e) (((e)&HUFFMAN_LUT_LEAF_FLAG) != 0)
#define HUFFMAN_LUT_LEAF_SYMBOL(e) ((uint8_t)((e)&0xFFu))
#define HUFFMAN_LUT_LEAF_LEN(e) ((uint8_t)(((e) >> 8) & 0x3Fu))
#define HUFFMAN_LUT_LINK_WIDTH(e) (((e) >> 24) & 0x7Fu)
#define HUFFMAN_LUT_LINK_BASE(e) ((e)&0x00FFFFFFu)

struct huffman_lut {
    uint32_t *entries; /* malloc'd; free with huffman_lut_clean_up */
    uint32_t count;
    uint8_t root_bits;
    uint8_t sub_bits;
    uint8_t min_len; /* shortest / longest code present (0 when the table is empty) */
    uint8_t max_len;
    uint8_t has_unknown_symbols; /* some symbol has num_bits == 0 */
    uint8_t is_complete;         /* Kraft sum == 1: every window matches a code */
};

enum {
    HUFFMAN_LUT_OK = 0,
    HUFFMAN_LUT_ERR_NOT_PREFIX_FREE = 1,
    HUFFMAN_LUT_ERR_BAD_LENGTH = 2,
    HUFFMAN_LUT_ERR_OOM = 3,
};

/* patterns/num_bits: 256 entries each; pattern bits above num_bits are ignored. */
int huffman_lut_build(
    struct huffman_lut *lut,
    const uint32_t *patterns,
    const uint8_t *num_bits,
    unsigned root_bits,
    unsigned sub_bits);

void huffman_lut_clean_up(struct huffman_lut *lut);

/* Reference lookup (host). Returns the code length, 0 when nothing matches. */
static inline uint8_t huffman_lut_decode(const struct huffman_lut *lut, uint32_t window, uint8_t *symbol) {
    uint32_t e = lut->entries[window >> (32 - lut->root_bits)];
    unsigned used = lut->root_bits;
    while (e != 0 && !HUFFMAN_LUT_IS_LEAF(e)) {
        const unsigned width = HUFFMAN_LUT_LINK_WIDTH(e);
        e = lut->entries[HUFFMAN_LUT_LINK_BASE(e) + ((uint32_t)(window << used) >> (32 - width))];
        used += width;
    }
    if (e == 0) {
        return 0;
    }
    *symbol = HUFFMAN_LUT_LEAF_SYMBOL(e);
    return HUFFMAN_LUT_LEAF_LEN(e);
}

#ifdef __cplusplus
}
#endif

#endif /* AWS_COMPRESSION_B200_HUFFMAN_LUT_H */
