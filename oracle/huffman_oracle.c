/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE. See huffman_oracle.h.
 *
 * CPU restatement of awslabs/aws-c-compression's Huffman codec. Every function names the
 * reference lines it follows (paths relative to /root/reference). It keeps the reference's
 * byte-at-a-time structure on purpose: it is the checker, not something to be fast.
 */
#include "huffman_oracle.h"

#include <stdlib.h>
#include <string.h>

enum { PATTERN_BITS = 32 }; /* source/huffman.c:10 MAX_PATTERN_BITS */

/* ---- symbol coder (what huffman_generator emits) ---- */

static int32_t s_new_node(struct oracle_table *t) {
    if (t->num_nodes == t->cap_nodes) {
        int32_t cap = t->cap_nodes ? t->cap_nodes * 2 : 512;
        struct oracle_trie_node *grown = realloc(t->nodes, (size_t)cap * sizeof(*grown));
        if (!grown) {
            return -1;
        }
        t->nodes = grown;
        t->cap_nodes = cap;
    }
    memset(&t->nodes[t->num_nodes], 0, sizeof(t->nodes[0]));
    return t->num_nodes++;
}

/* source/huffman_generator/generator.c:239-278: insert each code MSB-first into a binary trie;
 * symbols with num_bits == 0 are skipped (:243-245). Unlike the generator this reports prefix
 * collisions instead of asserting. */
int oracle_table_init(struct oracle_table *t, const uint32_t *patterns, const uint8_t *num_bits) {
    memset(t, 0, sizeof(*t));
    if (s_new_node(t) != 0) {
        return -1;
    }
    for (int sym = 0; sym < 256; ++sym) {
        t->enc[sym].pattern = patterns[sym];
        t->enc[sym].num_bits = num_bits[sym];
        const int len = num_bits[sym];
        if (len == 0) {
            continue;
        }
        if (len > PATTERN_BITS) {
            return -1;
        }
        int32_t cur = 0;
        for (int bit_idx = len - 1; bit_idx >= 0; --bit_idx) {
            const int b = (int)((patterns[sym] >> bit_idx) & 1u);
            const int32_t next = t->nodes[cur].child[b];
            if (bit_idx == 0) {
                if (next != 0) {
                    return -1; /* duplicate code or a longer code already runs through here */
                }
                t->nodes[cur].child[b] = -(sym + 1);
            } else if (next > 0) {
                cur = next;
            } else if (next < 0) {
                return -1; /* an existing code is a prefix of this one */
            } else {
                const int32_t fresh = s_new_node(t);
                if (fresh < 0) {
                    return -1;
                }
                t->nodes[cur].child[b] = fresh;
                cur = fresh;
            }
        }
    }
    return 0;
}

void oracle_table_clean_up(struct oracle_table *t) {
    free(t->nodes);
    memset(t, 0, sizeof(*t));
}

/* tests/test_huffman_static.c:269-273 (emitted by generator.c:313-321): plain table load. */
struct oracle_code oracle_encode_symbol(const struct oracle_table *t, uint8_t symbol) {
    return t->enc[symbol];
}

/* tests/test_huffman_static.c:276-2381 (emitted by generator.c:175-214): test one bit per level
 * from the top of the 32-bit window; a leaf child yields (symbol, length); a missing child is
 * "return 0; / * invalid node * /" (generator.c:156-158). */
uint8_t oracle_decode_symbol(const struct oracle_table *t, uint32_t bits, uint8_t *symbol) {
    int32_t cur = 0;
    for (int depth = 0; depth < PATTERN_BITS; ++depth) {
        const int b = (int)((bits >> (31 - depth)) & 1u);
        const int32_t next = t->nodes[cur].child[b];
        if (next == 0) {
            return 0;
        }
        if (next < 0) {
            *symbol = (uint8_t)(-next - 1);
            return (uint8_t)(depth + 1);
        }
        cur = next;
    }
    return 0;
}

/* ---- init / reset: source/huffman.c:12-46 ---- */

void oracle_encoder_init(struct oracle_encoder *e, const struct oracle_table *table) {
    memset(e, 0, sizeof(*e));
    e->table = table;
    e->eos_padding = 0xFF; /* huffman.c:19 */
}

void oracle_encoder_reset(struct oracle_encoder *e) {
    memset(&e->overflow_bits, 0, sizeof(e->overflow_bits)); /* huffman.c:26 */
}

void oracle_decoder_init(struct oracle_decoder *d, const struct oracle_table *table) {
    memset(d, 0, sizeof(*d));
    d->table = table;
}

void oracle_decoder_reset(struct oracle_decoder *d) {
    d->working_bits = 0; /* huffman.c:40-41 */
    d->num_bits = 0;
}

/* ---- encode ---- */

/* source/huffman.c:107-129: sum of code lengths, rounded up to bytes; unknown symbols add 0 and
 * pending overflow bits are ignored. */
size_t oracle_get_encoded_length(const struct oracle_encoder *e, const uint8_t *in, size_t in_len) {
    size_t bits = 0;
    for (size_t i = 0; i < in_len; ++i) {
        bits += oracle_encode_symbol(e->table, in[i]).num_bits;
    }
    return bits / 8 + (bits % 8 ? 1 : 0);
}

struct enc_cursor {
    struct oracle_encoder *encoder;
    uint8_t *out;
    size_t out_capacity;
    size_t *out_len;
    uint8_t working; /* byte being assembled */
    uint8_t bit_pos; /* free bits left in `working`, 8..1 */
};

/* source/huffman.c:59-105 encode_write_bit_pattern: feed one code into the byte assembler.
 * Returns ORACLE_OK, ORACLE_ERR_UNKNOWN_SYMBOL (:62-64) or ORACLE_ERR_SHORT_BUFFER with the
 * unwritten low bits parked in overflow_bits (:88-99). */
static int s_put_code(struct enc_cursor *c, struct oracle_code code) {
    if (code.num_bits == 0) {
        return ORACLE_ERR_UNKNOWN_SYMBOL;
    }
    uint8_t remaining = code.num_bits;
    while (remaining > 0) {
        const uint8_t take = remaining > c->bit_pos ? c->bit_pos : remaining;
        /* :70-71: drop the unused high bits of the 32-bit pattern plus what was already written */
        uint8_t cut = (uint8_t)((PATTERN_BITS - code.num_bits) + (code.num_bits - remaining));
        /* :76: left-align what is left, then slide it under the bits already in `working`.
         * The assignment to a uint8_t keeps only the byte being assembled. */
        c->working |= (uint8_t)((code.pattern << cut) >> (PATTERN_BITS - c->bit_pos));
        remaining = (uint8_t)(remaining - take);
        c->bit_pos = (uint8_t)(c->bit_pos - take);

        if (c->bit_pos == 0) {
            /* :81-86 (aws_byte_buf_write_u8 is bounds-checked) */
            if (*c->out_len < c->out_capacity) {
                c->out[(*c->out_len)++] = c->working;
            }
            c->bit_pos = 8;
            c->working = 0;
            if (*c->out_len == c->out_capacity) {
                /* :88-99 */
                c->encoder->overflow_bits.num_bits = remaining;
                if (remaining) {
                    cut = (uint8_t)(cut + take);
                    c->encoder->overflow_bits.pattern = (code.pattern << cut) >> (PATTERN_BITS - remaining);
                    return ORACLE_ERR_SHORT_BUFFER;
                }
            }
        }
    }
    return ORACLE_OK;
}

/* source/huffman.c:131-187 aws_huffman_encode */
int oracle_encode(
    struct oracle_encoder *e,
    const uint8_t *in,
    size_t in_len,
    size_t *in_consumed,
    uint8_t *out,
    size_t out_capacity,
    size_t *out_len) {

    struct enc_cursor c = {e, out, out_capacity, out_len, 0, 8}; /* :141-146 */
    size_t pos = 0;
    *in_consumed = 0;

    /* :149-159 flush what the previous call could not place */
    if (e->overflow_bits.num_bits) {
        if (*out_len == out_capacity) {
            return ORACLE_ERR_SHORT_BUFFER;
        }
        const int rc = s_put_code(&c, e->overflow_bits);
        if (rc) {
            return rc;
        }
        e->overflow_bits.num_bits = 0;
    }

    /* :161-173 */
    while (pos < in_len) {
        if (*out_len == out_capacity) {
            *in_consumed = pos;
            return ORACLE_ERR_SHORT_BUFFER;
        }
        const uint8_t sym = in[pos++]; /* the cursor moves before the code is written (:167) */
        const int rc = s_put_code(&c, oracle_encode_symbol(e->table, sym));
        if (rc) {
            *in_consumed = pos;
            return rc;
        }
    }
    *in_consumed = pos;

    /* :178-184 pad the last byte with the LOW bit_pos bits of eos_padding */
    if (c.bit_pos != 8) {
        struct oracle_code pad = {e->eos_padding, c.bit_pos};
        (void)s_put_code(&c, pad);
    }
    return ORACLE_OK;
}

/* ---- decode ---- */

/* source/huffman.c:213-286 aws_huffman_decode with :196-211 decode_fill_working_bits inlined */
int oracle_decode(
    struct oracle_decoder *d,
    const uint8_t *in,
    size_t in_len,
    size_t *in_consumed,
    uint8_t *out,
    size_t out_capacity,
    size_t *out_len) {

    size_t pos = 0;
    size_t bits_left = d->num_bits + in_len * 8; /* :228 */
    *in_consumed = 0;

    for (;;) {
        /* :196-211 top up the left-aligned 64-bit register to at least 32 bits */
        while (d->num_bits < PATTERN_BITS && pos < in_len) {
            d->working_bits |= (uint64_t)in[pos++] << (64 - 8 - d->num_bits);
            d->num_bits = (uint8_t)(d->num_bits + 8);
        }
        *in_consumed = pos;

        uint8_t symbol = 0;
        const uint8_t used = oracle_decode_symbol(d->table, (uint32_t)(d->working_bits >> 32), &symbol); /* :235-238 */

        if (used == 0) {
            /* :240-247 */
            return bits_left < PATTERN_BITS ? ORACLE_OK : ORACLE_ERR_UNKNOWN_SYMBOL;
        }
        if (used > bits_left) {
            return ORACLE_OK; /* :248-255 trailing partial code / padding */
        }
        if (*out_len == out_capacity) {
            return ORACLE_ERR_SHORT_BUFFER; /* :257-268 without growth */
        }
        bits_left -= used; /* :270-272 */
        d->working_bits <<= used;
        d->num_bits = (uint8_t)(d->num_bits - used);
        out[(*out_len)++] = symbol; /* :275 */
        if (bits_left == 0) {
            return ORACLE_OK; /* :278-280 */
        }
    }
}

/* ---- batch drivers (per item: fresh state, one call) ---- */

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
    uint8_t *overflow_num_bits) {

    uint64_t cursor = 0;
    for (size_t i = 0; i < n; ++i) {
        struct oracle_encoder e;
        oracle_encoder_init(&e, table);
        e.eos_padding = eos_padding;
        uint64_t base, cap;
        if (out_caps) {
            base = out_offsets[i];
            cap = out_caps[i];
        } else {
            base = cursor;
            cap = out_capacity > cursor ? out_capacity - cursor : 0;
            out_offsets[i] = cursor;
        }
        size_t used = 0, len = 0;
        const int rc = oracle_encode(
            &e, in + in_offsets[i], (size_t)(in_offsets[i + 1] - in_offsets[i]), &used, out + base, (size_t)cap, &len);
        cursor += len;
        if (out_lens) out_lens[i] = len;
        if (status) status[i] = rc;
        if (consumed) consumed[i] = used;
        if (overflow_pattern) overflow_pattern[i] = e.overflow_bits.num_bits ? e.overflow_bits.pattern : 0;
        if (overflow_num_bits) overflow_num_bits[i] = e.overflow_bits.num_bits;
    }
    if (!out_caps) {
        out_offsets[n] = cursor;
    }
}

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
    uint8_t *leftover_num_bits) {

    uint64_t cursor = 0;
    for (size_t i = 0; i < n; ++i) {
        struct oracle_decoder d;
        oracle_decoder_init(&d, table);
        uint64_t base, cap;
        if (out_caps) {
            base = out_offsets[i];
            cap = out_caps[i];
        } else {
            base = cursor;
            cap = out_capacity > cursor ? out_capacity - cursor : 0;
            out_offsets[i] = cursor;
        }
        size_t used = 0, len = 0;
        const int rc = oracle_decode(
            &d, in + in_offsets[i], (size_t)(in_offsets[i + 1] - in_offsets[i]), &used, out + base, (size_t)cap, &len);
        cursor += len;
        if (out_lens) out_lens[i] = len;
        if (status) status[i] = rc;
        if (consumed) consumed[i] = used;
        if (leftover_working_bits) leftover_working_bits[i] = d.working_bits;
        if (leftover_num_bits) leftover_num_bits[i] = d.num_bits;
    }
    if (!out_caps) {
        out_offsets[n] = cursor;
    }
}
