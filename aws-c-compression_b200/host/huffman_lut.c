/* See huffman_lut.h. Builds a binary trie of the codes (which also proves prefix-freeness), then
 * flattens it into fixed-width lookup levels. */
#include "huffman_lut.h"

#include <stdlib.h>
#include <string.h>

struct trie_node {
    int32_t child[2]; /* >0 inner node; <0 -(symbol+1); 0 none */
    uint8_t height;   /* longest path to a leaf below this node */
};

struct builder {
    struct trie_node *nodes;
    int32_t num_nodes;
    int32_t cap_nodes;
    uint32_t *entries;
    uint32_t num_entries;
    uint32_t cap_entries;
    const uint8_t *num_bits;
    unsigned sub_bits;
    int oom;
};

static int32_t s_add_node(struct builder *b) {
    if (b->num_nodes == b->cap_nodes) {
        const int32_t cap = b->cap_nodes ? b->cap_nodes * 2 : 1024;
        struct trie_node *grown = realloc(b->nodes, (size_t)cap * sizeof(*grown));
        if (!grown) {
            b->oom = 1;
            return -1;
        }
        b->nodes = grown;
        b->cap_nodes = cap;
    }
    memset(&b->nodes[b->num_nodes], 0, sizeof(struct trie_node));
    return b->num_nodes++;
}

static int64_t s_add_entries(struct builder *b, uint32_t count) {
    if (b->num_entries + count > b->cap_entries) {
        uint32_t cap = b->cap_entries ? b->cap_entries : 1024;
        while (cap < b->num_entries + count) {
            cap *= 2;
        }
        uint32_t *grown = realloc(b->entries, (size_t)cap * sizeof(uint32_t));
        if (!grown) {
            b->oom = 1;
            return -1;
        }
        b->entries = grown;
        b->cap_entries = cap;
    }
    const uint32_t base = b->num_entries;
    memset(b->entries + base, 0, (size_t)count * sizeof(uint32_t));
    b->num_entries += count;
    return base;
}

static uint8_t s_compute_heights(struct builder *b, int32_t node) {
    uint8_t h = 0;
    for (int c = 0; c < 2; ++c) {
        const int32_t next = b->nodes[node].child[c];
        uint8_t below = 0;
        if (next < 0) {
            below = 1;
        } else if (next > 0) {
            below = (uint8_t)(1 + s_compute_heights(b, next));
        }
        if (below > h) {
            h = below;
        }
    }
    b->nodes[node].height = h;
    return h;
}

/* Fills table [base, base + 2^width) for the sub-trie under `node`. `depth` bits of the index have
 * been fixed to `prefix` so far. */
static void s_fill(struct builder *b, int32_t node, uint32_t base, unsigned width, unsigned depth, uint32_t prefix) {
    for (uint32_t bit = 0; bit < 2 && !b->oom; ++bit) {
        const int32_t next = b->nodes[node].child[bit];
        const uint32_t idx = (prefix << 1) | bit;
        const unsigned fixed = depth + 1;
        if (next == 0) {
            continue; /* hole: entries stay 0 */
        }
        if (next < 0) {
            const uint32_t symbol = (uint32_t)(-next - 1);
            const uint32_t leaf = HUFFMAN_LUT_LEAF_FLAG | ((uint32_t)b->num_bits[symbol] << 8) | symbol;
            const uint32_t span = 1u << (width - fixed);
            for (uint32_t k = 0; k < span; ++k) {
                b->entries[base + (idx << (width - fixed)) + k] = leaf;
            }
        } else if (fixed < width) {
            s_fill(b, next, base, width, fixed, idx);
        } else {
            unsigned sub_width = b->nodes[next].height;
            if (sub_width > b->sub_bits) {
                sub_width = b->sub_bits;
            }
            const int64_t sub_base = s_add_entries(b, 1u << sub_width);
            if (sub_base < 0) {
                return;
            }
            b->entries[base + idx] = ((uint32_t)sub_width << 24) | (uint32_t)sub_base;
            s_fill(b, next, (uint32_t)sub_base, sub_width, 0, 0);
        }
    }
}

int huffman_lut_build(
    struct huffman_lut *lut,
    const uint32_t *patterns,
    const uint8_t *num_bits,
    unsigned root_bits,
    unsigned sub_bits) {

    memset(lut, 0, sizeof(*lut));
    if (root_bits < 1 || root_bits > 16 || sub_bits < 1 || sub_bits > 8) {
        return HUFFMAN_LUT_ERR_BAD_LENGTH;
    }
    struct builder b;
    memset(&b, 0, sizeof(b));
    b.num_bits = num_bits;
    b.sub_bits = sub_bits;
    int rc = HUFFMAN_LUT_OK;
    unsigned min_len = 0, max_len = 0;
    uint64_t kraft = 0; /* in units of 2^-32 */
    int unknown = 0;

    if (s_add_node(&b) != 0) {
        rc = HUFFMAN_LUT_ERR_OOM;
        goto done;
    }
    for (int sym = 0; sym < 256; ++sym) {
        const unsigned len = num_bits[sym];
        if (len == 0) {
            unknown = 1;
            continue;
        }
        if (len > 32) {
            rc = HUFFMAN_LUT_ERR_BAD_LENGTH;
            goto done;
        }
        if (min_len == 0 || len < min_len) {
            min_len = len;
        }
        if (len > max_len) {
            max_len = len;
        }
        kraft += 1ull << (32 - len);
        int32_t cur = 0;
        for (int bit_idx = (int)len - 1; bit_idx >= 0; --bit_idx) {
            const int bit = (int)((patterns[sym] >> bit_idx) & 1u);
            const int32_t next = b.nodes[cur].child[bit];
            if (bit_idx == 0) {
                if (next != 0) {
                    rc = HUFFMAN_LUT_ERR_NOT_PREFIX_FREE;
                    goto done;
                }
                b.nodes[cur].child[bit] = -(sym + 1);
            } else if (next < 0) {
                rc = HUFFMAN_LUT_ERR_NOT_PREFIX_FREE;
                goto done;
            } else if (next > 0) {
                cur = next;
            } else {
                const int32_t fresh = s_add_node(&b);
                if (fresh < 0) {
                    rc = HUFFMAN_LUT_ERR_OOM;
                    goto done;
                }
                b.nodes[cur].child[bit] = fresh;
                cur = fresh;
            }
        }
    }
    s_compute_heights(&b, 0);
    if (s_add_entries(&b, 1u << root_bits) < 0) {
        rc = HUFFMAN_LUT_ERR_OOM;
        goto done;
    }
    s_fill(&b, 0, 0, root_bits, 0, 0);
    if (b.oom || b.num_entries >= (1u << 24)) {
        rc = HUFFMAN_LUT_ERR_OOM;
        goto done;
    }

    lut->entries = b.entries;
    b.entries = NULL;
    lut->count = b.num_entries;
    lut->root_bits = (uint8_t)root_bits;
    lut->sub_bits = (uint8_t)sub_bits;
    lut->min_len = (uint8_t)min_len;
    lut->max_len = (uint8_t)max_len;
    lut->has_unknown_symbols = (uint8_t)unknown;
    lut->is_complete = (uint8_t)(kraft == (1ull << 32));

done:
    free(b.nodes);
    free(b.entries);
    return rc;
}

void huffman_lut_clean_up(struct huffman_lut *lut) {
    free(lut->entries);
    memset(lut, 0, sizeof(*lut));
}
