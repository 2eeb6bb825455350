/*
 * Code-table construction from symbol counts: the step BEFORE the path. The reference has no
 * tree builder at all — it only consumes given tables (.def files, source/huffman_generator/generator.c) —
 * so nothing here restates reference code; the algorithms are the published ones:
 *   lengths : package-merge (Larmore & Hirschberg 1990), the optimal length-limited prefix code
 *   codes   : canonical assignment, ordered by (length, symbol) — the convention of RFC 7541 Appendix B, whose
 *             longest code is all ones and belongs to EOS
 * plus a runtime, table-driven aws_huffman_symbol_coder over such a table (the generator's output needs a
 * compile step; this one does not) and a writer for the .def grammar the generator reads.
 */
#include <aws/compression/huffman_table_builder.h>

#include "huffman_lut.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_SYMS 257 /* 256 byte values + the optional EOS */

struct pm_item {
    unsigned __int128 weight;
    uint8_t uses[MAX_SYMS]; /* how many leaves of each symbol the item contains (over all levels so far) */
};

static int s_cmp_leaf(const void *a, const void *b) {
    const struct pm_item *x = a, *y = b;
    if (x->weight != y->weight) {
        return x->weight < y->weight ? -1 : 1;
    }
    return 0;
}

int aws_huffman_code_lengths_from_counts(
    const uint64_t counts[256],
    unsigned max_bits,
    bool cover_all_symbols,
    bool reserve_eos,
    uint8_t lengths[256],
    uint8_t *eos_length) {

    if (!counts || !lengths || max_bits < 1 || max_bits > 32) {
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    }
    memset(lengths, 0, 256);
    if (eos_length) {
        *eos_length = 0;
    }

    /* leaves: weights scaled by 2^16, + 1 so that symbols that never occur still sort (and cost) last */
    struct pm_item *leaves = calloc(MAX_SYMS, sizeof(struct pm_item));
    int sym_of[MAX_SYMS];
    size_t n = 0;
    if (!leaves) {
        return aws_raise_error(AWS_ERROR_OOM);
    }
    for (int s = 0; s < 256; ++s) {
        if (counts[s] || cover_all_symbols) {
            leaves[n].weight = ((unsigned __int128)counts[s] << 16) | 1u;
            sym_of[n] = s;
            ++n;
        }
    }
    if (reserve_eos) {
        leaves[n].weight = 0; /* lighter than everything: ends up with the longest code */
        sym_of[n] = 256;
        ++n;
    }
    if (n == 0) {
        free(leaves);
        return AWS_OP_SUCCESS;
    }
    if (n == 1) {
        if (sym_of[0] < 256) {
            lengths[sym_of[0]] = 1;
        } else if (eos_length) {
            *eos_length = 1;
        }
        free(leaves);
        return AWS_OP_SUCCESS;
    }
    if (max_bits < 32 && ((size_t)1 << max_bits) < n) {
        free(leaves);
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT); /* more symbols than codes of that length */
    }
    /* stable order: by weight, ties by symbol (qsort is not stable: fold the index into the key) */
    for (size_t i = 0; i < n; ++i) {
        leaves[i].weight = (leaves[i].weight << 9) | (unsigned)sym_of[i];
        leaves[i].uses[i] = 0;
    }
    /* sort (weight, symbol) pairs while keeping sym_of in step: sort an index permutation instead */
    {
        struct pm_item *tmp = malloc(n * sizeof(struct pm_item));
        if (!tmp) {
            free(leaves);
            return aws_raise_error(AWS_ERROR_OOM);
        }
        memcpy(tmp, leaves, n * sizeof(struct pm_item));
        qsort(tmp, n, sizeof(struct pm_item), s_cmp_leaf);
        int sorted_sym[MAX_SYMS];
        for (size_t i = 0; i < n; ++i) {
            sorted_sym[i] = (int)(unsigned)(tmp[i].weight & 0x1ff);
            tmp[i].weight >>= 9;
            memset(tmp[i].uses, 0, MAX_SYMS);
            tmp[i].uses[i] = 1; /* `uses` is indexed by sorted position */
        }
        memcpy(leaves, tmp, n * sizeof(struct pm_item));
        memcpy(sym_of, sorted_sym, n * sizeof(int));
        free(tmp);
    }

    /* package-merge: level max_bits holds the leaves; each level above merges the leaves with the packages
     * (pairs) of the level below; the answer is the first 2n - 2 items of level 1 */
    const size_t cap = 2 * n;
    struct pm_item *cur = malloc(cap * sizeof(struct pm_item));
    struct pm_item *next = malloc(cap * sizeof(struct pm_item));
    if (!cur || !next) {
        free(cur);
        free(next);
        free(leaves);
        return aws_raise_error(AWS_ERROR_OOM);
    }
    memcpy(cur, leaves, n * sizeof(struct pm_item));
    size_t cur_n = n;
    for (unsigned level = max_bits; level > 1; --level) {
        /* packages of `cur`, merged with the leaves */
        const size_t packages = cur_n / 2;
        size_t li = 0, pi = 0, out = 0;
        while (out < cap && (li < n || pi < packages)) {
            unsigned __int128 pw = 0;
            if (pi < packages) {
                pw = cur[2 * pi].weight + cur[2 * pi + 1].weight;
            }
            if (pi >= packages || (li < n && leaves[li].weight <= pw)) {
                next[out++] = leaves[li++];
            } else {
                next[out].weight = pw;
                for (size_t k = 0; k < n; ++k) {
                    next[out].uses[k] = (uint8_t)(cur[2 * pi].uses[k] + cur[2 * pi + 1].uses[k]);
                }
                ++out;
                ++pi;
            }
        }
        struct pm_item *swap = cur;
        cur = next;
        next = swap;
        cur_n = out;
    }
    uint8_t len_sorted[MAX_SYMS];
    memset(len_sorted, 0, sizeof(len_sorted));
    for (size_t i = 0; i < 2 * n - 2 && i < cur_n; ++i) {
        for (size_t k = 0; k < n; ++k) {
            len_sorted[k] = (uint8_t)(len_sorted[k] + cur[i].uses[k]);
        }
    }
    for (size_t k = 0; k < n; ++k) {
        if (sym_of[k] < 256) {
            lengths[sym_of[k]] = len_sorted[k];
        } else if (eos_length) {
            *eos_length = len_sorted[k];
        }
    }
    free(cur);
    free(next);
    free(leaves);
    return AWS_OP_SUCCESS;
}

int aws_huffman_canonical_codes(
    const uint8_t lengths[256],
    uint8_t eos_length,
    struct aws_huffman_code codes[256],
    struct aws_huffman_code *eos_code) {

    if (!lengths || !codes) {
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    }
    uint64_t kraft = 0; /* units of 2^-32 */
    for (int s = 0; s <= 256; ++s) {
        const unsigned len = s < 256 ? lengths[s] : eos_length;
        if (len > 32) {
            return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
        }
        if (len) {
            kraft += 1ull << (32 - len);
        }
    }
    if (kraft > (1ull << 32)) {
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT); /* no prefix code has these lengths */
    }
    uint64_t code = 0;
    unsigned prev = 0;
    for (unsigned len = 1; len <= 32; ++len) {
        for (int s = 0; s <= 256; ++s) {
            const unsigned mine = s < 256 ? lengths[s] : eos_length;
            if (mine != len) {
                continue;
            }
            code <<= (len - prev);
            prev = len;
            if (s < 256) {
                codes[s].pattern = (uint32_t)code;
                codes[s].num_bits = (uint8_t)len;
            } else if (eos_code) {
                eos_code->pattern = (uint32_t)code;
                eos_code->num_bits = (uint8_t)len;
            }
            ++code;
        }
    }
    for (int s = 0; s < 256; ++s) {
        if (!lengths[s]) {
            codes[s].pattern = 0;
            codes[s].num_bits = 0;
        }
    }
    if (eos_code && !eos_length) {
        eos_code->pattern = 0;
        eos_code->num_bits = 0;
    }
    return AWS_OP_SUCCESS;
}

int aws_huffman_code_table_from_counts(
    const uint64_t counts[256],
    unsigned max_bits,
    bool cover_all_symbols,
    bool reserve_eos,
    struct aws_huffman_code codes[256],
    struct aws_huffman_code *eos_code) {

    uint8_t lengths[256];
    uint8_t eos_length = 0;
    if (aws_huffman_code_lengths_from_counts(counts, max_bits, cover_all_symbols, reserve_eos, lengths, &eos_length)) {
        return AWS_OP_ERR;
    }
    return aws_huffman_canonical_codes(lengths, eos_length, codes, eos_code);
}

int aws_huffman_code_table_write_def(const struct aws_huffman_code codes[256], const char *path) {
    if (!codes || !path) {
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    }
    FILE *f = fopen(path, "w");
    if (!f) {
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    }
    /* the grammar of the reference's tests/test_huffman_static_table.def (generator.c:86-140):
     * HUFFMAN_CODE(symbol, "bits", 0xhex, length) */
    fprintf(f, "#ifndef HUFFMAN_CODE\n#error \"define HUFFMAN_CODE first\"\n#endif\n");
    for (int s = 0; s < 256; ++s) {
        const unsigned len = codes[s].num_bits;
        if (!len) {
            continue;
        }
        char bits[33];
        for (unsigned b = 0; b < len; ++b) {
            bits[b] = (codes[s].pattern >> (len - 1 - b)) & 1u ? '1' : '0';
        }
        bits[len] = 0;
        fprintf(f, "HUFFMAN_CODE(%d, \"%s\", 0x%x, %u)\n", s, bits, codes[s].pattern, len);
    }
    return fclose(f) == 0 ? AWS_OP_SUCCESS : aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
}

/* ---- a symbol coder over a table built at run time ------------------------------------------------------ */
static struct aws_huffman_code s_table_encode(uint8_t symbol, void *userdata) {
    const struct aws_huffman_table_coder *t = userdata;
    return t->codes[symbol];
}

static uint8_t s_table_decode(uint32_t bits, uint8_t *symbol, void *userdata) {
    const struct aws_huffman_table_coder *t = userdata;
    const struct huffman_lut lut = {
        .entries = t->lut_entries,
        .count = t->lut_count,
        .root_bits = t->lut_root_bits,
    };
    return huffman_lut_decode(&lut, bits, symbol);
}

int aws_huffman_table_coder_init(struct aws_huffman_table_coder *coder, const struct aws_huffman_code codes[256]) {
    if (!coder || !codes) {
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    }
    memset(coder, 0, sizeof(*coder));
    uint32_t patterns[256];
    uint8_t num_bits[256];
    for (int s = 0; s < 256; ++s) {
        if (codes[s].num_bits > 32) {
            return aws_raise_error(AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE);
        }
        coder->codes[s] = codes[s];
        num_bits[s] = codes[s].num_bits;
        patterns[s] = codes[s].num_bits >= 32 ? codes[s].pattern : (codes[s].pattern & ((1u << codes[s].num_bits) - 1u));
        coder->codes[s].pattern = patterns[s];
    }
    struct huffman_lut lut;
    const int rc = huffman_lut_build(&lut, patterns, num_bits, 10, 8);
    if (rc == HUFFMAN_LUT_ERR_OOM) {
        return aws_raise_error(AWS_ERROR_OOM);
    }
    if (rc != HUFFMAN_LUT_OK) {
        return aws_raise_error(AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE);
    }
    coder->lut_entries = lut.entries; /* ownership moves to the coder */
    coder->lut_count = lut.count;
    coder->lut_root_bits = lut.root_bits;
    coder->coder.encode = s_table_encode;
    coder->coder.decode = s_table_decode;
    coder->coder.userdata = coder;
    return AWS_OP_SUCCESS;
}

void aws_huffman_table_coder_clean_up(struct aws_huffman_table_coder *coder) {
    if (coder) {
        free(coder->lut_entries);
        memset(coder, 0, sizeof(*coder));
    }
}
