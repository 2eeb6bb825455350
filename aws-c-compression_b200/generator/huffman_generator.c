/*
 * huffman_generator: turns a .def code table into a C file exporting
 *     struct aws_huffman_symbol_coder *<name>_get_coder(void);
 *
 * Command line and input grammar are those of the reference tool
 * (reference source/huffman_generator/generator.c:216-226 and :42-105):
 *     huffman_generator <input.def> <output.c> <name>
 *     HUFFMAN_CODE(<symbol 0..255>, "<bit string>", <hex code>, <length>)
 * with '#' lines and C comments skipped. The emitted coder keeps the callback contract (encode:
 * table load, num_bits == 0 for symbols the table lacks; decode: the unique code that prefixes
 * the 32-bit window, 0 for a hole) but decodes with a two-level lookup table instead of a
 * bit-at-a-time goto tree, and also exports <name>_get_code_table() so the batched CUDA context
 * can be fed without probing the callback.
 *
 * Stricter than the reference: duplicate symbols, symbols > 255, lengths > 32 and tables that are
 * not prefix codes are rejected with a message (the reference only asserts, or accepts silently);
 * a bit string that disagrees with the hex column is reported as a warning.
 */
#include "../host/huffman_lut.h"

#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { ROOT_BITS = 9, SUB_BITS = 8 };

static uint32_t g_patterns[256];
static uint8_t g_num_bits[256];
static uint8_t g_seen[256];

static char *s_read_file(const char *path, size_t *size) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        return NULL;
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *text = malloc((size_t)n + 1);
    if (!text || fread(text, 1, (size_t)n, f) != (size_t)n) {
        fclose(f);
        free(text);
        return NULL;
    }
    fclose(f);
    text[n] = '\0';
    *size = (size_t)n;
    return text;
}

/* Blank out C comments and preprocessor lines in place so the scanner below never sees them. */
static void s_strip(char *text, size_t size) {
    int in_comment = 0;
    int line_start = 1;
    for (size_t i = 0; i < size; ++i) {
        if (in_comment) {
            if (text[i] == '*' && i + 1 < size && text[i + 1] == '/') {
                text[i] = text[i + 1] = ' ';
                ++i;
                in_comment = 0;
            } else if (text[i] != '\n') {
                text[i] = ' ';
            }
            continue;
        }
        if (text[i] == '/' && i + 1 < size && text[i + 1] == '*') {
            text[i] = text[i + 1] = ' ';
            ++i;
            in_comment = 1;
            continue;
        }
        if (line_start && text[i] == '#') {
            while (i < size && text[i] != '\n') {
                text[i++] = ' ';
            }
        }
        line_start = (i < size && text[i] == '\n');
    }
}

static const char *s_skip_space(const char *p) {
    while (*p && isspace((unsigned char)*p)) {
        ++p;
    }
    return p;
}

static int s_line_of(const char *text, const char *p) {
    int line = 1;
    for (; text < p; ++text) {
        line += *text == '\n';
    }
    return line;
}

static int s_parse(const char *path, char *text) {
    static const char KEYWORD[] = "HUFFMAN_CODE";
    const char *p = text;
    int entries = 0;
    while ((p = strstr(p, KEYWORD)) != NULL) {
        const char *at = p;
        p += sizeof(KEYWORD) - 1;
        p = s_skip_space(p);
        if (*p != '(') {
            fprintf(stderr, "%s:%d: expected '(' after HUFFMAN_CODE\n", path, s_line_of(text, at));
            return 1;
        }
        char *end = NULL;
        errno = 0;
        const char *num_begin = p + 1;
        long symbol = strtol(num_begin, &end, 0);
        p = s_skip_space(end);
        if (end == num_begin || errno || symbol < 0 || symbol > 255 || *p != ',') {
            fprintf(stderr, "%s:%d: bad symbol (must be 0..255)\n", path, s_line_of(text, at));
            return 1;
        }
        p = s_skip_space(p + 1);
        if (*p != '"') {
            fprintf(stderr, "%s:%d: expected a quoted bit string\n", path, s_line_of(text, at));
            return 1;
        }
        const char *bits_begin = ++p;
        while (*p == '0' || *p == '1') {
            ++p;
        }
        const char *bits_end = p;
        if (*p != '"') {
            fprintf(stderr, "%s:%d: bit string may hold only 0 and 1\n", path, s_line_of(text, at));
            return 1;
        }
        p = s_skip_space(p + 1);
        if (*p != ',') {
            fprintf(stderr, "%s:%d: expected ',' after the bit string\n", path, s_line_of(text, at));
            return 1;
        }
        errno = 0;
        num_begin = p + 1;
        unsigned long code = strtoul(num_begin, &end, 16);
        if (end == num_begin || errno || code > 0xFFFFFFFFul) {
            fprintf(stderr, "%s:%d: bad hex code\n", path, s_line_of(text, at));
            return 1;
        }
        p = s_skip_space(end);
        if (*p != ',') {
            fprintf(stderr, "%s:%d: expected ',' after the hex code\n", path, s_line_of(text, at));
            return 1;
        }
        num_begin = p + 1;
        long length = strtol(num_begin, &end, 0);
        p = s_skip_space(end);
        if (end == num_begin || length < 1 || length > 32 || *p != ')') {
            fprintf(stderr, "%s:%d: bad length (must be 1..32)\n", path, s_line_of(text, at));
            return 1;
        }
        if (g_seen[symbol]) {
            fprintf(stderr, "%s:%d: symbol %ld defined twice\n", path, s_line_of(text, at), symbol);
            return 1;
        }
        if (length < 32 && (code >> length) != 0) {
            fprintf(stderr, "%s:%d: code 0x%lx does not fit in %ld bits\n", path, s_line_of(text, at), code, length);
            return 1;
        }
        /* cross-check the human-readable column; the hex column wins (as in the reference) */
        int agrees = (bits_end - bits_begin) == length;
        for (long i = 0; agrees && i < length; ++i) {
            agrees = ((code >> (length - 1 - i)) & 1ul) == (unsigned long)(bits_begin[i] - '0');
        }
        if (!agrees) {
            fprintf(
                stderr,
                "%s:%d: warning: bit string of symbol %ld disagrees with 0x%lx/%ld; using the hex column\n",
                path,
                s_line_of(text, at),
                symbol,
                code,
                length);
        }
        g_seen[symbol] = 1;
        g_patterns[symbol] = (uint32_t)code;
        g_num_bits[symbol] = (uint8_t)length;
        ++entries;
    }
    if (entries == 0) {
        fprintf(stderr, "%s: no HUFFMAN_CODE entries found\n", path);
        return 1;
    }
    return 0;
}

static void s_emit(FILE *out, const char *name, const struct huffman_lut *lut) {
    fprintf(
        out,
        "/* GENERATED by huffman_generator from a .def table -- do not edit. */\n"
        "/* clang-format off */\n"
        "\n"
        "#include <aws/compression/huffman.h>\n"
        "\n"
        "/* symbol -> code; num_bits == 0 marks a symbol the table does not define */\n"
        "static const struct aws_huffman_code s_%s_codes[256] = {\n",
        name);
    for (int sym = 0; sym < 256; ++sym) {
        fprintf(out, "    {0x%xu, %u},", g_patterns[sym], g_num_bits[sym]);
        if (isprint(sym) && sym != '\\' && sym != '/' && sym != '*') {
            fprintf(out, " /* %3d '%c' */\n", sym, sym);
        } else {
            fprintf(out, " /* %3d */\n", sym);
        }
    }
    fprintf(
        out,
        "};\n"
        "\n"
        "/* Decode lookup: %u root entries indexed by the top %u window bits, then sub-tables.\n"
        " * leaf = 0x80000000 | len << 8 | symbol; link = width << 24 | base; 0 = no such code. */\n"
        "static const uint32_t s_%s_lut[%u] = {",
        1u << lut->root_bits,
        lut->root_bits,
        name,
        lut->count);
    for (uint32_t i = 0; i < lut->count; ++i) {
        fprintf(out, "%s0x%08xu,", (i % 8 == 0) ? "\n    " : " ", lut->entries[i]);
    }
    fprintf(
        out,
        "\n};\n"
        "\n"
        "static struct aws_huffman_code encode_symbol(uint8_t symbol, void *userdata) {\n"
        "    (void)userdata;\n"
        "    return s_%s_codes[symbol];\n"
        "}\n"
        "\n"
        "static uint8_t decode_symbol(uint32_t bits, uint8_t *symbol, void *userdata) {\n"
        "    (void)userdata;\n"
        "    uint32_t entry = s_%s_lut[bits >> %u];\n"
        "    unsigned used = %u;\n"
        "    while (entry != 0 && (entry & 0x80000000u) == 0) {\n"
        "        const unsigned width = entry >> 24;\n"
        "        entry = s_%s_lut[(entry & 0xFFFFFFu) + ((uint32_t)(bits << used) >> (32 - width))];\n"
        "        used += width;\n"
        "    }\n"
        "    if (entry == 0) {\n"
        "        return 0; /* hole */\n"
        "    }\n"
        "    *symbol = (uint8_t)entry;\n"
        "    return (uint8_t)((entry >> 8) & 0x3Fu);\n"
        "}\n"
        "\n"
        "struct aws_huffman_symbol_coder *%s_get_coder(void) {\n"
        "    static struct aws_huffman_symbol_coder coder = {\n"
        "        .encode = encode_symbol,\n"
        "        .decode = decode_symbol,\n"
        "        .userdata = NULL,\n"
        "    };\n"
        "    return &coder;\n"
        "}\n"
        "\n"
        "/* The raw 256-entry code table, for callers that want it without going through encode(). */\n"
        "const struct aws_huffman_code *%s_get_code_table(void) {\n"
        "    return s_%s_codes;\n"
        "}\n",
        name,
        name,
        32 - lut->root_bits,
        lut->root_bits,
        name,
        name,
        name,
        name);
}

int main(int argc, char **argv) {
    if (argc != 4) {
        fprintf(
            stderr,
            "generator expects 3 arguments: [input file] [output file] [encoding name]\n"
            "A function of the following signature will be exported:\n"
            "struct aws_huffman_symbol_coder *[encoding name]_get_coder()\n");
        return 1;
    }
    const char *input_path = argv[1];
    const char *output_path = argv[2];
    const char *name = argv[3];

    for (const char *c = name; *c; ++c) {
        if (!(isalnum((unsigned char)*c) || *c == '_') || (c == name && isdigit((unsigned char)*c))) {
            fprintf(stderr, "encoding name '%s' is not a C identifier\n", name);
            return 1;
        }
    }

    size_t size = 0;
    char *text = s_read_file(input_path, &size);
    if (!text) {
        printf("Failed to open file '%s' for read.", input_path);
        return 1;
    }
    s_strip(text, size);
    if (s_parse(input_path, text)) {
        free(text);
        return 1;
    }
    free(text);

    struct huffman_lut lut;
    const int rc = huffman_lut_build(&lut, g_patterns, g_num_bits, ROOT_BITS, SUB_BITS);
    if (rc == HUFFMAN_LUT_ERR_NOT_PREFIX_FREE) {
        fprintf(stderr, "%s: not a prefix code (two codes collide or one starts another)\n", input_path);
        return 1;
    }
    if (rc != HUFFMAN_LUT_OK) {
        fprintf(stderr, "%s: could not build the decode table (error %d)\n", input_path, rc);
        return 1;
    }

    FILE *out = fopen(output_path, "w");
    if (!out) {
        printf("Failed to open file '%s' for write.", output_path);
        huffman_lut_clean_up(&lut);
        return 1;
    }
    s_emit(out, name, &lut);
    fclose(out);
    huffman_lut_clean_up(&lut);
    return 0;
}
