#ifndef AWS_COMPRESSION_HUFFMAN_H
#define AWS_COMPRESSION_HUFFMAN_H
/*
 * Streaming Huffman codec: the C surface of awslabs/aws-c-compression, kept as a drop-in.
 * Struct layouts and prototypes match the reference's include/aws/compression/huffman.h
 * (x86-64: aws_huffman_code 8 B, symbol_coder 24 B, encoder 24 B, decoder 32 B); the
 * implementation (aws-c-compression_b200/host/huffman.c) is written from scratch.
 * The batched CUDA entry points live in <aws/compression/huffman_batch.h>.
 */
#include <aws/compression/compression.h>

#include <aws/common/byte_buf.h>

AWS_PUSH_SANE_WARNING_LEVEL

/* One code word. reference huffman.h:18-26 */
struct aws_huffman_code {
    uint32_t pattern; /* right-aligned: the code is the low num_bits bits */
    uint8_t num_bits; /* 1..32; 0 means "no code for this symbol" */
};

/* symbol -> code; return num_bits == 0 for a symbol the table does not know. reference huffman.h:37 */
typedef struct aws_huffman_code(aws_huffman_symbol_encoder_fn)(uint8_t symbol, void *userdata);

/* `bits` holds the next 32 stream bits, first bit in bit 31. Write the matched symbol and return
 * the code length; return 0 (and leave *symbol alone) when no code matches. reference huffman.h:48 */
typedef uint8_t(aws_huffman_symbol_decoder_fn)(uint32_t bits, uint8_t *symbol, void *userdata);

/* The table plugin. reference huffman.h:53-57 */
struct aws_huffman_symbol_coder {
    aws_huffman_symbol_encoder_fn *encode;
    aws_huffman_symbol_decoder_fn *decode;
    void *userdata;
};

/* Encoder parameters + resume state; caller-owned POD. reference huffman.h:63-70 */
struct aws_huffman_encoder {
    struct aws_huffman_symbol_coder *coder;
    uint8_t eos_padding; /* its LOW bits fill the last byte; 0xFF after init */

    struct aws_huffman_code overflow_bits; /* tail of the code that straddled a full output */
};

/* Decoder parameters + resume state; caller-owned POD. reference huffman.h:76-84 */
struct aws_huffman_decoder {
    struct aws_huffman_symbol_coder *coder;
    bool allow_growth;

    uint64_t working_bits; /* unread stream bits, left-aligned */
    uint8_t num_bits;      /* how many of them are valid */
};

AWS_EXTERN_C_BEGIN

/* reference huffman.h:92 */
AWS_COMPRESSION_API
void aws_huffman_encoder_init(struct aws_huffman_encoder *encoder, struct aws_huffman_symbol_coder *coder);

/* Forget pending overflow bits before starting a new stream. reference huffman.h:98 */
AWS_COMPRESSION_API
void aws_huffman_encoder_reset(struct aws_huffman_encoder *encoder);

/* reference huffman.h:104 */
AWS_COMPRESSION_API
void aws_huffman_decoder_init(struct aws_huffman_decoder *decoder, struct aws_huffman_symbol_coder *coder);

/* Forget buffered bits before starting a new stream. reference huffman.h:110 */
AWS_COMPRESSION_API
void aws_huffman_decoder_reset(struct aws_huffman_decoder *decoder);

/* ceil(sum of code lengths / 8); unknown symbols count 0 bits. reference huffman.h:121 */
AWS_COMPRESSION_API
size_t aws_huffman_get_encoded_length(struct aws_huffman_encoder *encoder, struct aws_byte_cursor to_encode);

/*
 * Packs the codes of to_encode MSB-first into output, advancing both. When every symbol has been
 * placed the last byte is padded with eos_padding and AWS_OP_SUCCESS is returned. When output
 * fills first: AWS_OP_ERR / AWS_ERROR_SHORT_BUFFER, output holds a byte-exact prefix and the
 * call may be repeated with more room. A symbol without a code: AWS_OP_ERR /
 * AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL. reference huffman.h:133-136
 */
AWS_COMPRESSION_API
int aws_huffman_encode(
    struct aws_huffman_encoder *encoder,
    struct aws_byte_cursor *to_encode,
    struct aws_byte_buf *output);

/*
 * Decodes symbols until the input (plus bits buffered by earlier calls) is used up. Input may be
 * split anywhere across calls. Full output: grows it (doubling) when allow_growth, otherwise
 * AWS_OP_ERR / AWS_ERROR_SHORT_BUFFER. A 32-bit window that matches no code:
 * AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL. reference huffman.h:149-152
 */
AWS_COMPRESSION_API
int aws_huffman_decode(
    struct aws_huffman_decoder *decoder,
    struct aws_byte_cursor *to_decode,
    struct aws_byte_buf *output);

/* Off by default. reference huffman.h:159 */
AWS_COMPRESSION_API
void aws_huffman_decoder_allow_growth(struct aws_huffman_decoder *decoder, bool allow_growth);

AWS_EXTERN_C_END
AWS_POP_SANE_WARNING_LEVEL

#endif /* AWS_COMPRESSION_HUFFMAN_H */
