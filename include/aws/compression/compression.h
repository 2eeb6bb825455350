#ifndef AWS_COMPRESSION_COMPRESSION_H
#define AWS_COMPRESSION_COMPRESSION_H
/*
 * Library bootstrap and error codes. Replaces the reference's
 * include/aws/compression/compression.h:13-37 (same package id, same first error value).
 */
#include <aws/compression/exports.h>

#include <aws/common/common.h>

AWS_PUSH_SANE_WARNING_LEVEL

#define AWS_C_COMPRESSION_PACKAGE_ID 3

enum aws_compression_error {
    /* reference compression.h:17 — a symbol with no code (encode) or a bit window matching no
     * code with at least 32 bits left (decode) */
    AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL = AWS_ERROR_ENUM_BEGIN_RANGE(AWS_C_COMPRESSION_PACKAGE_ID),

    /* New in the B200 build: the batched (CUDA) entry points could not run — no device, a CUDA
     * runtime error, or device memory exhausted. Never degraded to a CPU path. */
    AWS_ERROR_COMPRESSION_DEVICE_FAILURE,

    /* New in the B200 build: the symbol coder handed to aws_huffman_batch_ctx_new is not a prefix
     * code (two codes collide, or a code is longer than 32 bits). */
    AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE,

    /* New in the B200 build (hpack_string_batch.h): a Huffman-coded HPACK string literal whose padding is
     * a byte or longer, is not all ones, or contains the EOS symbol (RFC 7541 section 5.2). */
    AWS_ERROR_COMPRESSION_INVALID_PADDING,

    AWS_ERROR_END_COMPRESSION_RANGE = AWS_ERROR_ENUM_END_RANGE(AWS_C_COMPRESSION_PACKAGE_ID)
};

AWS_EXTERN_C_BEGIN

/* Registers this library's error strings (idempotent). reference compression.h:30 */
AWS_COMPRESSION_API
void aws_compression_library_init(struct aws_allocator *alloc);

/* Undoes aws_compression_library_init (idempotent). reference compression.h:37 */
AWS_COMPRESSION_API
void aws_compression_library_clean_up(void);

AWS_EXTERN_C_END
AWS_POP_SANE_WARNING_LEVEL

#endif /* AWS_COMPRESSION_COMPRESSION_H */
