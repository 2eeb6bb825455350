/*
 * Library bootstrap: registers the error strings of this package with aws-c-common's error
 * registry. Replaces the reference's source/compression.c:13-44 (plus the two error codes the
 * B200 build adds for the batched path).
 */
#include <aws/compression/compression.h>

#define COMPRESSION_ERROR_SLOT(CODE) ((CODE) - AWS_ERROR_ENUM_BEGIN_RANGE(AWS_C_COMPRESSION_PACKAGE_ID))

static struct aws_error_info s_error_table[] = {
    [COMPRESSION_ERROR_SLOT(AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL)] = AWS_DEFINE_ERROR_INFO(
        AWS_ERROR_COMPRESSION_UNKNOWN_SYMBOL,
        "Compression encountered an unknown symbol.",
        "aws-c-compression"),
    [COMPRESSION_ERROR_SLOT(AWS_ERROR_COMPRESSION_DEVICE_FAILURE)] = AWS_DEFINE_ERROR_INFO(
        AWS_ERROR_COMPRESSION_DEVICE_FAILURE,
        "The CUDA device needed by the batched codec is missing or reported an error.",
        "aws-c-compression"),
    [COMPRESSION_ERROR_SLOT(AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE)] = AWS_DEFINE_ERROR_INFO(
        AWS_ERROR_COMPRESSION_INVALID_CODE_TABLE,
        "The symbol coder does not describe a prefix code of at most 32 bits.",
        "aws-c-compression"),
    [COMPRESSION_ERROR_SLOT(AWS_ERROR_COMPRESSION_INVALID_PADDING)] = AWS_DEFINE_ERROR_INFO(
        AWS_ERROR_COMPRESSION_INVALID_PADDING,
        "A Huffman-coded HPACK string literal ends in invalid padding (RFC 7541 section 5.2).",
        "aws-c-compression"),
};

static struct aws_error_info_list s_error_info = {
    .error_list = s_error_table,
    .count = AWS_ARRAY_SIZE(s_error_table),
};

static bool s_registered = false;

void aws_compression_library_init(struct aws_allocator *alloc) {
    if (!s_registered) {
        s_registered = true;
        aws_common_library_init(alloc);
        aws_register_error_info(&s_error_info);
    }
}

void aws_compression_library_clean_up(void) {
    if (s_registered) {
        s_registered = false;
        aws_unregister_error_info(&s_error_info);
        aws_common_library_clean_up();
    }
}
