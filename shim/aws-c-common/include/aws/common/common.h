#ifndef AWS_COMMON_COMMON_H
#define AWS_COMMON_COMMON_H
/*
 * Minimal stand-in for the slice of aws-c-common that aws-c-compression touches.
 *
 * aws-c-common is an external, unpinned dependency of the reference
 * (/root/reference/CMakeLists.txt:6, builder.json:3-5) and is not present in this image.
 * This shim supplies only the surface listed in SURVEY.md Appendix C. When the real
 * aws-c-common is installed, drop this include root and link the real library instead:
 * struct layouts (aws_byte_buf, aws_byte_cursor) and numeric error codes follow the
 * real library so the ABI of libaws-c-compression stays the same.
 */

#include <assert.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
#    define AWS_EXTERN_C_BEGIN extern "C" {
#    define AWS_EXTERN_C_END }
#else
#    define AWS_EXTERN_C_BEGIN
#    define AWS_EXTERN_C_END
#endif

#define AWS_PUSH_SANE_WARNING_LEVEL
#define AWS_POP_SANE_WARNING_LEVEL

#define AWS_OP_SUCCESS (0)
#define AWS_OP_ERR (-1)

#ifdef NDEBUG
#    define AWS_ASSERT(cond) ((void)0)
#else
#    define AWS_ASSERT(cond) assert(cond)
#endif
#define AWS_PRECONDITION(cond) AWS_ASSERT(cond)
#define AWS_POSTCONDITION(cond) AWS_ASSERT(cond)
#define AWS_FATAL_ASSERT(cond)                                                                                         \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            abort();                                                                                                   \
        }                                                                                                              \
    } while (0)

#define AWS_ZERO_STRUCT(object) memset(&(object), 0, sizeof(object))
#define AWS_ZERO_ARRAY(array) memset((void *)(array), 0, sizeof(array))
#define AWS_ARRAY_SIZE(array) (sizeof(array) / sizeof((array)[0]))
#define AWS_VARIABLE_LENGTH_ARRAY(type, name, size) type name[(size) > 0 ? (size) : 1]

#define AWS_ERROR_ENUM_STRIDE_BITS 10
#define AWS_ERROR_ENUM_STRIDE (1U << AWS_ERROR_ENUM_STRIDE_BITS)
#define AWS_ERROR_ENUM_BEGIN_RANGE(x) ((x) * AWS_ERROR_ENUM_STRIDE)
#define AWS_ERROR_ENUM_END_RANGE(x) (((x) + 1) * AWS_ERROR_ENUM_STRIDE - 1)

#include <aws/common/allocator.h>
#include <aws/common/error.h>

AWS_EXTERN_C_BEGIN
void aws_common_library_init(struct aws_allocator *allocator);
void aws_common_library_clean_up(void);
AWS_EXTERN_C_END

#endif /* AWS_COMMON_COMMON_H */
