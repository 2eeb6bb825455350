#ifndef AWS_COMMON_ALLOCATOR_H
#define AWS_COMMON_ALLOCATOR_H
/* Shim: see common.h. Only what byte_buf growth (aws_huffman_decoder_allow_growth) needs. */

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

struct aws_allocator {
    void *(*mem_acquire)(struct aws_allocator *allocator, size_t size);
    void (*mem_release)(struct aws_allocator *allocator, void *ptr);
    void *(*mem_realloc)(struct aws_allocator *allocator, void *oldptr, size_t oldsize, size_t newsize);
    void *(*mem_calloc)(struct aws_allocator *allocator, size_t num, size_t size);
    void *impl;
};

struct aws_allocator *aws_default_allocator(void);
void *aws_mem_acquire(struct aws_allocator *allocator, size_t size);
void *aws_mem_calloc(struct aws_allocator *allocator, size_t num, size_t size);
void aws_mem_release(struct aws_allocator *allocator, void *ptr);
int aws_mem_realloc(struct aws_allocator *allocator, void **ptr, size_t oldsize, size_t newsize);

#ifdef __cplusplus
}
#endif

#endif /* AWS_COMMON_ALLOCATOR_H */
