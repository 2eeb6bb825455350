/*
 * Out-of-line part of the aws-c-common stand-in (see include/aws/common/common.h).
 * Error slot, error-info registry, default allocator and byte_buf growth.
 */
#include <aws/common/byte_buf.h>
#include <aws/common/common.h>

static _Thread_local int tl_last_error = 0;

int aws_last_error(void) {
    return tl_last_error;
}

void aws_reset_error(void) {
    tl_last_error = 0;
}

int aws_raise_error(int err) {
    tl_last_error = err;
    return AWS_OP_ERR;
}

/* Registry: one slot per 1024-wide package range, as in aws-c-common. */
enum { SHIM_MAX_PACKAGES = 32 };
static const struct aws_error_info_list *s_registry[SHIM_MAX_PACKAGES];

static const struct aws_error_info s_common_errors[] = {
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_SUCCESS, "Success.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_OOM, "Out of memory.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_NO_SPACE, "Out of space on disk.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_UNKNOWN, "Unknown error.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_SHORT_BUFFER, "Buffer is not large enough to hold result.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_OVERFLOW_DETECTED, "Fixed size value overflow was detected.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_UNSUPPORTED_OPERATION, "Unsupported operation.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_INVALID_BUFFER_SIZE, "Invalid buffer size.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_INVALID_INDEX, "Invalid index for list access.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_INVALID_ARGUMENT, "An invalid argument was passed to a function.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_UNIMPLEMENTED, "A function was called, but is not implemented.", "aws-c-common"),
    AWS_DEFINE_ERROR_INFO(AWS_ERROR_INVALID_STATE, "An invalid state was encountered.", "aws-c-common"),
};

static const struct aws_error_info *s_find(int err) {
    if (err < 0) {
        return NULL;
    }
    if (err < (int)AWS_ERROR_ENUM_STRIDE) {
        for (size_t i = 0; i < AWS_ARRAY_SIZE(s_common_errors); ++i) {
            if (s_common_errors[i].error_code == err) {
                return &s_common_errors[i];
            }
        }
        return NULL;
    }
    const size_t slot = (size_t)err >> AWS_ERROR_ENUM_STRIDE_BITS;
    const size_t index = (size_t)err & (AWS_ERROR_ENUM_STRIDE - 1);
    if (slot >= SHIM_MAX_PACKAGES || s_registry[slot] == NULL || index >= s_registry[slot]->count) {
        return NULL;
    }
    return &s_registry[slot]->error_list[index];
}

const char *aws_error_str(int err) {
    const struct aws_error_info *info = s_find(err);
    return info ? info->error_str : "Unknown Error Code";
}

const char *aws_error_name(int err) {
    const struct aws_error_info *info = s_find(err);
    return info ? info->literal_name : "Unknown Error Code";
}

const char *aws_error_lib_name(int err) {
    const struct aws_error_info *info = s_find(err);
    return info ? info->lib_name : "Unknown Error Code";
}

const char *aws_error_debug_str(int err) {
    const struct aws_error_info *info = s_find(err);
    return info ? info->formatted_name : "Unknown Error Code";
}

void aws_register_error_info(const struct aws_error_info_list *error_info) {
    AWS_FATAL_ASSERT(error_info && error_info->error_list && error_info->count);
    const size_t slot = (size_t)error_info->error_list[0].error_code >> AWS_ERROR_ENUM_STRIDE_BITS;
    AWS_FATAL_ASSERT(slot < SHIM_MAX_PACKAGES);
    s_registry[slot] = error_info;
}

void aws_unregister_error_info(const struct aws_error_info_list *error_info) {
    AWS_FATAL_ASSERT(error_info && error_info->error_list && error_info->count);
    const size_t slot = (size_t)error_info->error_list[0].error_code >> AWS_ERROR_ENUM_STRIDE_BITS;
    AWS_FATAL_ASSERT(slot < SHIM_MAX_PACKAGES);
    s_registry[slot] = NULL;
}

void aws_common_library_init(struct aws_allocator *allocator) {
    (void)allocator;
}

void aws_common_library_clean_up(void) {}

/* ---- allocator ---- */

static void *s_default_acquire(struct aws_allocator *allocator, size_t size) {
    (void)allocator;
    return malloc(size);
}

static void s_default_release(struct aws_allocator *allocator, void *ptr) {
    (void)allocator;
    free(ptr);
}

static void *s_default_realloc(struct aws_allocator *allocator, void *oldptr, size_t oldsize, size_t newsize) {
    (void)allocator;
    (void)oldsize;
    return realloc(oldptr, newsize);
}

static void *s_default_calloc(struct aws_allocator *allocator, size_t num, size_t size) {
    (void)allocator;
    return calloc(num, size);
}

static struct aws_allocator s_default_allocator = {
    .mem_acquire = s_default_acquire,
    .mem_release = s_default_release,
    .mem_realloc = s_default_realloc,
    .mem_calloc = s_default_calloc,
    .impl = NULL,
};

struct aws_allocator *aws_default_allocator(void) {
    return &s_default_allocator;
}

void *aws_mem_acquire(struct aws_allocator *allocator, size_t size) {
    void *mem = allocator->mem_acquire(allocator, size);
    if (mem == NULL) {
        aws_raise_error(AWS_ERROR_OOM);
    }
    return mem;
}

void *aws_mem_calloc(struct aws_allocator *allocator, size_t num, size_t size) {
    void *mem = allocator->mem_calloc ? allocator->mem_calloc(allocator, num, size)
                                      : allocator->mem_acquire(allocator, num * size);
    if (mem == NULL) {
        aws_raise_error(AWS_ERROR_OOM);
        return NULL;
    }
    if (!allocator->mem_calloc) {
        memset(mem, 0, num * size);
    }
    return mem;
}

void aws_mem_release(struct aws_allocator *allocator, void *ptr) {
    if (ptr != NULL) {
        allocator->mem_release(allocator, ptr);
    }
}

int aws_mem_realloc(struct aws_allocator *allocator, void **ptr, size_t oldsize, size_t newsize) {
    if (newsize == 0) {
        aws_mem_release(allocator, *ptr);
        *ptr = NULL;
        return AWS_OP_SUCCESS;
    }
    void *grown = NULL;
    if (allocator->mem_realloc) {
        grown = allocator->mem_realloc(allocator, *ptr, oldsize, newsize);
    } else {
        grown = allocator->mem_acquire(allocator, newsize);
        if (grown != NULL && *ptr != NULL) {
            memcpy(grown, *ptr, oldsize < newsize ? oldsize : newsize);
            allocator->mem_release(allocator, *ptr);
        }
    }
    if (grown == NULL) {
        return aws_raise_error(AWS_ERROR_OOM);
    }
    *ptr = grown;
    return AWS_OP_SUCCESS;
}

/* ---- byte_buf growth ---- */

int aws_byte_buf_init(struct aws_byte_buf *buf, struct aws_allocator *allocator, size_t capacity) {
    buf->buffer = (capacity == 0) ? NULL : (uint8_t *)aws_mem_acquire(allocator, capacity);
    if (capacity != 0 && buf->buffer == NULL) {
        AWS_ZERO_STRUCT(*buf);
        return AWS_OP_ERR;
    }
    buf->len = 0;
    buf->capacity = capacity;
    buf->allocator = allocator;
    return AWS_OP_SUCCESS;
}

void aws_byte_buf_clean_up(struct aws_byte_buf *buf) {
    if (buf->allocator && buf->buffer) {
        aws_mem_release(buf->allocator, buf->buffer);
    }
    buf->allocator = NULL;
    buf->buffer = NULL;
    buf->len = 0;
    buf->capacity = 0;
}

int aws_byte_buf_reserve(struct aws_byte_buf *buf, size_t requested_capacity) {
    if (buf->allocator == NULL || !aws_byte_buf_is_valid(buf)) {
        return aws_raise_error(AWS_ERROR_INVALID_ARGUMENT);
    }
    if (requested_capacity <= buf->capacity) {
        return AWS_OP_SUCCESS;
    }
    if (buf->buffer == NULL && buf->capacity == 0) {
        return aws_byte_buf_init(buf, buf->allocator, requested_capacity);
    }
    if (aws_mem_realloc(buf->allocator, (void **)&buf->buffer, buf->capacity, requested_capacity)) {
        return AWS_OP_ERR;
    }
    buf->capacity = requested_capacity;
    return AWS_OP_SUCCESS;
}

int aws_byte_buf_reserve_relative(struct aws_byte_buf *buf, size_t additional_length) {
    if (additional_length > SIZE_MAX - buf->len) {
        return aws_raise_error(AWS_ERROR_OVERFLOW_DETECTED);
    }
    return aws_byte_buf_reserve(buf, buf->len + additional_length);
}
