#ifndef AWS_COMMON_BYTE_BUF_H
#define AWS_COMMON_BYTE_BUF_H
/* Shim: see common.h. Struct layouts follow aws-c-common so the
 * aws-c-compression ABI is unchanged when the real library replaces this header. */

#include <aws/common/common.h>

#ifdef __cplusplus
extern "C" {
#endif

struct aws_byte_buf {
    size_t len;
    uint8_t *buffer;
    size_t capacity;
    struct aws_allocator *allocator;
};

struct aws_byte_cursor {
    size_t len;
    uint8_t *ptr;
};

static inline struct aws_byte_buf aws_byte_buf_from_empty_array(const void *bytes, size_t capacity) {
    struct aws_byte_buf buf;
    buf.buffer = (capacity > 0) ? (uint8_t *)bytes : NULL;
    buf.len = 0;
    buf.capacity = capacity;
    buf.allocator = NULL;
    return buf;
}

static inline struct aws_byte_buf aws_byte_buf_from_array(const void *bytes, size_t len) {
    struct aws_byte_buf buf;
    buf.buffer = (len > 0) ? (uint8_t *)bytes : NULL;
    buf.len = len;
    buf.capacity = len;
    buf.allocator = NULL;
    return buf;
}

static inline struct aws_byte_cursor aws_byte_cursor_from_array(const void *bytes, size_t len) {
    struct aws_byte_cursor cur;
    cur.ptr = (len > 0) ? (uint8_t *)bytes : NULL;
    cur.len = len;
    return cur;
}

static inline struct aws_byte_cursor aws_byte_cursor_from_buf(const struct aws_byte_buf *buf) {
    struct aws_byte_cursor cur;
    cur.ptr = buf->buffer;
    cur.len = buf->len;
    return cur;
}

static inline struct aws_byte_cursor aws_byte_cursor_from_c_str(const char *c_str) {
    return aws_byte_cursor_from_array(c_str, c_str ? strlen(c_str) : 0);
}

static inline bool aws_byte_cursor_is_valid(const struct aws_byte_cursor *cursor) {
    return cursor != NULL && (cursor->len == 0 || cursor->ptr != NULL);
}

static inline bool aws_byte_buf_is_valid(const struct aws_byte_buf *buf) {
    return buf != NULL && buf->len <= buf->capacity && (buf->capacity == 0 || buf->buffer != NULL);
}

/* Splits the first `len` bytes off the cursor; an empty cursor comes back when len is too large. */
static inline struct aws_byte_cursor aws_byte_cursor_advance(struct aws_byte_cursor *cursor, size_t len) {
    struct aws_byte_cursor head;
    if (len > cursor->len || cursor->len > (SIZE_MAX >> 1) || len > (SIZE_MAX >> 1)) {
        head.ptr = NULL;
        head.len = 0;
        return head;
    }
    head.ptr = cursor->ptr;
    head.len = len;
    cursor->ptr = (cursor->ptr == NULL) ? NULL : cursor->ptr + len;
    cursor->len -= len;
    return head;
}

static inline bool aws_byte_cursor_read_u8(struct aws_byte_cursor *cursor, uint8_t *var) {
    if (cursor->len == 0) {
        return false;
    }
    *var = *cursor->ptr;
    ++cursor->ptr;
    --cursor->len;
    return true;
}

static inline bool aws_byte_buf_write_u8(struct aws_byte_buf *buf, uint8_t c) {
    if (buf->len >= buf->capacity) {
        return false;
    }
    buf->buffer[buf->len++] = c;
    return true;
}

static inline bool aws_byte_buf_write(struct aws_byte_buf *buf, const uint8_t *src, size_t len) {
    if (len > buf->capacity - buf->len) {
        return false;
    }
    if (len > 0) {
        memcpy(buf->buffer + buf->len, src, len);
        buf->len += len;
    }
    return true;
}

static inline void aws_byte_buf_reset(struct aws_byte_buf *buf, bool zero_contents) {
    if (zero_contents && buf->buffer != NULL) {
        memset(buf->buffer, 0, buf->capacity);
    }
    buf->len = 0;
}

int aws_byte_buf_init(struct aws_byte_buf *buf, struct aws_allocator *allocator, size_t capacity);
void aws_byte_buf_clean_up(struct aws_byte_buf *buf);
/* Grows capacity to at least `requested_capacity` through the buffer's allocator. */
int aws_byte_buf_reserve(struct aws_byte_buf *buf, size_t requested_capacity);
/* Ensures `additional_length` free bytes after buf->len. */
int aws_byte_buf_reserve_relative(struct aws_byte_buf *buf, size_t additional_length);

#ifdef __cplusplus
}
#endif

#endif /* AWS_COMMON_BYTE_BUF_H */
