#ifndef AWS_COMMON_ERROR_H
#define AWS_COMMON_ERROR_H
/* Shim: see common.h. Thread-local last-error plumbing and the error-info registry. */

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Numeric values follow aws-c-common's enum aws_common_error (recalled, not verifiable here:
 * the library is absent from this image). Only SHORT_BUFFER is observable through the codec. */
enum aws_common_error {
    AWS_ERROR_SUCCESS = 0,
    AWS_ERROR_OOM = 1,
    AWS_ERROR_NO_SPACE = 2,
    AWS_ERROR_UNKNOWN = 3,
    AWS_ERROR_SHORT_BUFFER = 4,
    AWS_ERROR_OVERFLOW_DETECTED = 5,
    AWS_ERROR_UNSUPPORTED_OPERATION = 6,
    AWS_ERROR_INVALID_BUFFER_SIZE = 7,
    AWS_ERROR_INVALID_INDEX = 10,
    AWS_ERROR_INVALID_ARGUMENT = 34,
    AWS_ERROR_UNIMPLEMENTED = 37,
    AWS_ERROR_INVALID_STATE = 38,
};

struct aws_error_info {
    int error_code;
    const char *literal_name;
    const char *error_str;
    const char *lib_name;
    const char *formatted_name;
};

struct aws_error_info_list {
    const struct aws_error_info *error_list;
    uint16_t count;
};

#define AWS_DEFINE_ERROR_INFO(C, ES, LN)                                                                               \
    {                                                                                                                  \
        .literal_name = #C,                                                                                            \
        .error_code = (C),                                                                                             \
        .error_str = (ES),                                                                                             \
        .lib_name = (LN),                                                                                              \
        .formatted_name = LN ": " #C ", " ES,                                                                          \
    }

int aws_last_error(void);
void aws_reset_error(void);
/* Stores `err` in the calling thread's slot and returns AWS_OP_ERR. */
int aws_raise_error(int err);
const char *aws_error_str(int err);
const char *aws_error_name(int err);
const char *aws_error_lib_name(int err);
const char *aws_error_debug_str(int err);
void aws_register_error_info(const struct aws_error_info_list *error_info);
void aws_unregister_error_info(const struct aws_error_info_list *error_info);

#ifdef __cplusplus
}
#endif

#endif /* AWS_COMMON_ERROR_H */
