#ifndef ORACLE_REF_HARNESS_AWS_TEST_HARNESS_H
#define ORACLE_REF_HARNESS_AWS_TEST_HARNESS_H
/*
 * TEST INFRASTRUCTURE (oracle/): a stand-in for aws-c-common's <aws/testing/aws_test_harness.h>,
 * just enough to compile the reference's own tests/huffman_test.c and tests/library_test.c in
 * place (from /root/reference) into oracle/_ref/. Not part of the product.
 *
 * Each AWS_TEST_CASE registers itself in a linked list at load time; ref_test_main.c walks it.
 */
#include <aws/common/byte_buf.h>
#include <aws/common/common.h>
#include <stdio.h>

struct oracle_ref_test_case {
    const char *name;
    int (*fn)(struct aws_allocator *allocator, void *ctx);
    struct oracle_ref_test_case *next;
};
extern struct oracle_ref_test_case *g_oracle_ref_tests;

#define AWS_TEST_CASE(tname, tfn)                                                                                      \
    static int tfn(struct aws_allocator *allocator, void *ctx);                                                        \
    static struct oracle_ref_test_case s_case_##tname = {#tname, tfn, NULL};                                           \
    __attribute__((constructor)) static void s_register_##tname(void) {                                                \
        s_case_##tname.next = g_oracle_ref_tests;                                                                      \
        g_oracle_ref_tests = &s_case_##tname;                                                                          \
    }

#define ORACLE_FAIL_(what)                                                                                             \
    do {                                                                                                               \
        fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, what);                                                 \
        return AWS_OP_ERR;                                                                                             \
    } while (0)

#define ASSERT_SUCCESS(expr, ...)                                                                                      \
    do {                                                                                                               \
        if ((expr) != AWS_OP_SUCCESS)                                                                                  \
            ORACLE_FAIL_("ASSERT_SUCCESS(" #expr ")");                                                                 \
    } while (0)
#define ASSERT_TRUE(expr, ...)                                                                                         \
    do {                                                                                                               \
        if (!(expr))                                                                                                   \
            ORACLE_FAIL_("ASSERT_TRUE(" #expr ")");                                                                    \
    } while (0)
#define ASSERT_FALSE(expr, ...)                                                                                        \
    do {                                                                                                               \
        if (expr)                                                                                                      \
            ORACLE_FAIL_("ASSERT_FALSE(" #expr ")");                                                                   \
    } while (0)
#define ASSERT_UINT_EQUALS(expected, got, ...)                                                                         \
    do {                                                                                                               \
        if ((unsigned long long)(expected) != (unsigned long long)(got))                                               \
            ORACLE_FAIL_("ASSERT_UINT_EQUALS(" #expected ", " #got ")");                                               \
    } while (0)
#define ASSERT_INT_EQUALS(expected, got, ...)                                                                          \
    do {                                                                                                               \
        if ((long long)(expected) != (long long)(got))                                                                 \
            ORACLE_FAIL_("ASSERT_INT_EQUALS(" #expected ", " #got ")");                                                \
    } while (0)
#define ASSERT_STR_EQUALS(expected, got, ...)                                                                          \
    do {                                                                                                               \
        if (strcmp((expected), (got)) != 0)                                                                            \
            ORACLE_FAIL_("ASSERT_STR_EQUALS(" #expected ", " #got ")");                                                \
    } while (0)
#define ASSERT_BIN_ARRAYS_EQUALS(expected, expected_size, got, got_size, ...)                                          \
    do {                                                                                                               \
        if ((size_t)(expected_size) != (size_t)(got_size) ||                                                           \
            ((expected_size) != 0 && memcmp((expected), (got), (expected_size)) != 0))                                 \
            ORACLE_FAIL_("ASSERT_BIN_ARRAYS_EQUALS(" #expected ", " #got ")");                                         \
    } while (0)

#endif
