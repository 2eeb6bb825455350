/*
 * TEST INFRASTRUCTURE (oracle/): runs every AWS_TEST_CASE the reference's own test files
 * registered (see ref_harness/aws/testing/aws_test_harness.h). Built only into oracle/_ref/.
 */
#include <aws/testing/aws_test_harness.h>

struct oracle_ref_test_case *g_oracle_ref_tests = NULL;

int main(int argc, char **argv) {
    int failures = 0, ran = 0;
    for (struct oracle_ref_test_case *t = g_oracle_ref_tests; t; t = t->next) {
        if (argc > 1 && strcmp(argv[1], t->name) != 0) {
            continue;
        }
        aws_reset_error();
        int rc = t->fn(aws_default_allocator(), NULL);
        printf("%-48s %s\n", t->name, rc == AWS_OP_SUCCESS ? "ok" : "FAILED");
        failures += rc != AWS_OP_SUCCESS;
        ++ran;
    }
    printf("%d ran, %d failed\n", ran, failures);
    return failures ? 1 : (ran ? 0 : 2);
}
