/* The C ABI must be usable from plain C (cgo, JNI and N-API stubs compile it as C): include the header in a C99 translation
 * unit with -pedantic -Werror, take the address of a few entry points, and try to create a context.
 * Exit code: 0 = context created (a GPU is present), 3 = creation failed loudly with a message (no GPU), anything else = bug. */
#include <stdio.h>
#include <string.h>

#include "lbfgs_b200.h"

int main(void)
{
    lbfgs_b200_ctx* ctx = NULL;
    lbfgs_b200_status (*upd)(lbfgs_b200_hist*, const double*, const double*, const double*, const double*, double, double, double*,
                             int, int*, double*) = lbfgs_b200_hist_update_apply_Hv_f64;
    lbfgs_b200_status st;
    if (upd == NULL) return 2;
    printf("%s\n", lbfgs_b200_version());
    st = lbfgs_b200_ctx_create(&ctx, 0, NULL);
    if (st == LBFGS_B200_OK)
    {
        lbfgs_b200_ctx_destroy(ctx);
        return 0;
    }
    if (ctx != NULL) return 4;
    if (strlen(lbfgs_b200_last_error(NULL)) == 0) return 5;
    printf("no device: %s\n", lbfgs_b200_last_error(NULL));
    return 3;
}
