/* Minimal C client of the drop-in boundary: nothing but include/adpres_b200.h and the shared
 * library (what the Fortran ISO_C_BINDING shim links against).  Creates a context, reports the
 * library version and -- on a box without a GPU -- the loud failure of adp_create (there is no
 * CPU fallback).  Build:
 *   gcc -std=c99 -Wall -Iinclude examples/abi_probe.c -Ladpres_b200 -ladpres_b200 \
 *       -Wl,-rpath,$PWD/adpres_b200 -o /tmp/abi_probe                                  */
#include <stdio.h>
#include "adpres_b200.h"

int main(void)
{
    adp_ctx *ctx = NULL;
    int rc;
    printf("version: %s\n", adp_version());
    rc = adp_create(&ctx, 0);
    if (rc != ADP_OK) {
        printf("adp_create failed (%d): %s\n", rc, adp_last_error(NULL));
        return 3;
    }
    /* a context without geometry refuses to compute */
    rc = adp_matrix_setup(ctx, 1);
    printf("adp_matrix_setup on an empty context -> %d (%s)\n", rc, adp_last_error(ctx));
    adp_destroy(ctx);
    return rc < 0 ? 0 : 1;
}
