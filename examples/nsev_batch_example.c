/*
 * Minimal C caller of the drop-in library: one classic fnft_nsev call (same source as a
 * program written against FNFT's own headers) and one batched call.
 *
 *   gcc -std=c99 -Iinclude examples/nsev_batch_example.c -Lfnft_b200/lib -lfnft_b200 \
 *       -Wl,-rpath,$PWD/fnft_b200/lib -lm -o nsev_batch_example
 */
#include <stdio.h>
#include <stdlib.h>
#include "fnft_nsev.h"

int main(void)
{
    const FNFT_UINT D = 1024, M = 16, B = 4;
    FNFT_REAL T[2] = {-16.0, 16.0}, XI[2] = {-2.0, 2.0};
    FNFT_COMPLEX *q = malloc(B * D * sizeof(FNFT_COMPLEX));
    FNFT_COMPLEX *cs = malloc(B * M * sizeof(FNFT_COMPLEX));
    FNFT_INT rcs[4];
    for (FNFT_UINT b = 0; b < B; b++)
        for (FNFT_UINT i = 0; i < D; i++) {
            const FNFT_REAL t = T[0] + i * (T[1] - T[0]) / (D - 1);
            q[b * D + i] = (0.8 + 0.3 * b) / cosh(t) * cexp(0.5 * I * t);
        }
    fnft_nsev_opts_t opts = fnft_nsev_default_opts();
    /* classic single-signal entry point, continuous spectrum only */
    FNFT_INT rc = fnft_nsev(D, q, T, M, cs, XI, NULL, NULL, NULL, +1, &opts);
    if (rc != FNFT_SUCCESS) {
        printf("fnft_nsev failed: %d\n", rc);
        return EXIT_FAILURE;
    }
    printf("single: rho(xi0) = %+.6e %+.6ei\n", creal(cs[0]), cimag(cs[0]));
    /* batched entry point */
    rc = fnft_nsev_batch(B, D, q, T, M, cs, XI, NULL, 0, NULL, NULL, +1, &opts, rcs);
    if (rc != FNFT_SUCCESS) {
        printf("fnft_nsev_batch failed: %d\n", rc);
        return EXIT_FAILURE;
    }
    for (FNFT_UINT b = 0; b < B; b++)
        printf("batch %zu: rho(xi0) = %+.6e %+.6ei (rc %d)\n", b, creal(cs[b * M]), cimag(cs[b * M]),
               rcs[b]);
    free(q);
    free(cs);
    return EXIT_SUCCESS;
}
