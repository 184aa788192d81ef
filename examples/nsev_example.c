/*
 * BASELINE config 1 through the drop-in library: the scenario of the reference's
 * examples/fnft_nsev_example.c (rectangular pulse q = 2 on [-1, 1], 256 samples, 8 points of the
 * reflection coefficient on [-2, 2], bound states and norming constants with the DEFAULT options,
 * i.e. subsample-and-refine localization) written against the same public API.  The reference's
 * own example file builds against include/ unchanged as well.
 *
 *   gcc -std=c99 -Iinclude examples/nsev_example.c -Lfnft_b200/lib -lfnft_b200 \
 *       -Wl,-rpath,$PWD/fnft_b200/lib -lm -o nsev_example
 */
#include <stdio.h>
#include <stdlib.h>
#include "fnft_nsev.h"

int main(void)
{
    const FNFT_UINT D = 256, M = 8;
    FNFT_REAL T[2] = {-1.0, 1.0}, XI[2] = {-2.0, 2.0};
    FNFT_COMPLEX q[256], contspec[8];
    for (FNFT_UINT i = 0; i < D; i++)
        q[i] = 2.0;
    fnft_nsev_opts_t opts = fnft_nsev_default_opts();
    FNFT_UINT K = fnft_nsev_max_K(D, &opts); /* size of the arrays below, then number found */
    FNFT_COMPLEX *bound_states = malloc(K * sizeof(FNFT_COMPLEX));
    FNFT_COMPLEX *normconsts = malloc(K * sizeof(FNFT_COMPLEX));
    if (bound_states == NULL || normconsts == NULL)
        return EXIT_FAILURE;
    const FNFT_INT rc = fnft_nsev(D, q, T, M, contspec, XI, &K, bound_states, normconsts, +1, &opts);
    if (rc != FNFT_SUCCESS) {
        printf("fnft_nsev failed: %d\n", rc);
        return EXIT_FAILURE;
    }
    const FNFT_REAL eps_xi = (XI[1] - XI[0]) / (M - 1);
    printf("continuous spectrum\n");
    for (FNFT_UINT i = 0; i < M; i++)
        printf("  xi = %+.6f  rho = %+.9e %+.9ei\n", XI[0] + i * eps_xi, creal(contspec[i]), cimag(contspec[i]));
    printf("discrete spectrum: K = %zu\n", K);
    for (FNFT_UINT i = 0; i < K; i++)
        printf("  lambda = %+.9e %+.9ei  b = %+.9e %+.9ei\n", creal(bound_states[i]), cimag(bound_states[i]),
               creal(normconsts[i]), cimag(normconsts[i]));
    free(bound_states);
    free(normconsts);
    return EXIT_SUCCESS;
}
