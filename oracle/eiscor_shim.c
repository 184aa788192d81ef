/*
 * ORACLE-ONLY TOOLING (never linked into the product library).
 *
 * Stand-in for the Fortran entry point z_poly_roots_modified_ of the eiscor
 * package that the reference calls from
 * src/private/fnft__poly_roots_fasteigen.c:26-42.  No Fortran compiler exists
 * in this image, so the root finder is replaced by a dense companion-matrix
 * eigenvalue solve (LAPACK zhseqr) looked up at run time in the OpenBLAS copy
 * that ships inside scipy.  If that library cannot be found the shim reports
 * failure (info = 1), which the reference turns into an error return; every
 * code path of the hot path proper (continuous spectrum, Newton refinement,
 * kdvv, nsep grid search) never reaches this function.
 */
#include <complex.h>
#include <dlfcn.h>
#include <glob.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int (*zhseqr_fn)(int layout, char job, char compz, int n, int ilo,
                         int ihi, double complex *h, int ldh,
                         double complex *w, double complex *z, int ldz);

static zhseqr_fn find_zhseqr(void)
{
    static zhseqr_fn cached = NULL;
    static int tried = 0;
    if (tried)
        return cached;
    tried = 1;
    const char *env = getenv("FNFT_ORACLE_OPENBLAS");
    const char *patterns[] = {
        env ? env : "",
        "/opt/prime-rl/.venv/lib/python3.12/site-packages/scipy.libs/libscipy_openblas-*.so",
        "/opt/prime-rl/.venv/lib/python3*/site-packages/scipy.libs/libscipy_openblas*.so",
    };
    for (size_t k = 0; k < sizeof(patterns) / sizeof(patterns[0]); k++) {
        if (patterns[k][0] == '\0')
            continue;
        glob_t g;
        if (glob(patterns[k], 0, NULL, &g) != 0)
            continue;
        for (size_t i = 0; i < g.gl_pathc && cached == NULL; i++) {
            void *h = dlopen(g.gl_pathv[i], RTLD_NOW | RTLD_LOCAL);
            if (h == NULL)
                continue;
            cached = (zhseqr_fn)dlsym(h, "scipy_LAPACKE_zhseqr");
            if (cached == NULL)
                cached = (zhseqr_fn)dlsym(h, "LAPACKE_zhseqr");
        }
        globfree(&g);
        if (cached != NULL)
            break;
    }
    return cached;
}

/* coeffs: N+1 values, highest power first.  roots: N values. */
int32_t z_poly_roots_modified_(int32_t *N, double complex const *const coeffs,
                               double complex *const roots, int32_t *info)
{
    const int n = *N;
    int lead = 0;
    *info = 1;
    for (int i = 0; i < n; i++)
        roots[i] = 0.0;
    while (lead < n && coeffs[lead] == 0.0)
        lead++;
    const int m = n - lead; /* effective degree */
    if (m <= 0) {
        *info = 0;
        return 0;
    }
    zhseqr_fn zhseqr = find_zhseqr();
    if (zhseqr == NULL)
        return 0;
    double complex *H = calloc((size_t)m * (size_t)m, sizeof(double complex));
    if (H == NULL)
        return 0;
    /* column-major companion matrix, already upper Hessenberg */
    for (int k = 0; k < m; k++)
        H[(size_t)k * m] = -coeffs[lead + k + 1] / coeffs[lead];
    for (int k = 0; k + 1 < m; k++)
        H[(size_t)k * m + (k + 1)] = 1.0;
    int rc = zhseqr(102, 'E', 'N', m, 1, m, H, m, roots, NULL, 1);
    free(H);
    *info = (rc != 0);
    return 0;
}
