/*
 * fnft_b200 host library -- GPU implementations of the reference's private
 * building blocks, exported under their original names so that the reference's own
 * unit tests (which call fnft__* symbols directly) can be linked
 * against this library.  Each is the B = 1 case of the batched device pipeline.
 */
#include "fnft_internal.h"
#include <stdlib.h>

/* src/private/fnft__poly_fmult.c:40-43 */
FNFT_UINT fnft__poly_fmult2x2_numel(FNFT_UINT deg, FNFT_UINT n)
{
    return 4 * (deg + 1) * fnftb__nextpow2(n);
}

static fnftb_ctx *ctx_or_error(void)
{
    fnftb_ctx *ctx = fnftb__ctx();
    if (ctx == NULL)
        (void)E_OTHER("No usable CUDA device: the fnft_b200 hot path has no CPU fallback.");
    return ctx;
}

/* src/private/fnft__poly_fmult.c:381-546.  p: [4][n][deg+1]; result: [4][deg_out+1]
 * in the first 4*(deg_out+1) entries; *d is updated to deg_out = deg*n.  (The
 * reference also clobbers p; we leave it untouched.) */
FNFT_INT fnft__poly_fmult2x2(FNFT_UINT *const d, FNFT_UINT n, FNFT_COMPLEX *const p,
                             FNFT_COMPLEX *const result, FNFT_INT *const W_ptr)
{
    if (d == NULL)
        return E_INVALID_ARGUMENT(d);
    if (p == NULL)
        return E_INVALID_ARGUMENT(p);
    if (result == NULL)
        return E_INVALID_ARGUMENT(result);
    if (n == 0)
        return E_INVALID_ARGUMENT(n);
    if (*d == 0) /* degree-0 "polynomials": plain matrix product, not a hot path */
        return E_NOT_YET_IMPLEMENTED(d, Degree zero is not supported by the GPU tree.);
    fnftb_ctx *ctx = ctx_or_error();
    if (ctx == NULL)
        return FNFT_EC_OTHER;
    if (n == 1) {
        memmove(result, p, 4 * (*d + 1) * sizeof(FNFT_COMPLEX));
        if (W_ptr != NULL)
            *W_ptr = 0;
        return FNFT_SUCCESS;
    }
    if (fnftb_fmult2x2(ctx, *d, n, p, W_ptr != NULL) != 0)
        return E_DEVICE;
    int32_t W = 0;
    if (fnftb_get_transfer_matrix(ctx, result, &W) != 0)
        return E_DEVICE;
    *d = fnftb_result_degree(ctx);
    if (W_ptr != NULL)
        *W_ptr = W;
    return FNFT_SUCCESS;
}

/* src/private/fnft__poly_fmult.c:35-38 */
FNFT_UINT fnft__poly_fmult_numel(FNFT_UINT deg, FNFT_UINT n) { return (deg + 1) * fnftb__nextpow2(n); }

/*
 * Product of n scalar polynomials of degree *d (src/private/fnft__poly_fmult.c:152-237): p holds them back to
 * back, highest power first; on return the first *d + 1 entries hold the product and *d its degree.  The GPU
 * tree multiplies 2x2 matrices, so the polynomials ride as diag(p_i, 1): one launch sequence of the general
 * tree (tree_kernels.cuh), entry 11 of the result is the product.  With W_ptr the tree rescales by powers of
 * two like poly_rescale (:330-374); result * 2^W is the product.
 */
FNFT_INT fnft__poly_fmult(FNFT_UINT *const d, FNFT_UINT n, FNFT_COMPLEX *const p, FNFT_INT *const W_ptr)
{
    if (d == NULL)
        return E_INVALID_ARGUMENT(d);
    if (p == NULL)
        return E_INVALID_ARGUMENT(p);
    if (n == 0)
        return E_INVALID_ARGUMENT(n);
    const FNFT_UINT deg = *d, d1 = deg + 1;
    if (n == 1) {
        if (W_ptr != NULL)
            *W_ptr = 0;
        return FNFT_SUCCESS;
    }
    const FNFT_UINT numel = fnft__poly_fmult2x2_numel(deg, n);
    FNFT_COMPLEX *m = calloc(2 * numel, sizeof(FNFT_COMPLEX)); /* matrices, then the result */
    if (m == NULL)
        return E_NOMEM;
    memcpy(m, p, n * d1 * sizeof(FNFT_COMPLEX)); /* entry 11 of every matrix */
    for (FNFT_UINT i = 0; i < n; i++)
        m[3 * n * d1 + i * d1 + deg] = 1.0; /* entry 22 = the constant 1 */
    FNFT_UINT dout = deg;
    FNFT_INT ret_code = fnft__poly_fmult2x2(&dout, n, m, m + numel, W_ptr);
    if (ret_code == FNFT_SUCCESS) {
        memcpy(p, m + numel, (dout + 1) * sizeof(FNFT_COMPLEX));
        *d = dout;
    } else {
        ret_code = E_SUBROUTINE(ret_code);
    }
    free(m);
    return ret_code;
}

/* src/private/fnft__poly_chirpz.c:33-105 */
FNFT_INT fnft__poly_chirpz(const FNFT_UINT deg, FNFT_COMPLEX const *const p,
                           const FNFT_COMPLEX A, const FNFT_COMPLEX W, const FNFT_UINT M,
                           FNFT_COMPLEX *const result)
{
    if (p == NULL)
        return E_INVALID_ARGUMENT(p);
    if (M == 0)
        return E_INVALID_ARGUMENT(M);
    if (result == NULL)
        return E_INVALID_ARGUMENT(result);
    fnftb_ctx *ctx = ctx_or_error();
    if (ctx == NULL)
        return FNFT_EC_OTHER;
    if (fnftb_set_polynomial(ctx, deg, p) != 0)
        return E_DEVICE;
    fnftb_contspec_desc cd;
    memset(&cd, 0, sizeof(cd));
    cd.mode = FNFTB_MODE_RAW;
    cd.npoly = 1;
    cd.M = M;
    fnftb__logpolar(W, &cd.lwr, &cd.lwi);
    fnftb__logpolar(A, &cd.lar, &cd.lai);
    if (fnftb_contspec(ctx, &cd, result, M, 0, NULL) != 0)
        return E_DEVICE;
    return FNFT_SUCCESS;
}

/* src/private/fnft__poly_roots_fasteigen.c:29-48 */
FNFT_INT fnft__poly_roots_fasteigen(const FNFT_UINT deg, FNFT_COMPLEX const *const p,
                                    FNFT_COMPLEX *const roots)
{
    if (p == NULL)
        return E_INVALID_ARGUMENT(p);
    if (roots == NULL)
        return E_INVALID_ARGUMENT(roots);
    if (deg == 0)
        return FNFT_SUCCESS;
    fnftb_ctx *ctx = ctx_or_error();
    if (ctx == NULL)
        return FNFT_EC_OTHER;
    if (fnftb_set_polynomial(ctx, deg, p) != 0)
        return E_DEVICE;
    int32_t info[4];
    if (fnftb_poly_roots(ctx, 0, roots, info) != 0)
        return E_DEVICE;
    if (info[3] != 0) /* eiscor's info != 0 (:44-47) */
        return E_SUBROUTINE(FNFT_EC_OTHER);
    return FNFT_SUCCESS;
}

/* src/private/fnft__akns_fscatter.c:33-41 */
FNFT_UINT fnft__akns_fscatter_numel(FNFT_UINT D, fnft__akns_discretization_t discretization)
{
    const FNFT_UINT deg = fnftb__akns_degree(discretization);
    if (deg == 0)
        return 0;
    return fnft__poly_fmult2x2_numel(deg, D);
}

static FNFT_INT fscatter_common(const FNFT_UINT D, FNFT_COMPLEX const *q, FNFT_COMPLEX const *r,
                                int rmode, int kappa, const FNFT_REAL eps_t, FNFT_COMPLEX *result,
                                FNFT_UINT *deg_ptr, FNFT_INT *W_ptr,
                                fnft__akns_discretization_t akns)
{
    const FNFT_UINT deg0 = fnftb__akns_degree(akns);
    if (deg0 == 0)
        return E_INVALID_ARGUMENT(discretization);
    if (!fnftb__akns_on_gpu(akns))
        return E_NOT_YET_IMPLEMENTED(discretization, This splitting scheme has no GPU leaf kernel yet.);
    fnftb_ctx *ctx = ctx_or_error();
    if (ctx == NULL)
        return FNFT_EC_OTHER;
    if (fnftb_set_signals(ctx, 1, D, q, r, 0) != 0)
        return E_DEVICE;
    fnftb_scatter_desc sd;
    memset(&sd, 0, sizeof(sd));
    sd.rmode = rmode;
    sd.kappa = kappa;
    sd.scheme = (int)akns;
    sd.deg0 = (int)deg0;
    sd.normalize = (W_ptr != NULL);
    sd.eps_t = eps_t;
    if (fnftb_fscatter(ctx, &sd) != 0)
        return E_DEVICE;
    int32_t W = 0, status = 0;
    if (fnftb_get_transfer_matrix(ctx, result, &W) != 0)
        return E_DEVICE;
    if (fnftb_get_status(ctx, &status) != 0)
        return E_DEVICE;
    if (status == 1) /* src/private/fnft__akns_fscatter.c:124-126 */
        return E_OTHER("kappa == -1 but eps_t*|q[i]|>=1 ... decrease step size");
    if (status != 0)
        return E_INVALID_ARGUMENT(discretization);
    *deg_ptr = fnftb_result_degree(ctx);
    if (W_ptr != NULL)
        *W_ptr = W;
    return FNFT_SUCCESS;
}

/* src/private/fnft__akns_fscatter.c:64-925 */
FNFT_INT fnft__akns_fscatter(const FNFT_UINT D, FNFT_COMPLEX const *const q,
                             FNFT_COMPLEX const *const r, const FNFT_REAL eps_t,
                             FNFT_COMPLEX *const result, FNFT_UINT *const deg_ptr,
                             FNFT_INT *const W_ptr, fnft__akns_discretization_t discretization)
{
    if (D == 0)
        return E_INVALID_ARGUMENT(D);
    if (q == NULL)
        return E_INVALID_ARGUMENT(q);
    if (r == NULL)
        return E_INVALID_ARGUMENT(r);
    if (eps_t <= 0.0)
        return E_INVALID_ARGUMENT(eps_t);
    if (result == NULL)
        return E_INVALID_ARGUMENT(result);
    if (deg_ptr == NULL)
        return E_INVALID_ARGUMENT(deg_ptr);
    return fscatter_common(D, q, r, FNFTB_RMODE_EXPLICIT, 0, eps_t, result, deg_ptr, W_ptr,
                           discretization);
}

/* src/private/fnft__nse_fscatter.c:30-42 */
FNFT_UINT fnft__nse_fscatter_numel(FNFT_UINT D, fnft_nse_discretization_t discretization)
{
    const FNFT_UINT deg = fnftb__nse_degree(discretization);
    if (deg == 0)
        return 0;
    return fnft__poly_fmult2x2_numel(deg, D);
}

/* src/private/fnft__nse_fscatter.c:44-91 (r = -kappa*conj(q) is formed on the device) */
FNFT_INT fnft__nse_fscatter(const FNFT_UINT D, FNFT_COMPLEX const *const q, const FNFT_REAL eps_t,
                            const FNFT_INT kappa, FNFT_COMPLEX *const result,
                            FNFT_UINT *const deg_ptr, FNFT_INT *const W_ptr,
                            fnft_nse_discretization_t discretization)
{
    if (D == 0)
        return E_INVALID_ARGUMENT(D);
    if (q == NULL)
        return E_INVALID_ARGUMENT(q);
    if (eps_t <= 0.0)
        return E_INVALID_ARGUMENT(eps_t);
    if (abs(kappa) != 1)
        return E_INVALID_ARGUMENT(kappa);
    if (result == NULL)
        return E_INVALID_ARGUMENT(result);
    if (deg_ptr == NULL)
        return E_INVALID_ARGUMENT(deg_ptr);
    fnft__akns_discretization_t akns;
    FNFT_INT ret_code = fnftb__nse_to_akns(discretization, &akns);
    if (ret_code != FNFT_SUCCESS)
        return E_SUBROUTINE(ret_code);
    return fscatter_common(D, q, NULL, FNFTB_RMODE_NSE, kappa, eps_t, result, deg_ptr, W_ptr, akns);
}

/* src/private/fnft__kdv_fscatter.c:32-43 */
FNFT_UINT fnft__kdv_fscatter_numel(FNFT_UINT D, fnft_kdv_discretization_t discretization)
{
    fnft__akns_discretization_t akns;
    if (!fnftb__kdv_to_akns_quiet(discretization, &akns))
        return 0;
    return fnft__akns_fscatter_numel(D, akns);
}

/* src/private/fnft__kdv_fscatter.c:45-83 (r = -1 is formed on the device) */
FNFT_INT fnft__kdv_fscatter(const FNFT_UINT D, FNFT_COMPLEX const *const u, const FNFT_REAL eps_t,
                            FNFT_COMPLEX *const result, FNFT_UINT *const deg_ptr,
                            FNFT_INT *const W_ptr, fnft_kdv_discretization_t discretization)
{
    if (D == 0)
        return E_INVALID_ARGUMENT(D);
    if (u == NULL)
        return E_INVALID_ARGUMENT(u);
    if (eps_t <= 0.0)
        return E_INVALID_ARGUMENT(eps_t);
    if (result == NULL)
        return E_INVALID_ARGUMENT(result);
    if (deg_ptr == NULL)
        return E_INVALID_ARGUMENT(deg_ptr);
    fnft__akns_discretization_t akns;
    FNFT_INT ret_code = fnftb__kdv_to_akns(discretization, &akns);
    if (ret_code != FNFT_SUCCESS)
        return E_SUBROUTINE(ret_code);
    return fscatter_common(D, u, NULL, FNFTB_RMODE_KDV, 0, eps_t, result, deg_ptr, W_ptr, akns);
}

/* src/private/fnft__nse_scatter_bound_states.c:29-667, BO and CF4_2 only (the two
 * base methods the fast discretizations use, src/fnft_nsev.c:675-680).  The GPU
 * kernels form r = -conj(q) themselves; an explicit r is accepted only if NULL. */
FNFT_INT fnft__nse_scatter_bound_states(const FNFT_UINT D, FNFT_COMPLEX const *const q,
                                        FNFT_COMPLEX *r, FNFT_REAL const *const T, FNFT_UINT K,
                                        FNFT_COMPLEX *bound_states, FNFT_COMPLEX *a_vals,
                                        FNFT_COMPLEX *aprime_vals, FNFT_COMPLEX *b,
                                        fnft_nse_discretization_t discretization,
                                        FNFT_UINT skip_b_flag)
{
    if (D == 0)
        return E_INVALID_ARGUMENT(D);
    if (q == NULL)
        return E_INVALID_ARGUMENT(q);
    if (T == NULL)
        return E_INVALID_ARGUMENT(eps_t);
    if (K == 0)
        return E_INVALID_ARGUMENT(K);
    if (bound_states == NULL)
        return E_INVALID_ARGUMENT(bound_states);
    if (a_vals == NULL)
        return E_INVALID_ARGUMENT(a);
    if (aprime_vals == NULL)
        return E_INVALID_ARGUMENT(a_prime);
    if (b == NULL)
        return E_INVALID_ARGUMENT(b);
    int upsampling, wsel = 0;
    if (discretization == fnft_nse_discretization_BO)
        upsampling = 1;
    else if (discretization == fnft_nse_discretization_CF4_2)
        upsampling = 2;
    else if (discretization == fnft_nse_discretization_CF4_3) {
        upsampling = 3;
        wsel = 1;
    } else if (discretization == fnft_nse_discretization_ES4 || discretization == fnft_nse_discretization_TES4) {
        upsampling = 3; /* the samples are (q, q', q'') per grid point */
        wsel = (discretization == fnft_nse_discretization_ES4) ? 4 : 5;
    } else
        return E_NOT_YET_IMPLEMENTED(discretization, CF5_3 and CF6_4 need the explicit r samples of fnft_nsev.);
    if (D % upsampling != 0) /* fnft__nse_scatter_bound_states.c:231-246 */
        return E_ASSERTION_FAILED;
    if (r != NULL) {
        for (FNFT_UINT n = 0; n < D; n++)
            if (r[n] != -conj(q[n]))
                return E_NOT_YET_IMPLEMENTED(r, The GPU kernels assume r = -conj(q).);
    }
    fnftb_ctx *ctx = ctx_or_error();
    if (ctx == NULL)
        return FNFT_EC_OTHER;
    if (fnftb_set_signals(ctx, 1, D, q, NULL, 0) != 0 || fnftb_set_slow_weights(ctx, wsel) != 0)
        return E_DEVICE;
    const FNFT_UINT D_given = D / upsampling;
    fnftb_bound_desc bd;
    memset(&bd, 0, sizeof(bd));
    bd.upsampling = upsampling;
    bd.Kmax = (int)K;
    bd.T0 = T[0];
    bd.T1 = T[1];
    bd.eps_t = (T[1] - T[0]) / (D_given - 1);
    bd.bc = 0.5;
    bd.lweight = (upsampling == 2) ? 0.5 : 1.0;
    bd.scl = (wsel >= 4) ? 1.0 : 1.0 / upsampling; /* fnft__nse_scatter_bound_states.c:132,157,225,235,247 */
    int32_t Kc = (int32_t)K;
    (void)skip_b_flag; /* b is cheap next to the sweeps; always computed */
    if (fnftb_normconsts(ctx, &bd, &Kc, bound_states, a_vals, aprime_vals, b) != 0)
        return E_DEVICE;
    return FNFT_SUCCESS;
}
