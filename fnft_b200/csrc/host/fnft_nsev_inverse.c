/*
 * fnft_b200 host library -- inverse nonlinear Fourier transform, fnft_nsev_inverse
 * (include/fnft_nsev_inverse.h, src/fnft_nsev_inverse.c:26-1033) and the private symbols
 * fnft__nse_finvscatter (src/private/fnft__nse_finvscatter.c:243-366) and fnft__poly_specfact
 * (src/private/fnft__poly_specfact.c:25-147).
 *
 * Host: argument checks, option logic, the O(M) phase factors that the reference applies to the caller's
 * contspec array in place, sorting / residue conversion of the K bound states.  Device (../cuda/inverse_api.cu):
 * the FFTs, the spectral factorisation, the fast inverse scattering and the Darboux transforms.
 * No CPU fallback: without a GPU every entry point returns FNFT_EC_OTHER.
 */
#include "fnft_internal.h"
#include <complex.h>
#include <math.h>
#include <stdlib.h>

#define PI 3.14159265358979323846

static const fnft_nsev_inverse_opts_t inverse_defaults = {
    /* src/fnft_nsev_inverse.c:26-33 */
    .discretization = fnft_nse_discretization_2SPLIT2A,
    .contspec_type = fnft_nsev_inverse_cstype_REFLECTION_COEFFICIENT,
    .contspec_inversion_method = fnft_nsev_inverse_csmethod_DEFAULT,
    .discspec_type = fnft_nsev_inverse_dstype_NORMING_CONSTANTS,
    .max_iter = 100,
    .oversampling_factor = 8};

fnft_nsev_inverse_opts_t fnft_nsev_inverse_default_opts(void) { return inverse_defaults; }

/* src/fnft_nsev_inverse.c:40-66 */
FNFT_INT fnft_nsev_inverse_XI(const FNFT_UINT D, FNFT_REAL const *const T, const FNFT_UINT M, FNFT_REAL *const XI,
                              const fnft_nse_discretization_t discretization)
{
    if (D < 2)
        return E_INVALID_ARGUMENT(D);
    if (M == 0)
        return E_INVALID_ARGUMENT(M);
    if (XI == NULL)
        return E_INVALID_ARGUMENT(XI);
    if (T == NULL || !(T[0] < T[1]))
        return E_INVALID_ARGUMENT(T);
    const FNFT_UINT deg1 = fnftb__nse_degree(discretization) * fnftb__nse_upsampling(discretization);
    if (deg1 == 0)
        return E_INVALID_ARGUMENT(discretization);
    const FNFT_REAL eps_t = (T[1] - T[0]) / (D - 1);
    /* z = exp(2 pi i (M/2 + 1) / M) and z = -1, mapped by z_to_lambda (fnft__akns_discretization.c:225-240) */
    const FNFT_COMPLEX z0 = cexp(2.0 * PI * I * (FNFT_REAL)(M / 2 + 1) / (FNFT_REAL)M);
    const FNFT_COMPLEX z1 = -1.0;
    XI[0] = creal(clog(z0) / (2 * I * eps_t / (FNFT_REAL)deg1));
    XI[1] = creal(clog(z1) / (2 * I * eps_t / (FNFT_REAL)deg1));
    return FNFT_SUCCESS;
}

/* src/private/fnft__nse_finvscatter.c:243-366 */
FNFT_INT fnft__nse_finvscatter(const FNFT_UINT deg, FNFT_COMPLEX *const transfer_matrix, FNFT_COMPLEX *const q,
                               const FNFT_REAL eps_t, const FNFT_INT kappa,
                               const fnft_nse_discretization_t discretization)
{
    if (deg == 0)
        return E_INVALID_ARGUMENT(de);
    if (transfer_matrix == NULL)
        return E_INVALID_ARGUMENT(transfer_matrix);
    if (q == NULL)
        return E_INVALID_ARGUMENT(q);
    if (!(eps_t > 0.0))
        return E_INVALID_ARGUMENT(eps_t);
    if (kappa != -1 && kappa != 1)
        return E_INVALID_ARGUMENT(kappa);
    const FNFT_UINT d1 = fnftb__nse_degree(discretization);
    if (d1 == 0)
        return E_INVALID_ARGUMENT(discretization);
    const FNFT_UINT D = deg / d1;
    if (D < 2 || (D & (D - 1)) != 0)
        return E_OTHER("Number of samples D used to build the transfer matrix was no positive power of two.");
    /* the base case of the reference knows these two discretizations only (:165-196) */
    if (discretization != fnft_nse_discretization_2SPLIT2A && discretization != fnft_nse_discretization_2SPLIT2_MODAL)
        return E_INVALID_ARGUMENT(discretization);
    fnftb_ctx *ctx = fnftb__ctx();
    if (ctx == NULL)
        return E_OTHER("No usable CUDA device: the fnft_b200 hot path has no CPU fallback.");
    int32_t status = 0;
    if (fnftb_finvscatter(ctx, 1, deg, transfer_matrix, q, eps_t, kappa,
                          discretization == fnft_nse_discretization_2SPLIT2_MODAL, 0, &status) != 0)
        return E_DEVICE;
    if (status != 0)
        return E_OTHER("A reconstruced sample violates the condition |q[n]|<1.");
    return FNFT_SUCCESS;
}

/* src/private/fnft__poly_specfact.c:25-147 */
FNFT_INT fnft__poly_specfact(const FNFT_UINT deg, FNFT_COMPLEX const *const poly, FNFT_COMPLEX *const result,
                             const FNFT_UINT oversampling_factor, const FNFT_INT kappa)
{
    if (deg == 0)
        return E_INVALID_ARGUMENT(deg);
    if (poly == NULL)
        return E_INVALID_ARGUMENT(poly);
    if (result == NULL)
        return E_INVALID_ARGUMENT(result);
    if (oversampling_factor == 0)
        return E_INVALID_ARGUMENT(oversampling_factor);
    if (kappa < -1 || kappa > 1)
        return E_INVALID_ARGUMENT(kappa);
    fnftb_ctx *ctx = fnftb__ctx();
    if (ctx == NULL)
        return E_OTHER("No usable CUDA device: the fnft_b200 hot path has no CPU fallback.");
    int32_t warn = 0;
    if (fnftb_specfact(ctx, 1, deg, poly, result, oversampling_factor, kappa, &warn) != 0)
        return E_DEVICE;
    if (warn)
        WARN("Ill-posed spectral factorization problem.");
    return FNFT_SUCCESS;
}

/* sorts by descending imaginary part with the reference's exchange sort (src/fnft_nsev_inverse.c:741-752: the order
 * of equal imaginary parts matters for the recursion), rejects multiple bound states, converts residues */
static FNFT_INT prepare_discspec(FNFT_UINT K, FNFT_COMPLEX const *bound_states, FNFT_COMPLEX const *nc_or_res,
                                 FNFT_COMPLEX *bs, FNFT_COMPLEX *nc)
{
    for (FNFT_UINT i = 0; i < K; i++) {
        bs[i] = bound_states[i];
        nc[i] = nc_or_res[i];
    }
    for (FNFT_UINT i = 0; i < K; ++i)
        for (FNFT_UINT j = i + 1; j < K; ++j)
            if (cimag(bs[i]) < cimag(bs[j])) {
                FNFT_COMPLEX tmp = bs[i];
                bs[i] = bs[j];
                bs[j] = tmp;
                tmp = nc[i];
                nc[i] = nc[j];
                nc[j] = tmp;
            }
    for (FNFT_UINT i = 0; i + 1 < K; i++)
        if (bs[i + 1] == bs[i])
            return ERRMSG(FNFT_EC_SANITY_CHECK_FAILED,
                          "Sanity check failed (Bound_states should be simple (multiplicity should be 1).).");
    return FNFT_SUCCESS;
}

/* residues -> norming constants (:764-789); acoeff: a(lambda_i) of the seed potential, or NULL for 1 */
static void residues_to_normconsts(FNFT_UINT K, FNFT_COMPLEX const *bs, FNFT_COMPLEX *nc, FNFT_COMPLEX const *acoeff)
{
    for (FNFT_UINT i = 0; i < K; i++) {
        FNFT_COMPLEX tmp = acoeff ? acoeff[i] : 1.0;
        for (FNFT_UINT j = 0; j < K; j++)
            if (j != i)
                tmp = tmp * (bs[i] - bs[j]) / (bs[i] - conj(bs[j]));
        nc[i] = (nc[i] / (2 * I * cimag(bs[i]))) * tmp;
    }
}

static FNFT_INT inverse_checks(const FNFT_UINT M, FNFT_COMPLEX const *contspec, FNFT_REAL const *XI, FNFT_UINT const K,
                               FNFT_COMPLEX const *bound_states, FNFT_COMPLEX const *nc, const FNFT_UINT D,
                               FNFT_COMPLEX const *q, FNFT_REAL const *T, const FNFT_INT kappa,
                               fnft_nsev_inverse_opts_t const *o, FNFT_UINT B)
{
    /* src/fnft_nsev_inverse.c:134-172 */
    if (M > 0 && contspec == NULL)
        return E_INVALID_ARGUMENT(contspec);
    if (contspec != NULL && M % 2 != 0)
        return E_INVALID_ARGUMENT(M);
    if (contspec != NULL && M < D)
        return E_INVALID_ARGUMENT(M);
    if (D < 2 || (D & (D - 1)) != 0)
        return E_INVALID_ARGUMENT(D);
    if (q == NULL)
        return E_INVALID_ARGUMENT(q);
    if (T == NULL || !(T[0] < T[1]))
        return E_INVALID_ARGUMENT(T);
    if (kappa != +1 && kappa != -1)
        return E_INVALID_ARGUMENT(kappa);
    if (K > 0 && kappa != +1)
        return ERRMSG(FNFT_EC_SANITY_CHECK_FAILED,
                      "Sanity check failed (Discrete spectrum is present only in the focussing case(kappa=1).).");
    if (K > 0 && bound_states == NULL)
        return E_INVALID_ARGUMENT(bound_states);
    for (FNFT_UINT i = 0; i < B * K; i++)
        if (cimag(bound_states[i]) <= 0)
            return ERRMSG(FNFT_EC_SANITY_CHECK_FAILED,
                          "Sanity check failed (bound_states should be stricly in the upper-half complex-plane.).");
    if (K > 0 && nc == NULL)
        return E_INVALID_ARGUMENT(normconsts_or_residues);
    if (o->discretization != fnft_nse_discretization_2SPLIT2A &&
        o->discretization != fnft_nse_discretization_2SPLIT2_MODAL)
        return E_INVALID_ARGUMENT(opts_ptr->discretization);
    if (contspec == NULL && K == 0)
        return ERRMSG(FNFT_EC_SANITY_CHECK_FAILED, "Sanity check failed (Neither contspec nor discspec provided.).");
    if (XI == NULL && contspec != NULL && o->contspec_type != fnft_nsev_inverse_cstype_B_OF_TAU)
        return E_INVALID_ARGUMENT(XI);
    return FNFT_SUCCESS;
}

static FNFT_INT nsev_inverse_batch_impl(const FNFT_UINT B, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                                 FNFT_REAL const *const XI, FNFT_UINT const K,
                                 FNFT_COMPLEX const *const bound_states,
                                 FNFT_COMPLEX const *const normconsts_or_residues, const FNFT_UINT D,
                                 FNFT_COMPLEX *const q, FNFT_REAL const *const T, const FNFT_INT kappa,
                                 fnft_nsev_inverse_opts_t const *opts_ptr, FNFT_INT *ret_codes);

/* public entry: NVTX range around the call (FNFT_B200_NVTX=1, no-op otherwise) */
FNFT_INT fnft_nsev_inverse_batch(const FNFT_UINT B, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                                 FNFT_REAL const *const XI, FNFT_UINT const K,
                                 FNFT_COMPLEX const *const bound_states,
                                 FNFT_COMPLEX const *const normconsts_or_residues, const FNFT_UINT D,
                                 FNFT_COMPLEX *const q, FNFT_REAL const *const T, const FNFT_INT kappa,
                                 fnft_nsev_inverse_opts_t const *opts_ptr, FNFT_INT *ret_codes)
{
    fnftb_range_push("fnft_nsev_inverse_batch");
    const FNFT_INT rc = nsev_inverse_batch_impl(B, M, contspec, XI, K, bound_states, normconsts_or_residues, D, q, T, kappa, opts_ptr, ret_codes);
    fnftb_range_pop();
    return rc;
}

static FNFT_INT nsev_inverse_batch_impl(const FNFT_UINT B, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                                 FNFT_REAL const *const XI, FNFT_UINT const K,
                                 FNFT_COMPLEX const *const bound_states,
                                 FNFT_COMPLEX const *const normconsts_or_residues, const FNFT_UINT D,
                                 FNFT_COMPLEX *const q, FNFT_REAL const *const T, const FNFT_INT kappa,
                                 fnft_nsev_inverse_opts_t const *opts_ptr, FNFT_INT *ret_codes)
{
    if (B == 0)
        return E_INVALID_ARGUMENT(B);
    if (opts_ptr == NULL)
        opts_ptr = &inverse_defaults;
    FNFT_INT ret_code = inverse_checks(M, contspec, XI, K, bound_states, normconsts_or_residues, D, q, T, kappa,
                                       opts_ptr, B);
    if (ret_code != FNFT_SUCCESS)
        return ret_code;
    if (ret_codes != NULL)
        for (FNFT_UINT b = 0; b < B; b++)
            ret_codes[b] = FNFT_SUCCESS;
    fnftb_ctx *ctx = fnftb__ctx();
    if (ctx == NULL)
        return E_OTHER("No usable CUDA device: the fnft_b200 hot path has no CPU fallback.");

    const fnft_nse_discretization_t disc = opts_ptr->discretization;
    const int modal = (disc == fnft_nse_discretization_2SPLIT2_MODAL);
    const FNFT_REAL eps_t = (T[1] - T[0]) / (D - 1);
    const FNFT_UINT deg = D * fnftb__nse_degree(disc);
    FNFT_COMPLEX *bs = NULL, *nc = NULL, *acoeff = NULL;
    int32_t *flags = NULL;
    int contspec_flag = 0;

    if (contspec != NULL) {
        contspec_flag = 1;
        const fnft_nsev_inverse_cstype_t cst = opts_ptr->contspec_type;
        const fnft_nsev_inverse_csmethod_t method = opts_ptr->contspec_inversion_method;
        if (cst != fnft_nsev_inverse_cstype_REFLECTION_COEFFICIENT && cst != fnft_nsev_inverse_cstype_B_OF_XI &&
            cst != fnft_nsev_inverse_cstype_B_OF_TAU)
            return E_INVALID_ARGUMENT(opts_ptr->contspec_type);
        if (cst == fnft_nsev_inverse_cstype_REFLECTION_COEFFICIENT) {
            if (method == fnft_nsev_inverse_csmethod_TFMATRIX_CONTAINS_AB_FROM_ITER) {
                /* :388-395 */
                if (M != D)
                    return E_SUBROUTINE(E_INVALID_ARGUMENT(M));
                if (D != deg)
                    return E_SUBROUTINE(E_ASSERTION_FAILED);
                if (kappa != -1)
                    return E_SUBROUTINE(E_INVALID_ARGUMENT(kappa));
            } else if (method != fnft_nsev_inverse_csmethod_DEFAULT &&
                       method != fnft_nsev_inverse_csmethod_TFMATRIX_CONTAINS_REFL_COEFF)
                return E_SUBROUTINE(E_INVALID_ARGUMENT(opts_ptr->contspec_inversion_method));
        }
        if (cst == fnft_nsev_inverse_cstype_B_OF_TAU) {
            /* :640-646 */
            if (M != D)
                return E_INVALID_ARGUMENT(M);
            if (T[0] != -T[1])
                return E_INVALID_ARGUMENT(T);
            if (method != fnft_nsev_inverse_csmethod_DEFAULT)
                return E_INVALID_ARGUMENT(opts_ptr->contspec_inversion_method);
        } else {
            /* precompensation for the phase shifts of the Darboux transform (:1013-1033, reflection coefficient only)
             * and removal of the boundary-condition phase factors (:263-287); both modify contspec in place */
            FNFT_REAL phase_factor = 0.0;
            if (cst == fnft_nsev_inverse_cstype_REFLECTION_COEFFICIENT)
                ret_code = fnftb__nse_phase_factor_rho(eps_t, T[1], &phase_factor, disc);
            else
                ret_code = fnftb__nse_phase_factor_b(eps_t, D, T, &phase_factor, disc);
            if (ret_code != FNFT_SUCCESS)
                return E_SUBROUTINE(ret_code);
            const FNFT_REAL eps_xi = (XI[1] - XI[0]) / (M - 1);
            for (FNFT_UINT b = 0; b < B; b++) {
                FNFT_COMPLEX *cs = contspec + b * M;
                for (FNFT_UINT i = 0; i < M; i++) {
                    const FNFT_REAL xi = XI[0] + i * eps_xi;
                    if (cst == fnft_nsev_inverse_cstype_REFLECTION_COEFFICIENT)
                        for (FNFT_UINT k = 0; k < K; k++)
                            cs[i] *= (xi - bound_states[b * K + k]) / (xi - conj(bound_states[b * K + k]));
                    cs[i] *= cexp(-I * xi * phase_factor);
                }
            }
        }
        flags = calloc(B, sizeof(int32_t));
        if (flags == NULL)
            return E_NOMEM;
        if (cst == fnft_nsev_inverse_cstype_REFLECTION_COEFFICIENT &&
            method == fnft_nsev_inverse_csmethod_TFMATRIX_CONTAINS_AB_FROM_ITER) {
            /* Algorithm 1 of arXiv:1607.01305v2 (:375-510): the number of iterations depends on the signal, so the
             * signals of a batch are processed one after the other */
            for (FNFT_UINT b = 0; b < B; b++) {
                int32_t hit_max = 0, warn = 0;
                if (fnftb_inv_tm_ab_from_iter(ctx, D, contspec + b * M, kappa, opts_ptr->max_iter, &hit_max, &warn) != 0) {
                    ret_code = E_DEVICE;
                    goto leave_fun;
                }
                if (warn)
                    WARN("Ill-posed spectral factorization problem.");
                if (hit_max)
                    WARN("Maximum number of iterations reached when constructing transfer matrix.");
                if (fnftb_finvscatter_staged(ctx, 1, deg, q + b * D, eps_t, kappa, modal, flags + b) != 0) {
                    ret_code = E_DEVICE;
                    goto leave_fun;
                }
            }
        } else {
        if (fnftb_inv_tm_from_contspec(ctx, B, M, D, deg, contspec, (int)cst, kappa, eps_t,
                                       opts_ptr->oversampling_factor, flags) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        for (FNFT_UINT b = 0; b < B; b++)
            if (flags[b]) {
                WARN("Ill-posed spectral factorization problem.");
                break;
            }
        if (fnftb_finvscatter_staged(ctx, B, deg, q, eps_t, kappa, modal, flags) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        }
        for (FNFT_UINT b = 0; b < B; b++)
            if (flags[b]) {
                const FNFT_INT ec = E_SUBROUTINE(E_OTHER("A reconstruced sample violates the condition |q[n]|<1."));
                if (ret_codes != NULL)
                    ret_codes[b] = ec;
                if (ret_code == FNFT_SUCCESS)
                    ret_code = ec;
            }
        if (ret_code != FNFT_SUCCESS && B == 1)
            goto leave_fun;
    }

    if (K > 0) {
        bs = malloc(B * K * sizeof(FNFT_COMPLEX));
        nc = malloc(B * K * sizeof(FNFT_COMPLEX));
        if (bs == NULL || nc == NULL) {
            ret_code = E_NOMEM;
            goto leave_fun;
        }
        const fnft_nsev_inverse_csmethod_t method = opts_ptr->contspec_inversion_method;
        const int seed = contspec_flag ? (method != fnft_nsev_inverse_csmethod_USE_SEED_POTENTIAL_INSTEAD)
                                       : (method == fnft_nsev_inverse_csmethod_USE_SEED_POTENTIAL_INSTEAD);
        const int pure = (!contspec_flag && method != fnft_nsev_inverse_csmethod_USE_SEED_POTENTIAL_INSTEAD);
        if (!seed && !pure) {
            ret_code = E_INVALID_ARGUMENT(opts_ptr->contspec_inversion_method); /* :900-901 */
            goto leave_fun;
        }
        for (FNFT_UINT b = 0; b < B; b++) {
            FNFT_INT rc = prepare_discspec(K, bound_states + b * K, normconsts_or_residues + b * K, bs + b * K,
                                           nc + b * K);
            if (rc != FNFT_SUCCESS) {
                ret_code = E_SUBROUTINE(rc);
                goto leave_fun;
            }
            if (opts_ptr->discspec_type == fnft_nsev_inverse_dstype_RESIDUES) {
                if (contspec_flag) {
                    /* the non-solitonic part of the potential contributes to the residues (:767-774): a(lambda_i) of
                     * the seed potential, BO scheme, computed by the GPU bound-state kernels */
                    if (acoeff == NULL)
                        acoeff = malloc(3 * K * sizeof(FNFT_COMPLEX));
                    if (acoeff == NULL) {
                        ret_code = E_NOMEM;
                        goto leave_fun;
                    }
                    rc = fnft__nse_scatter_bound_states(D, q + b * D, NULL, T, K, bs + b * K, acoeff, acoeff + K,
                                                        acoeff + 2 * K, fnft_nse_discretization_BO, 1);
                    if (rc != FNFT_SUCCESS) {
                        ret_code = E_SUBROUTINE(rc);
                        goto leave_fun;
                    }
                    residues_to_normconsts(K, bs + b * K, nc + b * K, acoeff);
                } else {
                    residues_to_normconsts(K, bs + b * K, nc + b * K, NULL);
                }
            }
        }
        /* first sample with t >= 0 (:727-733); stays 0 when there is none */
        int zc = 0;
        for (FNFT_UINT i = 0; i < D; i++)
            if (T[0] + eps_t * i >= 0.0) {
                zc = (int)i;
                break;
            }
        if (fnftb_inv_add_solitons(ctx, B, K, D, bs, nc, q, T[0], T[1], zc, seed, 0) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
    }

leave_fun:
    free(bs);
    free(nc);
    free(acoeff);
    free(flags);
    return ret_code;
}

/* include/fnft_nsev_inverse.h:258-263, src/fnft_nsev_inverse.c:121-249 */
FNFT_INT fnft_nsev_inverse(const FNFT_UINT M, FNFT_COMPLEX *const contspec, FNFT_REAL const *const XI,
                           FNFT_UINT const K, FNFT_COMPLEX const *const bound_states,
                           FNFT_COMPLEX const *const normconsts_or_residues, const FNFT_UINT D,
                           FNFT_COMPLEX *const q, FNFT_REAL const *const T, const FNFT_INT kappa,
                           fnft_nsev_inverse_opts_t *opts_ptr)
{
    return fnft_nsev_inverse_batch(1, M, contspec, XI, K, bound_states, normconsts_or_residues, D, q, T, kappa,
                                   opts_ptr, NULL);
}
