/*
 * fnft_b200 host library -- discrete spectrum of fnft_nsev on the GPU.
 *
 * Host-side mirror of nsev_compute_boundstates (src/fnft_nsev.c:595-741, NEWTON
 * branch :665-685), nsev_refine_bound_states_newton (:973-1038, the loop itself runs
 * in the kernel) and nsev_compute_normconsts_or_residues (:895-970).  The cheap
 * order-dependent steps misc_filter / misc_merge stay on the host, as the reference
 * does them, so that K and the ordering of the results are reproduced exactly.
 */
#include "fnft_nsev_discrete.h"

void fnftb__filter_box(FNFT_UINT *N, FNFT_COMPLEX *vals, FNFT_REAL const *box)
{
    FNFT_UINT kept = 0;
    for (FNFT_UINT i = 0; i < *N; i++) {
        const FNFT_REAL re = creal(vals[i]), im = cimag(vals[i]);
        /* written so that NaNs fall outside the box */
        if (!(re >= box[0]) || !(re <= box[1]) || !(im >= box[2]) || !(im <= box[3]))
            continue;
        vals[kept++] = vals[i];
    }
    *N = kept;
}

void fnftb__merge(FNFT_UINT *N, FNFT_COMPLEX *vals, FNFT_REAL tol)
{
    if (*N == 0)
        return;
    FNFT_UINT kept = 1;
    for (FNFT_UINT i = 1; i < *N; i++) {
        /* compared against the ORIGINAL predecessors vals[0..i), like the reference */
        FNFT_REAL dist = -1.0;
        for (FNFT_UINT j = 0; j < i; j++) {
            dist = cabs(vals[j] - vals[i]);
            if (dist < tol)
                break;
        }
        if (dist < tol)
            continue;
        vals[kept++] = vals[i];
    }
    *N = kept;
}

FNFT_INT fnftb__nsev_discrete_chunk(fnftb_ctx *ctx, FNFT_UINT nb, FNFT_UINT D_eff, FNFT_UINT D_given,
                                    FNFT_REAL const *T, FNFT_REAL eps_t, FNFT_UINT *K,
                                    FNFT_UINT Kmax, FNFT_COMPLEX *bound_states,
                                    FNFT_COMPLEX *normconsts_or_residues,
                                    fnft_nsev_opts_t const *opts, FNFT_INT *ret_codes)
{
    FNFT_INT ret_code = FNFT_SUCCESS;
    int32_t *Kc = NULL, *flag = NULL;
    double *box3 = NULL;
    FNFT_COMPLEX *b_vals = NULL, *ap_vals = NULL;
    (void)D_given;

    const FNFT_UINT upsampling = D_eff / D_given;
    const FNFT_REAL degree1step = (FNFT_REAL)fnftb__nse_degree(opts->discretization);
    const FNFT_REAL map_coeff = (degree1step != 0) ? 2 / degree1step : 2.0;

    fnftb_bound_desc bd;
    memset(&bd, 0, sizeof(bd));
    bd.upsampling = (int)upsampling; /* BO for upsampling 1, CF4_2 for 2 (:675-680) */
    bd.Kmax = (int)Kmax;
    bd.T0 = T[0];
    bd.T1 = T[1];
    bd.eps_t = eps_t;
    bd.bc = 0.5;
    bd.lweight = (upsampling == 2) ? 0.5 : 1.0;
    bd.scl = (upsampling == 2) ? 0.5 : 1.0;
    bd.niter = (int)opts->niter;

    Kc = malloc(nb * sizeof(int32_t));
    flag = malloc(nb * Kmax * sizeof(int32_t));
    box3 = malloc(nb * sizeof(double));
    if (Kc == NULL || flag == NULL || box3 == NULL) {
        ret_code = E_NOMEM;
        goto leave_fun;
    }
    for (FNFT_UINT b = 0; b < nb; b++) {
        if (K[b] > Kmax) {
            ret_code = E_INVALID_ARGUMENT(K);
            goto leave_fun;
        }
        Kc[b] = (int32_t)K[b];
    }

    /* bounding box, src/fnft_nsev.c:628-659 */
    if (opts->bound_state_filtering == fnft_nsev_bsfilt_FULL) {
        bd.box1 = 0.9 * FNFT_PI / fabs(map_coeff * eps_t);
        bd.box0 = -bd.box1;
        bd.box2 = 0.0;
        bd.use_box3 = 1;
        if (fnftb_imbound(ctx, (int)upsampling, T[0], T[1], box3) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
    } else if (opts->bound_state_filtering == fnft_nsev_bsfilt_BASIC) {
        bd.box0 = -INFINITY;
        bd.box1 = INFINITY;
        bd.box2 = 0.0;
        for (FNFT_UINT b = 0; b < nb; b++)
            box3[b] = INFINITY;
    } else {
        bd.box0 = -INFINITY;
        bd.box1 = INFINITY;
        bd.box2 = -INFINITY;
        for (FNFT_UINT b = 0; b < nb; b++)
            box3[b] = INFINITY;
    }

    /* Newton refinement (skipped like the reference when niter == 0, :992) */
    if (opts->niter > 0) {
        if (fnftb_newton(ctx, &bd, Kc, bound_states, flag) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
    } else {
        memset(flag, 0, nb * Kmax * sizeof(int32_t));
    }

    /* filter + merge per signal, :717-724 */
    for (FNFT_UINT b = 0; b < nb; b++) {
        FNFT_UINT Kb = K[b];
        FNFT_INT rc_b = FNFT_SUCCESS;
        for (FNFT_UINT i = 0; i < Kb; i++)
            if (flag[b * Kmax + i] == FNFT_EC_DIV_BY_ZERO)
                rc_b = E_DIV_BY_ZERO; /* :1020-1021 */
        if (rc_b == FNFT_SUCCESS && opts->bound_state_filtering != fnft_nsev_bsfilt_NONE) {
            const FNFT_REAL box[4] = {bd.box0, bd.box1, bd.box2, box3[b]};
            fnftb__filter_box(&Kb, bound_states + b * Kmax, box);
            fnftb__merge(&Kb, bound_states + b * Kmax, sqrt(FNFT_EPSILON));
        }
        K[b] = Kb;
        Kc[b] = (int32_t)Kb;
        if (rc_b != FNFT_SUCCESS) {
            if (ret_codes != NULL)
                ret_codes[b] = rc_b;
            if (ret_code == FNFT_SUCCESS)
                ret_code = rc_b;
            Kc[b] = 0;
        }
    }
    if (ret_code != FNFT_SUCCESS && ret_codes == NULL)
        goto leave_fun;

    /* norming constants and / or residues, :895-970 */
    if (normconsts_or_residues != NULL) {
        int any = 0;
        for (FNFT_UINT b = 0; b < nb; b++)
            any |= (Kc[b] > 0);
        if (any) {
            b_vals = malloc(nb * Kmax * sizeof(FNFT_COMPLEX));
            ap_vals = malloc(nb * Kmax * sizeof(FNFT_COMPLEX));
            if (b_vals == NULL || ap_vals == NULL) {
                ret_code = E_NOMEM;
                goto leave_fun;
            }
            if (fnftb_normconsts(ctx, &bd, Kc, bound_states, NULL, ap_vals, b_vals) != 0) {
                ret_code = E_DEVICE;
                goto leave_fun;
            }
            const FNFT_UINT nlen = (opts->discspec_type == fnft_nsev_dstype_BOTH) ? 2 * Kmax : Kmax;
            for (FNFT_UINT b = 0; b < nb; b++) {
                FNFT_COMPLEX *out = normconsts_or_residues + b * nlen;
                const FNFT_UINT Kb = (FNFT_UINT)Kc[b];
                FNFT_UINT offset = 0;
                if (opts->discspec_type == fnft_nsev_dstype_NORMING_CONSTANTS ||
                    opts->discspec_type == fnft_nsev_dstype_BOTH) {
                    for (FNFT_UINT i = 0; i < Kb; i++)
                        out[i] = b_vals[b * Kmax + i];
                    offset = Kb; /* residues follow the Kb norming constants (:950-954) */
                }
                if (opts->discspec_type == fnft_nsev_dstype_RESIDUES ||
                    opts->discspec_type == fnft_nsev_dstype_BOTH) {
                    for (FNFT_UINT i = 0; i < Kb; i++) {
                        if (ap_vals[b * Kmax + i] == 0.0) {
                            const FNFT_INT ec = E_DIV_BY_ZERO; /* :960-961 */
                            if (ret_codes != NULL)
                                ret_codes[b] = ec;
                            if (ret_code == FNFT_SUCCESS)
                                ret_code = ec;
                            break;
                        }
                        out[offset + i] = b_vals[b * Kmax + i] / ap_vals[b * Kmax + i];
                    }
                } else if (opts->discspec_type != fnft_nsev_dstype_NORMING_CONSTANTS) {
                    ret_code = E_INVALID_ARGUMENT(opts->discspec_type);
                    goto leave_fun;
                }
            }
        }
    }

leave_fun:
    free(Kc);
    free(flag);
    free(box3);
    free(b_vals);
    free(ap_vals);
    return ret_code;
}
