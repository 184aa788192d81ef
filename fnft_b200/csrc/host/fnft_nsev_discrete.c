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

/* bounding box of src/fnft_nsev.c:628-659: box[0..2] common, box3[b] per signal */
static FNFT_INT nsev_bounding_boxes(fnftb_ctx *ctx, FNFT_UINT nb, FNFT_UINT upsampling, FNFT_REAL const *T,
                                    FNFT_REAL eps_t, fnft_nsev_opts_t const *opts, double *box,
                                    double *box3, int *use_box3)
{
    const FNFT_REAL degree1step = (FNFT_REAL)fnftb__nse_degree(opts->discretization);
    const FNFT_REAL map_coeff = (degree1step != 0) ? 2 / degree1step : 2.0;
    *use_box3 = 0;
    if (opts->bound_state_filtering == fnft_nsev_bsfilt_FULL) {
        box[1] = 0.9 * FNFT_PI / fabs(map_coeff * eps_t);
        box[0] = -box[1];
        box[2] = 0.0;
        *use_box3 = 1;
        if (fnftb_imbound(ctx, (int)upsampling, T[0], T[1], box3) != 0)
            return E_DEVICE;
    } else if (opts->bound_state_filtering == fnft_nsev_bsfilt_BASIC) {
        box[0] = -INFINITY;
        box[1] = INFINITY;
        box[2] = 0.0;
        for (FNFT_UINT b = 0; b < nb; b++)
            box3[b] = INFINITY;
    } else {
        box[0] = -INFINITY;
        box[1] = INFINITY;
        box[2] = -INFINITY;
        for (FNFT_UINT b = 0; b < nb; b++)
            box3[b] = INFINITY;
    }
    return FNFT_SUCCESS;
}

FNFT_INT fnftb__nsev_fasteig_chunk(fnftb_ctx *ctx, FNFT_UINT nb, FNFT_UINT D_eff, FNFT_UINT D_given,
                                   FNFT_REAL const *T, FNFT_INT kappa, int have_tm, FNFT_UINT *K,
                                   FNFT_UINT Kmax, FNFT_COMPLEX *bound_states,
                                   fnft_nsev_opts_t const *opts)
{
    FNFT_INT ret_code = FNFT_SUCCESS;
    FNFT_COMPLEX *roots = NULL;
    int32_t *info = NULL;
    double *box3 = NULL;
    const FNFT_UINT upsampling = D_eff / D_given;
    const FNFT_REAL eps_t = (T[1] - T[0]) / (D_given - 1);
    fnft__akns_discretization_t akns;
    if (fnftb__nse_to_akns(opts->discretization, &akns) != FNFT_SUCCESS)
        return E_INVALID_ARGUMENT(opts->discretization);
    const FNFT_UINT deg0 = fnftb__akns_degree(akns);
    if (deg0 == 0) /* slow discretizations only support Newton (include/fnft_nse_discretization_t.h) */
        return E_INVALID_ARGUMENT(opts->bound_state_localization);
    const FNFT_UINT deg = deg0 * D_eff;

    if (!have_tm) { /* nse_fscatter of fnft_nsev_base, src/fnft_nsev.c:527 */
        fnftb_scatter_desc sd;
        memset(&sd, 0, sizeof(sd));
        sd.rmode = FNFTB_RMODE_NSE;
        sd.kappa = kappa;
        sd.scheme = (int)akns;
        sd.deg0 = (int)deg0;
        sd.normalize = opts->normalization_flag ? 1 : 0;
        sd.eps_t = eps_t;
        if (fnftb_fscatter(ctx, &sd) != 0)
            return E_DEVICE;
    }
    /* poly_roots_fasteigen(deg, transfer_matrix, buffer), :702 -- the roots stay on the device */
    if (fnftb_poly_roots(ctx, 0, NULL, NULL) != 0)
        return E_DEVICE;
    box3 = malloc(nb * sizeof(double));
    info = malloc(nb * sizeof(int32_t));
    if (box3 == NULL || info == NULL) {
        ret_code = E_NOMEM;
        goto leave_fun;
    }
    double box[4];
    int use_box3;
    ret_code = nsev_bounding_boxes(ctx, nb, upsampling, T, eps_t, opts, box, box3, &use_box3);
    CHECK_RETCODE(ret_code, leave_fun);
    /* z -> lambda (src/private/fnft__akns_discretization.c:225-240) and the box filter (:717-720) on
     * the device; merge (order dependent, :721-723) here.  Room for every root is only needed when
     * nothing is filtered. */
    const FNFT_REAL degree1step = (FNFT_REAL)(deg0 * upsampling);
    const int filtering = (opts->bound_state_filtering != fnft_nsev_bsfilt_NONE);
    FNFT_UINT stride = filtering ? (deg < 4096 ? deg : 4096) : deg;
    box[3] = INFINITY; /* per-signal upper bounds come from box3 */
    for (;;) {
        roots = malloc(nb * stride * sizeof(FNFT_COMPLEX));
        if (roots == NULL) {
            ret_code = E_NOMEM;
            goto leave_fun;
        }
        if (fnftb_roots_lambda(ctx, 2 * eps_t / degree1step, filtering ? box : NULL, filtering && use_box3,
                               roots, stride, info) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        FNFT_UINT mx = 0;
        for (FNFT_UINT b = 0; b < nb; b++)
            if ((FNFT_UINT)info[b] > mx)
                mx = (FNFT_UINT)info[b];
        if (mx <= stride)
            break;
        free(roots); /* more survivors than expected: once more with room for all of them */
        roots = NULL;
        stride = mx;
    }
    for (FNFT_UINT b = 0; b < nb; b++) {
        FNFT_COMPLEX *buf = roots + b * stride;
        FNFT_UINT Kb = (FNFT_UINT)info[b];
        if (filtering) {
            if (!use_box3) { /* BASIC: the common box only */
                const FNFT_REAL bx[4] = {box[0], box[1], box[2], INFINITY};
                fnftb__filter_box(&Kb, buf, bx);
            }
            fnftb__merge(&Kb, buf, sqrt(FNFT_EPSILON));
        }
        if (Kb > Kmax) { /* :728-731 */
            WARN("Found more than *K_ptr bound states. Returning as many as possible.");
            Kb = Kmax;
        }
        memcpy(bound_states + b * Kmax, buf, Kb * sizeof(FNFT_COMPLEX));
        K[b] = Kb;
    }

leave_fun:
    free(roots);
    free(info);
    free(box3);
    return ret_code;
}

FNFT_INT fnftb__nsev_discrete_chunk(fnftb_ctx *ctx, FNFT_UINT nb, FNFT_UINT D_eff, FNFT_UINT D_given,
                                    FNFT_REAL const *T, FNFT_REAL eps_t, FNFT_UINT *K,
                                    FNFT_UINT Kmax, FNFT_COMPLEX *bound_states,
                                    FNFT_COMPLEX *normconsts_or_residues,
                                    fnft_nsev_opts_t const *opts, FNFT_INT *ret_codes, int skip_newton)
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
    bd.upsampling = (int)upsampling; /* BO for upsampling 1, CF4_2 for 2 (:675-680), CF4_3 itself for 3 */
    bd.Kmax = (int)Kmax;
    bd.T0 = T[0];
    bd.T1 = T[1];
    bd.eps_t = eps_t;
    bd.bc = 0.5;
    bd.lweight = (upsampling == 2) ? 0.5 : 1.0;
    /* fnft__nse_scatter_bound_states.c:225,235,247; ES4 / TES4: scl_factor = 1 (:132,157) */
    bd.scl = (opts->discretization == fnft_nse_discretization_ES4 ||
              opts->discretization == fnft_nse_discretization_TES4)
                 ? 1.0
                 : 1.0 / upsampling;
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
    if (opts->niter > 0 && !skip_newton) {
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
        if (rc_b == FNFT_SUCCESS && opts->bound_state_filtering != fnft_nsev_bsfilt_NONE && !skip_newton) {
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
