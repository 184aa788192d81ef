/*
 * fnft_b200 host library -- fnft_nsep and fnft_nsep_batch (GRIDSEARCH localization).
 *
 * Host-side mirror of /root/reference/src/fnft_nsep.c: argument checks (:99-115),
 * step size eps_t = (T1-T0)/D (:119), de-rotation by Lam_shift (:118-128), shift of a
 * MANUAL bounding box (:132-136), and of gridsearch (:222-436): preprocessing (:259),
 * nse_fscatter (:279), automatic bounding box (:837-864), angular range PHI (:289-295),
 * Floquet polynomials and the three grid searches (:319-326,355-360,400), coordinate
 * transform + filtering (:335-345) and the truncation warnings (:347-353,377-386).
 *
 * fnft_nsep_loc_SUBSAMPLE_AND_REFINE and fnft_nsep_loc_MIXED (the reference default) mirror
 * subsample_and_refine (:441-706): the roots of the Floquet polynomials of a subsampled signal
 * come from the GPU root finder (poly_roots.cuh, in place of eiscor), refine_mainspec /
 * refine_auxspec (:708-835) run as one-warp-per-point kernels on the full signal
 * (nsep_refine.cuh); the order-dependent filtering and the assembly of the user arrays stay
 * on the host like in the reference.
 */
#include "fnft_internal.h"
#include "fnft_nsev_discrete.h"

static const fnft_nsep_opts_t nsep_defaults = {
    .localization = fnft_nsep_loc_MIXED,
    .filtering = fnft_nsep_filt_AUTO,
    .bounding_box = {-INFINITY, INFINITY, -INFINITY, INFINITY},
    .max_evals = 20,
    .discretization = fnft_nse_discretization_2SPLIT2A,
    .normalization_flag = 1,
    .floquet_range = {-1, 1},
    .points_per_spine = 2,
    .Dsub = 0,
    .tol = -1};

fnft_nsep_opts_t fnft_nsep_default_opts(void) { return nsep_defaults; }

/* src/private/fnft__misc.c:205-226 */
/* FNFT_B200_NSEP_TIMING=1: host wall time of the phases of fnft_nsep_batch (each closed by a device sync) */
#include <stdio.h>
#include <time.h>
static int nsep_timing_on(void)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("FNFT_B200_NSEP_TIMING");
        v = (e && e[0]) ? atoi(e) : 0;
    }
    return v;
}
static void nsep_tick(fnftb_ctx *ctx, const char *label)
{
    static struct timespec last;
    if (!nsep_timing_on())
        return;
    fnftb_ctx_sync(ctx);
    struct timespec now;
    clock_gettime(CLOCK_MONOTONIC, &now);
    if (label != NULL)
        fprintf(stderr, "[nsep timing] %-28s %8.3f ms\n", label,
                (now.tv_sec - last.tv_sec) * 1e3 + (now.tv_nsec - last.tv_nsec) * 1e-6);
    last = now;
}

static void filter_nonreal(FNFT_UINT *N, FNFT_COMPLEX *vals, FNFT_REAL tol_im)
{
    FNFT_UINT kept = 0;
    for (FNFT_UINT i = 0; i < *N; i++) {
        if (!(fabs(cimag(vals[i])) > tol_im))
            continue;
        vals[kept++] = vals[i];
    }
    *N = kept;
}

/*
 * subsample_and_refine (src/fnft_nsep.c:441-706) for the nb signals of a chunk.  On entry the
 * de-rotated signals are saved in slot 1 of the context and the fully preprocessed ones in
 * slot 0.  main_out / aux_out: host rows of stride Kmax / Mmax (NULL = not wanted); K1 / M1
 * receive the numbers of points written.  box: bounding box (MANUAL: already shifted; AUTO:
 * recomputed here from the subsampled step size, :548 and returned).
 */
static FNFT_INT nsep_subsample_refine_chunk(fnftb_ctx *ctx, FNFT_UINT nb, FNFT_UINT D, FNFT_UINT upsampling,
                                            FNFT_UINT deg0, fnft__akns_discretization_t akns,
                                            FNFT_REAL eps_t, FNFT_INT kappa, fnft_nsep_opts_t const *opts,
                                            FNFT_REAL *box, int skip_real, FNFT_COMPLEX *main_out,
                                            FNFT_UINT Kmax, FNFT_UINT *K1, FNFT_COMPLEX *aux_out,
                                            FNFT_UINT Mmax, FNFT_UINT *M1, FNFT_REAL Lam_shift,
                                            int *warned, FNFT_INT *ret_codes)
{
    FNFT_INT ret_code = FNFT_SUCCESS;
    FNFT_COMPLEX *roots = NULL;
    int32_t *info = NULL, *Kn = NULL, *flag = NULL;
    char *full = NULL;

    /* number of samples of the subsampled signal, :485-497 and the clamps / rounding of
     * nse_discretization_preprocess_signal (fnft__nse_discretization.c:424-431) */
    FNFT_UINT Dsub = opts->Dsub;
    if (Dsub == 0)
        Dsub = (FNFT_UINT)pow(2.0, ceil(0.5 * log2(D * log2(D) * log2(D))));
    else
        Dsub = (FNFT_UINT)pow(2.0, round(log2(Dsub)));
    if (Dsub < 2)
        Dsub = 2;
    if (Dsub > D)
        Dsub = D;
    const FNFT_UINT nskip = (FNFT_UINT)round((FNFT_REAL)D / Dsub);
    Dsub = (FNFT_UINT)round((FNFT_REAL)D / nskip);
    const FNFT_UINT nskip_per_step = D / Dsub;
    if ((Dsub - 1) * nskip + nskip_per_step != D) /* :494-498 */
        return E_ASSERTION_FAILED;

    if (fnftb_signals_restore(ctx, 1) != 0)
        return E_DEVICE;
    if ((upsampling == 2 ? fnftb_resample_4split4_sub(ctx, eps_t, nskip, Dsub, NULL)
                         : fnftb_subsample(ctx, nskip, Dsub)) != 0)
        return E_DEVICE;
    const FNFT_REAL eps_t_sub = nskip_per_step * eps_t; /* :525 */
    fnftb_scatter_desc sd;
    memset(&sd, 0, sizeof(sd));
    sd.rmode = FNFTB_RMODE_NSE;
    sd.kappa = kappa;
    sd.scheme = (int)akns;
    sd.deg0 = (int)deg0;
    sd.normalize = opts->normalization_flag ? 1 : 0;
    sd.eps_t = eps_t_sub;
    if (fnftb_fscatter(ctx, &sd) != 0)
        return E_DEVICE;
    const FNFT_UINT deg = deg0 * Dsub * upsampling;

    const FNFT_REAL degree1step = (FNFT_REAL)deg0;
    const FNFT_REAL map_coeff = 2 / degree1step;
    if (opts->filtering == fnft_nsep_filt_AUTO) { /* update_bounding_box_if_auto(eps_t_sub, ...), :548 */
        box[1] = 0.9 * FNFT_PI / (fabs(map_coeff) * eps_t_sub);
        box[0] = -box[1];
        box[3] = -log(0.1) / (fabs(map_coeff) * eps_t_sub);
        box[2] = -box[3];
    }
    const FNFT_REAL tol_im = (box[1] - box[0]) / (32 * (D - 1));                 /* :549-550 */
    const FNFT_REAL refine_tol = (opts->tol < 0) ? sqrt(FNFT_EPSILON) : opts->tol; /* :507-510 */
    const FNFT_REAL lam_den = 2 * eps_t_sub / (degree1step * upsampling);

    roots = malloc(nb * deg * sizeof(FNFT_COMPLEX));
    info = malloc(nb * 4 * sizeof(int32_t));
    Kn = malloc(nb * sizeof(int32_t));
    flag = malloc(nb * deg * sizeof(int32_t));
    full = calloc(nb, 1);
    if (roots == NULL || info == NULL || Kn == NULL || flag == NULL || full == NULL) {
        ret_code = E_NOMEM;
        goto leave_fun;
    }
    fnftb_refine_desc rd;
    memset(&rd, 0, sizeof(rd));
    rd.upsampling = (int)upsampling; /* BO or CF4_2, :500-504 */
    rd.kappa = kappa;
    rd.Kstride = (int)deg;
    rd.max_evals = (int)opts->max_evals;
    rd.eps_t = eps_t;
    rd.tol = refine_tol;

/* z_to_lambda (fnft__akns_discretization.c:225-240) and the box filter on the device, only the
 * survivors are copied (row stride deg); misc_filter_nonreal here */
#define NSEP_POSTPROCESS_ROOTS(do_nonreal)                                                   \
    if (fnftb_roots_lambda(ctx, lam_den, opts->filtering != fnft_nsep_filt_NONE ? box : NULL, 0, roots, deg, \
                           Kn) != 0) {                                                        \
        ret_code = E_DEVICE;                                                                  \
        goto leave_fun;                                                                       \
    }                                                                                         \
    for (FNFT_UINT b = 0; b < nb; b++) {                                                      \
        FNFT_COMPLEX *buf = roots + b * deg;                                                  \
        FNFT_UINT Kb = (FNFT_UINT)Kn[b];                                                      \
        if (do_nonreal)                                                                       \
            filter_nonreal(&Kb, buf, tol_im);                                                 \
        Kn[b] = (int32_t)Kb;                                                                  \
    }
#define NSEP_CHECK_FLAGS()                                                                   \
    for (FNFT_UINT b = 0; b < nb; b++)                                                        \
        for (FNFT_UINT i = 0; i < (FNFT_UINT)Kn[b]; i++)                                      \
            if (flag[b * deg + i] == FNFT_EC_DIV_BY_ZERO) {                                   \
                const FNFT_INT ec = E_DIV_BY_ZERO; /* :741-742, :820-821 */                   \
                if (ret_codes != NULL)                                                        \
                    ret_codes[b] = ec;                                                        \
                if (ret_code == FNFT_SUCCESS)                                                 \
                    ret_code = ec;                                                            \
                Kn[b] = 0;                                                                    \
                break;                                                                        \
            }

    if (main_out != NULL) { /* :553-640 */
        const FNFT_REAL rhs_0 = opts->floquet_range[0], rhs_1 = opts->floquet_range[1];
        const FNFT_UINT nvals = opts->points_per_spine;
        FNFT_REAL rhs_step = rhs_1 - rhs_0;
        if (nvals > 1)
            rhs_step /= nvals - 1;
        for (FNFT_UINT nval = 0; nval < nvals; nval++) {
            const FNFT_REAL rhs = 2.0 * (rhs_0 + nval * rhs_step);
            if (fnftb_nsep_floquet_roots(ctx, rhs, NULL, NULL) != 0) {
                ret_code = E_DEVICE;
                goto leave_fun;
            }
            NSEP_POSTPROCESS_ROOTS(skip_real)
            rd.mode = 0;
            rd.rhs = -rhs;
            if (fnftb_signals_restore(ctx, 0) != 0 || fnftb_nsep_refine(ctx, &rd, Kn, roots, flag) != 0) {
                ret_code = E_DEVICE;
                goto leave_fun;
            }
            NSEP_CHECK_FLAGS()
            if (ret_code != FNFT_SUCCESS && ret_codes == NULL)
                goto leave_fun;
            for (FNFT_UINT b = 0; b < nb; b++) {
                if (full[b])
                    continue; /* user-provided array is full, :637-638 */
                FNFT_COMPLEX *buf = roots + b * deg;
                FNFT_UINT K_new = (FNFT_UINT)Kn[b];
                if (opts->filtering != fnft_nsep_filt_NONE)
                    fnftb__filter_box(&K_new, buf, box);
                if (skip_real)
                    filter_nonreal(&K_new, buf, tol_im);
                if (K1[b] + K_new > Kmax) {
                    if (!warned[0]) {
                        WARN("Found more than *K_ptr main spectrum points. Returning as many as possible.");
                        warned[0] = 1;
                    }
                    K_new = Kmax - K1[b];
                    full[b] = 1;
                }
                for (FNFT_UINT i = 0; i < K_new; i++)
                    main_out[b * Kmax + K1[b] + i] = buf[i] + Lam_shift;
                K1[b] += K_new;
            }
        }
    }
    if (aux_out != NULL) { /* :642-691 */
        if (fnftb_poly_roots(ctx, 1, NULL, NULL) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        NSEP_POSTPROCESS_ROOTS(0)
        rd.mode = 1;
        rd.rhs = 0.0;
        if (fnftb_signals_restore(ctx, 0) != 0 || fnftb_nsep_refine(ctx, &rd, Kn, roots, flag) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        NSEP_CHECK_FLAGS()
        if (ret_code != FNFT_SUCCESS && ret_codes == NULL)
            goto leave_fun;
        for (FNFT_UINT b = 0; b < nb; b++) {
            FNFT_COMPLEX *buf = roots + b * deg;
            FNFT_UINT Mb = (FNFT_UINT)Kn[b];
            if (opts->filtering != fnft_nsep_filt_NONE)
                fnftb__filter_box(&Mb, buf, box);
            if (skip_real)
                filter_nonreal(&Mb, buf, tol_im);
            if (Mb > Mmax) {
                if (!warned[1]) {
                    WARN("Found more than *M_ptr aux spectrum points. Returning as many as possible.");
                    warned[1] = 1;
                }
                Mb = Mmax;
            }
            for (FNFT_UINT i = 0; i < Mb; i++)
                aux_out[b * Mmax + i] = buf[i] + Lam_shift;
            M1[b] = Mb;
        }
    }
#undef NSEP_POSTPROCESS_ROOTS
#undef NSEP_CHECK_FLAGS

leave_fun:
    free(roots);
    free(info);
    free(Kn);
    free(flag);
    free(full);
    return ret_code;
}

/* shared by the single-signal and the batched entry point; box_out receives the
 * bounding box actually used (the reference writes it back into *opts) */
static FNFT_INT nsep_core(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                          FNFT_REAL const *const T, FNFT_REAL const phase_shift, FNFT_UINT *const K,
                          const FNFT_UINT Kmax, FNFT_COMPLEX *const main_spec,
                          FNFT_UINT *const Mcount, const FNFT_UINT Mmax,
                          FNFT_COMPLEX *const aux_spec, const FNFT_INT kappa,
                          fnft_nsep_opts_t const *opts, FNFT_INT *const ret_codes,
                          FNFT_REAL *box_out)
{
    FNFT_INT ret_code = FNFT_SUCCESS;
    int32_t *status = NULL;
    uint64_t *Kc = NULL, *Mc = NULL;
    FNFT_UINT *K1 = NULL, *M1 = NULL;
    FNFT_COMPLEX *main_tmp = NULL, *aux_tmp = NULL;

    if (B == 0)
        return E_INVALID_ARGUMENT(B);
    if (D < 2 || (D & (D - 1)) != 0) /* src/fnft_nsep.c:99 */
        return E_INVALID_ARGUMENT(D);
    if (q == NULL)
        return E_INVALID_ARGUMENT(q);
    if (T == NULL || T[0] >= T[1])
        return E_INVALID_ARGUMENT(T);
    if (abs(kappa) != 1)
        return E_INVALID_ARGUMENT(kappa);
    if (K == NULL)
        return E_INVALID_ARGUMENT(K_ptr);
    if (Mcount == NULL)
        return E_INVALID_ARGUMENT(M_ptr);
    if (opts == NULL)
        opts = &nsep_defaults;
    if (opts->filtering != fnft_nsep_filt_NONE && main_spec == NULL && aux_spec != NULL)
        return E_INVALID_ARGUMENT(main_spec.Filtering of the auxiliary spectrum is not possible if the main spectrum is not computed.);
    const fnft_nsep_loc_t loc = opts->localization;
    if (loc != fnft_nsep_loc_GRIDSEARCH && loc != fnft_nsep_loc_SUBSAMPLE_AND_REFINE &&
        loc != fnft_nsep_loc_MIXED)
        return E_INVALID_ARGUMENT(opts_ptr->discretization); /* sic, src/fnft_nsep.c:191 */
    const int do_sub = (loc != fnft_nsep_loc_GRIDSEARCH), do_grid = (loc != fnft_nsep_loc_SUBSAMPLE_AND_REFINE);

    fnft__akns_discretization_t akns;
    ret_code = fnftb__nse_to_akns(opts->discretization, &akns);
    if (ret_code != FNFT_SUCCESS)
        return E_INVALID_ARGUMENT(opts->discretization);
    const FNFT_UINT deg0 = fnftb__akns_degree(akns);
    const FNFT_UINT upsampling = fnftb__akns_upsampling(akns);
    if (deg0 == 0 || !fnftb__akns_on_gpu(akns) || upsampling > 2)
        return E_NOT_YET_IMPLEMENTED(opts->discretization,
                                     This discretization has no GPU leaf kernel yet.);

    fnftb_ctx *ctx = fnftb__ctx();
    if (ctx == NULL)
        return E_OTHER("No usable CUDA device: the fnft_b200 hot path has no CPU fallback.");

    const FNFT_REAL Lam_shift = phase_shift / (-2 * (T[1] - T[0]));
    const FNFT_REAL eps_t = (T[1] - T[0]) / D;
    const FNFT_UINT D_eff = D * upsampling;
    const FNFT_REAL degree1step = (FNFT_REAL)deg0;
    const FNFT_REAL map_coeff = 2 / degree1step;

    /* bounding box: MANUAL boxes are shifted (:132-136), AUTO boxes computed (:856-863) */
    FNFT_REAL box[4] = {opts->bounding_box[0], opts->bounding_box[1], opts->bounding_box[2],
                        opts->bounding_box[3]};
    if (opts->filtering == fnft_nsep_filt_MANUAL) {
        box[0] -= Lam_shift;
        box[1] -= Lam_shift;
    } else if (opts->filtering == fnft_nsep_filt_AUTO) {
        box[1] = 0.9 * FNFT_PI / (fabs(map_coeff) * eps_t);
        box[0] = -box[1];
        box[3] = -log(0.1) / (fabs(map_coeff) * eps_t);
        box[2] = -box[3];
    }
    if (box_out != NULL)
        memcpy(box_out, box, sizeof(box));
    FNFT_REAL PHI[2] = {map_coeff * eps_t * box[0], map_coeff * eps_t * box[1]};
    if (PHI[0] > PHI[1]) {
        const FNFT_REAL tmp = PHI[0];
        PHI[0] = PHI[1];
        PHI[1] = tmp;
    }
    /* argument check of poly_roots_fftgridsearch (fftgridsearch.c:53-55) */
    if (do_grid && (!(PHI[0] < PHI[1]) || PHI[0] == -INFINITY || PHI[1] == INFINITY))
        return E_SUBROUTINE(E_INVALID_ARGUMENT(PHI));

    fnftb_scatter_desc sd;
    memset(&sd, 0, sizeof(sd));
    sd.rmode = FNFTB_RMODE_NSE;
    sd.kappa = kappa;
    sd.scheme = (int)akns;
    sd.deg0 = (int)deg0;
    sd.normalize = opts->normalization_flag ? 1 : 0;
    sd.eps_t = eps_t;

    fnftb_nsep_desc nd;
    memset(&nd, 0, sizeof(nd));
    nd.PHI0 = PHI[0];
    nd.PHI1 = PHI[1];
    nd.lam_den = 2 * eps_t / (degree1step * upsampling);
    nd.filtering = (opts->filtering != fnft_nsep_filt_NONE);
    memcpy(nd.box, box, sizeof(box));
    nd.lam_shift = Lam_shift;
    nd.Kmax = Kmax;
    nd.Mmax = Mmax;

    size_t chunk = fnftb_nsep_chunk(ctx, D_eff, (int)deg0, fnftb__workspace_limit());
    if (chunk > B)
        chunk = B;
    status = malloc(chunk * sizeof(int32_t));
    Kc = malloc(chunk * sizeof(uint64_t));
    Mc = malloc(chunk * sizeof(uint64_t));
    if (status == NULL || Kc == NULL || Mc == NULL) {
        ret_code = E_NOMEM;
        goto leave_fun;
    }
    int warned_main = 0, warned_aux = 0;
    int warned_sub[2] = {0, 0};
    K1 = calloc(chunk, sizeof(FNFT_UINT));
    M1 = calloc(chunk, sizeof(FNFT_UINT));
    if (K1 == NULL || M1 == NULL) {
        ret_code = E_NOMEM;
        goto leave_fun;
    }
    if (do_sub && do_grid) { /* MIXED: the grid search results are appended on the host */
        main_tmp = main_spec ? malloc(chunk * Kmax * sizeof(FNFT_COMPLEX)) : NULL;
        aux_tmp = aux_spec ? malloc(chunk * Mmax * sizeof(FNFT_COMPLEX)) : NULL;
        if ((main_spec && main_tmp == NULL) || (aux_spec && aux_tmp == NULL)) {
            ret_code = E_NOMEM;
            goto leave_fun;
        }
    }
    for (FNFT_UINT b0 = 0; b0 < B; b0 += chunk) {
        const FNFT_UINT nb = (B - b0 < chunk) ? (B - b0) : chunk;
        nsep_tick(ctx, b0 == 0 ? "(setup)" : "(chunk epilogue)");
        if (fnftb_set_signals(ctx, nb, D, q + b0 * D, NULL, 0) != 0 ||
            fnftb_nsep_derotate(ctx, Lam_shift, T[0], eps_t) != 0 || fnftb_signals_save(ctx, 1) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        if (upsampling == 2) {
            int32_t *warn = malloc(nb * sizeof(int32_t));
            if (fnftb_resample_4split4(ctx, eps_t, warn) != 0) {
                free(warn);
                ret_code = E_DEVICE;
                goto leave_fun;
            }
            for (FNFT_UINT b = 0; warn != NULL && b < nb; b++) {
                if (warn[b]) {
                    WARN("Signal does not appear to be bandlimited. Interpolation step may be inaccurate. Try to reduce the step size, or switch to a discretization that does not require interpolation");
                    break;
                }
            }
            free(warn);
        }
        if (fnftb_signals_save(ctx, 0) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        memset(K1, 0, nb * sizeof(FNFT_UINT));
        memset(M1, 0, nb * sizeof(FNFT_UINT));
        if (do_sub) {
            /* MIXED computes only the non-real points here (skip_real_flag, :155-167); no non-real
             * main spectrum exists in the defocusing case */
            const int skip_real = do_grid;
            FNFT_COMPLEX *main_dst = (main_spec != NULL && !(do_grid && kappa == -1)) ? main_spec + b0 * Kmax : NULL;
            FNFT_REAL box_sub[4];
            memcpy(box_sub, box, sizeof(box));
            if (opts->filtering == fnft_nsep_filt_AUTO && !do_grid) {
                /* nothing: recomputed inside from the subsampled step size */
            }
            ret_code = nsep_subsample_refine_chunk(ctx, nb, D, upsampling, deg0, akns, eps_t, kappa, opts, box_sub,
                                                   skip_real, main_dst, Kmax, K1, aux_spec ? aux_spec + b0 * Mmax : NULL,
                                                   Mmax, M1, Lam_shift, warned_sub, ret_codes ? ret_codes + b0 : NULL);
            if (ret_code != FNFT_SUCCESS && ret_codes == NULL) {
                ret_code = E_SUBROUTINE(ret_code);
                goto leave_fun;
            }
            if (!do_grid) {
                if (box_out != NULL)
                    memcpy(box_out, box_sub, sizeof(box_sub));
                for (FNFT_UINT b = 0; b < nb; b++) {
                    K[b0 + b] = K1[b];
                    Mcount[b0 + b] = M1[b];
                }
                continue;
            }
            if (fnftb_signals_restore(ctx, 0) != 0) {
                ret_code = E_DEVICE;
                goto leave_fun;
            }
        }
        nsep_tick(ctx, "stage + preprocess");
        if (fnftb_fscatter(ctx, &sd) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        nsep_tick(ctx, "fscatter");
        memset(Kc, 0, nb * sizeof(uint64_t));
        memset(Mc, 0, nb * sizeof(uint64_t));
        FNFT_COMPLEX *gs_main = main_spec ? (do_sub ? main_tmp : main_spec + b0 * Kmax) : NULL;
        FNFT_COMPLEX *gs_aux = aux_spec ? (do_sub ? aux_tmp : aux_spec + b0 * Mmax) : NULL;
        if (fnftb_nsep_gridsearch(ctx, &nd, Kc, gs_main, Mc, gs_aux, status) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        nsep_tick(ctx, "gridsearch");
        for (FNFT_UINT b = 0; b < nb; b++) {
            if (do_sub) {
                /* MIXED: real points from the grid search behind the non-real ones, in what is
                 * left of the user arrays (:171-185) */
                FNFT_UINT K2 = (FNFT_UINT)Kc[b], M2 = (FNFT_UINT)Mc[b];
                if (main_spec != NULL) {
                    if (K1[b] + K2 > Kmax) {
                        K2 = Kmax - K1[b];
                        status[b] |= 16;
                    }
                    memcpy(main_spec + (b0 + b) * Kmax + K1[b], main_tmp + b * Kmax, K2 * sizeof(FNFT_COMPLEX));
                }
                if (aux_spec != NULL) {
                    if (M1[b] + M2 > Mmax) {
                        M2 = Mmax - M1[b];
                        status[b] |= 32;
                    }
                    memcpy(aux_spec + (b0 + b) * Mmax + M1[b], aux_tmp + b * Mmax, M2 * sizeof(FNFT_COMPLEX));
                }
                Kc[b] = K1[b] + K2;
                Mc[b] = M1[b] + M2;
            }
            K[b0 + b] = (FNFT_UINT)Kc[b];
            Mcount[b0 + b] = (FNFT_UINT)Mc[b];
            if (status[b] == 1) {
                const FNFT_INT ec = E_OTHER("Found more roots than memory is available.");
                if (ret_codes != NULL)
                    ret_codes[b0 + b] = ec;
                if (ret_code == FNFT_SUCCESS)
                    ret_code = ec;
                continue;
            }
            if ((status[b] & 16) && !warned_main) {
                WARN("Found more than *K_ptr main spectrum points. Returning as many as possible.");
                warned_main = 1;
            }
            if ((status[b] & 32) && !warned_aux) {
                WARN("Found more than *M_ptr aux spectrum points. Returning as many as possible.");
                warned_aux = 1;
            }
        }
        if (ret_code != FNFT_SUCCESS && ret_codes == NULL)
            goto leave_fun;
    }

leave_fun:
    free(status);
    free(Kc);
    free(Mc);
    free(K1);
    free(M1);
    free(main_tmp);
    free(aux_tmp);
    return ret_code;
}

/* one shard of a batch that fnft_b200_set_devices spreads over several GPUs (fnft_runtime.c) */
typedef struct {
    FNFT_UINT B, D, Kmax, Mmax;
    FNFT_COMPLEX const *q;
    FNFT_REAL const *T;
    FNFT_REAL phase_shift;
    FNFT_UINT *K, *Mcount;
    FNFT_COMPLEX *main_spec, *aux_spec;
    FNFT_INT kappa;
    fnft_nsep_opts_t const *opts;
    FNFT_INT *ret_codes;
    FNFT_INT rc[16];
} nsep_job;

static void nsep_shard(void *arg, int shard, int nshards)
{
    nsep_job *j = (nsep_job *)arg;
    FNFT_UINT b0, b1;
    fnftb__shard_range(j->B, shard, nshards, &b0, &b1);
    j->rc[shard] = fnft_nsep_batch(b1 - b0, j->D, j->q + b0 * j->D, j->T, j->phase_shift,
                                   j->K ? j->K + b0 : NULL, j->Kmax,
                                   j->main_spec ? j->main_spec + b0 * j->Kmax : NULL,
                                   j->Mcount ? j->Mcount + b0 : NULL, j->Mmax,
                                   j->aux_spec ? j->aux_spec + b0 * j->Mmax : NULL, j->kappa, j->opts,
                                   j->ret_codes ? j->ret_codes + b0 : NULL);
}

static FNFT_INT nsep_batch_impl(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                         FNFT_REAL const *const T, FNFT_REAL const phase_shift,
                         FNFT_UINT *const K, const FNFT_UINT Kmax, FNFT_COMPLEX *const main_spec,
                         FNFT_UINT *const Mcount, const FNFT_UINT Mmax,
                         FNFT_COMPLEX *const aux_spec, const FNFT_INT kappa,
                         fnft_nsep_opts_t const *opts, FNFT_INT *const ret_codes);

/* public entry: NVTX range around the call (FNFT_B200_NVTX=1, no-op otherwise) */
FNFT_INT fnft_nsep_batch(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                         FNFT_REAL const *const T, FNFT_REAL const phase_shift,
                         FNFT_UINT *const K, const FNFT_UINT Kmax, FNFT_COMPLEX *const main_spec,
                         FNFT_UINT *const Mcount, const FNFT_UINT Mmax,
                         FNFT_COMPLEX *const aux_spec, const FNFT_INT kappa,
                         fnft_nsep_opts_t const *opts, FNFT_INT *const ret_codes)
{
    fnftb_range_push("fnft_nsep_batch");
    const FNFT_INT rc = nsep_batch_impl(B, D, q, T, phase_shift, K, Kmax, main_spec, Mcount, Mmax, aux_spec, kappa, opts, ret_codes);
    fnftb_range_pop();
    return rc;
}

static FNFT_INT nsep_batch_impl(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                         FNFT_REAL const *const T, FNFT_REAL const phase_shift,
                         FNFT_UINT *const K, const FNFT_UINT Kmax, FNFT_COMPLEX *const main_spec,
                         FNFT_UINT *const Mcount, const FNFT_UINT Mmax,
                         FNFT_COMPLEX *const aux_spec, const FNFT_INT kappa,
                         fnft_nsep_opts_t const *opts, FNFT_INT *const ret_codes)
{
    /* the periodic transform assembles variable-length point lists on the host: host buffers only */
    if (fnftb__device_pointers())
        return E_NOT_YET_IMPLEMENTED(device pointers, fnft_nsep_batch takes host buffers;
                                     call fnft_b200_set_device_pointers(0) first.);
    const int nshards = (q != NULL) ? fnftb__fanout_shards(B) : 1;
    if (nshards > 1) { /* several GPUs: every shard is this same call on its own device */
        nsep_job j = {B, D, Kmax, Mmax, q, T, phase_shift, K, Mcount, main_spec, aux_spec, kappa, opts, ret_codes,
                      {0}};
        if (fnftb__fanout_run(nshards, nsep_shard, &j) != 0)
            return E_OTHER("Could not start the per-device worker threads.");
        for (int i = 0; i < nshards; i++)
            if (j.rc[i] != FNFT_SUCCESS)
                return E_SUBROUTINE(j.rc[i]);
        return FNFT_SUCCESS;
    }
    if (ret_codes != NULL)
        for (FNFT_UINT b = 0; b < B; b++)
            ret_codes[b] = FNFT_SUCCESS;
    return nsep_core(B, D, q, T, phase_shift, K, Kmax, main_spec, Mcount, Mmax, aux_spec, kappa,
                     opts, ret_codes, NULL);
}

/* include/fnft_nsep.h:263-267, src/fnft_nsep.c:82-218 */
FNFT_INT fnft_nsep(const FNFT_UINT D, FNFT_COMPLEX const *const q, FNFT_REAL const *const T,
                   FNFT_REAL const phase_shift, FNFT_UINT *const K_ptr,
                   FNFT_COMPLEX *const main_spec, FNFT_UINT *const M_ptr,
                   FNFT_COMPLEX *const aux_spec, FNFT_REAL *const sheet_indices,
                   const FNFT_INT kappa, fnft_nsep_opts_t *opts)
{
    if (K_ptr == NULL)
        return E_INVALID_ARGUMENT(K_ptr);
    if (M_ptr == NULL)
        return E_INVALID_ARGUMENT(M_ptr);
    if (sheet_indices != NULL)
        return E_NOT_YET_IMPLEMENTED(sheet_indices, Pass sheet_indices = "NULL".);
    FNFT_REAL box[4];
    FNFT_UINT Kc = 0, Mc = 0;
    const FNFT_INT rc = nsep_core(1, D, q, T, phase_shift, &Kc, *K_ptr, main_spec, &Mc, *M_ptr,
                                  aux_spec, kappa, opts, NULL, box);
    if (rc == FNFT_SUCCESS) {
        *K_ptr = Kc;
        *M_ptr = Mc;
        /* the reference leaves the automatically chosen box in *opts (:858-863) */
        if (opts != NULL && opts->filtering == fnft_nsep_filt_AUTO)
            memcpy(opts->bounding_box, box, sizeof(box));
    }
    return rc;
}
