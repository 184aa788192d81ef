/*
 * fnft_b200 host library -- fnft_nsev and fnft_nsev_batch.
 *
 * Host-side mirror of /root/reference/src/fnft_nsev.c: argument checks (:162-220),
 * step size (:237), option logic, and the sequence
 *   fscatter -> continuous spectrum -> bound states -> norming constants
 * of fnft_nsev_base (:458-565).  All numerical work is done by CUDA kernels reached
 * through fnftb_device.h; the single-signal entry point is the B = 1 case of the
 * batched one.
 */
#include "fnft_internal.h"
#include "fnft_nsev_discrete.h"

static const fnft_nsev_opts_t nsev_defaults = {
    .bound_state_filtering = fnft_nsev_bsfilt_FULL,
    .bound_state_localization = fnft_nsev_bsloc_SUBSAMPLE_AND_REFINE,
    .niter = 10,
    .Dsub = 0,
    .discspec_type = fnft_nsev_dstype_NORMING_CONSTANTS,
    .contspec_type = fnft_nsev_cstype_REFLECTION_COEFFICIENT,
    .normalization_flag = 1,
    .discretization = fnft_nse_discretization_2SPLIT4B,
    .richardson_extrapolation_flag = 0};

/* src/fnft_nsev.c:42-45 */
fnft_nsev_opts_t fnft_nsev_default_opts(void) { return nsev_defaults; }

/* src/fnft_nsev.c:51-57 */
FNFT_UINT fnft_nsev_max_K(const FNFT_UINT D, fnft_nsev_opts_t const *const opts)
{
    const fnft_nse_discretization_t d = opts ? opts->discretization : nsev_defaults.discretization;
    return fnftb__nse_degree(d) * D;
}

static FNFT_UINT contspec_len(fnft_nsev_cstype_t t, FNFT_UINT M)
{
    switch (t) {
    case fnft_nsev_cstype_REFLECTION_COEFFICIENT: return M;
    case fnft_nsev_cstype_AB: return 2 * M;
    case fnft_nsev_cstype_BOTH: return 3 * M;
    default: return 0;
    }
}

/*
 * Continuous spectrum of the signals currently staged + scattered in ctx.
 * Mirrors nsev_compute_contspec (src/fnft_nsev.c:744-891) for the polynomial
 * (fast) discretizations: chirp constants :822-827, epilogue :846-876.
 */
static FNFT_INT nsev_contspec_chunk_eps(fnftb_ctx *ctx, FNFT_UINT D_given, FNFT_REAL const *T, FNFT_REAL eps_t,
                                        FNFT_UINT M, FNFT_REAL const *XI, fnft_nsev_opts_t const *opts,
                                        FNFT_COMPLEX *out, int on_device, int32_t *status, int unit_circle)
{
    const fnft_nse_discretization_t disc = opts->discretization;
    const FNFT_REAL step_div = (FNFT_REAL)(fnftb__nse_degree(disc) * fnftb__nse_upsampling(disc));
    const FNFT_REAL eps_xi = (XI[1] - XI[0]) / (M - 1);
    fnftb_contspec_desc cd;
    memset(&cd, 0, sizeof(cd));
    cd.mode = FNFTB_MODE_NSEV;
    cd.cstype = (int)opts->contspec_type;
    cd.npoly = 2;
    cd.ent[0] = 0; /* H11 = a-polynomial */
    cd.ent[1] = 2; /* H21 = b-polynomial */
    cd.M = M;
    /* z = exp(2i*lambda*eps_t/(deg*up)): V from eps_xi, A from -XI[0]; formed as
     * rounded complex doubles exactly like the reference, then taken apart */
    const FNFT_COMPLEX V = cexp(2 * I * eps_xi * eps_t / step_div);
    const FNFT_COMPLEX A = cexp(2 * I * (-XI[0]) * eps_t / step_div);
    fnftb__logpolar(V, &cd.lwr, &cd.lwi);
    fnftb__logpolar(A, &cd.lar, &cd.lai);
    /* The rounded V and A are a few 1e-17 off the unit circle, and the reference evaluates at A V^-m as rounded
     * (|z_m|^deg = 1 + O(m deg 1e-17)).  The chaining of segments relies on |z| = 1 (second row of a piece from
     * the conjugates of the first), so there the points are put exactly on the circle. */
    if (unit_circle)
        cd.lwr = cd.lar = 0.0;
    cd.xi0 = XI[0];
    cd.eps_xi = eps_xi;
    FNFT_INT ret_code;
    ret_code = fnftb__nse_phase_factor_rho(eps_t, T[1], &cd.ph_rho, disc);
    if (ret_code != FNFT_SUCCESS)
        return E_SUBROUTINE(ret_code);
    ret_code = fnftb__nse_phase_factor_a(eps_t, D_given, T, &cd.ph_a, disc);
    if (ret_code != FNFT_SUCCESS)
        return E_SUBROUTINE(ret_code);
    ret_code = fnftb__nse_phase_factor_b(eps_t, D_given, T, &cd.ph_b, disc);
    if (ret_code != FNFT_SUCCESS)
        return E_SUBROUTINE(ret_code);
    if (fnftb_contspec(ctx, &cd, out, contspec_len(opts->contspec_type, M), on_device, status) != 0)
        return E_DEVICE;
    return FNFT_SUCCESS;
}

static FNFT_INT nsev_contspec_chunk(fnftb_ctx *ctx, FNFT_UINT D_given, FNFT_REAL const *T,
                                    FNFT_UINT M, FNFT_REAL const *XI, fnft_nsev_opts_t const *opts,
                                    FNFT_COMPLEX *out, int on_device, int32_t *status)
{
    return nsev_contspec_chunk_eps(ctx, D_given, T, (T[1] - T[0]) / (D_given - 1), M, XI, opts, out, on_device,
                                   status, 0);
}

/*
 * Continuous spectrum of nb signals that are LONGER than one product tree can hold (transfer matrix of degree
 * > 2^18; the reference multiplies polynomials of any length, src/private/fnft__poly_fmult.c:404-445).  The
 * signals are cut into nseg pieces of nearly equal length; piece s (samples n0 .. n1-1, window
 * [T0 + n0 eps_t, T0 + (n1-1) eps_t], same step) goes through the normal path -- leaves, product tree,
 * chirp-z, contspec type AB -- which gives its scattering coefficients (a_s, b_s) on the xi grid, and the
 * pieces are chained on the device in order of increasing time (fnftb_seg_compose).  On the real axis this is
 * the same product of transfer matrices, taken point by point instead of coefficient by coefficient.
 */
static FNFT_INT nsev_contspec_segmented(fnftb_ctx *ctx, FNFT_UINT nb, FNFT_UINT D, FNFT_UINT Dseg_max,
                                        FNFT_COMPLEX const *q, FNFT_REAL const *T, FNFT_REAL eps_t, FNFT_UINT M,
                                        FNFT_REAL const *XI, FNFT_INT kappa, fnft_nsev_opts_t const *opts,
                                        fnftb_scatter_desc const *sd, FNFT_COMPLEX *out, int on_device,
                                        int32_t *status)
{
    const FNFT_UINT nseg = (D + Dseg_max - 1) / Dseg_max;
    const FNFT_UINT base = D / nseg, extra = D % nseg;
    fnft_nsev_opts_t o = *opts;
    o.contspec_type = fnft_nsev_cstype_AB;
    FNFT_COMPLEX *cur = (FNFT_COMPLEX *)fnftb_seg_buffer(ctx, nb, M, 2);
    if (cur == NULL)
        return E_DEVICE;
    FNFT_UINT n0 = 0;
    for (FNFT_UINT s = 0; s < nseg; s++) {
        const FNFT_UINT len = base + (s < extra ? 1 : 0);
        const FNFT_REAL Ts[2] = {T[0] + n0 * eps_t, T[0] + (n0 + len - 1) * eps_t};
        if (fnftb_set_signals_strided(ctx, nb, len, q + n0, D, on_device) != 0)
            return E_DEVICE;
        if (fnftb_fscatter(ctx, sd) != 0)
            return E_DEVICE;
        FNFT_INT ret_code = nsev_contspec_chunk_eps(ctx, len, Ts, eps_t, M, XI, &o, cur, 1, NULL, 1);
        if (ret_code != FNFT_SUCCESS)
            return E_SUBROUTINE(ret_code);
        if (fnftb_seg_compose(ctx, nb, M, (int)kappa, s == 0) != 0)
            return E_DEVICE;
        n0 += len;
    }
    if (fnftb_seg_finish(ctx, nb, M, (int)opts->contspec_type, out, contspec_len(opts->contspec_type, M), on_device,
                         status) != 0)
        return E_DEVICE;
    return FNFT_SUCCESS;
}

/*
 * Continuous spectrum of the staged signals for the slow discretizations BO / CF4_2
 * (nsev_compute_contspec with deg == 0, src/fnft_nsev.c:794-814, epilogue :836-876).
 */
static FNFT_INT nsev_slow_contspec_chunk(fnftb_ctx *ctx, FNFT_UINT D_given, FNFT_REAL const *T, FNFT_UINT M,
                                         FNFT_REAL const *XI, FNFT_INT kappa, int upsampling,
                                         fnft_nsev_opts_t const *opts, FNFT_COMPLEX *out, int on_device,
                                         int32_t *status)
{
    const fnft_nse_discretization_t disc = opts->discretization;
    const FNFT_REAL eps_t = (T[1] - T[0]) / (D_given - 1);
    fnftb_contspec_desc cd;
    memset(&cd, 0, sizeof(cd));
    cd.mode = FNFTB_MODE_NSEV;
    cd.cstype = (int)opts->contspec_type;
    cd.M = M;
    cd.xi0 = XI[0];
    cd.eps_xi = (XI[1] - XI[0]) / (M - 1);
    FNFT_INT ret_code;
    ret_code = fnftb__nse_phase_factor_rho(eps_t, T[1], &cd.ph_rho, disc);
    if (ret_code != FNFT_SUCCESS)
        return E_SUBROUTINE(ret_code);
    ret_code = fnftb__nse_phase_factor_a(eps_t, D_given, T, &cd.ph_a, disc);
    if (ret_code != FNFT_SUCCESS)
        return E_SUBROUTINE(ret_code);
    ret_code = fnftb__nse_phase_factor_b(eps_t, D_given, T, &cd.ph_b, disc);
    if (ret_code != FNFT_SUCCESS)
        return E_SUBROUTINE(ret_code);
    if (fnftb_slow_contspec(ctx, &cd, upsampling, kappa, eps_t, out, contspec_len(opts->contspec_type, M),
                            on_device, status) != 0)
        return E_DEVICE;
    return FNFT_SUCCESS;
}

/*
 * One pass of fnft_nsev over a batch: preprocessing + fnft_nsev_base (src/fnft_nsev.c:266-309,
 * 458-565).  Dsub_req != 0 asks for the pass on the subsampled signals
 * (nse_discretization_preprocess_signal with *Dsub_ptr = Dsub_req; second pass of the
 * Richardson extrapolation, :374-391); the time window and step size follow (:382-384).
 */
/* FNFT_B200_NSEV_TIMING=1: host wall time of the phases of one chunk of fnft_nsev_batch (each closed by a device sync) */
#include <stdio.h>
#include <time.h>
static int nsev_timing_on(void)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("FNFT_B200_NSEV_TIMING");
        v = (e && e[0]) ? atoi(e) : 0;
    }
    return v;
}
static void nsev_tick(fnftb_ctx *ctx, const char *label)
{
    static __thread struct timespec last;
    if (!nsev_timing_on())
        return;
    fnftb_ctx_sync(ctx);
    struct timespec now;
    clock_gettime(CLOCK_MONOTONIC, &now);
    if (label != NULL)
        fprintf(stderr, "[nsev timing] %-36s %8.3f ms\n", label,
                (now.tv_sec - last.tv_sec) * 1e3 + (now.tv_nsec - last.tv_nsec) * 1e-6);
    last = now;
}

static FNFT_INT nsev_pass(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                          FNFT_REAL const *const T_full, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                          FNFT_REAL const *const XI, FNFT_UINT *const K, const FNFT_UINT Kmax,
                          FNFT_COMPLEX *const bound_states,
                          FNFT_COMPLEX *const normconsts_or_residues, const FNFT_INT kappa,
                          fnft_nsev_opts_t const *opts, FNFT_INT *const ret_codes,
                          const FNFT_UINT Dsub_req, FNFT_REAL *const eps_t_pass)
{
    FNFT_INT ret_code = FNFT_SUCCESS;
    int32_t *status = NULL;

    /* argument checks of src/fnft_nsev.c:162-180 */
    if (B == 0)
        return E_INVALID_ARGUMENT(B);
    if (D < 2)
        return E_INVALID_ARGUMENT(D);
    if (q == NULL)
        return E_INVALID_ARGUMENT(q);
    if (T_full == NULL || T_full[0] >= T_full[1])
        return E_INVALID_ARGUMENT(T);
    if (contspec != NULL) {
        if (XI == NULL || XI[0] >= XI[1])
            return E_INVALID_ARGUMENT(XI);
    }
    if (abs(kappa) != 1)
        return E_INVALID_ARGUMENT(kappa);
    if (bound_states != NULL) {
        if (K == NULL)
            return E_INVALID_ARGUMENT(K_ptr);
    }
    if (opts == NULL)
        opts = &nsev_defaults;
    if (contspec != NULL && M > 0 && contspec_len(opts->contspec_type, M) == 0)
        return E_INVALID_ARGUMENT(opts->contspec_type);

    const fnft_nse_discretization_t disc = opts->discretization;
    fnft__akns_discretization_t akns;
    ret_code = fnftb__nse_to_akns(disc, &akns);
    if (ret_code != FNFT_SUCCESS)
        return E_INVALID_ARGUMENT(opts->discretization);
    const FNFT_UINT deg0 = fnftb__akns_degree(akns);
    const FNFT_UINT upsampling = fnftb__akns_upsampling(akns);
    /* "slow" discretizations (no polynomial transfer matrix): BO and CF4_2 run on the GPU, the
     * continuous spectrum as one product of step matrices per spectral point (slow_scatter.cuh) */
    const int slow = (deg0 == 0);
    /* weight selector of the commutator-free schemes with more than two exponentials (bo_l_at) */
    /* ... and of the exponential schemes on (q, q', q''): 4 = ES4, 5 = TES4 (es_step, bound_kernels.cuh) */
    const int cf_wsel = (akns == fnft__akns_discretization_CF4_3)   ? 1
                        : (akns == fnft__akns_discretization_CF5_3) ? 2
                        : (akns == fnft__akns_discretization_CF6_4) ? 3
                        : (akns == fnft__akns_discretization_ES4)   ? 4
                        : (akns == fnft__akns_discretization_TES4)  ? 5
                                                                    : 0;
    const int es = (cf_wsel >= 4);
    if (slow && akns != fnft__akns_discretization_BO && akns != fnft__akns_discretization_CF4_2 && cf_wsel == 0)
        return E_INVALID_ARGUMENT(opts->discretization);
    if (!slow && !fnftb__akns_on_gpu(akns))
        return E_NOT_YET_IMPLEMENTED(opts->discretization,
                                     This splitting scheme has no GPU leaf kernel yet.);
    /* src/fnft_nsev.c:209-219: the slow discretizations only support Newton localization */
    if (slow && kappa == +1 && opts->bound_state_localization != fnft_nsev_bsloc_NEWTON)
        return E_INVALID_ARGUMENT(opts->bound_state_localization);
    if (upsampling > 4 || (upsampling > 2 && cf_wsel == 0))
        return E_NOT_YET_IMPLEMENTED(opts->discretization, Unsupported upsampling factor.);
    const int want_contspec = (contspec != NULL && M > 0);
    const int want_discspec = (kappa == +1 && bound_states != NULL);
    const fnft_nsev_bsloc_t bsloc = opts->bound_state_localization;
    if (want_discspec && bsloc != fnft_nsev_bsloc_NEWTON && bsloc != fnft_nsev_bsloc_FAST_EIGENVALUE &&
        bsloc != fnft_nsev_bsloc_SUBSAMPLE_AND_REFINE)
        return E_INVALID_ARGUMENT(opts->bound_state_localization);
    if (want_discspec && Kmax == 0)
        return E_INVALID_ARGUMENT(Kmax);

    fnftb_ctx *ctx = fnftb__ctx();
    if (ctx == NULL)
        return E_OTHER("No usable CUDA device: the fnft_b200 hot path has no CPU fallback.");
    const int devptr = fnftb__device_pointers();

    const FNFT_REAL eps_t_full = (T_full[1] - T_full[0]) / (D - 1); /* src/fnft_nsev.c:237 */
    /* subsampled pass: src/private/fnft__nse_discretization.c:424-431,623-625 */
    FNFT_UINT D_given = D, nskip_pass = 1;
    if (Dsub_req != 0) {
        FNFT_UINT Dsub = Dsub_req < 2 ? 2 : (Dsub_req > D ? D : Dsub_req);
        nskip_pass = (FNFT_UINT)round((FNFT_REAL)D / Dsub);
        D_given = (FNFT_UINT)round((FNFT_REAL)D / nskip_pass);
    }
    const FNFT_REAL T[2] = {T_full[0], Dsub_req != 0 ? T_full[0] + ((D_given - 1) * nskip_pass) * eps_t_full : T_full[1]};
    const FNFT_REAL eps_t = (Dsub_req != 0) ? (T[1] - T[0]) / (D_given - 1) : eps_t_full;
    if (eps_t_pass != NULL)
        *eps_t_pass = eps_t;
    const FNFT_UINT D_eff = D_given * upsampling;
    const FNFT_UINT cs_len = want_contspec ? contspec_len(opts->contspec_type, M) : 0;

    /* signals whose transfer matrix is longer than one product tree: continuous spectrum by segments */
    const FNFT_UINT Dtree_max = slow ? 0 : (FNFT_UINT)fnftb_tree_max_samples((int)akns, (int)deg0);
    const int segmented = (!slow && want_contspec && D_eff > Dtree_max);
    if (!slow && D_eff > Dtree_max) {
        if (want_contspec && (upsampling != 1 || Dsub_req != 0))
            return E_NOT_YET_IMPLEMENTED(D, Signals this long are only supported without resampling and Richardson extrapolation.);
        /* Newton and the default SUBSAMPLE_AND_REFINE (fast eigenvalues of the SUBSAMPLED signal, then Newton)
         * never need the long polynomial; FAST_EIGENVALUE on the full signal does */
        if (want_discspec && bsloc == fnft_nsev_bsloc_FAST_EIGENVALUE)
            return E_NOT_YET_IMPLEMENTED(D, Signals this long need bound_state_localization NEWTON or SUBSAMPLE_AND_REFINE.);
    }
    const FNFT_UINT D_ws = segmented ? Dtree_max : D_eff; /* what the tree workspace has to hold at a time */

    /* the bound-state kernels keep eigenvalue, flag, a, a' and b per entry of the [nb][Kmax] arrays */
    size_t chunk = fnftb_max_chunk_ex(ctx, D_ws, slow ? 1 : (int)deg0, (want_contspec && !slow) ? M : 0, 2,
                                      (want_discspec ? Kmax * (4 * sizeof(FNFT_COMPLEX) + sizeof(int32_t)) : 0) +
                                          (segmented ? (4 * M + (want_discspec ? D : 0)) * sizeof(FNFT_COMPLEX) : 0),
                                      fnftb__workspace_limit());
    if (chunk > B)
        chunk = B;
    /* Continuous spectrum only, host buffers: overlap the host<->device copies of
     * neighbouring chunks with the kernels (fnftb_pipeline_*); at least 8 chunks for
     * large batches so that only a small first copy-in / last copy-out stay exposed. */
    const int piped = (!devptr && want_contspec && !want_discspec && B >= 16 && fnftb__pipe_chunks() > 0 &&
                       Dsub_req == 0 && !slow && !segmented);
    if (piped) {
        const size_t nch = (size_t)fnftb__pipe_chunks();
        size_t c8 = (B + nch - 1) / nch;
        if (c8 < 64)
            c8 = 64;
        if (chunk > c8)
            chunk = c8;
    }
    if (!devptr) {
        status = malloc(2 * chunk * sizeof(int32_t));
        if (status == NULL)
            return E_NOMEM;
    }
    if (ret_codes != NULL)
        for (FNFT_UINT b = 0; b < B; b++)
            ret_codes[b] = FNFT_SUCCESS;

    fnftb_scatter_desc sd;
    memset(&sd, 0, sizeof(sd));
    sd.rmode = FNFTB_RMODE_NSE;
    sd.kappa = kappa;
    sd.scheme = (int)akns;
    sd.deg0 = (int)deg0;
    sd.normalize = opts->normalization_flag ? 1 : 0;
    sd.eps_t = eps_t;
    /* fnftb_contspec may read (a, b) straight from the level buffer and skip the final tree kernel.
     * FAST_EIGENVALUE reuses the transfer matrix of the chunk for the root finder, so it must exist. */
    sd.defer_final = !(want_discspec && opts->bound_state_localization == fnft_nsev_bsloc_FAST_EIGENVALUE);

    if (piped && fnftb_pipeline_begin(ctx, B) != 0) {
        ret_code = E_DEVICE;
        goto leave_fun;
    }
    FNFT_UINT step = chunk;
    for (FNFT_UINT b0 = 0; b0 < B; b0 += step) {
        step = piped ? fnftb__pipe_step(b0, B, chunk) : chunk; /* tapered at both ends */
        const FNFT_UINT nb = (B - b0 < step) ? (B - b0) : step;
        int32_t *st_cur = status;
        int staged_above = 0;
        nsev_tick(ctx, NULL);
        if (nb > 0 && want_discspec && bsloc == fnft_nsev_bsloc_SUBSAMPLE_AND_REFINE && Dsub_req == 0) {
            staged_above = 1;
            /* First step of the mixed method (src/fnft_nsev.c:276-296): initial guesses from
             * the fast eigenvalue method on a subsampled signal. */
            FNFT_UINT Dsub = opts->Dsub;
            if (Dsub == 0) /* the user wants us to determine Dsub */
                Dsub = (FNFT_UINT)sqrt(D * log2(D) * log2(D));
            if (Dsub < 2) /* src/private/fnft__nse_discretization.c:424-431 */
                Dsub = 2;
            if (Dsub > D)
                Dsub = D;
            const FNFT_UINT nskip = (FNFT_UINT)round((FNFT_REAL)D / Dsub);
            Dsub = (FNFT_UINT)round((FNFT_REAL)D / nskip);
            if (fnftb_set_signals(ctx, nb, D, q + b0 * D, NULL, devptr) != 0) {
                ret_code = E_DEVICE;
                goto leave_fun;
            }
            const int rc_sub = (upsampling == 2)   ? fnftb_resample_4split4_sub(ctx, eps_t_full, nskip, Dsub, NULL)
                               : es                ? fnftb_preprocess_es4(ctx, cf_wsel, eps_t_full, nskip, Dsub)
                               : (upsampling >= 3) ? fnftb_resample_cf_sub(ctx, cf_wsel, (int)kappa, eps_t_full, nskip, Dsub, NULL)
                                                   : fnftb_subsample(ctx, nskip, Dsub);
            if (rc_sub != 0) {
                ret_code = E_DEVICE;
                goto leave_fun;
            }
            nsev_tick(ctx, "copy in + sub-sampling");
            const FNFT_REAL Tsub[2] = {T[0], T[0] + ((Dsub - 1) * nskip) * eps_t_full}; /* :290-291 */
            ret_code = fnftb__nsev_fasteig_chunk(ctx, nb, Dsub * upsampling, Dsub, Tsub, kappa, 0, K + b0,
                                                 Kmax, bound_states + b0 * Kmax, opts);
            CHECK_RETCODE(ret_code, leave_fun);
            nsev_tick(ctx, "sub-sampled fast eigenvalue pass");
        }
        if (nb > 0 && segmented) {
            ret_code = nsev_contspec_segmented(ctx, nb, D, Dtree_max, q + b0 * D, T, eps_t, M, XI, kappa, opts, &sd,
                                               contspec + b0 * cs_len, devptr, st_cur);
            CHECK_RETCODE(ret_code, leave_fun);
        }
        if (nb > 0 && (!segmented || want_discspec)) {
            /* preprocessing (src/fnft_nsev.c:272): plain copy for upsampling 1; the
             * 4SPLIT4 schemes resample on the device.  The chunk is still on the device when the
             * sub-sampled pass above has just uploaded it. */
            if (!(staged_above && !devptr && !segmented && fnftb_signals_staged(ctx, nb, D, q + b0 * D)) &&
                fnftb_set_signals(ctx, nb, D, q + b0 * D, NULL, devptr) != 0) {
                ret_code = E_DEVICE;
                goto leave_fun;
            }
            if (upsampling == 1 && Dsub_req != 0) {
                if (fnftb_subsample(ctx, nskip_pass, D_given) != 0) {
                    ret_code = E_DEVICE;
                    goto leave_fun;
                }
            }
            if (es) { /* finite differences on the (sub-sampled) grid, src/private/fnft__nse_discretization.c:609-631 */
                if (fnftb_preprocess_es4(ctx, cf_wsel, eps_t_full, nskip_pass, D_given) != 0) {
                    ret_code = E_DEVICE;
                    goto leave_fun;
                }
            } else if (upsampling >= 2) {
                int32_t *warn = devptr ? NULL : malloc(nb * sizeof(int32_t));
                if ((upsampling == 2 ? fnftb_resample_4split4_sub(ctx, eps_t_full, nskip_pass, D_given, warn)
                                     : fnftb_resample_cf_sub(ctx, cf_wsel, (int)kappa, eps_t_full, nskip_pass, D_given,
                                                             warn)) != 0) {
                    free(warn);
                    ret_code = E_DEVICE;
                    goto leave_fun;
                }
                if (warn != NULL) {
                    for (FNFT_UINT b = 0; b < nb; b++) {
                        if (warn[b]) { /* src/private/fnft__misc.c:379-380 */
                            WARN("Signal does not appear to be bandlimited. Interpolation step may be inaccurate. Try to reduce the step size, or switch to a discretization that does not require interpolation");
                            break;
                        }
                    }
                    free(warn);
                }
            }

            /* transfer matrix: nse_fscatter (src/fnft_nsev.c:527).  The Newton path never
             * reads it, so it is only built when a continuous spectrum is wanted. */
            if (segmented) {
                /* done above */
            } else if (want_contspec && slow) {
                ret_code = nsev_slow_contspec_chunk(ctx, D_given, T, M, XI, kappa, (int)upsampling, opts,
                                                    contspec + b0 * cs_len, devptr, st_cur);
                CHECK_RETCODE(ret_code, leave_fun);
            } else if (want_contspec) {
                if (fnftb_fscatter(ctx, &sd) != 0) {
                    ret_code = E_DEVICE;
                    goto leave_fun;
                }
                ret_code = nsev_contspec_chunk(ctx, D_given, T, M, XI, opts, contspec + b0 * cs_len, devptr,
                                               st_cur);
                CHECK_RETCODE(ret_code, leave_fun);
            }
        }
        if (!piped)
            nsev_tick(ctx, "staging + continuous spectrum");
        /* status of this chunk (pipelined mode: all chunks at once after the loop) */
        if (want_contspec && !devptr && !piped) {
            const FNFT_UINT cb0 = b0, cnb = nb;
            const int32_t *st_chk = st_cur;
            for (FNFT_UINT b = 0; b < cnb; b++) {
                if (st_chk[b] == FNFT_EC_DIV_BY_ZERO) {
                    const FNFT_INT ec = E_DIV_BY_ZERO; /* src/fnft_nsev.c:850-852 */
                    if (ret_codes != NULL)
                        ret_codes[cb0 + b] = ec;
                    if (ret_code == FNFT_SUCCESS)
                        ret_code = ec;
                }
            }
            if (ret_code != FNFT_SUCCESS && ret_codes == NULL)
                goto leave_fun;
        }

        if (want_discspec) {
            const int fast = (bsloc == fnft_nsev_bsloc_FAST_EIGENVALUE);
            if (fast) { /* src/fnft_nsev.c:687-711; the transfer matrix of this chunk is reused */
                ret_code = fnftb__nsev_fasteig_chunk(ctx, nb, D_eff, D_given, T, kappa, want_contspec, K + b0,
                                                     Kmax, bound_states + b0 * Kmax, opts);
                CHECK_RETCODE(ret_code, leave_fun);
            }
            FNFT_INT rc2 = fnftb__nsev_discrete_chunk(ctx, nb, D_eff, D_given, T, eps_t, K + b0, Kmax,
                                                      bound_states + b0 * Kmax,
                                                      normconsts_or_residues == NULL
                                                          ? NULL
                                                          : normconsts_or_residues +
                                                                b0 * Kmax *
                                                                    (opts->discspec_type == fnft_nsev_dstype_BOTH ? 2 : 1),
                                                      opts, ret_codes ? ret_codes + b0 : NULL, fast);
            if (rc2 != FNFT_SUCCESS && ret_code == FNFT_SUCCESS)
                ret_code = rc2;
            if (rc2 != FNFT_SUCCESS && ret_codes == NULL)
                goto leave_fun;
            nsev_tick(ctx, "Newton + norming constants");
        } else if (K != NULL && !devptr) {
            for (FNFT_UINT b = 0; b < nb; b++)
                K[b0 + b] = 0; /* src/fnft_nsev.c:558-560 */
        }
    }

    if (piped) {
        /* pipelined mode: the whole batch has been enqueued without a single host-side wait; now
         * wait for the streams and look at the per-signal status of all chunks (src/fnft_nsev.c:850-852) */
        const int32_t *st_all = NULL;
        if (fnftb_pipeline_end(ctx) != 0 || fnftb_pipeline_status(ctx, &st_all) != 0 || st_all == NULL) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        for (FNFT_UINT b = 0; b < B; b++) {
            if (st_all[b] == FNFT_EC_DIV_BY_ZERO) {
                const FNFT_INT ec = E_DIV_BY_ZERO;
                if (ret_codes != NULL)
                    ret_codes[b] = ec;
                if (ret_code == FNFT_SUCCESS)
                    ret_code = ec;
            }
        }
    }

leave_fun:
    if (piped)
        (void)fnftb_pipeline_end(ctx);
    free(status);
    return ret_code;
}

/* one shard of a batch that fnft_b200_set_devices spreads over several GPUs (fnft_runtime.c) */
typedef struct {
    FNFT_UINT B, D, M, Kmax;
    FNFT_COMPLEX const *q;
    FNFT_REAL const *T, *XI;
    FNFT_COMPLEX *contspec, *bound_states, *normconsts_or_residues;
    FNFT_UINT *K;
    FNFT_INT kappa;
    fnft_nsev_opts_t const *opts;
    FNFT_INT *ret_codes;
    FNFT_INT rc[16];
} nsev_job;

static void nsev_shard(void *arg, int shard, int nshards)
{
    nsev_job *j = (nsev_job *)arg;
    FNFT_UINT b0, b1;
    fnftb__shard_range(j->B, shard, nshards, &b0, &b1);
    const FNFT_UINT cs_len = (j->contspec != NULL) ? contspec_len(j->opts->contspec_type, j->M) : 0;
    const FNFT_UINT nlen = (j->opts->discspec_type == fnft_nsev_dstype_BOTH) ? 2 * j->Kmax : j->Kmax;
    j->rc[shard] = fnft_nsev_batch(b1 - b0, j->D, j->q + b0 * j->D, j->T, j->M,
                                   j->contspec ? j->contspec + b0 * cs_len : NULL, j->XI,
                                   j->K ? j->K + b0 : NULL, j->Kmax,
                                   j->bound_states ? j->bound_states + b0 * j->Kmax : NULL,
                                   j->normconsts_or_residues ? j->normconsts_or_residues + b0 * nlen : NULL,
                                   j->kappa, j->opts, j->ret_codes ? j->ret_codes + b0 : NULL);
}

static FNFT_INT nsev_batch_impl(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                         FNFT_REAL const *const T, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                         FNFT_REAL const *const XI, FNFT_UINT *const K, const FNFT_UINT Kmax,
                         FNFT_COMPLEX *const bound_states,
                         FNFT_COMPLEX *const normconsts_or_residues, const FNFT_INT kappa,
                         fnft_nsev_opts_t const *opts, FNFT_INT *const ret_codes);

/* public entry: NVTX range around the call (FNFT_B200_NVTX=1, no-op otherwise) */
FNFT_INT fnft_nsev_batch(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                         FNFT_REAL const *const T, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                         FNFT_REAL const *const XI, FNFT_UINT *const K, const FNFT_UINT Kmax,
                         FNFT_COMPLEX *const bound_states,
                         FNFT_COMPLEX *const normconsts_or_residues, const FNFT_INT kappa,
                         fnft_nsev_opts_t const *opts, FNFT_INT *const ret_codes)
{
    fnftb_range_push("fnft_nsev_batch");
    const FNFT_INT rc = nsev_batch_impl(B, D, q, T, M, contspec, XI, K, Kmax, bound_states, normconsts_or_residues, kappa, opts, ret_codes);
    fnftb_range_pop();
    return rc;
}

static FNFT_INT nsev_batch_impl(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                         FNFT_REAL const *const T, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                         FNFT_REAL const *const XI, FNFT_UINT *const K, const FNFT_UINT Kmax,
                         FNFT_COMPLEX *const bound_states,
                         FNFT_COMPLEX *const normconsts_or_residues, const FNFT_INT kappa,
                         fnft_nsev_opts_t const *opts, FNFT_INT *const ret_codes)
{
    if (opts == NULL)
        opts = &nsev_defaults;
    const int nshards = (q != NULL) ? fnftb__fanout_shards(B) : 1;
    if (nshards > 1) { /* several GPUs: every shard is this same call on its own device */
        nsev_job j = {B, D, M, Kmax, q, T, XI, contspec, bound_states, normconsts_or_residues, K, kappa, opts,
                      ret_codes, {0}};
        if (fnftb__fanout_run(nshards, nsev_shard, &j) != 0)
            return E_OTHER("Could not start the per-device worker threads.");
        for (int i = 0; i < nshards; i++)
            if (j.rc[i] != FNFT_SUCCESS)
                return E_SUBROUTINE(j.rc[i]);
        return FNFT_SUCCESS;
    }
    if (opts->richardson_extrapolation_flag != 1)
        return nsev_pass(B, D, q, T, M, contspec, XI, K, Kmax, bound_states, normconsts_or_residues, kappa,
                         opts, ret_codes, 0, NULL);

    /*
     * Richardson extrapolation (src/fnft_nsev.c:250-263, 316-442): a second approximation from
     * every other sample, bound states refined by Newton from the first pass's values, then
     * (scl*first - second)/(scl - 1) with scl = (eps_t_sub/eps_t)^order.  The two passes run on
     * the GPU; the O(M + K^2) combination is done here like in the reference.
     */
    FNFT_INT ret_code = FNFT_SUCCESS;
    FNFT_COMPLEX *cs_sub = NULL, *bs_sub = NULL, *nc_sub = NULL, *nc_res = NULL;
    FNFT_UINT *K_sub = NULL;
    if (fnftb__device_pointers())
        return E_NOT_YET_IMPLEMENTED(richardson_extrapolation_flag, Richardson extrapolation needs host buffers.);
    const int want_cs = (contspec != NULL && M > 0);
    const int want_ds = (kappa == +1 && bound_states != NULL);
    const FNFT_UINT cs_len = want_cs ? contspec_len(opts->contspec_type, M) : 0;
    if (want_cs && cs_len == 0)
        return E_INVALID_ARGUMENT(opts->contspec_type);
    fnft_nsev_opts_t o = *opts;
    const fnft_nsev_dstype_t ds_type_opt = opts->discspec_type;
    FNFT_COMPLEX *nc_main = normconsts_or_residues;
    if (want_ds && normconsts_or_residues != NULL && ds_type_opt == fnft_nsev_dstype_RESIDUES) {
        /* a' is extrapolated, so norming constants and residues are both needed (:252-262) */
        o.discspec_type = fnft_nsev_dstype_BOTH;
        nc_res = malloc(B * 2 * Kmax * sizeof(FNFT_COMPLEX));
        if (nc_res == NULL)
            return E_NOMEM;
        nc_main = nc_res;
    }
    const FNFT_UINT nc_stride = (o.discspec_type == fnft_nsev_dstype_BOTH) ? 2 * Kmax : Kmax;
    FNFT_REAL eps_t = 0, eps_t_sub = 0;
    ret_code = nsev_pass(B, D, q, T, M, contspec, XI, K, Kmax, bound_states, nc_main, kappa, &o, ret_codes, 0,
                         &eps_t);
    CHECK_RETCODE(ret_code, leave_fun);

    const FNFT_UINT method_order = fnftb__nse_method_order(opts->discretization);
    if (method_order == 0) {
        ret_code = E_INVALID_ARGUMENT(discretization);
        goto leave_fun;
    }
    if (want_cs) {
        cs_sub = malloc(B * cs_len * sizeof(FNFT_COMPLEX));
        if (cs_sub == NULL) {
            ret_code = E_NOMEM;
            goto leave_fun;
        }
    }
    int any_bs = 0;
    if (want_ds) {
        K_sub = malloc(B * sizeof(FNFT_UINT));
        bs_sub = malloc(B * Kmax * sizeof(FNFT_COMPLEX));
        nc_sub = (normconsts_or_residues != NULL) ? malloc(B * nc_stride * sizeof(FNFT_COMPLEX)) : NULL;
        if (K_sub == NULL || bs_sub == NULL || (normconsts_or_residues != NULL && nc_sub == NULL)) {
            ret_code = E_NOMEM;
            goto leave_fun;
        }
        for (FNFT_UINT b = 0; b < B; b++) {
            K_sub[b] = K[b];
            any_bs |= (K[b] != 0);
            memcpy(bs_sub + b * Kmax, bound_states + b * Kmax, K[b] * sizeof(FNFT_COMPLEX));
        }
    }
    o.bound_state_localization = fnft_nsev_bsloc_NEWTON; /* :387-388 */
    ret_code = nsev_pass(B, D, q, T, M, cs_sub, XI, (want_ds && any_bs) ? K_sub : NULL, Kmax,
                         (want_ds && any_bs) ? bs_sub : NULL, (want_ds && any_bs) ? nc_sub : NULL, kappa, &o,
                         NULL, D / 2 /* Dsub = CEIL(D/2) with the integer quotient, :374 */, &eps_t_sub);
    CHECK_RETCODE(ret_code, leave_fun);

    const FNFT_REAL scl_num = pow(eps_t_sub / eps_t, (FNFT_REAL)method_order);
    const FNFT_REAL scl_den = scl_num - 1.0;
    if (want_cs) { /* :398-407 */
        const FNFT_REAL dxi = (XI[1] - XI[0]) / (M - 1);
        for (FNFT_UINT b = 0; b < B; b++) {
            FNFT_COMPLEX *c1 = contspec + b * cs_len;
            FNFT_COMPLEX const *c2 = cs_sub + b * cs_len;
            for (FNFT_UINT i = 0; i < M; i++)
                if (fabs(XI[0] + dxi * i) < 0.9 * FNFT_PI / (2.0 * eps_t_sub))
                    for (FNFT_UINT j = 0; j < cs_len; j += M)
                        c1[i + j] = (scl_num * c1[i + j] - c2[i + j]) / scl_den;
        }
    }
    if (want_ds && any_bs) { /* :409-441 */
        for (FNFT_UINT b = 0; b < B; b++) {
            const FNFT_UINT Kb = K[b], Ks = K_sub[b];
            FNFT_COMPLEX *bs1 = bound_states + b * Kmax;
            FNFT_COMPLEX const *bs2 = bs_sub + b * Kmax;
            FNFT_COMPLEX *n1 = nc_main ? nc_main + b * nc_stride : NULL;
            FNFT_COMPLEX *n2 = nc_sub ? nc_sub + b * nc_stride : NULL;
            for (FNFT_UINT i = 0; i < Kb && Ks != 0; i++) {
                FNFT_UINT loc = Ks;
                FNFT_REAL thres = eps_t;
                for (FNFT_UINT j = 0; j < Ks; j++) {
                    const FNFT_REAL err = cabs(bs1[i] - bs2[j]) / cabs(bs1[i]);
                    if (err < thres) {
                        thres = err;
                        loc = j;
                    }
                }
                if (loc < Ks) {
                    bs1[i] = (scl_num * bs1[i] - bs2[loc]) / scl_den;
                    if (n1 != NULL && (ds_type_opt == fnft_nsev_dstype_RESIDUES ||
                                       ds_type_opt == fnft_nsev_dstype_BOTH)) {
                        /* a' = b / residue on both grids, Richardson step on a', new residue */
                        n1[Kb + i] = n1[i] / n1[Kb + i];
                        n2[Ks + loc] = n2[loc] / n2[Ks + loc];
                        n1[Kb + i] = (scl_num * n1[Kb + i] - n2[loc + Ks]) / scl_den;
                        n1[Kb + i] = n1[i] / n1[Kb + i];
                    }
                }
            }
            if (n1 != NULL && ds_type_opt == fnft_nsev_dstype_RESIDUES)
                memcpy(normconsts_or_residues + b * Kmax, n1 + Kb, Kb * sizeof(FNFT_COMPLEX));
        }
    }

leave_fun:
    free(cs_sub);
    free(bs_sub);
    free(nc_sub);
    free(nc_res);
    free(K_sub);
    return ret_code;
}

/* include/fnft_nsev.h:371-376, src/fnft_nsev.c:133-453 */
FNFT_INT fnft_nsev(const FNFT_UINT D, FNFT_COMPLEX *const q, FNFT_REAL const *const T,
                   const FNFT_UINT M, FNFT_COMPLEX *const contspec, FNFT_REAL const *const XI,
                   FNFT_UINT *const K_ptr, FNFT_COMPLEX *const bound_states,
                   FNFT_COMPLEX *const normconsts_or_residues, const FNFT_INT kappa,
                   fnft_nsev_opts_t *opts)
{
    /* K_ptr doubles as "size of the user arrays" (in) and "number found" (out) */
    if (bound_states != NULL && K_ptr == NULL)
        return E_INVALID_ARGUMENT(K_ptr);
    const FNFT_UINT Kmax = (bound_states != NULL && K_ptr != NULL) ? *K_ptr : 0;
    if (bound_states != NULL && kappa == +1 && Kmax == 0) {
        /* nothing to refine and nowhere to store: mirror K = 0 result */
        FNFT_INT rc = fnft_nsev_batch(1, D, q, T, M, contspec, XI, NULL, 0, NULL, NULL, kappa, opts,
                                      NULL);
        return rc;
    }
    const int devptr = fnftb__device_pointers();
    if (devptr)
        fnft_b200_set_device_pointers(0); /* the classic entry point takes host memory */
    const FNFT_INT rc = fnft_nsev_batch(1, D, q, T, M, contspec, XI, K_ptr, Kmax, bound_states,
                                        normconsts_or_residues, kappa, opts, NULL);
    if (devptr)
        fnft_b200_set_device_pointers(1);
    return rc;
}
