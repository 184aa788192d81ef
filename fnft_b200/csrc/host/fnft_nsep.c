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
 * Only fnft_nsep_loc_GRIDSEARCH runs here.  SUBSAMPLE_AND_REFINE and MIXED (the
 * reference default) need the eiscor root finder and the slow scattering with
 * derivatives and return FNFT_EC_NOT_YET_IMPLEMENTED.
 */
#include "fnft_internal.h"

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
    if (opts->localization != fnft_nsep_loc_GRIDSEARCH)
        return E_NOT_YET_IMPLEMENTED(opts->localization,
                                     The GPU build implements fnft_nsep_loc_GRIDSEARCH only.);

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
    if (!(PHI[0] < PHI[1]) || PHI[0] == -INFINITY || PHI[1] == INFINITY)
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
    for (FNFT_UINT b0 = 0; b0 < B; b0 += chunk) {
        const FNFT_UINT nb = (B - b0 < chunk) ? (B - b0) : chunk;
        if (fnftb_set_signals(ctx, nb, D, q + b0 * D, NULL, 0) != 0 ||
            fnftb_nsep_derotate(ctx, Lam_shift, T[0], eps_t) != 0) {
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
        if (fnftb_fscatter(ctx, &sd) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        memset(Kc, 0, nb * sizeof(uint64_t));
        memset(Mc, 0, nb * sizeof(uint64_t));
        if (fnftb_nsep_gridsearch(ctx, &nd, Kc, main_spec ? main_spec + b0 * Kmax : NULL, Mc,
                                  aux_spec ? aux_spec + b0 * Mmax : NULL, status) != 0) {
            ret_code = E_DEVICE;
            goto leave_fun;
        }
        for (FNFT_UINT b = 0; b < nb; b++) {
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
    return ret_code;
}

FNFT_INT fnft_nsep_batch(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                         FNFT_REAL const *const T, FNFT_REAL const phase_shift,
                         FNFT_UINT *const K, const FNFT_UINT Kmax, FNFT_COMPLEX *const main_spec,
                         FNFT_UINT *const Mcount, const FNFT_UINT Mmax,
                         FNFT_COMPLEX *const aux_spec, const FNFT_INT kappa,
                         fnft_nsep_opts_t const *opts, FNFT_INT *const ret_codes)
{
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
