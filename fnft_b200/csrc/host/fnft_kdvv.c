/*
 * fnft_b200 host library -- fnft_kdvv and fnft_kdvv_batch.
 *
 * Host-side mirror of /root/reference/src/fnft_kdvv.c: argument checks (:74-92),
 * kdv_fscatter with r = -1 (:108, src/private/fnft__kdv_fscatter.c:74-77) and
 * tf2contspec_negxi (:126-209): chirp-z of H12 and H22 on the NEGATED xi grid with
 *   V = exp(-2i*eps_xi*eps_t/deg),  A = exp(+2i*XI[0]*eps_t/deg)          (:169-170)
 *   R = exp(2i*xi*(T1 + bc*eps_t)) * H12 / (2i*xi*H22 - H12)               (:198-203)
 * Like the reference, no preprocessing/upsampling step exists for KdV, so
 * kdv_discretization_4SPLIT4B evaluates the same leaf as 2SPLIT4B
 * (src/private/fnft__akns_fscatter.c:402-403).
 *
 * Difference by design: the reference never normalises here (W_ptr = NULL, :73,109)
 * and can overflow for long/strong signals; we always normalise -- the power of two
 * cancels in the ratio R.
 */
#include "fnft_internal.h"

static const fnft_kdvv_opts_t kdvv_defaults = {.discretization = fnft_kdv_discretization_2SPLIT8B};

fnft_kdvv_opts_t fnft_kdvv_default_opts(void) { return kdvv_defaults; }

/* one shard of a batch that fnft_b200_set_devices spreads over several GPUs (fnft_runtime.c) */
typedef struct {
    FNFT_UINT B, D, M;
    FNFT_COMPLEX const *u;
    FNFT_REAL const *T, *XI;
    FNFT_COMPLEX *contspec;
    fnft_kdvv_opts_t const *opts;
    FNFT_INT *ret_codes;
    FNFT_INT rc[16];
} kdvv_job;

static void kdvv_shard(void *arg, int shard, int nshards)
{
    kdvv_job *j = (kdvv_job *)arg;
    FNFT_UINT b0, b1;
    fnftb__shard_range(j->B, shard, nshards, &b0, &b1);
    j->rc[shard] = fnft_kdvv_batch(b1 - b0, j->D, j->u + b0 * j->D, j->T, j->M, j->contspec + b0 * j->M, j->XI,
                                   j->opts, j->ret_codes ? j->ret_codes + b0 : NULL);
}

static FNFT_INT kdvv_batch_impl(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const u,
                         FNFT_REAL const *const T, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                         FNFT_REAL const *const XI, fnft_kdvv_opts_t const *opts,
                         FNFT_INT *const ret_codes);

/* public entry: NVTX range around the call (FNFT_B200_NVTX=1, no-op otherwise) */
FNFT_INT fnft_kdvv_batch(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const u,
                         FNFT_REAL const *const T, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                         FNFT_REAL const *const XI, fnft_kdvv_opts_t const *opts,
                         FNFT_INT *const ret_codes)
{
    fnftb_range_push("fnft_kdvv_batch");
    const FNFT_INT rc = kdvv_batch_impl(B, D, u, T, M, contspec, XI, opts, ret_codes);
    fnftb_range_pop();
    return rc;
}

static FNFT_INT kdvv_batch_impl(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const u,
                         FNFT_REAL const *const T, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                         FNFT_REAL const *const XI, fnft_kdvv_opts_t const *opts,
                         FNFT_INT *const ret_codes)
{
    FNFT_INT ret_code = FNFT_SUCCESS;
    const int nshards = (u != NULL && contspec != NULL) ? fnftb__fanout_shards(B) : 1;
    if (nshards > 1) { /* several GPUs: every shard is this same call on its own device */
        kdvv_job j = {B, D, M, u, T, XI, contspec, opts, ret_codes, {0}};
        if (fnftb__fanout_run(nshards, kdvv_shard, &j) != 0)
            return E_OTHER("Could not start the per-device worker threads.");
        for (int i = 0; i < nshards; i++)
            if (j.rc[i] != FNFT_SUCCESS)
                return E_SUBROUTINE(j.rc[i]);
        return FNFT_SUCCESS;
    }
    if (B == 0)
        return E_INVALID_ARGUMENT(B);
    if (D < 2)
        return E_INVALID_ARGUMENT(D);
    if (u == NULL)
        return E_INVALID_ARGUMENT(u);
    if (T == NULL || T[0] >= T[1])
        return E_INVALID_ARGUMENT(T);
    if (contspec == NULL)
        return E_INVALID_ARGUMENT(contspec);
    if (XI == NULL || XI[0] >= XI[1])
        return E_INVALID_ARGUMENT(XI);
    if (M == 0)
        return E_INVALID_ARGUMENT(M);
    if (opts == NULL)
        opts = &kdvv_defaults;

    fnft__akns_discretization_t akns;
    ret_code = fnftb__kdv_to_akns(opts->discretization, &akns);
    if (ret_code != FNFT_SUCCESS)
        return E_INVALID_ARGUMENT(opts->discretization);
    const FNFT_UINT deg0 = fnftb__akns_degree(akns);
    if (deg0 == 0 || !fnftb__akns_on_gpu(akns))
        return E_NOT_YET_IMPLEMENTED(opts->discretization,
                                     This discretization has no GPU leaf kernel yet.);

    fnftb_ctx *ctx = fnftb__ctx();
    if (ctx == NULL)
        return E_OTHER("No usable CUDA device: the fnft_b200 hot path has no CPU fallback.");
    const int devptr = fnftb__device_pointers();

    const FNFT_REAL eps_t = (T[1] - T[0]) / (D - 1);
    const FNFT_REAL eps_xi = (XI[1] - XI[0]) / (M - 1);
    const FNFT_REAL degree1step = (FNFT_REAL)deg0;
    const FNFT_REAL boundary_coeff = fnftb__akns_boundary_coeff(akns);

    fnftb_scatter_desc sd;
    memset(&sd, 0, sizeof(sd));
    sd.rmode = FNFTB_RMODE_KDV;
    sd.kappa = 0;
    sd.scheme = (int)akns;
    sd.deg0 = (int)deg0;
    sd.normalize = 1;
    sd.eps_t = eps_t;

    fnftb_contspec_desc cd;
    memset(&cd, 0, sizeof(cd));
    cd.mode = FNFTB_MODE_KDVV;
    cd.npoly = 2;
    cd.ent[0] = 1; /* H12 */
    cd.ent[1] = 3; /* H22 */
    cd.M = M;
    const FNFT_COMPLEX V = cexp(-2.0 * I * eps_xi * eps_t / degree1step);
    const FNFT_COMPLEX A = cexp(2.0 * I * XI[0] * eps_t / degree1step);
    fnftb__logpolar(V, &cd.lwr, &cd.lwi);
    fnftb__logpolar(A, &cd.lar, &cd.lai);
    cd.xi0 = XI[0];
    cd.eps_xi = eps_xi;
    cd.kdv_ph = T[1] + boundary_coeff * eps_t;
    cd.kdv_sqrtz =
        (opts->discretization == fnft_kdv_discretization_2SPLIT2A) ? eps_t / degree1step : 0.0;

    /* signals whose transfer matrix is longer than one product tree can hold: by segments (below) */
    const FNFT_UINT Dtree_max = (FNFT_UINT)fnftb_tree_max_samples((int)akns, (int)deg0);
    const int segmented = (D > Dtree_max);
    size_t chunk = segmented ? fnftb_max_chunk_ex(ctx, Dtree_max, (int)deg0, M, 2, 6 * M * sizeof(FNFT_COMPLEX),
                                                  fnftb__workspace_limit())
                             : fnftb_max_chunk(ctx, D, (int)deg0, M, 2, fnftb__workspace_limit());
    if (chunk > B)
        chunk = B;
    if (ret_codes != NULL)
        for (FNFT_UINT b = 0; b < B; b++)
            ret_codes[b] = FNFT_SUCCESS;

    if (segmented) {
        /* The reference multiplies polynomials of any length (src/private/fnft__poly_fmult.c:404-445).  Here the
         * signal is cut into pieces the tree can multiply; the four entries of every piece's transfer matrix are
         * evaluated at the same points z_m (chirp-z, raw values: the power-of-two scale of a piece is common to
         * its entries and drops out of the ratio below), the second column of the product is chained piece by
         * piece on the device, and the epilogue of src/fnft_kdvv.c:186-203 follows. */
        const FNFT_UINT nseg = (D + Dtree_max - 1) / Dtree_max;
        const FNFT_UINT base = D / nseg, extra = D % nseg;
        fnftb_contspec_desc raw = cd;
        raw.mode = FNFTB_MODE_RAW;
        for (FNFT_UINT b0 = 0; b0 < B; b0 += chunk) {
            const FNFT_UINT nb = (B - b0 < chunk) ? (B - b0) : chunk;
            FNFT_COMPLEX *cur = (FNFT_COMPLEX *)fnftb_seg_buffer(ctx, nb, M, 4);
            if (cur == NULL)
                return E_DEVICE;
            FNFT_UINT n0 = 0;
            for (FNFT_UINT s = 0; s < nseg; s++) {
                const FNFT_UINT len = base + (s < extra ? 1 : 0);
                if (fnftb_set_signals_strided(ctx, nb, len, u + b0 * D + n0, D, devptr) != 0 ||
                    fnftb_fscatter(ctx, &sd) != 0)
                    return E_DEVICE;
                raw.ent[0] = 1; /* H12, H22 */
                raw.ent[1] = 3;
                if (fnftb_contspec(ctx, &raw, cur, 4 * M, 1, NULL) != 0)
                    return E_DEVICE;
                if (s > 0) { /* the first piece only contributes its second column */
                    raw.ent[0] = 0; /* H11, H21 */
                    raw.ent[1] = 2;
                    if (fnftb_contspec(ctx, &raw, cur + 2 * M, 4 * M, 1, NULL) != 0)
                        return E_DEVICE;
                }
                if (fnftb_seg_compose_general(ctx, nb, M, s == 0) != 0)
                    return E_DEVICE;
                n0 += len;
            }
            if (fnftb_seg_finish_kdv(ctx, nb, M, cd.xi0, cd.eps_xi, cd.kdv_ph, cd.kdv_sqrtz, contspec + b0 * M, M,
                                     devptr) != 0)
                return E_DEVICE;
        }
        return FNFT_SUCCESS;
    }

    /* host buffers: overlap the copies of neighbouring chunks with the kernels */
    const int piped = (!devptr && B >= 16 && fnftb__pipe_chunks() > 0);
    if (piped) {
        const size_t nch = (size_t)fnftb__pipe_chunks();
        size_t c8 = (B + nch - 1) / nch;
        if (c8 < 64)
            c8 = 64;
        if (chunk > c8)
            chunk = c8;
        if (fnftb_pipeline_begin(ctx, 0) != 0)
            return E_DEVICE;
    }
    FNFT_UINT step = chunk;
    for (FNFT_UINT b0 = 0; b0 < B; b0 += step) {
        step = piped ? fnftb__pipe_step(b0, B, chunk) : chunk; /* tapered at both ends */
        const FNFT_UINT nb = (B - b0 < step) ? (B - b0) : step;
        if (fnftb_set_signals(ctx, nb, D, u + b0 * D, NULL, devptr) != 0 ||
            fnftb_fscatter(ctx, &sd) != 0 ||
            fnftb_contspec(ctx, &cd, contspec + b0 * M, M, devptr, NULL) != 0) {
            ret_code = E_DEVICE;
            break;
        }
    }
    if (piped && fnftb_pipeline_end(ctx) != 0 && ret_code == FNFT_SUCCESS)
        ret_code = E_DEVICE;
    return ret_code;
}

/* include/fnft_kdvv.h:104-109, src/fnft_kdvv.c:59-122 */
FNFT_INT fnft_kdvv(const FNFT_UINT D, FNFT_COMPLEX *const u, FNFT_REAL const *const T,
                   const FNFT_UINT M, FNFT_COMPLEX *const contspec, FNFT_REAL const *const XI,
                   FNFT_UINT *const K_ptr, FNFT_COMPLEX *const bound_states,
                   FNFT_COMPLEX *const normconsts_or_residues, fnft_kdvv_opts_t *opts)
{
    if (D < 2)
        return E_INVALID_ARGUMENT(D);
    if (u == NULL)
        return E_INVALID_ARGUMENT(u);
    if (T == NULL || T[0] >= T[1])
        return E_INVALID_ARGUMENT(T);
    if (contspec == NULL)
        return E_INVALID_ARGUMENT(contspec);
    if (XI == NULL || XI[0] >= XI[1])
        return E_INVALID_ARGUMENT(XI);
    if (K_ptr != NULL)
        return E_NOT_YET_IMPLEMENTED(K_ptr, Please pass "NULL".);
    if (bound_states != NULL)
        return E_NOT_YET_IMPLEMENTED(bound_states, Please pass "NULL".);
    if (normconsts_or_residues != NULL)
        return E_NOT_YET_IMPLEMENTED(normconsts_or_residues, Please pass "NULL".);
    const int devptr = fnftb__device_pointers();
    if (devptr)
        fnft_b200_set_device_pointers(0);
    const FNFT_INT rc = fnft_kdvv_batch(1, D, u, T, M, contspec, XI, opts, NULL);
    if (devptr)
        fnft_b200_set_device_pointers(1);
    return rc;
}
