/*
 * fnft_b200 -- thin C-ABI between the C host library (csrc/host) and the CUDA
 * kernels (csrc/cuda).  Plain pointers and sizes only; every function returns 0 on
 * success or a nonzero code whose text is available from fnftb_last_error().
 *
 * A context owns one device, one stream and grow-only device workspaces.  The host
 * code stages a chunk of signals, runs the fast scattering on it (leaves + product
 * tree, result kept on the device) and then asks for spectra / coefficients.
 */
#ifndef FNFTB_DEVICE_H
#define FNFTB_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fnftb_ctx fnftb_ctx;

/* values of rmode */
#define FNFTB_RMODE_NSE 0      /* r = -kappa*conj(q) */
#define FNFTB_RMODE_KDV 1      /* r = -1 */
#define FNFTB_RMODE_EXPLICIT 2 /* r given */

/* chirp-z epilogue modes */
#define FNFTB_MODE_RAW 0
#define FNFTB_MODE_NSEV 1
#define FNFTB_MODE_KDVV 2

typedef struct {
    int rmode;      /* FNFTB_RMODE_* */
    int kappa;      /* +1 / -1 (NSE only) */
    int scheme;     /* value of fnft__akns_discretization_t */
    int deg0;       /* polynomial degree of one step */
    int normalize;  /* 1: power-of-two rescaling with exponent W */
    double eps_t;   /* step size entering the leaves */
    int defer_final; /* 1: the caller only wants fnftb_contspec next; the [B][4][deg+1] transfer
                      * matrix is then built lazily (or never, on the chirp-z fast path) */
} fnftb_scatter_desc;

typedef struct {
    int mode;       /* FNFTB_MODE_* */
    int cstype;     /* NSEV: 0 rho, 1 a|b, 2 rho|a|b */
    int npoly;      /* 1 or 2 */
    int ent[2];     /* which entries (0=11,1=12,2=21,3=22) are evaluated */
    size_t M;       /* number of output points */
    double lwr, lwi;   /* ln|W|, arg W of the chirp step */
    double lar, lai;   /* ln|A|, arg A of the chirp start */
    double xi0, eps_xi;
    double ph_rho, ph_a, ph_b;
    double kdv_ph, kdv_sqrtz;
} fnftb_contspec_desc;

int fnftb_device_count(void);
const char *fnftb_last_error(void);
unsigned long long fnftb_launch_count(void);
/* per-launch CUDA-event timing: enable, run, then read "name count total_ms" lines */
void fnftb_profile_enable(int on);
const char *fnftb_profile_report(void);

int fnftb_ctx_create(fnftb_ctx **out, int device);
void fnftb_ctx_destroy(fnftb_ctx *ctx);
int fnftb_ctx_device(const fnftb_ctx *ctx);
int fnftb_ctx_sync(fnftb_ctx *ctx);
/* the context's stream as a cudaStream_t cast to void* (for event timing) */
void *fnftb_ctx_stream(fnftb_ctx *ctx);

/* Pipelined host transfers: between begin and end, fnftb_set_signals (host q) and
 * fnftb_contspec (host out) run their copies on separate streams and do not wait; the
 * caller alternates slots 0, 1, 0, ... with its chunks and calls fnftb_pipeline_wait(slot)
 * before reading the outputs of the chunk that used that slot; the per-signal status of
 * that chunk is returned in a pinned buffer owned by the context (the status_host argument
 * of fnftb_contspec is ignored in this mode: a pageable destination would serialise). */
int fnftb_pipeline_begin(fnftb_ctx *ctx, size_t total_signals);
int fnftb_pipeline_wait(fnftb_ctx *ctx, int slot, const int32_t **status);
int fnftb_pipeline_end(fnftb_ctx *ctx);
/* With total_signals > 0 the per-signal status of ALL chunks accumulates in one pinned array
 * (chunk after chunk, in call order), so the host does not have to wait for a chunk before it
 * issues the next ones: it enqueues the whole batch, calls fnftb_pipeline_end (which waits for the
 * three streams) and then reads the array through fnftb_pipeline_status. */
int fnftb_pipeline_status(fnftb_ctx *ctx, const int32_t **status);

/* Largest number of signals one fscatter+contspec pass may hold given the
 * workspace budget (bytes; 0 = default budget). */
size_t fnftb_max_chunk(const fnftb_ctx *ctx, size_t D, int deg0, size_t M, int npoly,
                       size_t budget_bytes);
/* same with additional workspace per signal that the caller knows about (bound-state arrays of Kmax
 * entries: eigenvalues, flags, a, a', b -- 68 bytes per entry) */
size_t fnftb_max_chunk_ex(const fnftb_ctx *ctx, size_t D, int deg0, size_t M, int npoly,
                          size_t extra_per_signal, size_t budget_bytes);

/* Stage B signals of D samples.  q (and r if rmode is EXPLICIT) are host pointers
 * (copied) or, if on_device != 0, device pointers that are used in place. */
int fnftb_set_signals(fnftb_ctx *ctx, size_t B, size_t D, const void *q, const void *r,
                      int on_device);
/* One shot: 1 (and the batch staged again) when the signal buffer still holds the host batch that the previous
 * fnftb_set_signals call uploaded -- saves the second host-to-device copy of SUBSAMPLE_AND_REFINE. */
int fnftb_signals_staged(fnftb_ctx *ctx, size_t B, size_t D, const void *q);
/* NVTX ranges (FNFT_B200_NVTX=1): every kernel launch is wrapped in a range named like its entry of the profile
 * report; the host layer adds one range per public call and per phase of a chunk. */
void fnftb_range_push(const char *name);
void fnftb_range_pop(void);


/* 4SPLIT4 preprocessing of the staged signals: replaces them by the 2*D resampled and
 * weighted samples (device resident).  warn_host[B] (may be NULL) gets 1 where the
 * signal does not look band-limited. */
int fnftb_resample_4split4(fnftb_ctx *ctx, double eps_t, int32_t *warn_host);

/* Same with subsampling (src/private/fnft__nse_discretization.c:428-431,483-500): the shifts
 * are scaled by nskip and only the samples 0, nskip, 2*nskip, ... (Dsub of them) are kept. */
int fnftb_resample_4split4_sub(fnftb_ctx *ctx, double eps_t, size_t nskip, size_t Dsub,
                               int32_t *warn_host);
/* CF4_3 preprocessing (src/private/fnft__nse_discretization.c:505-531): shifts by -/+ sqrt(3/20)*eps_t
 * (scaled by nskip), the unshifted samples as middle node, 3x3 Gauss-node weights; the staged
 * signals are replaced by the 3*Dsub weighted samples. */
int fnftb_resample_cf4_3_sub(fnftb_ctx *ctx, double eps_t, size_t nskip, size_t Dsub,
                             int32_t *warn_host);
/* CF5_3 (wsel 2, 3 exponentials per step) and CF6_4 (wsel 3, 4 per step), :532-604: shifts by
 * -/+ sqrt(15)/10*eps_t, complex weights applied to q and (unconjugated) to r = -kappa*conj(q); the
 * explicit r samples stay on the device next to the weighted q.  wsel 1 = CF4_3. */
int fnftb_resample_cf_sub(fnftb_ctx *ctx, int wsel, int kappa, double eps_t, size_t nskip, size_t Dsub,
                          int32_t *warn_host);
/* Declares the staged signals to be CF4_3-preprocessed samples (wsel 1; 0 = BO / CF4_2), for callers
 * of the private API that pass preprocessed samples themselves. */
int fnftb_set_slow_weights(fnftb_ctx *ctx, int wsel);
/* ES4 (wsel 4) / TES4 (wsel 5) preprocessing of the staged signals: (q, q', q'') per sub-sampled grid point by
 * central differences; the staged signals become [B][3*Dsub] (device resident) */
int fnftb_preprocess_es4(fnftb_ctx *ctx, int wsel, double eps_t, size_t nskip, size_t Dsub);
/* Plain subsampling of the staged signals (fnft__nse_discretization.c:463-470): keeps the
 * samples 0, nskip, ..., (Dsub-1)*nskip (device resident). */
int fnftb_subsample(fnftb_ctx *ctx, size_t nskip, size_t Dsub);

/* All roots of entry `ent` (0 = a(z)) of every transfer matrix held in the context
 * (fnft__poly_roots_fasteigen, src/private/fnft__poly_roots_fasteigen.c:29-48; Aberth-Ehrlich
 * iteration on the device, poly_roots.cuh).  roots_host: [B][deg]; info_host (may be NULL):
 * [B][4] = {leading zeros, effective degree, sweeps, roots that did not converge}. */
int fnftb_poly_roots(fnftb_ctx *ctx, int ent, void *roots_host, int32_t *info_host);

/* z -> lambda = log(z)/(i*lam_den) and box filter (order preserving) of the roots found by the LAST
 * fnftb_poly_roots / fnftb_nsep_floquet_roots call, on the device; only the survivors are copied:
 * lam_host[b*stride + i], i < count_host[b] (counts may exceed stride: then only stride values are
 * stored).  box == NULL: no filtering.  use_box3: the upper imaginary bound of signal b is the value
 * computed by the last fnftb_imbound call instead of box[3]. */
int fnftb_roots_lambda(fnftb_ctx *ctx, double lam_den, const double *box, int use_box3, void *lam_host,
                       size_t stride, int32_t *count_host);

/* leaves + product tree for the staged signals */
int fnftb_fscatter(fnftb_ctx *ctx, const fnftb_scatter_desc *desc);

/* product of n given 2x2 matrix polynomials of degree deg (host, reference layout
 * [4][n][deg+1]); leaves the result in the context like fnftb_fscatter (B = 1). */
int fnftb_fmult2x2(fnftb_ctx *ctx, size_t deg, size_t n, const void *p_host, int normalize);

/* degree of the transfer matrices currently held */
size_t fnftb_result_degree(const fnftb_ctx *ctx);

/* copy transfer matrices [B][4][deg+1] and exponents W[B] to the host */
int fnftb_get_transfer_matrix(fnftb_ctx *ctx, void *tm_host, int32_t *W_host);

/* per-signal status of the last fscatter / contspec (0 = ok) */
int fnftb_get_status(fnftb_ctx *ctx, int32_t *status_host);

/* replace entry-row `ent` of signal 0's transfer matrix by a host polynomial, or
 * load a standalone polynomial as a 1-signal / 1-entry "matrix" (for chirp-z of
 * arbitrary polynomials) */
int fnftb_set_polynomial(fnftb_ctx *ctx, size_t deg, const void *p_host);

/* chirp-z of the selected entries + epilogue.  out: [B][out_sstride] complex
 * (host, or device if on_device).  status_host (may be NULL): [B] int32. */
int fnftb_contspec(fnftb_ctx *ctx, const fnftb_contspec_desc *desc, void *out,
                   size_t out_sstride, int on_device, int32_t *status_host);

/* general 2x2 chaining of raw transfer-matrix values (fnft_kdvv): acc = second column of the product so far */
int fnftb_seg_compose_general(fnftb_ctx *ctx, size_t B, size_t M, int first);
/* epilogue of src/fnft_kdvv.c:186-203 on the chained (H12, H22) */
int fnftb_seg_finish_kdv(fnftb_ctx *ctx, size_t B, size_t M, double xi0, double eps_xi, double kdv_ph, double kdv_sqrtz,
                         void *out, size_t out_sstride, int on_device);
/* ---- continuous spectrum by segments (signals longer than one product tree) ------------ */
/* Largest number of samples (a power of two) whose transfer matrix one product tree can hold for this scheme. */
size_t fnftb_tree_max_samples(int scheme, int deg0);
/* Stage B pieces of Dseg samples, read with a row stride of `stride` samples from q (host or device pointer). */
int fnftb_set_signals_strided(fnftb_ctx *ctx, size_t B, size_t Dseg, const void *q, size_t stride, int on_device);
/* Device buffer [B][nent*M] for the values of the current piece (target of fnftb_contspec with on_device = 1):
 * nent = 2: (a, b) of fnft_nsev; nent = 4: [H12 | H22 | H11 | H21] of fnft_kdvv (FNFTB_MODE_RAW). */
void *fnftb_seg_buffer(fnftb_ctx *ctx, size_t B, size_t M, int nent);
/* acc = cur (first != 0) or acc = [a_s, -kappa b_s*; b_s, a_s*] acc, point by point on the real xi grid. */
int fnftb_seg_compose(fnftb_ctx *ctx, size_t B, size_t M, int kappa, int first);
/* Epilogue of src/fnft_nsev.c:846-876 on the chained (a, b): cstype 0 rho, 1 a|b, 2 rho|a|b. */
int fnftb_seg_finish(fnftb_ctx *ctx, size_t B, size_t M, int cstype, void *out, size_t out_sstride, int on_device,
                     int32_t *status_host);

/* Continuous spectrum of the staged signals with the slow discretizations BO (upsampling 1),
 * CF4_2 and CF4_3 (upsampling 2 and 3; the staged signals are the resampled ones): one product of D step matrices
 * per spectral point, src/fnft_nsev.c:794-814 + epilogue :836-876.  Uses mode / cstype / M / xi0 /
 * eps_xi / ph_* of the descriptor. */
int fnftb_slow_contspec(fnftb_ctx *ctx, const fnftb_contspec_desc *desc, int upsampling, int kappa,
                        double eps_t, void *out, size_t out_sstride, int on_device, int32_t *status_host);

/* ---- periodic NFT (grid search) --------------------------------------------------- */
typedef struct {
    double PHI0, PHI1;   /* angular range of the search on the unit circle */
    double lam_den;      /* lambda = log(z)/(i*lam_den), lam_den = 2*eps_t/(deg1*upsampling) */
    int filtering;       /* 0: none, 1: keep only values inside box */
    double box[4];
    double lam_shift;    /* added to every returned value */
    size_t Kmax, Mmax;   /* capacities of the per-signal output rows */
} fnftb_nsep_desc;

/* q[i] *= exp(2i*lam_shift*(T0 + eps_t*i)) on the staged signals (src/fnft_nsep.c:127-128) */
int fnftb_nsep_derotate(fnftb_ctx *ctx, double lam_shift, double T0, double eps_t);
size_t fnftb_nsep_chunk(const fnftb_ctx *ctx, size_t D_eff, int deg0, size_t budget_bytes);
int fnftb_nsep_gridsearch(fnftb_ctx *ctx, const fnftb_nsep_desc *desc, uint64_t *K_host,
                          void *main_host, uint64_t *M_host, void *aux_host, int32_t *status_host);

/* Remember / re-select the currently staged signals (pointer + length; slot 0 or 1), so that
 * a subsampled copy can be scattered in between (fnft_nsep subsample-and-refine). */
int fnftb_signals_save(fnftb_ctx *ctx, int slot);
int fnftb_signals_restore(fnftb_ctx *ctx, int slot);

/* Roots of the Floquet polynomial a(z) + a#(z) - rhs*2^-W of every transfer matrix held
 * (src/fnft_nsep.c:561-586).  roots_host: [B][deg]; info_host as for fnftb_poly_roots. */
int fnftb_nsep_floquet_roots(fnftb_ctx *ctx, double rhs, void *roots_host, int32_t *info_host);

/* Newton refinement of main (mode 0: trace + rhs) or auxiliary (mode 1: b) spectrum points on
 * the staged signals, src/fnft_nsep.c:708-835.  lam_host: [B][Kstride] in/out. */
typedef struct {
    int upsampling;   /* 1: BO, 2: CF4_2 */
    int kappa;
    int Kstride;
    int mode;
    int max_evals;
    double eps_t, rhs, tol;
} fnftb_refine_desc;
int fnftb_nsep_refine(fnftb_ctx *ctx, const fnftb_refine_desc *desc, const int32_t *K_host,
                      void *lam_host, int32_t *flag_host);

/* ---- bound states (Newton on the BO / CF4_2 recurrence) ------------------------ */
typedef struct {
    int upsampling;   /* 1: BO, 2: CF4_2, 3: CF4_3 */
    int Kmax;         /* stride of the per-signal eigenvalue arrays */
    double T0, T1, eps_t, bc;
    double lweight;   /* 1 (BO), 0.5 (CF4_2), 1 (CF4_3: per-sample weights inside the kernels) */
    double scl;       /* factor of a' */
    int niter;
    double box0, box1, box2; /* re_min, re_max, im_min */
    int use_box3;     /* 1: per-signal im_max computed by fnftb_imbound */
} fnftb_bound_desc;

/* per-signal 1.5*0.25*||q||^2 of the staged signals (kept on the device; copied to
 * box3_host [B] if not NULL) */
int fnftb_imbound(fnftb_ctx *ctx, int upsampling, double T0, double T1, double *box3_host);
/* Newton refinement of lam_host[B][Kmax] (in/out), K_host[B] guesses per signal;
 * flag_host[B][Kmax] (may be NULL): 3 = division by zero */
int fnftb_newton(fnftb_ctx *ctx, const fnftb_bound_desc *desc, const int32_t *K_host,
                 void *lam_host, int32_t *flag_host);
/* a, a', b at the given eigenvalues (each output [B][Kmax], may be NULL) */
int fnftb_normconsts(fnftb_ctx *ctx, const fnftb_bound_desc *desc, const int32_t *K_host,
                     const void *lam_host, void *a_host, void *ap_host, void *b_host);

/* ---- inverse transform (inverse_api.cu) ------------------------------------------------- */
/* fast inverse scattering of B transfer matrices [B][4][deg+1] -> q [B][deg] (deg = power of two <= 32768, degree-1
 * discretizations; modal = 1: 2SPLIT2_MODAL, 0: 2SPLIT2A); status_host[b] = 1 where |Q| >= 1 (kappa = -1) */
int fnftb_finvscatter(fnftb_ctx *ctx, size_t B, size_t deg, const void *tm, void *q, double eps_t, int kappa,
                      int modal, int on_device, int32_t *status_host);
/* Darboux transforms: K solitons per signal (bs sorted by descending imaginary part, nc norming constants, host
 * [B][K]) added to zero (seed = 0) or to the potential in q (seed = 1); q [B][D] */
int fnftb_inv_add_solitons(fnftb_ctx *ctx, size_t B, size_t K, size_t D, const void *bs, const void *nc, void *q,
                           double T0, double T1, int zc_point, int seed, int q_on_device);
/* transfer matrices [B][4][deg+1] (kept on the device for fnftb_finvscatter_staged) from continuous spectra, and
 * spectral factorisation; see inverse_api.cu */
int fnftb_inv_tm_from_contspec(fnftb_ctx *ctx, size_t B, size_t M, size_t D, size_t deg, const void *contspec_host,
                               int cstype, int kappa, double eps_t, size_t oversampling, int32_t *warn_host);
int fnftb_inv_tm_ab_from_iter(fnftb_ctx *ctx, size_t D, const void *contspec_host, int kappa, size_t max_iter,
                              int32_t *hit_max, int32_t *warn_host);
int fnftb_finvscatter_staged(fnftb_ctx *ctx, size_t B, size_t deg, void *q_host, double eps_t, int kappa, int modal,
                             int32_t *status_host);
int fnftb_specfact(fnftb_ctx *ctx, size_t B, size_t deg, const void *poly_host, void *result_host, size_t oversampling,
                   int kappa, int32_t *warn_host);

/* DFMA throughput of the context's device in TFLOP/s (probe kernel, ~10 ms); 0 on failure */
double fnftb_probe_fp64_tflops(fnftb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
