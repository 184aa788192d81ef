/*
 * fnft_b200 -- B200-native (sm_100a) implementation of the fast forward nonlinear
 * Fourier transform hot path of FNFT 0.4.1, exposed through FNFT's own C ABI plus
 * batched entry points.
 *
 * This header is the drop-in boundary.  Every declaration cites the reference
 * interface it replaces (paths relative to the FNFT source tree).  Type names,
 * enumerator names AND values, struct layouts, argument order, ownership rules and
 * error codes are those of the reference, so a program compiled against FNFT's
 * headers links and runs unchanged against libfnft_b200.so.  The header files
 * fnft.h, fnft_nsev.h, fnft_kdvv.h, fnft_nsep.h, ... in this directory are thin
 * forwarders to this file.
 *
 * All computation of the hot path runs on the GPU.  There is no CPU fallback: when
 * no CUDA device is usable the entry points return FNFT_EC_OTHER and print the
 * reason through the fnft_errwarn channel.
 */
#ifndef FNFT_B200_H
#define FNFT_B200_H

#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#ifndef __cplusplus
#include <complex.h>
#else
#include <complex>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- numeric types: include/fnft_numtypes.h:40-62 ----------------------------- */
typedef double FNFT_REAL;
#ifndef __cplusplus
typedef double complex FNFT_COMPLEX;
#else
typedef std::complex<double> FNFT_COMPLEX;
#endif
typedef int32_t FNFT_INT;
typedef size_t FNFT_UINT;
#define FNFT_EPSILON DBL_EPSILON
#define FNFT_NAN NAN
#define FNFT_INF INFINITY
#define FNFT_PI acos(-1.0)
/* thin aliases of the C math / complex functions (include/fnft_numtypes.h:64-110), kept
 * so that sources written against the reference headers compile unchanged */
#define FNFT_FABS(X) fabs(X)
#define FNFT_SQRT(X) sqrt(X)
#define FNFT_COS(X) cos(X)
#define FNFT_SIN(X) sin(X)
#define FNFT_SINH(X) sinh(X)
#define FNFT_COSH(X) cosh(X)
#define FNFT_ATAN(X) atan(X)
#define FNFT_ATANH(X) atanh(X)
#define FNFT_LOG(X) log(X)
#define FNFT_LOG2(X) log2(X)
#define FNFT_POW(X, Y) pow(X, Y)
#define FNFT_GAMMA(X) tgamma(X)
#define FNFT_FLOOR(X) floor(X)
#define FNFT_ROUND(X) round(X)
#define FNFT_CEIL(X) ceil(X)
#define FNFT_HYPOT(X, Y) hypot(X, Y)
#define FNFT_CREAL(X) creal(X)
#define FNFT_CIMAG(X) cimag(X)
#define FNFT_CABS(X) cabs(X)
#define FNFT_CARG(X) carg(X)
#define FNFT_CONJ(X) conj(X)
#define FNFT_CPOW(X, Y) cpow(X, Y)
#define FNFT_CEXP(X) cexp(X)
#define FNFT_CLOG(X) clog(X)
#define FNFT_CSQRT(X) csqrt(X)
#define FNFT_CSINH(X) csinh(X)
#define FNFT_CCOSH(X) ccosh(X)
#define FNFT_CSIN(X) csin(X)
#define FNFT_CCOS(X) ccos(X)

/* ---- error codes and message channel: include/fnft_errwarn.h:44-108 ------------ */
typedef FNFT_INT (*fnft_printf_ptr_t)(const char *, ...);
#define FNFT_SUCCESS 0
#define FNFT_EC_NOMEM 1
#define FNFT_EC_INVALID_ARGUMENT 2
#define FNFT_EC_DIV_BY_ZERO 3
#define FNFT_EC_TEST_FAILED 4
#define FNFT_EC_OTHER 5
#define FNFT_EC_NOT_YET_IMPLEMENTED 6
#define FNFT_EC_SANITY_CHECK_FAILED 7
#define FNFT_EC_ASSERTION_FAILED 8
void fnft_errwarn_setprintf(fnft_printf_ptr_t printf_ptr);   /* fnft_errwarn.h:101 */
fnft_printf_ptr_t fnft_errwarn_getprintf(void);               /* fnft_errwarn.h:108 */

/* ---- version: include/fnft_version.h:47 ----------------------------------------- */
#define FNFT_VERSION_MAJOR 0
#define FNFT_VERSION_MINOR 4
#define FNFT_VERSION_PATCH 1
#define FNFT_VERSION_SUFFIX ""
#define FNFT_VERSION_SUFFIX_MAXLEN 8
FNFT_INT fnft_version(FNFT_UINT *major, FNFT_UINT *minor, FNFT_UINT *patch,
                      char suffix[FNFT_VERSION_SUFFIX_MAXLEN + 1]);

/* ---- discretizations: include/fnft_nse_discretization_t.h:104-133 -------------- */
typedef enum {
    fnft_nse_discretization_2SPLIT2_MODAL,
    fnft_nse_discretization_BO,
    fnft_nse_discretization_2SPLIT1A,
    fnft_nse_discretization_2SPLIT1B,
    fnft_nse_discretization_2SPLIT2A,
    fnft_nse_discretization_2SPLIT2B,
    fnft_nse_discretization_2SPLIT2S,
    fnft_nse_discretization_2SPLIT3A,
    fnft_nse_discretization_2SPLIT3B,
    fnft_nse_discretization_2SPLIT3S,
    fnft_nse_discretization_2SPLIT4A,
    fnft_nse_discretization_2SPLIT4B, /* = 11, default of fnft_nsev */
    fnft_nse_discretization_2SPLIT5A,
    fnft_nse_discretization_2SPLIT5B,
    fnft_nse_discretization_2SPLIT6A,
    fnft_nse_discretization_2SPLIT6B,
    fnft_nse_discretization_2SPLIT7A,
    fnft_nse_discretization_2SPLIT7B,
    fnft_nse_discretization_2SPLIT8A,
    fnft_nse_discretization_2SPLIT8B,
    fnft_nse_discretization_4SPLIT4A,
    fnft_nse_discretization_4SPLIT4B, /* = 21 */
    fnft_nse_discretization_CF4_2,
    fnft_nse_discretization_CF4_3,
    fnft_nse_discretization_CF5_3,
    fnft_nse_discretization_CF6_4,
    fnft_nse_discretization_ES4,
    fnft_nse_discretization_TES4
} fnft_nse_discretization_t;

/* include/fnft_kdv_discretization_t.h:96-122 */
typedef enum {
    fnft_kdv_discretization_2SPLIT1A,
    fnft_kdv_discretization_2SPLIT1B,
    fnft_kdv_discretization_2SPLIT2A,
    fnft_kdv_discretization_2SPLIT2B,
    fnft_kdv_discretization_2SPLIT2S,
    fnft_kdv_discretization_2SPLIT3A,
    fnft_kdv_discretization_2SPLIT3B,
    fnft_kdv_discretization_2SPLIT3S,
    fnft_kdv_discretization_2SPLIT4A,
    fnft_kdv_discretization_2SPLIT4B, /* = 9 */
    fnft_kdv_discretization_2SPLIT5A,
    fnft_kdv_discretization_2SPLIT5B,
    fnft_kdv_discretization_2SPLIT6A,
    fnft_kdv_discretization_2SPLIT6B,
    fnft_kdv_discretization_2SPLIT7A,
    fnft_kdv_discretization_2SPLIT7B,
    fnft_kdv_discretization_2SPLIT8A,
    fnft_kdv_discretization_2SPLIT8B,
    fnft_kdv_discretization_4SPLIT4A,
    fnft_kdv_discretization_4SPLIT4B, /* = 19 */
    fnft_kdv_discretization_BO,
    fnft_kdv_discretization_CF4_2,
    fnft_kdv_discretization_CF4_3,
    fnft_kdv_discretization_CF5_3,
    fnft_kdv_discretization_CF6_4
} fnft_kdv_discretization_t;

/* include/private/fnft__akns_discretization_t.h:43-72 */
typedef enum {
    fnft__akns_discretization_2SPLIT2_MODAL,
    fnft__akns_discretization_2SPLIT1A,
    fnft__akns_discretization_2SPLIT1B,
    fnft__akns_discretization_2SPLIT2A,
    fnft__akns_discretization_2SPLIT2B,
    fnft__akns_discretization_2SPLIT2S,
    fnft__akns_discretization_2SPLIT3A,
    fnft__akns_discretization_2SPLIT3B,
    fnft__akns_discretization_2SPLIT3S,
    fnft__akns_discretization_2SPLIT4A,
    fnft__akns_discretization_2SPLIT4B,
    fnft__akns_discretization_2SPLIT5A,
    fnft__akns_discretization_2SPLIT5B,
    fnft__akns_discretization_2SPLIT6A,
    fnft__akns_discretization_2SPLIT6B,
    fnft__akns_discretization_2SPLIT7A,
    fnft__akns_discretization_2SPLIT7B,
    fnft__akns_discretization_2SPLIT8A,
    fnft__akns_discretization_2SPLIT8B,
    fnft__akns_discretization_BO,
    fnft__akns_discretization_4SPLIT4A,
    fnft__akns_discretization_4SPLIT4B,
    fnft__akns_discretization_CF4_2,
    fnft__akns_discretization_CF4_3,
    fnft__akns_discretization_CF5_3,
    fnft__akns_discretization_CF6_4,
    fnft__akns_discretization_ES4,
    fnft__akns_discretization_TES4
} fnft__akns_discretization_t;

/* ---- fnft_nsev: include/fnft_nsev.h:51-55,91-95,108-112,130-134,198-208 --------- */
typedef enum {
    fnft_nsev_bsfilt_NONE,
    fnft_nsev_bsfilt_BASIC,
    fnft_nsev_bsfilt_FULL
} fnft_nsev_bsfilt_t;

typedef enum {
    fnft_nsev_bsloc_FAST_EIGENVALUE,
    fnft_nsev_bsloc_NEWTON,
    fnft_nsev_bsloc_SUBSAMPLE_AND_REFINE
} fnft_nsev_bsloc_t;

typedef enum {
    fnft_nsev_dstype_NORMING_CONSTANTS,
    fnft_nsev_dstype_RESIDUES,
    fnft_nsev_dstype_BOTH
} fnft_nsev_dstype_t;

typedef enum {
    fnft_nsev_cstype_REFLECTION_COEFFICIENT,
    fnft_nsev_cstype_AB,
    fnft_nsev_cstype_BOTH
} fnft_nsev_cstype_t;

/* 48 bytes on LP64, identical member order to the reference */
typedef struct {
    fnft_nsev_bsfilt_t bound_state_filtering;
    fnft_nsev_bsloc_t bound_state_localization;
    FNFT_UINT niter;
    FNFT_UINT Dsub;
    fnft_nsev_dstype_t discspec_type;
    fnft_nsev_cstype_t contspec_type;
    FNFT_INT normalization_flag;
    fnft_nse_discretization_t discretization;
    FNFT_UINT richardson_extrapolation_flag;
} fnft_nsev_opts_t;

fnft_nsev_opts_t fnft_nsev_default_opts(void);                         /* fnft_nsev.h:226 */
FNFT_UINT fnft_nsev_max_K(const FNFT_UINT D, fnft_nsev_opts_t const *const opts); /* :241 */

/* include/fnft_nsev.h:371-376 -- same arguments, ownership and return codes */
FNFT_INT fnft_nsev(const FNFT_UINT D, FNFT_COMPLEX *const q, FNFT_REAL const *const T,
                   const FNFT_UINT M, FNFT_COMPLEX *const contspec, FNFT_REAL const *const XI,
                   FNFT_UINT *const K_ptr, FNFT_COMPLEX *const bound_states,
                   FNFT_COMPLEX *const normconsts_or_residues, const FNFT_INT kappa,
                   fnft_nsev_opts_t *opts);

/* ---- fnft_kdvv: include/fnft_kdvv.h:46-48,58,104-109 ---------------------------- */
typedef struct {
    fnft_kdv_discretization_t discretization;
} fnft_kdvv_opts_t;
fnft_kdvv_opts_t fnft_kdvv_default_opts(void);
FNFT_INT fnft_kdvv(const FNFT_UINT D, FNFT_COMPLEX *const u, FNFT_REAL const *const T,
                   const FNFT_UINT M, FNFT_COMPLEX *const contspec, FNFT_REAL const *const XI,
                   FNFT_UINT *const K_ptr, FNFT_COMPLEX *const bound_states,
                   FNFT_COMPLEX *const normconsts_or_residues, fnft_kdvv_opts_t *opts);

/* ---- fnft_nsep: include/fnft_nsep.h:47-51,60-64,140-151,172,263-267 ------------- */
typedef enum {
    fnft_nsep_loc_SUBSAMPLE_AND_REFINE,
    fnft_nsep_loc_GRIDSEARCH,
    fnft_nsep_loc_MIXED
} fnft_nsep_loc_t;

typedef enum {
    fnft_nsep_filt_NONE,
    fnft_nsep_filt_MANUAL,
    fnft_nsep_filt_AUTO
} fnft_nsep_filt_t;

/* 96 bytes on LP64 */
typedef struct {
    fnft_nsep_loc_t localization;
    fnft_nsep_filt_t filtering;
    FNFT_REAL bounding_box[4];
    FNFT_UINT max_evals;
    fnft_nse_discretization_t discretization;
    FNFT_INT normalization_flag;
    FNFT_REAL floquet_range[2];
    FNFT_UINT points_per_spine;
    FNFT_UINT Dsub;
    FNFT_REAL tol;
} fnft_nsep_opts_t;
fnft_nsep_opts_t fnft_nsep_default_opts(void);
FNFT_INT fnft_nsep(const FNFT_UINT D, FNFT_COMPLEX const *const q, FNFT_REAL const *const T,
                   FNFT_REAL const phase_shift, FNFT_UINT *const K_ptr,
                   FNFT_COMPLEX *const main_spec, FNFT_UINT *const M_ptr,
                   FNFT_COMPLEX *const aux_spec, FNFT_REAL *const sheet_indices,
                   const FNFT_INT kappa, fnft_nsep_opts_t *opts);

/* ---- private symbols the reference's own unit tests link against ---------------- */
/* include/private/fnft__poly_fmult.h:160-230, src/private/fnft__poly_fmult.c:40-43,381 */
FNFT_UINT fnft__poly_fmult2x2_numel(FNFT_UINT deg, FNFT_UINT n);
FNFT_INT fnft__poly_fmult2x2(FNFT_UINT *const d, FNFT_UINT n, FNFT_COMPLEX *const p,
                             FNFT_COMPLEX *const result, FNFT_INT *const W_ptr);
/* include/private/fnft__poly_roots_fasteigen.h:45, src/private/fnft__poly_roots_fasteigen.c:29-48:
 * all deg roots of p (deg+1 coefficients, highest power first).  The reference calls eiscor's
 * companion-pencil QR; this library runs an Aberth-Ehrlich iteration on the GPU. */
FNFT_INT fnft__poly_roots_fasteigen(const FNFT_UINT deg, FNFT_COMPLEX const *const p,
                                    FNFT_COMPLEX *const roots);

/* include/private/fnft__poly_chirpz.h:66, src/private/fnft__poly_chirpz.c:33 */
FNFT_INT fnft__poly_chirpz(const FNFT_UINT deg, FNFT_COMPLEX const *const p,
                           const FNFT_COMPLEX A, const FNFT_COMPLEX W, const FNFT_UINT M,
                           FNFT_COMPLEX *const result);
/* include/private/fnft__akns_fscatter.h:55,92, src/private/fnft__akns_fscatter.c:33,64 */
FNFT_UINT fnft__akns_fscatter_numel(FNFT_UINT D, fnft__akns_discretization_t discretization);
FNFT_INT fnft__akns_fscatter(const FNFT_UINT D, FNFT_COMPLEX const *const q,
                             FNFT_COMPLEX const *const r, const FNFT_REAL eps_t,
                             FNFT_COMPLEX *const result, FNFT_UINT *const deg_ptr,
                             FNFT_INT *const W_ptr, fnft__akns_discretization_t discretization);
/* include/private/fnft__nse_fscatter.h:48,84, src/private/fnft__nse_fscatter.c:30,44 */
FNFT_UINT fnft__nse_fscatter_numel(FNFT_UINT D, fnft_nse_discretization_t discretization);
FNFT_INT fnft__nse_fscatter(const FNFT_UINT D, FNFT_COMPLEX const *const q, const FNFT_REAL eps_t,
                            const FNFT_INT kappa, FNFT_COMPLEX *const result,
                            FNFT_UINT *const deg_ptr, FNFT_INT *const W_ptr,
                            fnft_nse_discretization_t discretization);
/* include/private/fnft__kdv_fscatter.h:52,108, src/private/fnft__kdv_fscatter.c:32,45 */
FNFT_UINT fnft__kdv_fscatter_numel(FNFT_UINT D, fnft_kdv_discretization_t discretization);
FNFT_INT fnft__kdv_fscatter(const FNFT_UINT D, FNFT_COMPLEX const *const u, const FNFT_REAL eps_t,
                            FNFT_COMPLEX *const result, FNFT_UINT *const deg_ptr,
                            FNFT_INT *const W_ptr, fnft_kdv_discretization_t discretization);
/* include/private/fnft__nse_scatter.h:118, src/private/fnft__nse_scatter_bound_states.c:29 */
FNFT_INT fnft__nse_scatter_bound_states(const FNFT_UINT D, FNFT_COMPLEX const *const q,
                                        FNFT_COMPLEX *r, FNFT_REAL const *const T, FNFT_UINT K,
                                        FNFT_COMPLEX *bound_states, FNFT_COMPLEX *a_vals,
                                        FNFT_COMPLEX *aprime_vals, FNFT_COMPLEX *b,
                                        fnft_nse_discretization_t discretization,
                                        FNFT_UINT skip_b_flag);

/* include/private/fnft__errwarn.h:36,125, src/private/fnft__errwarn.c:28-45: what FNFT__ERRMSG / FNFT__WARN expand to */
FNFT_INT fnft__errmsg_aux(const FNFT_INT ec, const char *func, const FNFT_INT line, const char *msg);
void fnft__warn_aux(const char *func, const FNFT_INT line, const char *msg);
/* include/private/fnft__poly_fmult.h:162,183, src/private/fnft__poly_fmult.c:35-38,152-237: product of n scalar
 * polynomials (runs on the GPU tree as diag(p, 1) matrices) */
FNFT_UINT fnft__poly_fmult_numel(FNFT_UINT deg, FNFT_UINT n);
FNFT_INT fnft__poly_fmult(FNFT_UINT *const d, FNFT_UINT n, FNFT_COMPLEX *const p, FNFT_INT *const W_ptr);
/* include/private/fnft__poly_eval.h:57,93, src/private/fnft__poly_eval.c:24-91 (host, O(deg) per point) */
FNFT_INT fnft__poly_eval(const FNFT_UINT deg, FNFT_COMPLEX const *const p, const FNFT_UINT nz,
                         FNFT_COMPLEX *const z);
FNFT_INT fnft__poly_evalderiv(const FNFT_UINT deg, FNFT_COMPLEX const *const p, const FNFT_UINT nz,
                              FNFT_COMPLEX *const z, FNFT_COMPLEX *const deriv);
/* include/private/fnft__misc.h:42-219, src/private/fnft__misc.c:28-324: helpers the reference's test programs
 * call (comparison metrics, filters); host code on caller memory.  Not exported: the single-shift resampling
 * routine of that header (:241; the band-limited shift only exists as GPU preprocessing) and the pair-product
 * routines of fnft__poly_fmult.h (:39,82,136), whose signatures carry Kiss FFT plan handles
 * (include/private/fnft__fft_wrapper_plan_t.h). */
void fnft__misc_print_buf(const FNFT_INT len, FNFT_COMPLEX const *const buf, char const *const varname);
FNFT_REAL fnft__misc_rel_err(const FNFT_INT len, FNFT_COMPLEX const *const vec_numer,
                             FNFT_COMPLEX const *const vec_exact);
FNFT_REAL fnft__misc_hausdorff_dist(const FNFT_UINT lenA, FNFT_COMPLEX const *const vecA, const FNFT_UINT lenB,
                                    FNFT_COMPLEX const *const vecB);
FNFT_COMPLEX fnft__misc_sech(FNFT_COMPLEX Z);
FNFT_REAL fnft__misc_l2norm2(const FNFT_UINT N, FNFT_COMPLEX const *const Z, const FNFT_REAL a, const FNFT_REAL b);
FNFT_INT fnft__misc_filter(FNFT_UINT *const N_ptr, FNFT_COMPLEX *const vals, FNFT_COMPLEX *const rearrange_as_well,
                           FNFT_REAL const *const bounding_box);
FNFT_INT fnft__misc_filter_inv(FNFT_UINT *const N_ptr, FNFT_COMPLEX *const vals,
                               FNFT_COMPLEX *const rearrange_as_well, FNFT_REAL const *const bounding_box);
FNFT_INT fnft__misc_filter_nonreal(FNFT_UINT *N_ptr, FNFT_COMPLEX *const vals, const FNFT_REAL tol_im);
FNFT_INT fnft__misc_merge(FNFT_UINT *N_ptr, FNFT_COMPLEX *const vals, FNFT_REAL tol);
FNFT_INT fnft__misc_downsample(const FNFT_UINT D, FNFT_COMPLEX const *const q, FNFT_UINT *const Dsub_ptr,
                               FNFT_COMPLEX **qsub_ptr, FNFT_UINT *const first_last_index);
FNFT_COMPLEX fnft__misc_CSINC(FNFT_COMPLEX x);
FNFT_UINT fnft__misc_nextpowerof2(const FNFT_UINT number);

/* ---- NEW: batched entry points (not in the reference) ------------ */
/*
 * B independent signals, row-major: q[b*D + n].  T, XI, M, kappa and opts are shared
 * by the batch.  contspec[b*len + i] with len = M, 2M or 3M according to
 * opts->contspec_type (pass NULL to skip).  Bound states: K[b] holds on entry the
 * number of initial guesses stored in bound_states[b*Kmax ...] and on exit the
 * number found; normconsts_or_residues[b*nlen + i] with nlen = Kmax (2*Kmax for
 * dstype_BOTH).  Pass bound_states = NULL to skip the discrete spectrum.  opts is NOT
 * modified (NULL = defaults).  ret_codes[b] (may be NULL) receives the per-signal
 * return code; the function result is the first nonzero code or FNFT_SUCCESS.
 */
FNFT_INT fnft_nsev_batch(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                         FNFT_REAL const *const T, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                         FNFT_REAL const *const XI, FNFT_UINT *const K, const FNFT_UINT Kmax,
                         FNFT_COMPLEX *const bound_states,
                         FNFT_COMPLEX *const normconsts_or_residues, const FNFT_INT kappa,
                         fnft_nsev_opts_t const *opts, FNFT_INT *const ret_codes);

FNFT_INT fnft_kdvv_batch(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const u,
                         FNFT_REAL const *const T, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                         FNFT_REAL const *const XI, fnft_kdvv_opts_t const *opts,
                         FNFT_INT *const ret_codes);

/*
 * main_spec[b*Kmax + i], aux_spec[b*Mmax + i]; K[b] / Mcount[b] receive the counts.
 */
FNFT_INT fnft_nsep_batch(const FNFT_UINT B, const FNFT_UINT D, FNFT_COMPLEX const *const q,
                         FNFT_REAL const *const T, FNFT_REAL const phase_shift,
                         FNFT_UINT *const K, const FNFT_UINT Kmax, FNFT_COMPLEX *const main_spec,
                         FNFT_UINT *const Mcount, const FNFT_UINT Mmax,
                         FNFT_COMPLEX *const aux_spec, const FNFT_INT kappa,
                         fnft_nsep_opts_t const *opts, FNFT_INT *const ret_codes);

/* ---- NEW: runtime control ------------------------------------------------------ */
/* Number of usable CUDA devices (0 => every transform returns FNFT_EC_OTHER). */
FNFT_INT fnft_b200_device_count(void);
/* Select the device used by the calling thread's subsequent calls (default 0, or
 * the value of the environment variable FNFT_B200_DEVICE). */
FNFT_INT fnft_b200_set_device(FNFT_INT device);
/* Several GPUs behind ONE batched call: give the calling thread a set of n
 * devices (n = 0 or 1 switches the fan-out off; an id may repeat = several contexts on one GPU).  A subsequent fnft_nsev_batch /
 * fnft_kdvv_batch / fnft_nsep_batch with HOST buffers splits its batch into n contiguous shards,
 * runs shard i on devices[i] (persistent worker threads, one context per device) and lands all
 * results in the caller's [B][...] arrays; signals are independent, so nothing is exchanged between
 * the devices.  The environment variable FNFT_B200_DEVICES ("all" or "0,1,2,3") sets the default.
 * fnft_b200_get_devices returns the number of devices in the set (and copies up to `capacity` ids). */
FNFT_INT fnft_b200_set_devices(FNFT_INT n, FNFT_INT const *devices);
FNFT_INT fnft_b200_get_devices(FNFT_INT *devices, FNFT_INT capacity);
/* Declare that q / contspec pointers passed to the *_batch functions by this thread
 * are DEVICE pointers on the selected device (1) or host pointers (0, default).
 * With device pointers the calls are asynchronous; use fnft_b200_synchronize. */
FNFT_INT fnft_b200_set_device_pointers(FNFT_INT flag);
FNFT_INT fnft_b200_synchronize(void);
/* Workspace budget in bytes for the calling thread's context (0 = automatic). */
FNFT_INT fnft_b200_set_workspace_limit(FNFT_UINT bytes);
/* CUDA stream (cudaStream_t as void*) the calling thread's context launches on. */
void *fnft_b200_stream(void);
/* Kernels launched by this process so far. */
unsigned long long fnft_b200_launch_count(void);
/* Per-launch CUDA-event timing of the library's kernels (for benchmarking): enable,
 * run, then read a text report with one "kernel_name launches total_ms" line per
 * kernel; reading clears the records. */
void fnft_b200_profile_enable(FNFT_INT on);
const char *fnft_b200_profile_report(void);
/* Releases the calling thread's GPU context (and the workers of fnft_b200_set_devices). */
void fnft_b200_release(void);
/* Measured DFMA throughput (TFLOP/s) of the calling thread's device: the FP64 roofline denominator. */
double fnft_b200_probe_fp64_tflops(void);

/* ---- inverse NFT (include/fnft_nsev_inverse.h) --------------------------------------- */
typedef enum {
    fnft_nsev_inverse_cstype_REFLECTION_COEFFICIENT,
    fnft_nsev_inverse_cstype_B_OF_XI,
    fnft_nsev_inverse_cstype_B_OF_TAU
} fnft_nsev_inverse_cstype_t; /* include/fnft_nsev_inverse.h:54-58 */
typedef enum {
    fnft_nsev_inverse_dstype_NORMING_CONSTANTS,
    fnft_nsev_inverse_dstype_RESIDUES
} fnft_nsev_inverse_dstype_t; /* :72-75 */
typedef enum {
    fnft_nsev_inverse_csmethod_DEFAULT,
    fnft_nsev_inverse_csmethod_TFMATRIX_CONTAINS_REFL_COEFF,
    fnft_nsev_inverse_csmethod_TFMATRIX_CONTAINS_AB_FROM_ITER,
    fnft_nsev_inverse_csmethod_USE_SEED_POTENTIAL_INSTEAD
} fnft_nsev_inverse_csmethod_t; /* :105-110 */
typedef struct {
    fnft_nse_discretization_t discretization;
    fnft_nsev_inverse_cstype_t contspec_type;
    fnft_nsev_inverse_csmethod_t contspec_inversion_method;
    fnft_nsev_inverse_dstype_t discspec_type;
    FNFT_UINT max_iter;
    FNFT_UINT oversampling_factor;
} fnft_nsev_inverse_opts_t; /* :151-158, 32 bytes on LP64 */
fnft_nsev_inverse_opts_t fnft_nsev_inverse_default_opts(void); /* :168 */
FNFT_INT fnft_nsev_inverse_XI(const FNFT_UINT D, FNFT_REAL const *const T, const FNFT_UINT M, FNFT_REAL *const XI,
                              const fnft_nse_discretization_t discretization); /* :187-190 */
/* include/fnft_nsev_inverse.h:258-263, src/fnft_nsev_inverse.c:121-249.  Like the reference, contspec is modified
 * in place (boundary phase factors, Blaschke precompensation). */
FNFT_INT fnft_nsev_inverse(const FNFT_UINT M, FNFT_COMPLEX *const contspec, FNFT_REAL const *const XI,
                           FNFT_UINT const K, FNFT_COMPLEX const *const bound_states,
                           FNFT_COMPLEX const *const normconsts_or_residues, const FNFT_UINT D,
                           FNFT_COMPLEX *const q, FNFT_REAL const *const T, const FNFT_INT kappa,
                           fnft_nsev_inverse_opts_t *opts_ptr);
/* NEW: B independent inverse transforms with shared M, XI, K, D, T, kappa and options; row-major arrays
 * contspec[b*M + i] (may be NULL), bound_states[b*K + i], normconsts_or_residues[b*K + i], q[b*D + n];
 * ret_codes[b] (may be NULL) receives the per-signal code. */
FNFT_INT fnft_nsev_inverse_batch(const FNFT_UINT B, const FNFT_UINT M, FNFT_COMPLEX *const contspec,
                                 FNFT_REAL const *const XI, FNFT_UINT const K,
                                 FNFT_COMPLEX const *const bound_states,
                                 FNFT_COMPLEX const *const normconsts_or_residues, const FNFT_UINT D,
                                 FNFT_COMPLEX *const q, FNFT_REAL const *const T, const FNFT_INT kappa,
                                 fnft_nsev_inverse_opts_t const *opts_ptr, FNFT_INT *ret_codes);
/* include/private/fnft__nse_finvscatter.h:61-63, src/private/fnft__nse_finvscatter.c:243-366 */
FNFT_INT fnft__nse_finvscatter(const FNFT_UINT deg, FNFT_COMPLEX *const transfer_matrix, FNFT_COMPLEX *const q,
                               const FNFT_REAL eps_t, const FNFT_INT kappa,
                               const fnft_nse_discretization_t discretization);
/* include/private/fnft__poly_specfact.h:62-66, src/private/fnft__poly_specfact.c:25-147 */
FNFT_INT fnft__poly_specfact(const FNFT_UINT deg, FNFT_COMPLEX const *const poly, FNFT_COMPLEX *const result,
                             const FNFT_UINT oversampling_factor, const FNFT_INT kappa);

#ifdef __cplusplus
}
#endif

/* ---- optional short names, as in the reference headers -------------------------- */
#ifdef FNFT_ENABLE_SHORT_NAMES
#define REAL FNFT_REAL
#define COMPLEX FNFT_COMPLEX
#define INT FNFT_INT
#define UINT FNFT_UINT
#define SUCCESS FNFT_SUCCESS
#define nsev_bsfilt_NONE fnft_nsev_bsfilt_NONE
#define nsev_bsfilt_BASIC fnft_nsev_bsfilt_BASIC
#define nsev_bsfilt_FULL fnft_nsev_bsfilt_FULL
#define nsev_bsloc_FAST_EIGENVALUE fnft_nsev_bsloc_FAST_EIGENVALUE
#define nsev_bsloc_NEWTON fnft_nsev_bsloc_NEWTON
#define nsev_bsloc_SUBSAMPLE_AND_REFINE fnft_nsev_bsloc_SUBSAMPLE_AND_REFINE
#define nsev_dstype_NORMING_CONSTANTS fnft_nsev_dstype_NORMING_CONSTANTS
#define nsev_dstype_RESIDUES fnft_nsev_dstype_RESIDUES
#define nsev_dstype_BOTH fnft_nsev_dstype_BOTH
#define nsev_cstype_REFLECTION_COEFFICIENT fnft_nsev_cstype_REFLECTION_COEFFICIENT
#define nsev_cstype_AB fnft_nsev_cstype_AB
#define nsev_cstype_BOTH fnft_nsev_cstype_BOTH
#define nsev_opts_t fnft_nsev_opts_t
#define kdvv_opts_t fnft_kdvv_opts_t
#define nsep_opts_t fnft_nsep_opts_t
#define nse_discretization_t fnft_nse_discretization_t
#define kdv_discretization_t fnft_kdv_discretization_t
#define akns_discretization_t fnft__akns_discretization_t
#endif

#endif /* FNFT_B200_H */
