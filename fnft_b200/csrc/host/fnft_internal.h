/*
 * fnft_b200 host library -- internal declarations shared by the C sources.
 * The host side is plain C99 (as the reference library is); it reaches the GPU only
 * through the thin C-ABI of ../cuda/fnftb_device.h.
 */
#ifndef FNFT_B200_INTERNAL_H
#define FNFT_B200_INTERNAL_H

#include "fnft_b200.h"
#include "../cuda/fnftb_device.h"
#include <string.h>

/* ---- error / warning channel (format of src/private/fnft__errwarn.c:28-45) ---- */
FNFT_INT fnftb__errmsg(const FNFT_INT ec, const char *func, const FNFT_INT line, const char *msg);
void fnftb__warn(const char *func, const FNFT_INT line, const char *msg);
#define ERRMSG(ec, msg) fnftb__errmsg((ec), __func__, __LINE__, (msg))
#define WARN(msg) fnftb__warn(__func__, __LINE__, (msg))
#define E_NOMEM ERRMSG(FNFT_EC_NOMEM, "Out of memory.")
#define E_INVALID_ARGUMENT(name) ERRMSG(FNFT_EC_INVALID_ARGUMENT, "Invalid argument " #name ".")
#define E_SUBROUTINE(ec) ERRMSG(-abs(ec), "Subroutine failure.")
#define E_DIV_BY_ZERO ERRMSG(FNFT_EC_DIV_BY_ZERO, "Division by zero.")
#define E_OTHER(msg) ERRMSG(FNFT_EC_OTHER, (msg))
#define E_NOT_YET_IMPLEMENTED(name, msg) \
    ERRMSG(FNFT_EC_NOT_YET_IMPLEMENTED, "Not yet implemented (" #name "). " #msg)
#define E_ASSERTION_FAILED ERRMSG(FNFT_EC_ASSERTION_FAILED, "Assertion failed.")
#define CHECK_RETCODE(ret_code, label)           \
    {                                            \
        if ((ret_code) != FNFT_SUCCESS) {        \
            (ret_code) = E_SUBROUTINE(ret_code); \
            goto label;                          \
        }                                        \
    }
/* turns a failure of the device layer into an FNFT error (message forwarded) */
FNFT_INT fnftb__device_error(const char *func, const FNFT_INT line);
#define E_DEVICE fnftb__device_error(__func__, __LINE__)

/* ---- per-thread runtime state (fnft_runtime.c) --------------------------------- */
fnftb_ctx *fnftb__ctx(void);          /* NULL (after printing why) if no GPU */
int fnftb__device_pointers(void);     /* flag set by fnft_b200_set_device_pointers */
size_t fnftb__workspace_limit(void);
/* several devices behind one batched call (fnft_runtime.c) */
typedef void (*fnftb_shard_fn)(void *arg, int shard, int nshards);
int fnftb__fanout_shards(FNFT_UINT B); /* 1 = run in the calling thread */
int fnftb__fanout_run(int nshards, fnftb_shard_fn fn, void *arg);
void fnftb__shard_range(FNFT_UINT B, int shard, int nshards, FNFT_UINT *b0, FNFT_UINT *b1);
void fnftb__fanout_shutdown(void);
int fnftb__pipe_chunks(void);
FNFT_UINT fnftb__pipe_step(FNFT_UINT b0, FNFT_UINT B, FNFT_UINT chunk);

/* ln|z| and arg z of a complex double, accurate to far below one ulp of |z|-1 */
void fnftb__logpolar(FNFT_COMPLEX z, double *ln_abs, double *arg);

/* ---- discretization tables (fnft_discretization.c) ----------------------------- */
FNFT_UINT fnftb__akns_degree(fnft__akns_discretization_t d);
FNFT_UINT fnftb__akns_upsampling(fnft__akns_discretization_t d);
FNFT_UINT fnftb__akns_method_order(fnft__akns_discretization_t d);
FNFT_REAL fnftb__akns_boundary_coeff(fnft__akns_discretization_t d);
FNFT_INT fnftb__nse_to_akns(fnft_nse_discretization_t d, fnft__akns_discretization_t *out);
FNFT_INT fnftb__kdv_to_akns(fnft_kdv_discretization_t d, fnft__akns_discretization_t *out);
int fnftb__kdv_to_akns_quiet(fnft_kdv_discretization_t d, fnft__akns_discretization_t *out); /* 1 = known, no message */
/* 1 if the leaf construction of this scheme exists as a CUDA kernel */
int fnftb__akns_on_gpu(fnft__akns_discretization_t d);
FNFT_UINT fnftb__nse_degree(fnft_nse_discretization_t d);
FNFT_UINT fnftb__nse_upsampling(fnft_nse_discretization_t d);
FNFT_UINT fnftb__nse_method_order(fnft_nse_discretization_t d);
FNFT_REAL fnftb__nse_boundary_coeff(fnft_nse_discretization_t d);
FNFT_INT fnftb__nse_phase_factor_rho(FNFT_REAL eps_t, FNFT_REAL T1, FNFT_REAL *out,
                                     fnft_nse_discretization_t d);
FNFT_INT fnftb__nse_phase_factor_a(FNFT_REAL eps_t, FNFT_UINT D, FNFT_REAL const *T, FNFT_REAL *out,
                                   fnft_nse_discretization_t d);
FNFT_INT fnftb__nse_phase_factor_b(FNFT_REAL eps_t, FNFT_UINT D, FNFT_REAL const *T, FNFT_REAL *out,
                                   fnft_nse_discretization_t d);

FNFT_UINT fnftb__nextpow2(FNFT_UINT v);

#endif
