/* fnft_b200 host library -- discrete spectrum of fnft_nsev (internal). */
#ifndef FNFT_B200_NSEV_DISCRETE_H
#define FNFT_B200_NSEV_DISCRETE_H
#include "fnft_internal.h"

/*
 * Bound states by Newton refinement + norming constants / residues for the nb
 * signals currently staged in ctx (D_eff effective samples each).  K[b] in: number
 * of guesses in bound_states[b*Kmax..], out: number kept.  ret_codes may be NULL.
 * skip_newton: the values are final (FAST_EIGENVALUE): no refinement, no second filter.
 */
FNFT_INT fnftb__nsev_discrete_chunk(fnftb_ctx *ctx, FNFT_UINT nb, FNFT_UINT D_eff, FNFT_UINT D_given,
                                    FNFT_REAL const *T, FNFT_REAL eps_t, FNFT_UINT *K,
                                    FNFT_UINT Kmax, FNFT_COMPLEX *bound_states,
                                    FNFT_COMPLEX *normconsts_or_residues,
                                    fnft_nsev_opts_t const *opts, FNFT_INT *ret_codes, int skip_newton);

/*
 * fnft_nsev_bsloc_FAST_EIGENVALUE for the nb signals staged in ctx (src/fnft_nsev.c:687-724):
 * transfer matrix (unless have_tm) -> all roots of a(z) on the device -> z_to_lambda ->
 * filter -> merge.  T is the time window of the staged samples (the subsampled window for
 * the first stage of SUBSAMPLE_AND_REFINE).  K[b] out: number of bound states stored in
 * bound_states[b*Kmax..].
 */
FNFT_INT fnftb__nsev_fasteig_chunk(fnftb_ctx *ctx, FNFT_UINT nb, FNFT_UINT D_eff, FNFT_UINT D_given,
                                   FNFT_REAL const *T, FNFT_INT kappa, int have_tm, FNFT_UINT *K,
                                   FNFT_UINT Kmax, FNFT_COMPLEX *bound_states,
                                   fnft_nsev_opts_t const *opts);

/* order-preserving box filter and merge (src/private/fnft__misc.c:114-157,228-259) */
void fnftb__filter_box(FNFT_UINT *N, FNFT_COMPLEX *vals, FNFT_REAL const *box);
void fnftb__merge(FNFT_UINT *N, FNFT_COMPLEX *vals, FNFT_REAL tol);

#endif
