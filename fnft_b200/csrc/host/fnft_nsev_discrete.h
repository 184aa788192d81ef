/* fnft_b200 host library -- discrete spectrum of fnft_nsev (internal). */
#ifndef FNFT_B200_NSEV_DISCRETE_H
#define FNFT_B200_NSEV_DISCRETE_H
#include "fnft_internal.h"

/*
 * Bound states by Newton refinement + norming constants / residues for the nb
 * signals currently staged in ctx (D_eff effective samples each).  K[b] in: number
 * of guesses in bound_states[b*Kmax..], out: number kept.  ret_codes may be NULL.
 */
FNFT_INT fnftb__nsev_discrete_chunk(fnftb_ctx *ctx, FNFT_UINT nb, FNFT_UINT D_eff, FNFT_UINT D_given,
                                    FNFT_REAL const *T, FNFT_REAL eps_t, FNFT_UINT *K,
                                    FNFT_UINT Kmax, FNFT_COMPLEX *bound_states,
                                    FNFT_COMPLEX *normconsts_or_residues,
                                    fnft_nsev_opts_t const *opts, FNFT_INT *ret_codes);

/* order-preserving box filter and merge (src/private/fnft__misc.c:114-157,228-259) */
void fnftb__filter_box(FNFT_UINT *N, FNFT_COMPLEX *vals, FNFT_REAL const *box);
void fnftb__merge(FNFT_UINT *N, FNFT_COMPLEX *vals, FNFT_REAL tol);

#endif
