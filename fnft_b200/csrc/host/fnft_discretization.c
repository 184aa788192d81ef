/*
 * fnft_b200 host library -- discretization tables and boundary phase factors.
 *
 * Restates, as lookup tables keyed by the enum values, what the reference spreads
 * over switch statements in
 *   src/private/fnft__akns_discretization.c:29-198   (degree, boundary coefficient,
 *                                                     upsampling factor, order)
 *   src/private/fnft__nse_discretization.c:109-202   (nse -> akns enum)
 *   src/private/fnft__kdv_discretization.c:98-193    (kdv -> akns enum)
 *   src/private/fnft__nse_discretization.c:240-379   (phase factors rho / a / b)
 */
#include "fnft_internal.h"

#define A(x) fnft__akns_discretization_##x

/* one row per value of fnft__akns_discretization_t, in enum order */
static const struct {
    unsigned char degree, upsampling, order, on_gpu;
} akns_tab[] = {
    /* 2SPLIT2_MODAL */ {1, 1, 2, 1},
    /* 2SPLIT1A */ {1, 1, 2, 1},
    /* 2SPLIT1B */ {1, 1, 2, 1},
    /* 2SPLIT2A */ {1, 1, 2, 1},
    /* 2SPLIT2B */ {1, 1, 2, 1},
    /* 2SPLIT2S */ {1, 1, 2, 1},
    /* 2SPLIT3A */ {3, 1, 2, 1},
    /* 2SPLIT3B */ {3, 1, 2, 1},
    /* 2SPLIT3S */ {2, 1, 2, 1},
    /* 2SPLIT4A */ {4, 1, 2, 1},
    /* 2SPLIT4B */ {2, 1, 2, 1},
    /* 2SPLIT5A */ {15, 1, 2, 1},
    /* 2SPLIT5B */ {15, 1, 2, 1},
    /* 2SPLIT6A */ {12, 1, 2, 1},
    /* 2SPLIT6B */ {6, 1, 2, 1},
    /* 2SPLIT7A */ {105, 1, 2, 1},
    /* 2SPLIT7B */ {105, 1, 2, 1},
    /* 2SPLIT8A */ {24, 1, 2, 1},
    /* 2SPLIT8B */ {12, 1, 2, 1},
    /* BO */ {0, 1, 2, 0},
    /* 4SPLIT4A */ {4, 2, 4, 1},
    /* 4SPLIT4B */ {2, 2, 4, 1},
    /* CF4_2 */ {0, 2, 4, 0},
    /* CF4_3 */ {0, 3, 4, 0},
    /* CF5_3 */ {0, 3, 5, 0},
    /* CF6_4 */ {0, 4, 6, 0},
    /* ES4 */ {0, 3, 4, 0},
    /* TES4 */ {0, 3, 4, 0},
};
#define AKNS_COUNT ((int)(sizeof(akns_tab) / sizeof(akns_tab[0])))

static int akns_valid(fnft__akns_discretization_t d) { return (int)d >= 0 && (int)d < AKNS_COUNT; }

FNFT_UINT fnftb__akns_degree(fnft__akns_discretization_t d)
{
    return akns_valid(d) ? akns_tab[d].degree : 0;
}
FNFT_UINT fnftb__akns_upsampling(fnft__akns_discretization_t d)
{
    return akns_valid(d) ? akns_tab[d].upsampling : 0;
}
FNFT_UINT fnftb__akns_method_order(fnft__akns_discretization_t d)
{
    return akns_valid(d) ? akns_tab[d].order : 0;
}
FNFT_REAL fnftb__akns_boundary_coeff(fnft__akns_discretization_t d)
{
    return akns_valid(d) ? 0.5 : NAN;
}
int fnftb__akns_on_gpu(fnft__akns_discretization_t d) { return akns_valid(d) && akns_tab[d].on_gpu; }

/* nse enum order: MODAL, BO, 1A, 1B, 2A, 2B, 2S, 3A, 3B, 3S, 4A, 4B, 5A, 5B, 6A, 6B,
 * 7A, 7B, 8A, 8B, 4SPLIT4A, 4SPLIT4B, CF4_2, CF4_3, CF5_3, CF6_4, ES4, TES4 */
static const fnft__akns_discretization_t nse_map[] = {
    A(2SPLIT2_MODAL), A(BO),       A(2SPLIT1A), A(2SPLIT1B), A(2SPLIT2A), A(2SPLIT2B), A(2SPLIT2S),
    A(2SPLIT3A),      A(2SPLIT3B), A(2SPLIT3S), A(2SPLIT4A), A(2SPLIT4B), A(2SPLIT5A), A(2SPLIT5B),
    A(2SPLIT6A),      A(2SPLIT6B), A(2SPLIT7A), A(2SPLIT7B), A(2SPLIT8A), A(2SPLIT8B), A(4SPLIT4A),
    A(4SPLIT4B),      A(CF4_2),    A(CF4_3),    A(CF5_3),    A(CF6_4),    A(ES4),      A(TES4)};
#define NSE_COUNT ((int)(sizeof(nse_map) / sizeof(nse_map[0])))

/* kdv enum order: 1A, 1B, 2A, 2B, 2S, 3A, 3B, 3S, 4A, 4B, 5A, 5B, 6A, 6B, 7A, 7B, 8A, 8B,
 * 4SPLIT4A, 4SPLIT4B, BO, CF4_2, CF4_3, CF5_3, CF6_4 */
static const fnft__akns_discretization_t kdv_map[] = {
    A(2SPLIT1A), A(2SPLIT1B), A(2SPLIT2A), A(2SPLIT2B), A(2SPLIT2S), A(2SPLIT3A), A(2SPLIT3B),
    A(2SPLIT3S), A(2SPLIT4A), A(2SPLIT4B), A(2SPLIT5A), A(2SPLIT5B), A(2SPLIT6A), A(2SPLIT6B),
    A(2SPLIT7A), A(2SPLIT7B), A(2SPLIT8A), A(2SPLIT8B), A(4SPLIT4A), A(4SPLIT4B), A(BO),
    A(CF4_2),    A(CF4_3),    A(CF5_3),    A(CF6_4)};
#define KDV_COUNT ((int)(sizeof(kdv_map) / sizeof(kdv_map[0])))

FNFT_INT fnftb__nse_to_akns(fnft_nse_discretization_t d, fnft__akns_discretization_t *out)
{
    if ((int)d < 0 || (int)d >= NSE_COUNT)
        return E_INVALID_ARGUMENT(discretization);
    *out = nse_map[d];
    return FNFT_SUCCESS;
}

FNFT_INT fnftb__kdv_to_akns(fnft_kdv_discretization_t d, fnft__akns_discretization_t *out)
{
    if ((int)d < 0 || (int)d >= KDV_COUNT)
        return E_INVALID_ARGUMENT(discretization);
    *out = kdv_map[d];
    return FNFT_SUCCESS;
}

/* the same without an error message (the _numel helpers report an unknown discretization as 0 elements and must not
 * touch the process-wide printf hook to stay quiet: another thread's messages would go with it) */
int fnftb__kdv_to_akns_quiet(fnft_kdv_discretization_t d, fnft__akns_discretization_t *out)
{
    if ((int)d < 0 || (int)d >= KDV_COUNT)
        return 0;
    *out = kdv_map[d];
    return 1;
}

FNFT_UINT fnftb__nse_degree(fnft_nse_discretization_t d)
{
    return ((int)d >= 0 && (int)d < NSE_COUNT) ? fnftb__akns_degree(nse_map[d]) : 0;
}
FNFT_UINT fnftb__nse_upsampling(fnft_nse_discretization_t d)
{
    return ((int)d >= 0 && (int)d < NSE_COUNT) ? fnftb__akns_upsampling(nse_map[d]) : 0;
}
/* src/private/fnft__nse_discretization.c (method_order via the akns table, :158-198) */
FNFT_UINT fnftb__nse_method_order(fnft_nse_discretization_t d)
{
    return ((int)d >= 0 && (int)d < NSE_COUNT) ? fnftb__akns_method_order(nse_map[d]) : 0;
}
FNFT_REAL fnftb__nse_boundary_coeff(fnft_nse_discretization_t d)
{
    return ((int)d >= 0 && (int)d < NSE_COUNT) ? 0.5 : NAN;
}

static int is_2a_or_modal(fnft_nse_discretization_t d)
{
    return d == fnft_nse_discretization_2SPLIT2A || d == fnft_nse_discretization_2SPLIT2_MODAL;
}

/* src/private/fnft__nse_discretization.c:240-256 */
FNFT_INT fnftb__nse_phase_factor_rho(FNFT_REAL eps_t, FNFT_REAL T1, FNFT_REAL *out,
                                     fnft_nse_discretization_t d)
{
    const FNFT_REAL bc = fnftb__nse_boundary_coeff(d);
    if (isnan(bc))
        return E_INVALID_ARGUMENT(nse_discretization);
    *out = -2.0 * (T1 + eps_t * bc);
    if (is_2a_or_modal(d))
        *out += eps_t / (FNFT_REAL)fnftb__nse_degree(d);
    return FNFT_SUCCESS;
}

/* src/private/fnft__nse_discretization.c:263-313: polynomial schemes carry the extra
 * -eps_t*D term, the slow schemes (degree 0) do not. */
FNFT_INT fnftb__nse_phase_factor_a(FNFT_REAL eps_t, FNFT_UINT D, FNFT_REAL const *T, FNFT_REAL *out,
                                   fnft_nse_discretization_t d)
{
    const FNFT_REAL bc = fnftb__nse_boundary_coeff(d);
    if (isnan(bc))
        return E_INVALID_ARGUMENT(nse_discretization);
    const FNFT_REAL span = (T[1] + eps_t * bc) - (T[0] - eps_t * bc);
    *out = (fnftb__nse_degree(d) != 0) ? -eps_t * D + span : span;
    return FNFT_SUCCESS;
}

/* src/private/fnft__nse_discretization.c:320-379 */
FNFT_INT fnftb__nse_phase_factor_b(FNFT_REAL eps_t, FNFT_UINT D, FNFT_REAL const *T, FNFT_REAL *out,
                                   fnft_nse_discretization_t d)
{
    const FNFT_REAL bc = fnftb__nse_boundary_coeff(d);
    if (isnan(bc))
        return E_INVALID_ARGUMENT(nse_discretization);
    const FNFT_REAL base = -(T[1] + eps_t * bc) - (T[0] - eps_t * bc);
    const FNFT_UINT deg = fnftb__nse_degree(d);
    if (deg == 0)
        *out = base;
    else if (is_2a_or_modal(d))
        *out = -eps_t * D + base + eps_t / (FNFT_REAL)deg;
    else
        *out = -eps_t * D + base;
    return FNFT_SUCCESS;
}

/* src/private/fnft__misc.c:316-324 */
FNFT_UINT fnftb__nextpow2(FNFT_UINT v)
{
    if (v == 0)
        return 0;
    FNFT_UINT r = 1;
    while (r < v)
        r *= 2;
    return r;
}
