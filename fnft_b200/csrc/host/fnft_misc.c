/*
 * fnft_b200 host library -- the small helper routines of the reference's private API that its test
 * programs (under test/fnft__poly and test/fnft__akns_fscatter) call next to the numerical kernels:
 *   fnft__misc_*      include/private/fnft__misc.h:42-241, src/private/fnft__misc.c:28-324
 *   fnft__poly_eval*  include/private/fnft__poly_eval.h:57,93, src/private/fnft__poly_eval.c:24-91
 * They are O(n) bookkeeping on caller memory (comparison metrics, filters, a Horner evaluation of a handful
 * of points) and run on the host like in the reference; nothing on the transform path calls them.
 * fnft__misc_resample is not exported: the band-limited shift is part of the GPU preprocessing
 * (resample_kernels.cuh) and has no host implementation here.
 */
#include "fnft_internal.h"
#include "fnft_nsev_discrete.h"
#include <stdio.h>
#include <stdlib.h>

/* "name = [re+imj, ...];" in MATLAB / Python syntax, 13 significant digits */
void fnft__misc_print_buf(const FNFT_INT len, FNFT_COMPLEX const *const buf, char const *const varname)
{
    printf("%s = [", varname);
    for (FNFT_INT k = 0; k < len; k++)
        printf(k + 1 < len ? "%1.12e+%1.12ej, " : "%1.12e+%1.12ej", creal(buf[k]), cimag(buf[k]));
    printf("];\n");
}

/* sum |numer - exact| / sum |exact| */
FNFT_REAL fnft__misc_rel_err(const FNFT_INT len, FNFT_COMPLEX const *const vec_numer,
                             FNFT_COMPLEX const *const vec_exact)
{
    double diff = 0.0, ref = 0.0;
    for (FNFT_INT k = 0; k < len; k++) {
        diff += cabs(vec_numer[k] - vec_exact[k]);
        ref += cabs(vec_exact[k]);
    }
    return diff / ref;
}

/* directed distance: max over a in A of the distance from a to the set B */
static double directed_dist(FNFT_UINT na, FNFT_COMPLEX const *a, FNFT_UINT nb, FNFT_COMPLEX const *b)
{
    double worst = -1.0;
    for (FNFT_UINT i = 0; i < na; i++) {
        double nearest = INFINITY;
        for (FNFT_UINT j = 0; j < nb; j++) {
            const double dist = cabs(a[i] - b[j]);
            if (dist < nearest)
                nearest = dist;
        }
        if (nearest > worst)
            worst = nearest;
    }
    return worst;
}

FNFT_REAL fnft__misc_hausdorff_dist(const FNFT_UINT lenA, FNFT_COMPLEX const *const vecA, const FNFT_UINT lenB,
                                    FNFT_COMPLEX const *const vecB)
{
    const double ab = directed_dist(lenA, vecA, lenB, vecB), ba = directed_dist(lenB, vecB, lenA, vecA);
    return ab > ba ? ab : ba;
}

FNFT_COMPLEX fnft__misc_sech(FNFT_COMPLEX Z) { return 2.0 / (cexp(Z) + cexp(-Z)); }

/* trapezoidal rule for |Z|^2 with the reference's step (b - a)/N */
FNFT_REAL fnft__misc_l2norm2(const FNFT_UINT N, FNFT_COMPLEX const *const Z, const FNFT_REAL a, const FNFT_REAL b)
{
    if (N < 2 || a >= b)
        return NAN;
    const FNFT_REAL h = (b - a) / N;
    FNFT_REAL m = cabs(Z[0]);
    FNFT_REAL acc = 0.5 * h * m * m;
    for (FNFT_UINT k = 1; k + 1 < N; k++) {
        m = cabs(Z[k]);
        acc += h * m * m;
    }
    m = cabs(Z[N - 1]);
    return acc + 0.5 * h * m * m;
}

static int box_is_valid(FNFT_REAL const *box)
{
    return box != NULL && (box[0] <= box[1]) && (box[2] <= box[3]); /* false for NaNs */
}

/* stable compaction of vals (and of the companion array) under a predicate on the value */
static FNFT_INT compact_box(FNFT_UINT *N_ptr, FNFT_COMPLEX *vals, FNFT_COMPLEX *companion, FNFT_REAL const *box,
                            int keep_inside)
{
    FNFT_UINT kept = 0;
    for (FNFT_UINT k = 0; k < *N_ptr; k++) {
        const FNFT_REAL re = creal(vals[k]), im = cimag(vals[k]);
        int keep;
        if (keep_inside) /* closed box; a NaN is never inside */
            keep = (re >= box[0]) && (re <= box[1]) && (im >= box[2]) && (im <= box[3]);
        else /* everything that is not strictly inside the open box, NaNs included */
            keep = !(re > box[0]) || !(re < box[1]) || !(im > box[2]) || !(im < box[3]);
        if (!keep)
            continue;
        vals[kept] = vals[k];
        if (companion != NULL)
            companion[kept] = companion[k];
        kept++;
    }
    *N_ptr = kept;
    return FNFT_SUCCESS;
}

FNFT_INT fnft__misc_filter(FNFT_UINT *const N_ptr, FNFT_COMPLEX *const vals, FNFT_COMPLEX *const rearrange_as_well,
                           FNFT_REAL const *const bounding_box)
{
    if (N_ptr == NULL)
        return E_INVALID_ARGUMENT(N_ptr);
    if (vals == NULL)
        return E_INVALID_ARGUMENT(vals);
    if (!box_is_valid(bounding_box))
        return E_INVALID_ARGUMENT(bounding_box);
    return compact_box(N_ptr, vals, rearrange_as_well, bounding_box, 1);
}

FNFT_INT fnft__misc_filter_inv(FNFT_UINT *const N_ptr, FNFT_COMPLEX *const vals,
                               FNFT_COMPLEX *const rearrange_as_well, FNFT_REAL const *const bounding_box)
{
    if (N_ptr == NULL)
        return E_INVALID_ARGUMENT(N_ptr);
    if (vals == NULL)
        return E_INVALID_ARGUMENT(vals);
    if (!box_is_valid(bounding_box))
        return E_INVALID_ARGUMENT(bounding_box);
    return compact_box(N_ptr, vals, rearrange_as_well, bounding_box, 0);
}

FNFT_INT fnft__misc_filter_nonreal(FNFT_UINT *N_ptr, FNFT_COMPLEX *const vals, const FNFT_REAL tol_im)
{
    if (N_ptr == NULL)
        return E_INVALID_ARGUMENT(N_ptr);
    if (vals == NULL)
        return E_INVALID_ARGUMENT(vals);
    if (!(tol_im >= 0))
        return E_INVALID_ARGUMENT(tol_im);
    FNFT_UINT kept = 0;
    for (FNFT_UINT k = 0; k < *N_ptr; k++)
        if (fabs(cimag(vals[k])) > tol_im)
            vals[kept++] = vals[k];
    *N_ptr = kept;
    return FNFT_SUCCESS;
}

/* Order-dependent like the reference (src/private/fnft__misc.c:228-259): a value is dropped when it is closer
 * than tol to one of its predecessors in the ARRAY AS IT IS BEING COMPACTED (positions below the write cursor
 * already hold survivors, the rest the original values) -- the same routine the bound-state code uses. */
FNFT_INT fnft__misc_merge(FNFT_UINT *N_ptr, FNFT_COMPLEX *const vals, FNFT_REAL tol)
{
    if (N_ptr == NULL)
        return E_INVALID_ARGUMENT(N_ptr);
    if (*N_ptr == 0)
        return FNFT_SUCCESS;
    if (vals == NULL)
        return E_INVALID_ARGUMENT(vals);
    if (tol < 0.0)
        return E_INVALID_ARGUMENT(tol);
    fnftb__merge(N_ptr, vals, tol);
    return FNFT_SUCCESS;
}

FNFT_INT fnft__misc_downsample(const FNFT_UINT D, FNFT_COMPLEX const *const q, FNFT_UINT *const Dsub_ptr,
                               FNFT_COMPLEX **qsub_ptr, FNFT_UINT *const first_last_index)
{
    if (q == NULL)
        return E_INVALID_ARGUMENT(q);
    if (D <= 2)
        return E_INVALID_ARGUMENT(D);
    if (qsub_ptr == NULL)
        return E_INVALID_ARGUMENT(qsub_ptr);
    if (Dsub_ptr == NULL)
        return E_INVALID_ARGUMENT(Dsub_ptr);
    if (first_last_index == NULL)
        return E_INVALID_ARGUMENT(first_last_index);
    /* same rounding as the preprocessing of the transforms (fnft_nsev.c of this library) */
    FNFT_UINT want = *Dsub_ptr;
    want = want < 2 ? 2 : (want > D ? D : want);
    const FNFT_UINT stride = (FNFT_UINT)round((FNFT_REAL)D / want);
    const FNFT_UINT count = (FNFT_UINT)round((FNFT_REAL)D / stride);
    FNFT_COMPLEX *out = malloc(count * sizeof(FNFT_COMPLEX));
    if (out == NULL)
        return E_NOMEM;
    for (FNFT_UINT k = 0; k < count; k++)
        out[k] = q[k * stride];
    first_last_index[0] = 0;
    first_last_index[1] = (count - 1) * stride;
    *qsub_ptr = out;
    *Dsub_ptr = count;
    return FNFT_SUCCESS;
}

/* sin(x)/x; below |x| = 1e-8 the reference switches to cos(x/sqrt(3)) = 1 - x^2/6 + O(x^4) */
FNFT_COMPLEX fnft__misc_CSINC(FNFT_COMPLEX x)
{
    return (cabs(x) >= 1.0e-8) ? csin(x) / x : ccos(x / csqrt(3));
}

FNFT_UINT fnft__misc_nextpowerof2(const FNFT_UINT number) { return number == 0 ? 0 : fnftb__nextpow2(number); }

/* ---- Horner evaluation (src/private/fnft__poly_eval.c:24-91): p holds deg+1 coefficients, highest power
 * first; z is overwritten by p(z).  Outside the unit disc the reversed polynomial is evaluated in 1/z and the
 * power z^deg restored afterwards, which keeps the intermediate values bounded. */
static void horner_pair(FNFT_UINT deg, FNFT_COMPLEX const *p, FNFT_COMPLEX z, FNFT_COMPLEX *val, FNFT_COMPLEX *der)
{
    FNFT_COMPLEX v, d = 0.0;
    if (cabs(z) <= 1.0) {
        v = p[0];
        for (FNFT_UINT k = 1; k <= deg; k++) {
            d = v + d * z;
            v = p[k] + v * z;
        }
        *val = v;
        *der = d;
        return;
    }
    /* r(u) = sum_k p[deg-k] u^k ... evaluated at u = 1/z: p(z) = z^deg r(u), p'(z) = z^(deg-1) (deg r(u) - u r'(u)) */
    v = p[deg];
    for (FNFT_UINT k = 1; k <= deg; k++) {
        d = v + d / z;
        v = p[deg - k] + v / z;
    }
    *val = v * cpow(z, deg);
    *der = (deg == 0) ? 0.0 : cpow(z, (FNFT_REAL)deg - 1) * (deg * v - d / z);
}

FNFT_INT fnft__poly_eval(const FNFT_UINT deg, FNFT_COMPLEX const *const p, const FNFT_UINT nz, FNFT_COMPLEX *const z)
{
    if (p == NULL)
        return E_INVALID_ARGUMENT(p);
    if (z == NULL)
        return E_INVALID_ARGUMENT(z);
    for (FNFT_UINT i = 0; i < nz; i++) {
        FNFT_COMPLEX v, d;
        horner_pair(deg, p, z[i], &v, &d);
        z[i] = v;
    }
    return FNFT_SUCCESS;
}

FNFT_INT fnft__poly_evalderiv(const FNFT_UINT deg, FNFT_COMPLEX const *const p, const FNFT_UINT nz,
                              FNFT_COMPLEX *const z, FNFT_COMPLEX *const deriv)
{
    if (p == NULL)
        return E_INVALID_ARGUMENT(p);
    if (z == NULL)
        return E_INVALID_ARGUMENT(z);
    if (deriv == NULL)
        return E_INVALID_ARGUMENT(deriv);
    for (FNFT_UINT i = 0; i < nz; i++) {
        FNFT_COMPLEX v, d;
        horner_pair(deg, p, z[i], &v, &d);
        z[i] = v;
        deriv[i] = d;
    }
    return FNFT_SUCCESS;
}
