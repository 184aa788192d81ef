// fnft_b200 -- Newton refinement of main / auxiliary spectrum points of the periodic NFT.
//
// Replaces refine_mainspec / refine_auxspec   (/root/reference/src/fnft_nsep.c:708-835) together
// with the monodromy matrix and its lambda-derivative they evaluate,
//   fnft__nse_scatter_matrix                  (/root/reference/src/private/fnft__nse_scatter_matrix.c:33-110)
//   fnft__akns_scatter_matrix, BO / CF4_2     (/root/reference/src/private/fnft__akns_scatter_matrix.c:112-205),
// and builds the Floquet polynomial  p(z) = a(z) + a#(z) - rhs*2^-W  (src/fnft_nsep.c:561-583).
//
// One warp per spectrum point, as in bound_warp.cuh: the 4x4 block-triangular step matrices
// [[U, 0], [U', U]] are multiplied up per lane over a contiguous chunk of the D samples and the
// 32 partial products are combined with an ordered shuffle reduction; the scalar Newton logic
// (line search over the root order m = 1, 2 for the main spectrum) is executed redundantly by all
// lanes on broadcast values, so the warp never diverges.
#pragma once
#ifndef FNFTB_EMUL
#include "bound_warp.cuh"

struct NsepRefineArgs {
    const cplx *q;     // [B][D] effective (preprocessed) samples of the full signal
    int B, D;
    int upsampling;    // 1: BO, 2: CF4_2
    int kappa;         // r = -kappa*conj(q)
    int Kstride;
    const int *K;      // [B] points per signal
    cplx *lam;         // [B][Kstride] in/out
    int *flag;         // [B][Kstride]: 3 = division by zero (f' = 0)
    double eps_t;
    double lweight;    // l = lam*lweight per sample (1 BO, 0.5 CF4_2)
    double scl;        // factor of the derivative (1 BO, 0.5 CF4_2)
    double rhs;        // main spectrum: f = a + atilde + rhs
    double tol;
    int max_evals;
    int mode;          // 0: main spectrum (trace), 1: auxiliary spectrum (entry 12)
};

// product (with derivative) of the steps of samples [lo, hi), r = -kappa*conj(q)
DEV BoMat bo_chunk_kappa(const cplx *q, int lo, int hi, cplx l, double h, int kappa)
{
    BoMat P;
    P.m[0] = make_cplx(1.0, 0.0);
    P.m[1] = czero();
    P.m[2] = czero();
    P.m[3] = make_cplx(1.0, 0.0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        P.d[i] = czero();
    const double ks = -(double)kappa;
    for (int n = lo; n < hi; ++n) {
        const cplx qn = __ldg(&q[n]);
        const cplx rn = make_cplx(ks * qn.x, -ks * qn.y);
        cplx U[4], Ud[4], t[4], td[4];
        bo_step<true>(qn, rn, l, h, U, Ud);
        bo_mm(Ud, P.m, td);
        bo_mm_acc(U, P.d, td);
        bo_mm(U, P.m, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            P.d[i] = td[i];
            P.m[i] = t[i];
        }
    }
    return P;
}

// f and f' at lam, identical on all lanes
DEV void nsep_eval(const NsepRefineArgs &a, const cplx *q, int lo, int hi, int lane, cplx lam, cplx *f, cplx *fp)
{
    const cplx l = cscale(lam, a.lweight);
    BoMat P = bo_chunk_kappa(q, lo, hi, l, a.eps_t, a.kappa);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        BoMat Hn;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            Hn.m[e] = shfl_down_c(P.m[e], off);
            Hn.d[e] = shfl_down_c(P.d[e], off);
        }
        if ((lane & (2 * off - 1)) == 0)
            P = bo_compose(Hn, P);
    }
    cplx fv, fpv;
    if (a.mode == 0) {
        fv = make_cplx(P.m[0].x + P.m[3].x + a.rhs, P.m[0].y + P.m[3].y);
        fpv = cscale(cadd(P.d[0], P.d[3]), a.scl);
    } else {
        fv = P.m[1];
        fpv = cscale(P.d[1], a.scl);
    }
    *f = shfl_c(fv, 0);
    *fp = shfl_c(fpv, 0);
}

__global__ void __launch_bounds__(128) k_nsep_refine(const NsepRefineArgs a)
{
    const int lane = threadIdx.x & 31;
    const long long gid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gid >= (long long)a.B * a.Kstride)
        return;
    const int s = (int)(gid / a.Kstride), i = (int)(gid % a.Kstride);
    if (i >= a.K[s] || a.max_evals <= 0)
        return;
    const cplx *q = a.q + (size_t)s * a.D;
    int lo, hi;
    bo_chunk_bounds(a.D, a.upsampling, lane, &lo, &hi);
    cplx lam = a.lam[gid];
    int status = 0;
    if (a.mode == 0) {
        // Newton's method for roots of order m in {1, 2}, src/fnft_nsep.c:724-786
        cplx next_f, next_fp;
        nsep_eval(a, q, lo, hi, lane, lam, &next_f, &next_fp);
        for (int nevals = 1; nevals <= a.max_evals;) {
            const cplx f = next_f, fp = next_fp;
            if (fp.x == 0.0 && fp.y == 0.0) {
                status = 3;
                break;
            }
            const cplx incr = cdiv(f, fp);
            double min_abs = INFINITY;
            int best_m = 1;
            for (int m = 1; m <= 2; ++m) {
                const cplx l2 = make_cplx(lam.x - m * incr.x, lam.y - m * incr.y);
                cplx tf, tfp;
                nsep_eval(a, q, lo, hi, lane, l2, &tf, &tfp);
                ++nevals;
                const double cur = hypot(tf.x, tf.y);
                if (cur < min_abs) {
                    min_abs = cur;
                    best_m = m;
                    next_f = tf;
                    next_fp = tfp;
                    if (cur < a.tol)
                        break;
                }
            }
            lam = make_cplx(lam.x - best_m * incr.x, lam.y - best_m * incr.y);
            if (min_abs < a.tol) {
                if (next_fp.x == 0.0 && next_fp.y == 0.0) {
                    status = 3;
                    break;
                }
                lam = csub(lam, cdiv(next_f, next_fp));
                break;
            }
        }
    } else {
        // plain Newton on b(lam), src/fnft_nsep.c:801-832
        for (int nevals = 0; nevals < a.max_evals;) {
            cplx f, fp;
            nsep_eval(a, q, lo, hi, lane, lam, &f, &fp);
            ++nevals;
            if (fp.x == 0.0 && fp.y == 0.0) {
                status = 3;
                break;
            }
            lam = csub(lam, cdiv(f, fp));
            if (hypot(f.x, f.y) < a.tol)
                break;
        }
    }
    if (lane == 0) {
        a.lam[gid] = lam;
        a.flag[gid] = status;
    }
}

// p_i = t11_i + conj(t11_{deg-i}),  p_{deg/2} -= rhs*2^-W   (src/fnft_nsep.c:561-583)
struct FloquetRhsArgs {
    const cplx *tm;  // [B][4][deg+1]
    const int *W;    // [B]
    cplx *P;         // [B][deg+1]
    int B, deg;
    double rhs;
};
__global__ void k_nsep_floquet_rhs(const FloquetRhsArgs a)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int d1 = a.deg + 1;
    if (gid >= (long long)a.B * d1)
        return;
    const int s = (int)(gid / d1), i = (int)(gid % d1);
    const cplx *t11 = a.tm + (size_t)s * 4 * d1;
    cplx v = cadd(t11[i], cconj(t11[a.deg - i]));
    if (i == a.deg / 2)
        v.x -= a.rhs * ldexp(1.0, -a.W[s]);
    a.P[gid] = v;
}
#endif
