// TEST TOOLING: host emulation of the CUDA block programs (see
// fnft_b200/csrc/cuda/common.cuh).  Built by tests/emul/build.sh with
//   g++ -O2 -DFNFTB_EMUL -shared -fPIC
// and driven from pytest through ctypes.  It lets the index arithmetic of the
// kernels be checked in the GPU-less container; it is never part of the product.
#define FNFTB_EMUL 1
#include "../../fnft_b200/csrc/cuda/tree_driver.cuh"
#include "../../fnft_b200/csrc/cuda/chirpz_driver.cuh"
#include "../../fnft_b200/csrc/cuda/twiddle.h"
#include <vector>

static std::vector<double> g_tw;
static TwTable get_tw()
{
    const int twn = 4096;
    if (g_tw.empty()) {
        g_tw.resize(2 * twn);
        fnftb_fill_twiddles(g_tw.data(), twn);
    }
    TwTable T;
    T.tw = (const cplx *)g_tw.data();
    T.twn = twn;
    T.log2twn = 12;
    return T;
}

struct Work {
    std::vector<cplx> lev0, lev1, gbuf, colbuf;
    std::vector<double> mx0, mx1;
    std::vector<int> W, status;
    TreeWork w;
    Work(size_t B, size_t npad, size_t deg0)
        : lev0(tree_lev_elems(B, npad, deg0)), lev1(tree_lev_elems(B, npad, deg0)),
          gbuf(tree_gbuf_elems(B, npad, deg0)), colbuf(tree_gbuf_elems(B, npad, deg0)), mx0(B * npad), mx1(B * npad), W(B), status(B)
    {
        w.lev[0] = lev0.data();
        w.lev[1] = lev1.data();
        w.mx[0] = mx0.data();
        w.mx[1] = mx1.data();
        w.gbuf = gbuf.data();
        w.colbuf = colbuf.data();
        w.W = W.data();
        w.status = status.data();
        w.tt[0] = w.tt[1] = nullptr;
        w.tws = nullptr;
    }
};

extern "C" {

// in-place smem FFT of nfft arrays of length n (dir = -1 forward, +1 inverse).
// The forward leaves digit-reversed order; perm_out[pos] = frequency index.
int emul_fft(double *data, int nfft, int n, int dir, int nt, int *perm_out)
{
    TwTable T = get_tw();
    FftPlan P = make_fft_plan(n);
    std::vector<cplx> S((size_t)nfft * n);
    for (int f = 0; f < nfft; ++f)
        for (int i = 0; i < n; ++i)
            S[(size_t)f * n + swz(i)] = make_cplx(data[2 * ((size_t)f * n + i)],
                                                  data[2 * ((size_t)f * n + i) + 1]);
    cplx *Sp = S.data();
    if (dir < 0) {
        FNFTB_SMEM_FFT_FWD(Sp, nfft, P, nt, T);
    } else {
        FNFTB_SMEM_FFT_INV(Sp, nfft, P, nt, T);
    }
    for (int f = 0; f < nfft; ++f)
        for (int i = 0; i < n; ++i) {
            data[2 * ((size_t)f * n + i)] = S[(size_t)f * n + swz(i)].x;
            data[2 * ((size_t)f * n + i) + 1] = S[(size_t)f * n + swz(i)].y;
        }
    if (perm_out) {
        for (int pos = 0; pos < n; ++pos) {
            // pos = sum j_p * s_p ; k = j_0 + r_0*(j_1 + r_1*(...))
            int rem = pos, k = 0, mult = 1, stride = n;
            for (int p = 0; p < P.npass; ++p) {
                stride /= P.radix[p];
                const int j = rem / stride;
                rem %= stride;
                k += j * mult;
                mult *= P.radix[p];
            }
            perm_out[pos] = k;
        }
    }
    return P.npass;
}

int emul_fscatter(const double *q, const double *r, int B, int D, int deg0, int rmode, int kappa,
                  int scheme, double eps_t, int normalize, double *tm, int *W, int use_direct,
                  int smem_n)
{
    const size_t npad = next_pow2_sz((size_t)D);
    Work wk((size_t)B, npad, (size_t)tree_leaf_degree(scheme, deg0));
    int rc = tree_fscatter(wk.w, (const cplx *)q, (const cplx *)r, B, D, deg0, rmode, kappa, scheme,
                           eps_t, normalize, (cplx *)tm, get_tw(), NULL, use_direct, smem_n);
    for (int s = 0; s < B; ++s) {
        W[s] = wk.W[s];
        if (wk.status[s])
            rc = 100 + wk.status[s];
    }
    return rc;
}

int emul_fmult2x2(const double *p, int n, int deg0, int normalize, double *tm, int *W,
                  int use_direct, int smem_n)
{
    const size_t npad = next_pow2_sz((size_t)n);
    Work wk(1, npad, (size_t)deg0);
    int rc = tree_fmult2x2(wk.w, (const cplx *)p, n, deg0, normalize, (cplx *)tm, get_tw(), NULL,
                           use_direct, smem_n);
    W[0] = wk.W[0];
    return rc;
}


// chirp-z of npoly polynomials per signal; see CzArgs for the meaning of the fields
int emul_chirpz(const double *tm, long tm_sstride, int ent0, int ent1, int npoly, int deg, int B,
                int M, double lwr, double lwi, double lar, double lai, int mode, int cstype,
                double *out, long out_sstride, const int *W, double xi0, double eps_xi,
                double ph_rho, double ph_a, double ph_b, double kdv_ph, double kdv_sqrtz,
                int *status, int row_n)
{
    CzArgs a;
    memset(&a, 0, sizeof(a));
    a.tm = (const cplx *)tm;
    a.tm_sstride = (size_t)tm_sstride;
    a.ent[0] = ent0;
    a.ent[1] = ent1;
    a.npoly = npoly;
    a.deg = deg;
    a.B = B;
    a.M = M;
    a.lwr = lwr;
    a.lwi = lwi;
    a.lar = lar;
    a.lai = lai;
    a.mode = mode;
    a.cstype = cstype;
    a.out = (cplx *)out;
    a.out_sstride = (size_t)out_sstride;
    a.W = W;
    a.xi0 = xi0;
    a.eps_xi = eps_xi;
    a.ph_rho = ph_rho;
    a.ph_a = ph_a;
    a.ph_b = ph_b;
    a.kdv_ph = kdv_ph;
    a.kdv_sqrtz = kdv_sqrtz;
    a.status = status;
    a.T = get_tw();
    const CzGeom g = cz_geometry(deg, M, row_n);
    std::vector<cplx> ybuf(cz_ybuf_elems(g, (size_t)B, npoly)), vhat((size_t)g.L);
    a.ybuf = ybuf.data();
    a.vhat = vhat.data();
    std::vector<cplx> tables(cz_table_elems(g, deg, M));
    return cz_run(a, tables.data(), NULL, row_n);
}

}  // extern "C"
