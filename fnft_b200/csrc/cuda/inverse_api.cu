// fnft_b200 -- device side of the inverse NFT (C-ABI: fnftb_device.h, "inverse transform"); kernels in
// inverse_kernels.cuh.  Its own translation unit: the context (device_api.cu) lends its stream, the general
// 2x2 pair product and the any-length DFT through ctx_hooks.h.
#include "ctx_hooks.h"
#include "inverse_kernels.cuh"
#include "launch.cuh"

#include <atomic>
#include <cmath>
#include <cstring>
#include <vector>
extern std::atomic<unsigned long long> g_fnftb_launch_count;

#define ICU(call)                                                                              \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fnftb__fail((int)_e, cudaGetErrorString(_e), __FILE__, __LINE__);           \
    } while (0)
#define IRC(call)                                                                              \
    do {                                                                                       \
        int _rc = (call);                                                                      \
        if (_rc != 0)                                                                          \
            return _rc;                                                                        \
    } while (0)

namespace {
struct IBuf {
    void *p = nullptr;
    size_t cap = 0;
};
// grow-only workspace of the inverse transform, owned by the context
struct InvWork {
    IBuf T, q, status;               // staged transfer matrices [B][4][deg+1], result [B][D], status [B]
    IBuf T1s[8], T1i[8], T2i[8];     // per recursion level (degree deg >> level)
    IBuf bs, nc, phi, psi;           // discrete spectrum
    IBuf a, b, c;                    // scratch of the continuous-spectrum path
};
void inv_free(void *p)
{
    InvWork *w = (InvWork *)p;
    IBuf *all[] = {&w->T, &w->q, &w->status, &w->bs, &w->nc, &w->phi, &w->psi, &w->a, &w->b, &w->c};
    for (IBuf *b : all)
        if (b->p)
            cudaFree(b->p);
    for (int i = 0; i < 8; ++i) {
        if (w->T1s[i].p)
            cudaFree(w->T1s[i].p);
        if (w->T1i[i].p)
            cudaFree(w->T1i[i].p);
        if (w->T2i[i].p)
            cudaFree(w->T2i[i].p);
    }
    delete w;
}
InvWork *inv_work(fnftb_ctx *c)
{
    void (**dtor)(void *) = nullptr;
    void **slot = fnftb__inv_slot(c, &dtor);
    if (*slot == nullptr) {
        *slot = new InvWork();
        *dtor = inv_free;
    }
    return (InvWork *)*slot;
}
int iensure(IBuf &b, size_t bytes)
{
    if (bytes <= b.cap)
        return 0;
    if (b.p)
        ICU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    ICU(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return 0;
}

struct FinvCall {
    fnftb_ctx *c;
    InvWork *w;
    cudaStream_t st;
    size_t B;
    int kappa, modal;
    double eps_t;
    int *status;
};

int launch_block(const FinvCall &f, InvPoly T, size_t n, InvPoly Ti, cplx *q, size_t qss)
{
    InvBlockArgs a;
    a.T = T;
    a.Ti = Ti;
    a.q = q;
    a.q_sstride = qss;
    a.status = f.status;
    a.n = (int)n;
    a.kappa = f.kappa;
    a.modal = f.modal;
    a.eps_t = f.eps_t;
    const size_t smem = 2 * 8 * (n + 2) * sizeof(cplx);
    IRC(fnftb_smem_optin((const void *)k_inv_block<256>, smem));
    k_inv_block<256><<<(unsigned)f.B, 256, smem, f.st>>>(a);
    ++g_fnftb_launch_count;
    ICU(cudaGetLastError());
    return 0;
}

// result (degree 2d, or its coefficients [i0, i0 + count)) of A * Bm, both taken as degree-d polynomials
int pair_product(const FinvCall &f, InvPoly A, int zlead, InvPoly Bm, size_t d, InvPoly dst, size_t i0, size_t count)
{
    cplx *lev0 = nullptr;
    IRC(fnftb__pair2x2_prepare(f.c, f.B, d, &lev0));
    InvGatherArgs g;
    g.lev0 = lev0;
    g.A = A;
    g.B = Bm;
    g.d = (int)d;
    g.zlead = zlead;
    g.total = (long long)f.B * 8 * (long long)(d + 1);
    k_inv_gather<<<(unsigned)((g.total + 255) / 256), 256, 0, f.st>>>(g);
    ++g_fnftb_launch_count;
    ICU(cudaGetLastError());
    const cplx *res = nullptr;
    IRC(fnftb__pair2x2_run(f.c, f.B, d, &res));
    InvScatterArgs s;
    s.res = res;
    s.dst = dst;
    s.rlen = (int)(2 * d + 1);
    s.i0 = (int)i0;
    s.count = (int)count;
    s.total = (long long)f.B * 4 * (long long)count;
    k_inv_scatter<<<(unsigned)((s.total + 255) / 256), 256, 0, f.st>>>(s);
    ++g_fnftb_launch_count;
    ICU(cudaGetLastError());
    return 0;
}

// src/private/fnft__nse_finvscatter.c:70-203 with blocks of FNFTB_INV_BLOCK samples as the base case
int finv_node(const FinvCall &f, InvPoly T, size_t d, InvPoly Ti, cplx *q, size_t qss, int level)
{
    if (d <= FNFTB_INV_BLOCK)
        return launch_block(f, T, d, Ti, q, qss);
    if (level >= 8)
        return fnftb__fail(-6, "inverse scattering: degree too large", __FILE__, __LINE__);
    const size_t h = d / 2;
    const size_t bytes = f.B * 4 * (h + 1) * sizeof(cplx);
    IRC(iensure(f.w->T1s[level], bytes));
    IRC(iensure(f.w->T1i[level], bytes));
    IRC(iensure(f.w->T2i[level], bytes));
    const InvPoly T1s = {(cplx *)f.w->T1s[level].p, 4 * (h + 1), h + 1};
    const InvPoly T1i = {(cplx *)f.w->T1i[level].p, 4 * (h + 1), h + 1};
    const InvPoly T2i = {(cplx *)f.w->T2i[level].p, 4 * (h + 1), h + 1};
    // Step 1 (:118-130): T2i(z) and q[h .. d-1] from the low-order half of T(z)
    const InvPoly Tlow = {T.p + h, T.sstride, T.estride};
    IRC(finv_node(f, Tlow, h, T2i, q + h, qss, level + 1));
    // Step 2 (:134-142): T1(z) = T2i(z) T(z); the recursion only reads its coefficients d .. d + h
    IRC(pair_product(f, T2i, (int)h, T, d, T1s, d, h + 1));
    // Step 3 (:144-156): T1i(z) and q[0 .. h-1]
    IRC(finv_node(f, T1s, h, T1i, q, qss, level + 1));
    // Step 4 (:160-173): Ti(z) = T1i(z) T2i(z)
    if (Ti.p)
        IRC(pair_product(f, T1i, 0, T2i, h, Ti, 0, d + 1));
    return 0;
}
}  // namespace

extern "C" {

// Fast inverse scattering of B transfer matrices of degree deg (a power of two >= 2; degree-1 discretizations:
// modal = 1 for 2SPLIT2_MODAL, 0 for 2SPLIT2A).  tm: [B][4][deg+1], q: [B][deg]; host pointers unless on_device.
// status_host[b] (may be NULL): 1 where a reconstructed sample violates |Q| < 1 (defocusing case).
int fnftb_finvscatter(fnftb_ctx *c, size_t B, size_t deg, const void *tm, void *q, double eps_t, int kappa,
                      int modal, int on_device, int32_t *status_host)
{
    if (!c || !tm || !q || B == 0 || deg < 2 || (deg & (deg - 1)) != 0)
        return fnftb__fail(-2, "invalid argument", __FILE__, __LINE__);
    if (deg > ((size_t)1 << 15))
        return fnftb__fail(-6, "inverse scattering: more than 32768 samples", __FILE__, __LINE__);
    IRC(fnftb__activate(c));
    InvWork *w = inv_work(c);
    cudaStream_t st = fnftb__stream(c);
    const size_t tbytes = B * 4 * (deg + 1) * sizeof(cplx), qbytes = B * deg * sizeof(cplx);
    const cplx *Tdev = (const cplx *)tm;
    cplx *qdev = (cplx *)q;
    if (!on_device) {
        IRC(iensure(w->T, tbytes));
        IRC(iensure(w->q, qbytes));
        ICU(cudaMemcpyAsync(w->T.p, tm, tbytes, cudaMemcpyHostToDevice, st));
        Tdev = (const cplx *)w->T.p;
        qdev = (cplx *)w->q.p;
    }
    IRC(iensure(w->status, B * sizeof(int)));
    ICU(cudaMemsetAsync(w->status.p, 0, B * sizeof(int), st));
    FinvCall f;
    f.c = c;
    f.w = w;
    f.st = st;
    f.B = B;
    f.kappa = kappa;
    f.modal = modal;
    f.eps_t = eps_t;
    f.status = (int *)w->status.p;
    const InvPoly T = {(cplx *)Tdev, 4 * (deg + 1), deg + 1};
    const InvPoly none = {nullptr, 0, 0};
    IRC(finv_node(f, T, deg, none, qdev, deg, 0));
    if (!on_device)
        ICU(cudaMemcpyAsync(q, qdev, qbytes, cudaMemcpyDeviceToHost, st));
    if (status_host)
        ICU(cudaMemcpyAsync(status_host, w->status.p, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    ICU(cudaStreamSynchronize(st));
    return 0;
}

// Adds K solitons per signal (bound states sorted by descending imaginary part, norming constants) by Darboux
// transforms.  seed = 0: pure multi-soliton, q is written; seed = 1: q holds the seed potential and is updated.
// bs, nc: host [B][K]; q: [B][D], host pointer unless q_on_device.
int fnftb_inv_add_solitons(fnftb_ctx *c, size_t B, size_t K, size_t D, const void *bs, const void *nc, void *q,
                           double T0, double T1, int zc_point, int seed, int q_on_device)
{
    if (!c || !bs || !nc || !q || B == 0 || K == 0 || D < 2)
        return fnftb__fail(-2, "invalid argument", __FILE__, __LINE__);
    if (K > 128)
        return fnftb__fail(-6, "more than 128 bound states per signal", __FILE__, __LINE__);
    IRC(fnftb__activate(c));
    InvWork *w = inv_work(c);
    cudaStream_t st = fnftb__stream(c);
    const size_t kb = B * K * sizeof(cplx), qbytes = B * D * sizeof(cplx);
    IRC(iensure(w->bs, kb));
    IRC(iensure(w->nc, kb));
    ICU(cudaMemcpyAsync(w->bs.p, bs, kb, cudaMemcpyHostToDevice, st));
    ICU(cudaMemcpyAsync(w->nc.p, nc, kb, cudaMemcpyHostToDevice, st));
    cplx *qdev = (cplx *)q;
    if (!q_on_device) {
        IRC(iensure(w->q, qbytes));
        qdev = (cplx *)w->q.p;
        if (seed)
            ICU(cudaMemcpyAsync(qdev, q, qbytes, cudaMemcpyHostToDevice, st));
    }
    const double eps_t = (T1 - T0) / (double)(D - 1);
    const long long tot = (long long)B * (long long)D;
    if (!seed) {
        InvCdtArgs a;
        a.bs = (const cplx *)w->bs.p;
        a.nc = (const cplx *)w->nc.p;
        a.q = qdev;
        a.B = (int)B;
        a.K = (int)K;
        a.D = (int)D;
        a.zc = zc_point;
        a.T0 = T0;
        a.eps_t = eps_t;
        const unsigned grid = (unsigned)((tot + 127) / 128);
        if (K <= 8)
            k_inv_cdt_pure<8><<<grid, 128, 0, st>>>(a);
        else if (K <= 32)
            k_inv_cdt_pure<32><<<grid, 128, 0, st>>>(a);
        else
            k_inv_cdt_pure<128><<<grid, 128, 0, st>>>(a);
        ++g_fnftb_launch_count;
        ICU(cudaGetLastError());
    } else {
        const size_t eb = B * 2 * K * D * sizeof(cplx);
        IRC(iensure(w->phi, eb));
        IRC(iensure(w->psi, eb));
        InvEigArgs e;
        e.bs = (const cplx *)w->bs.p;
        e.q = qdev;
        e.phi = (cplx *)w->phi.p;
        e.psi = (cplx *)w->psi.p;
        e.B = (int)B;
        e.K = (int)K;
        e.D = (int)D;
        e.T0 = T0;
        e.T1 = T1;
        const long long ne = (long long)B * (long long)K * 2;
        k_inv_eigenfunctions<<<(unsigned)((ne + 63) / 64), 64, 0, st>>>(e);
        ++g_fnftb_launch_count;
        ICU(cudaGetLastError());
        InvDarbouxArgs d;
        d.bs = e.bs;
        d.nc = (const cplx *)w->nc.p;
        d.phi = e.phi;
        d.psi = e.psi;
        d.q = qdev;
        d.B = (int)B;
        d.K = (int)K;
        d.D = (int)D;
        const unsigned grid = (unsigned)((tot + 127) / 128);
        if (K <= 8)
            k_inv_darboux<8><<<grid, 128, 0, st>>>(d);
        else if (K <= 32)
            k_inv_darboux<32><<<grid, 128, 0, st>>>(d);
        else
            k_inv_darboux<128><<<grid, 128, 0, st>>>(d);
        ++g_fnftb_launch_count;
        ICU(cudaGetLastError());
    }
    if (!q_on_device)
        ICU(cudaMemcpyAsync(q, qdev, qbytes, cudaMemcpyDeviceToHost, st));
    ICU(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"

namespace {
size_t next_fast_size(size_t n)  // kiss_fft_next_fast_size, src/3rd_party/kiss_fft/kiss_fft.c:396-408
{
    for (;; ++n) {
        size_t m = n;
        while (m % 2 == 0)
            m /= 2;
        while (m % 3 == 0)
            m /= 3;
        while (m % 5 == 0)
            m /= 5;
        if (m <= 1)
            return n;
    }
}
inline unsigned blocks(long long tot) { return (unsigned)((tot + 255) / 256); }

// batched poly_specfact on device arrays: poly / result entry i of signal s at [s * stride + i]
int specfact_dev(fnftb_ctx *c, InvWork *w, size_t B, size_t deg, const cplx *poly, size_t pstride, cplx *result,
                 size_t rstride, size_t oversampling, int kappa, int *warn_dev)
{
    cudaStream_t st = fnftb__stream(c);
    const size_t Ms = next_fast_size((deg + 1) * oversampling);
    const size_t bytes = B * Ms * sizeof(cplx);
    IRC(iensure(w->a, bytes));
    IRC(iensure(w->b, bytes));
    IRC(iensure(w->c, bytes));
    cplx *ba = (cplx *)w->a.p, *bb = (cplx *)w->b.p, *bc = (cplx *)w->c.p;
    SfArgs a;
    memset(&a, 0, sizeof(a));
    a.B = (long long)B;
    a.Ms = (int)Ms;
    a.deg = (int)deg;
    a.kappa = kappa;
    a.warn = warn_dev;
    const long long tot = (long long)B * (long long)Ms;
    // Step 1: P on the oversampled grid, x = log-magnitude
    a.in = poly;
    a.in_sstride = pstride;
    a.out = ba;
    k_sf_load<<<blocks(tot), 256, 0, st>>>(a);
    IRC(fnftb__dft(c, B, Ms, ba, bb, -1));
    a.in = bb;
    a.out = bc;   // x
    a.out2 = ba;  // x reversed
    k_sf_log<<<blocks(tot), 256, 0, st>>>(a);
    // Step 2: Hilbert transform y of x
    IRC(fnftb__dft(c, B, Ms, ba, bb, -1));
    a.in = bb;
    a.out = ba;
    k_sf_hilbert<<<blocks(tot), 256, 0, st>>>(a);
    IRC(fnftb__dft(c, B, Ms, ba, bb, +1));  // y
    // Step 3: exp(x - i y) / M back to coefficients
    a.in = bc;
    a.in2 = bb;
    a.out = ba;
    k_sf_exp<<<blocks(tot), 256, 0, st>>>(a);
    IRC(fnftb__dft(c, B, Ms, ba, bb, +1));
    a.in = bb;
    a.out = result;
    a.out_sstride = rstride;
    k_sf_store<<<blocks((long long)B * (long long)(deg + 1)), 256, 0, st>>>(a);
    g_fnftb_launch_count += 5;
    ICU(cudaGetLastError());
    return 0;
}
}  // namespace

extern "C" {

// fnft__poly_specfact for B polynomials (host arrays [B][deg+1])
int fnftb_specfact(fnftb_ctx *c, size_t B, size_t deg, const void *poly_host, void *result_host, size_t oversampling,
                   int kappa, int32_t *warn_host)
{
    if (!c || !poly_host || !result_host || B == 0 || deg == 0 || oversampling == 0 || kappa < -1 || kappa > 1)
        return fnftb__fail(-2, "invalid argument", __FILE__, __LINE__);
    IRC(fnftb__activate(c));
    InvWork *w = inv_work(c);
    cudaStream_t st = fnftb__stream(c);
    const size_t bytes = B * (deg + 1) * sizeof(cplx);
    IRC(iensure(w->T, 2 * bytes));
    IRC(iensure(w->status, B * sizeof(int)));
    ICU(cudaMemsetAsync(w->status.p, 0, B * sizeof(int), st));
    cplx *pin = (cplx *)w->T.p, *pout = pin + B * (deg + 1);
    ICU(cudaMemcpyAsync(pin, poly_host, bytes, cudaMemcpyHostToDevice, st));
    IRC(specfact_dev(c, w, B, deg, pin, deg + 1, pout, deg + 1, oversampling, kappa, (int *)w->status.p));
    ICU(cudaMemcpyAsync(result_host, pout, bytes, cudaMemcpyDeviceToHost, st));
    if (warn_host)
        ICU(cudaMemcpyAsync(warn_host, w->status.p, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    ICU(cudaStreamSynchronize(st));
    return 0;
}

// Transfer matrices [B][4][deg+1] from samples of the continuous spectrum (src/fnft_nsev_inverse.c:302-678); they
// stay on the device for fnftb_finvscatter_staged.  contspec_host: [B][M], boundary phase factors already removed
// (cstype 0 / 1).  cstype 0: reflection coefficient with A(z) = 1 (TFMATRIX_CONTAINS_REFL_COEFF), 1: b(xi) with
// spectral factorisation, 2: B(tau) (M == D == deg).  warn_host[b] = 1: ill-posed factorisation (:109-110).
int fnftb_inv_tm_from_contspec(fnftb_ctx *c, size_t B, size_t M, size_t D, size_t deg, const void *contspec_host,
                               int cstype, int kappa, double eps_t, size_t oversampling, int32_t *warn_host)
{
    if (!c || !contspec_host || B == 0 || M < 2 || deg == 0 || cstype < 0 || cstype > 2)
        return fnftb__fail(-2, "invalid argument", __FILE__, __LINE__);
    IRC(fnftb__activate(c));
    InvWork *w = inv_work(c);
    cudaStream_t st = fnftb__stream(c);
    const size_t len = deg + 1;
    IRC(iensure(w->T, B * 4 * len * sizeof(cplx)));
    IRC(iensure(w->status, B * sizeof(int)));
    ICU(cudaMemsetAsync(w->status.p, 0, B * sizeof(int), st));
    const size_t cbytes = B * M * sizeof(cplx);
    IRC(iensure(w->phi, cbytes));
    IRC(iensure(w->psi, cbytes));
    cplx *cs = (cplx *)w->phi.p, *tmp = (cplx *)w->psi.p;
    ICU(cudaMemcpyAsync(cs, contspec_host, cbytes, cudaMemcpyHostToDevice, st));
    InvTmArgs a;
    memset(&a, 0, sizeof(a));
    a.B = (long long)B;
    a.M = (int)M;
    a.D = (int)D;
    a.deg = (int)deg;
    a.kappa = kappa;
    cplx *T = (cplx *)w->T.p;
    if (cstype == 2) {
        if (M != D || deg != D)
            return fnftb__fail(-2, "B(tau): M, D and the degree must agree", __FILE__, __LINE__);
        a.cs = cs;
        a.out = T;
        a.scale = 2.0 * eps_t;  // degree1step = 1 for 2SPLIT2A / 2SPLIT2_MODAL
        k_inv_tm_btau_b<<<blocks((long long)B * (long long)D), 256, 0, st>>>(a);
        IRC(specfact_dev(c, w, B, D - 1, T + 2 * len + 1, 4 * len, T + 1, 4 * len, oversampling, kappa,
                         (int *)w->status.p));
        k_inv_tm_btau_rest<<<blocks((long long)B * (long long)D), 256, 0, st>>>(a);
        g_fnftb_launch_count += 2;
    } else {
        a.cs = cs;
        a.out = tmp;
        k_inv_cs_reorder<<<blocks((long long)B * (long long)M), 256, 0, st>>>(a);
        IRC(fnftb__dft(c, B, M, tmp, cs, -1));  // cs now holds b_coeffs
        a.out = T;
        k_inv_tm_from_b<<<blocks((long long)B * (long long)len), 256, 0, st>>>(a, cstype == 0 ? 1 : 0);
        g_fnftb_launch_count += 2;
        if (cstype == 1) {
            IRC(specfact_dev(c, w, B, deg, T + 2 * len, 4 * len, T, 4 * len, oversampling, kappa,
                             (int *)w->status.p));
            k_inv_tm_mirror_a<<<blocks((long long)B * (long long)len), 256, 0, st>>>(a);
            ++g_fnftb_launch_count;
        }
    }
    ICU(cudaGetLastError());
    if (warn_host) {
        ICU(cudaMemcpyAsync(warn_host, w->status.p, B * sizeof(int), cudaMemcpyDeviceToHost, st));
        ICU(cudaStreamSynchronize(st));
    }
    return 0;
}

// One signal, defocusing case: A(z) and B(z) by the iteration of src/fnft_nsev_inverse.c:375-510 (M = D = deg);
// contspec_host: [D] with the boundary phase factors removed.  The transfer matrix stays on the device.
// *hit_max = 1 when max_iter iterations ran without meeting a stopping criterion (:486-487).
int fnftb_inv_tm_ab_from_iter(fnftb_ctx *c, size_t D, const void *contspec_host, int kappa, size_t max_iter,
                              int32_t *hit_max, int32_t *warn_host)
{
    if (!c || !contspec_host || D < 2 || (D & (D - 1)) != 0)
        return fnftb__fail(-2, "invalid argument", __FILE__, __LINE__);
    IRC(fnftb__activate(c));
    InvWork *w = inv_work(c);
    cudaStream_t st = fnftb__stream(c);
    const size_t len = D + 1, vb = D * sizeof(cplx);
    IRC(iensure(w->T, 4 * len * sizeof(cplx)));
    IRC(iensure(w->status, 4 * sizeof(double)));
    IRC(iensure(w->phi, 4 * vb));
    IRC(iensure(w->psi, 2 * vb));
    cplx *cs = (cplx *)w->phi.p, *cr = cs + D, *t0 = cr + D, *t1 = t0 + D;
    cplx *bco = (cplx *)w->psi.p, *aco = bco + D;
    int *warn = (int *)w->status.p;
    double *sum = (double *)w->status.p + 1;
    ICU(cudaMemsetAsync(w->status.p, 0, 4 * sizeof(double), st));
    ICU(cudaMemcpyAsync(cs, contspec_host, vb, cudaMemcpyHostToDevice, st));
    InvIterArgs a;
    memset(&a, 0, sizeof(a));
    a.D = (int)D;
    a.kappa = kappa;
    a.sum = sum;
    const unsigned grid = blocks((long long)D);
    a.in = cs;
    a.out = cr;
    k_it_reorder<<<grid, 256, 0, st>>>(a);
    double prev_change = INFINITY, prev_diff = INFINITY;
    size_t iter = 0;
    for (; iter < max_iter; iter++) {
        a.in = cr;
        a.out = t0;
        k_it_prep<<<grid, 256, 0, st>>>(a);
        IRC(fnftb__dft(c, 1, D, t0, t1, -1));
        a.in = t1;
        a.out = bco;
        k_it_flip<<<grid, 256, 0, st>>>(a);
        IRC(specfact_dev(c, w, 1, D - 1, bco, D, aco, D, 32, kappa, warn));
        // IFFT of a_coeffs reversed: the descending-order input of fnftb__dft is a_coeffs itself
        IRC(fnftb__dft(c, 1, D, aco, t1, +1));
        ICU(cudaMemsetAsync(sum, 0, sizeof(double), st));
        a.in = t1;
        a.in2 = cs;
        a.out = cr;
        k_it_phase<<<grid, 256, 0, st>>>(a);
        g_fnftb_launch_count += 3;
        double cur = 0.0;
        ICU(cudaMemcpyAsync(&cur, sum, sizeof(double), cudaMemcpyDeviceToHost, st));
        ICU(cudaStreamSynchronize(st));
        const double diff = fabs(cur - prev_change);
        if (diff < 10 * 2.220446049250313e-16)
            break;
        prev_change = cur;
        if (diff > 0.9 * prev_diff)
            break;
        prev_diff = diff;
    }
    if (hit_max)
        *hit_max = (iter == max_iter) ? 1 : 0;
    a.in = aco;
    a.in2 = bco;
    a.out = (cplx *)w->T.p;
    k_it_build<<<grid, 256, 0, st>>>(a);
    ICU(cudaGetLastError());
    if (warn_host)
        ICU(cudaMemcpyAsync(warn_host, warn, sizeof(int), cudaMemcpyDeviceToHost, st));
    ICU(cudaStreamSynchronize(st));
    return 0;
}

// fnftb_finvscatter on the transfer matrices left on the device by fnftb_inv_tm_from_contspec
int fnftb_finvscatter_staged(fnftb_ctx *c, size_t B, size_t deg, void *q_host, double eps_t, int kappa, int modal,
                             int32_t *status_host)
{
    if (!c || !q_host)
        return fnftb__fail(-2, "invalid argument", __FILE__, __LINE__);
    InvWork *w = inv_work(c);
    if (w->T.cap < B * 4 * (deg + 1) * sizeof(cplx))
        return fnftb__fail(-2, "no transfer matrices staged", __FILE__, __LINE__);
    IRC(fnftb__activate(c));
    IRC(iensure(w->q, B * deg * sizeof(cplx)));
    const int rc = fnftb_finvscatter(c, B, deg, w->T.p, w->q.p, eps_t, kappa, modal, 1, status_host);
    if (rc)
        return rc;
    cudaStream_t st = fnftb__stream(c);
    ICU(cudaMemcpyAsync(q_host, w->q.p, B * deg * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    ICU(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
