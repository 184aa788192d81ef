// fnft_b200 -- signal preprocessing for the fourth-order splitting schemes
// (4SPLIT4A/B, CF4_2): two band-limited shifts of the samples by -/+ sqrt(3)/6 * eps_t
// and the 2x2 Gauss-node weighting.
//
// Replaces fnft__misc_resample           (/root/reference/src/private/fnft__misc.c:326-407)
// and the 4SPLIT4 branch of
//   fnft__nse_discretization_preprocess_signal
//                                          (/root/reference/src/private/fnft__nse_discretization.c:474-503)
// with weights from fnft__akns_discretization_method_weights
//                                          (/root/reference/src/private/fnft__akns_discretization.c:284-298).
// One CTA per signal; the length-D transforms live in shared memory (D a power of
// two, D <= 4096 in this version).
#pragma once
#include "fft_core.cuh"

struct ResampleArgs {
    const cplx *q;  // [B][D]
    cplx *out;      // [B][2*D]
    int *warn;      // [B] 1 if the spectrum does not look band-limited
    int B, D;
    int nskip, Dsub;  // keep the samples 0, nskip, ..., (Dsub-1)*nskip; shifts scaled by nskip
    double eps_t;
    FftPlan plan;
    TwTable T;
};

HD size_t resample_smem_bytes(int D, int nt) { return sizeof(cplx) * 2 * (size_t)D + sizeof(double) * 3 * nt; }

BLK void blk_resample_4split4(const ResampleArgs &a, blk3 bid, int nt, void *smem)
{
    const int D = a.D;
    cplx *S1 = (cplx *)smem;
    cplx *S2 = S1 + D;
    double *red = (double *)(S2 + D);
    const int s = bid.x;
    const cplx *q = a.q + (size_t)s * D;
    FOR_THREADS(tid, nt)
    {
        for (int i = tid; i < D; i += nt)
            S1[swz(i)] = q[i];
    }
    BLOCK_SYNC();
    FNFTB_SMEM_FFT_FWD(S1, 1, a.plan, nt, a.T);
    BLOCK_SYNC();
    // band-limitation check (fnft__misc.c:368-380): trapezoidal |X|^2 sums over the
    // two 5% bands next to the Nyquist bin versus the whole spectrum
    const int Dlp = D / 20;
    FOR_THREADS(tid, nt)
    {
        double lo = 0.0, hi = 0.0, all = 0.0;
        for (int pos = tid; pos < D; pos += nt) {
            const int k = plan_freq_of_pos(a.plan, pos);
            const cplx x = S1[swz(pos)];
            const double m2 = cabs2(x);
            all += ((k == 0 || k == D - 1) ? 0.5 : 1.0) * m2;
            if (Dlp >= 2) {
                const int j1 = k - (D / 2 - 1 - Dlp);
                if (j1 >= 0 && j1 < Dlp)
                    lo += ((j1 == 0 || j1 == Dlp - 1) ? 0.5 : 1.0) * m2;
                const int j2 = k - (D / 2 + 1);
                if (j2 >= 0 && j2 < Dlp)
                    hi += ((j2 == 0 || j2 == Dlp - 1) ? 0.5 : 1.0) * m2;
            }
        }
        red[tid] = lo;
        red[nt + tid] = hi;
        red[2 * nt + tid] = all;
    }
    BLOCK_SYNC();
    FOR_THREADS(tid, nt)
    {
        if (tid == 0) {
            double lo = 0.0, hi = 0.0, all = 0.0;
            for (int t = 0; t < nt; ++t) {
                lo += red[t];
                hi += red[nt + t];
                all += red[2 * nt + t];
            }
            // every sum carries the same step h = eps_t, which cancels in the ratio
            const double ratio = sqrt(lo + hi) / sqrt(all);
            a.warn[s] = (Dlp >= 2 && ratio > 1.4901161193847656e-08 /* sqrt(eps) */) ? 1 : 0;
        }
    }
    // phase shifts: delta = -/+ eps_t*sqrt(3)/6, freq = k/(D*eps_t) (k signed)
    FOR_THREADS(tid, nt)
    {
        const double scl = (double)D * a.eps_t;
        const double delta = a.eps_t * (1.7320508075688772 / 6.0) * (double)a.nskip;
        for (int pos = tid; pos < D; pos += nt) {
            const int k = plan_freq_of_pos(a.plan, pos);
            const double freq = (k < D / 2) ? (double)k / scl : ((double)k - (double)D) / scl;
            double sn, cs;
            SINCOS(2.0 * 3.141592653589793 * delta * freq, &sn, &cs);
            const cplx x = S1[swz(pos)];
            S1[swz(pos)] = cmul(x, make_cplx(cs, -sn));  // shift by -delta
            S2[swz(pos)] = cmul(x, make_cplx(cs, sn));   // shift by +delta
        }
    }
    BLOCK_SYNC();
    FNFTB_SMEM_FFT_INV(S1, 2, a.plan, nt, a.T);
    BLOCK_SYNC();
    FOR_THREADS(tid, nt)
    {
        const double sf = 1.7320508075688772 / 6.0;
        const double w0 = 0.25 + sf, w1 = 0.25 - sf;
        const double invD = 1.0 / (double)D;
        cplx *o = a.out + (size_t)s * 2 * a.Dsub;
        for (int isub = tid; isub < a.Dsub; isub += nt) {
            const int i = isub * a.nskip;
            const cplx q1 = cscale(S1[swz(i)], invD), q2 = cscale(S2[swz(i)], invD);
            o[2 * isub] = make_cplx(w0 * q1.x + w1 * q2.x, w0 * q1.y + w1 * q2.y);
            o[2 * isub + 1] = make_cplx(w1 * q1.x + w0 * q2.x, w1 * q1.y + w0 * q2.y);
        }
    }
}

// plain subsampling (fnft__nse_discretization.c:463-470)
struct SubsampleArgs {
    const cplx *q;  // [B][D]
    cplx *out;      // [B][Dsub]
    int B, D, nskip, Dsub;
};

BLK void blk_subsample(const SubsampleArgs &a, blk3 bid, int nt, void * /*smem*/)
{
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        if (gid < (long long)a.B * a.Dsub) {
            const int s = (int)(gid / a.Dsub), isub = (int)(gid % a.Dsub);
            a.out[gid] = a.q[(size_t)s * a.D + (size_t)isub * a.nskip];
        }
    }
}

// ES4 / TES4 preprocessing (fnft__nse_discretization.c:609-631): every (sub-sampled) grid point becomes the three
// effective samples (q, q', q'') with central differences over the sub-sampled grid, zero outside the window
struct Es4Args {
    const cplx *q;  // [B][D]
    cplx *out;      // [B][3 * Dsub]
    int B, D, nskip, Dsub;
    double eps_sub;  // eps_t * nskip
};

BLK void blk_es4_preprocess(const Es4Args &a, blk3 bid, int nt, void * /*smem*/)
{
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        if (gid < (long long)a.B * a.Dsub) {
            const int s = (int)(gid / a.Dsub), i = (int)(gid % a.Dsub);
            const cplx *qs = a.q + (size_t)s * a.D;
            const cplx c = qs[(size_t)i * a.nskip];
            const cplx m = (i > 0) ? qs[(size_t)(i - 1) * a.nskip] : czero();
            const cplx p = (i + 1 < a.Dsub) ? qs[(size_t)(i + 1) * a.nskip] : czero();
            const double d1 = 2.0 * a.eps_sub, d2 = a.eps_sub * a.eps_sub;
            cplx *o = a.out + (size_t)gid * 3;
            o[0] = c;
            o[1] = make_cplx((p.x - m.x) / d1, (p.y - m.y) / d1);
            o[2] = make_cplx(((p.x - 2.0 * c.x) + m.x) / d2, ((p.y - 2.0 * c.y) + m.y) / d2);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// General lengths (any D >= 4, not limited by shared memory): the length-D transforms are
// evaluated as chirp-z transforms with the batched Bluestein machinery of chirpz_driver.cuh
// (X_k = sum_n q_n w^(nk) is the polynomial sum_n q_n z^n at z = w^k), so the arithmetic is
// again power-of-two FFTs while the result is the length-D DFT the reference computes with
// Kiss FFT's mixed radices (fnft__misc.c:358-397).  Kernels around the two chirp-z calls:
//   blk_rs_reverse : coefficients highest power first, as the chirp-z kernels read them
//   blk_rs_shift   : band-limitation check (:368-380) and the two phase ramps (:382-392)
//   blk_rs_weights : 1/D, Gauss weights, subsampling (fnft__nse_discretization.c:483-500)
// ---------------------------------------------------------------------------------------------
struct RsArgs {
    const cplx *in;
    cplx *out;
    int *warn;
    int B, D, nskip, Dsub;
    double eps_t;
    // CF4_3 (up == 3, fnft__nse_discretization.c:505-531): shifts by -/+ sqrt(3/20)*eps_t, the
    // unshifted samples q0 as middle node and the 3x3 weight matrix w3 (row-major, real)
    int up;
    const cplx *q0;  // [B][D]
    double w3[9];
    // CF5_3 (wsel 2, up 3) and CF6_4 (wsel 3, up 4), :532-604: shifts by -/+ sqrt(15)/10*eps_t, complex
    // weights bo_cf_w applied to q and, unconjugated, to r = -kappa*conj(q): r is no longer
    // -kappa*conj(q_preprocessed) and is written to rout [B][up*Dsub]
    int wsel, kappa;
    cplx *rout;
};

BLK void blk_rs_reverse(const RsArgs &a, blk3 bid, int nt, void * /*smem*/)
{
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        if (gid < (long long)a.B * a.D) {
            const int s = (int)(gid / a.D), i = (int)(gid % a.D);
            a.out[gid] = a.in[(size_t)s * a.D + (size_t)(a.D - 1 - i)];
        }
    }
}

// in: X [B][D] (natural order); out: [B][2][D], entry j = spectrum shifted by -/+ delta, stored
// reversed (highest "power" first) for the inverse chirp-z.  One CTA per signal.
BLK void blk_rs_shift(const RsArgs &a, blk3 bid, int nt, void *smem)
{
    const int D = a.D, s = bid.x;
    double *red = (double *)smem;
    const cplx *X = a.in + (size_t)s * D;
    cplx *Y = a.out + (size_t)s * 2 * D;
    const int Dlp = D / 20;
    FOR_THREADS(tid, nt)
    {
        double lo = 0.0, hi = 0.0, all = 0.0;
        const double scl = (double)D * a.eps_t;
        const double delta = a.eps_t * (double)a.nskip *
                             (a.wsel >= 2 ? sqrt(15.0) / 10.0 : (a.up == 3 ? sqrt(3.0 / 20.0) : 1.7320508075688772 / 6.0));
        for (int k = tid; k < D; k += nt) {
            const cplx x = X[k];
            const double m2 = cabs2(x);
            all += ((k == 0 || k == D - 1) ? 0.5 : 1.0) * m2;
            if (Dlp >= 2) {
                const int j1 = k - (D / 2 - 1 - Dlp);
                if (j1 >= 0 && j1 < Dlp)
                    lo += ((j1 == 0 || j1 == Dlp - 1) ? 0.5 : 1.0) * m2;
                const int j2 = k - (D / 2 + 1);
                if (j2 >= 0 && j2 < Dlp)
                    hi += ((j2 == 0 || j2 == Dlp - 1) ? 0.5 : 1.0) * m2;
            }
            const double freq = (k < D / 2) ? (double)k / scl : ((double)k - (double)D) / scl;
            double sn, cs;
            SINCOS(2.0 * 3.141592653589793 * delta * freq, &sn, &cs);
            Y[D - 1 - k] = cmul(x, make_cplx(cs, -sn));      // shift by -delta
            Y[D + D - 1 - k] = cmul(x, make_cplx(cs, sn));   // shift by +delta
        }
        red[tid] = lo;
        red[nt + tid] = hi;
        red[2 * nt + tid] = all;
    }
    BLOCK_SYNC();
    FOR_THREADS(tid, nt)
    {
        if (tid == 0) {
            double lo = 0.0, hi = 0.0, all = 0.0;
            for (int t = 0; t < nt; ++t) {
                lo += red[t];
                hi += red[nt + t];
                all += red[2 * nt + t];
            }
            const double ratio = sqrt(lo + hi) / sqrt(all);
            a.warn[s] = (Dlp >= 2 && ratio > 1.4901161193847656e-08 /* sqrt(eps) */) ? 1 : 0;
        }
    }
}

// in: [B][2][D] unnormalised inverse transforms; out: [B][2*Dsub]
BLK void blk_rs_weights(const RsArgs &a, blk3 bid, int nt, void * /*smem*/)
{
    FOR_THREADS(tid, nt)
    {
        const long long gid = (long long)bid.x * nt + tid;
        if (gid < (long long)a.B * a.Dsub) {
            const int s = (int)(gid / a.Dsub), isub = (int)(gid % a.Dsub);
            const int i = isub * a.nskip;
            const double sf = 1.7320508075688772 / 6.0;
            const double w0 = 0.25 + sf, w1 = 0.25 - sf;
            const double invD = 1.0 / (double)a.D;
            const cplx q1 = cscale(a.in[(size_t)s * 2 * a.D + i], invD);
            const cplx q2 = cscale(a.in[(size_t)s * 2 * a.D + a.D + i], invD);
            if (a.wsel >= 2) {
                const cplx qn[3] = {q1, a.q0[(size_t)s * a.D + i], q2};
                const double ks = -(double)a.kappa;
                cplx *oq = a.out + ((size_t)s * a.Dsub + (size_t)isub) * a.up;
                cplx *orr = a.rout + ((size_t)s * a.Dsub + (size_t)isub) * a.up;
                for (int r = 0; r < a.up; ++r) {
                    cplx accq = czero(), accr = czero();
                    for (int m = 0; m < 3; ++m) {
                        const cplx w = bo_cf_w(a.wsel, r, m);
                        cfma(accq, w, qn[m]);
                        cfma(accr, w, make_cplx(ks * qn[m].x, -ks * qn[m].y));
                    }
                    oq[r] = accq;
                    orr[r] = accr;
                }
                continue;
            }
            if (a.up == 3) {
                const cplx qm = a.q0[(size_t)s * a.D + i];
                cplx *o3 = a.out + (size_t)s * 3 * a.Dsub + 3 * (size_t)isub;
                for (int r = 0; r < 3; ++r) {
                    const double *w = a.w3 + 3 * r;
                    o3[r] = make_cplx(w[0] * q1.x + w[1] * qm.x + w[2] * q2.x, w[0] * q1.y + w[1] * qm.y + w[2] * q2.y);
                }
                continue;
            }
            cplx *o = a.out + (size_t)s * 2 * a.Dsub + 2 * (size_t)isub;
            o[0] = make_cplx(w0 * q1.x + w1 * q2.x, w0 * q1.y + w1 * q2.y);
            o[1] = make_cplx(w1 * q1.x + w0 * q2.x, w1 * q1.y + w0 * q2.y);
        }
    }
}
