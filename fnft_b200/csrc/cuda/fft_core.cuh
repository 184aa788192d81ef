// fnft_b200 -- complex-double FFT building blocks (replacement for the Kiss FFT
// the reference vendors: /root/reference/src/3rd_party/kiss_fft/kiss_fft.c:21-408,
// reached through include/private/fnft__fft_wrapper.h:43-199).
//
// Conventions (same as the reference wrapper, fnft__fft_wrapper.h:79-80,102):
//   forward  X[k] = sum_n x[n] exp(-2*pi*i*k*n/N)
//   inverse  x[n] = sum_k X[k] exp(+2*pi*i*k*n/N)      (UNNORMALISED)
//
// Design: power-of-two lengths only (callers zero-pad / wrap-correct, linear
// convolution does not care about the FFT length).  A transform that fits in
// shared memory is done in place as a sequence of radix-16/8/4/2 passes:
//   forward = decimation in frequency, natural order in -> digit-reversed out
//   inverse = decimation in time,     digit-reversed in -> natural order out
// so no reordering pass is ever needed for convolutions.  Elements are stored
// with an XOR swizzle of the low three index bits so that every pass (strided
// or contiguous) is free of shared-memory bank conflicts for 16-byte accesses.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------
// compile-time roots of unity, exp(2*pi*i*k/64) via a quadrant table
// ---------------------------------------------------------------------------
HD constexpr double fnftb_cos64(int k)
{
    // cos(2*pi*k/64), k in [0, 64); first quadrant tabulated (nearest doubles)
    constexpr double t[17] = {
                              1.0,
                              0.9951847266721969,
                              0.9807852804032304,
                              0.9569403357322088,
                              0.9238795325112867,
                              0.881921264348355,
                              0.8314696123025452,
                              0.773010453362737,
                              0.7071067811865476,
                              0.6343932841636455,
                              0.5555702330196023,
                              0.4713967368259978,
                              0.38268343236508984,
                              0.29028467725446233,
                              0.19509032201612833,
                              0.09801714032956077,
                              0.0};
    k &= 63;
    if (k > 32)
        k = 64 - k;  // cos is even
    return (k <= 16) ? t[k] : -t[32 - k];
}
HD constexpr double fnftb_sin64(int k) { return fnftb_cos64(k - 16); }

// v *= exp(DIR * 2*pi*i * K / R) with trivial cases folded at compile time
template <int R, int K, int DIR>
HD cplx mul_root(cplx v)
{
    constexpr int k64 = (K * (64 / R)) & 63;
    if constexpr (k64 == 0) {
        return v;
    } else if constexpr (k64 == 16) {
        return (DIR > 0) ? cmuli(v) : cmulmi(v);
    } else if constexpr (k64 == 32) {
        return cneg(v);
    } else if constexpr (k64 == 48) {
        return (DIR > 0) ? cmulmi(v) : cmuli(v);
    } else {
        constexpr double c = fnftb_cos64(k64);
        constexpr double s = (DIR > 0) ? fnftb_sin64(k64) : -fnftb_sin64(k64);
        return make_cplx(v.x * c - v.y * s, v.x * s + v.y * c);
    }
}

// ---------------------------------------------------------------------------
// register-resident DFT of length R in {1,2,4,...,64}, natural order in/out
// (radix-2 decimation-in-time recursion, fully unrolled by the compiler)
// ---------------------------------------------------------------------------
template <int R, int DIR, int K>
struct DftCombine {
    HD static void run(cplx *v, const cplx *e, const cplx *o)
    {
        const cplx t = mul_root<R, K, DIR>(o[K]);
        v[K] = cadd(e[K], t);
        v[K + R / 2] = csub(e[K], t);
        if constexpr (K + 1 < R / 2)
            DftCombine<R, DIR, K + 1>::run(v, e, o);
    }
};

template <int R, int DIR>
struct Dft {
    HD static void run(cplx *v)
    {
        if constexpr (R == 1) {
            return;
        } else if constexpr (R == 2) {
            const cplx a = v[0], b = v[1];
            v[0] = cadd(a, b);
            v[1] = csub(a, b);
        } else {
            cplx e[R / 2], o[R / 2];
#pragma unroll
            for (int j = 0; j < R / 2; ++j) {
                e[j] = v[2 * j];
                o[j] = v[2 * j + 1];
            }
            Dft<R / 2, DIR>::run(e);
            Dft<R / 2, DIR>::run(o);
            DftCombine<R, DIR, 0>::run(v, e, o);
        }
    }
};

// ---------------------------------------------------------------------------
// shared-memory index swizzle (see header comment)
// ---------------------------------------------------------------------------
HD int swz(int i) { return i ^ ((i >> 3) & 7); }

// Twiddle table: tw[k] = exp(-2*pi*i*k/twn), k < twn (global memory, read-only).
struct TwTable {
    const cplx *tw;
    int twn;
    int log2twn;
};

// w = exp(-2*pi*i * m / len) from the table (len divides twn), DIR selects conj
template <int DIR>
HD cplx tw_lookup(const TwTable &T, int m, int log2len)
{
    const cplx w = LDG(&T.tw[m << (T.log2twn - log2len)]);
    return (DIR > 0) ? cconj(w) : w;
}

// compile-time log2 of the radix
template <int R>
struct Log2R {
    static const int value = (R <= 1) ? 0 : 1 + Log2R<R / 2>::value;
};
template <>
struct Log2R<1> {
    static const int value = 0;
};
template <>
struct Log2R<0> {
    static const int value = 0;
};

// ---------------------------------------------------------------------------
// one radix-R butterfly of an in-place pass over the array S (one FFT)
//   u    : butterfly index within this FFT, 0 <= u < n/R
//   s    : element stride of this pass (power of two), sub-FFT length = R*s
// forward (DIF):  y_j = DFT_R(x)_j * w_{R*s}^{j*o}
// inverse (DIT):  y   = IDFT_R( x_j * conj(w_{R*s}^{j*o}) )
// ---------------------------------------------------------------------------
template <int R, int DIR>
HD void fft_pass_butterfly(cplx *S, int u, int log2s, const TwTable &T)
{
    const int s = 1 << log2s;
    const int g = u >> log2s;
    const int o = u & (s - 1);
    const int base = (g << (Log2R<R>::value + log2s)) + o;
    const int len = Log2R<R>::value + log2s;  // log2 of the sub-transform length
    cplx v[R];
#pragma unroll
    for (int j = 0; j < R; ++j)
        v[j] = S[swz(base + j * s)];
    if constexpr (DIR > 0) {
        if (o != 0) {
#pragma unroll
            for (int j = 1; j < R; ++j)
                v[j] = cmul(v[j], tw_lookup<+1>(T, j * o, len));
        }
    }
    Dft<R, DIR>::run(v);
    if constexpr (DIR < 0) {
        if (o != 0) {
#pragma unroll
            for (int j = 1; j < R; ++j)
                v[j] = cmul(v[j], tw_lookup<-1>(T, j * o, len));
        }
    }
#pragma unroll
    for (int j = 0; j < R; ++j)
        S[swz(base + j * s)] = v[j];
}

// One pass over `nfft` transforms stored back to back (stride n) in shared memory.
template <int R, int DIR>
HD void fft_pass_all(cplx *S, int nfft, int log2n, int log2s, int tid, int nt,
                     const TwTable &T)
{
    const int log2bpf = log2n - Log2R<R>::value;  // butterflies per FFT (log2)
    const int nb = nfft << log2bpf;
    for (int b = tid; b < nb; b += nt) {
        const int f = b >> log2bpf;
        const int u = b & ((1 << log2bpf) - 1);
        fft_pass_butterfly<R, DIR>(S + ((size_t)f << log2n), u, log2s, T);
    }
}

template <int DIR>
HD void fft_pass_dispatch(int R, cplx *S, int nfft, int log2n, int log2s, int tid,
                          int nt, const TwTable &T)
{
    switch (R) {
    case 16: fft_pass_all<16, DIR>(S, nfft, log2n, log2s, tid, nt, T); break;
    case 8: fft_pass_all<8, DIR>(S, nfft, log2n, log2s, tid, nt, T); break;
    case 4: fft_pass_all<4, DIR>(S, nfft, log2n, log2s, tid, nt, T); break;
    case 2: fft_pass_all<2, DIR>(S, nfft, log2n, log2s, tid, nt, T); break;
    default: break;
    }
}

// Radix plan: as many radix-16 passes as possible, the remainder (8/4/2) LAST so
// that every earlier pass has stride >= 8 elements (bank-conflict-free together
// with swz()).  n = 1 gives an empty plan.
// Radix plan.  The LAST pass always has radix 4 (n >= 4): at stride 1 a pass carries no
// twiddles, which lets the convolution kernels fuse "last forward pass + pointwise
// product + first inverse pass" in registers (see fused_pointwise_* in the tree
// kernels).  The bits before it are covered by radix-16 passes plus one 8/4/2 pass
// (radix-8 passes only if max_radix < 16).
HD constexpr FftPlan make_fft_plan(int n, int max_radix = 16)
{
    FftPlan P{};
    P.n = n;
    P.log2n = 0;
    P.npass = 0;
    for (int i = 0; i < FNFTB_MAX_PASSES; ++i)
        P.radix[i] = 0;
    while ((1 << P.log2n) < n)
        ++P.log2n;
    if (P.log2n == 0)
        return P;
    if (P.log2n == 1) {
        P.radix[P.npass++] = 2;
        return P;
    }
    int rem = P.log2n - 2;
    if (max_radix < 8) {
        while (rem >= 2) {
            P.radix[P.npass++] = 4;
            rem -= 2;
        }
    } else if (max_radix < 16) {
        while (rem >= 3) {
            P.radix[P.npass++] = 8;
            rem -= 3;
        }
    } else {
        // 16 * 2 -> 8 * 4 so that no radix-2 pass appears next to radix-16 ones
        while (rem >= 4 && rem != 5 && rem != 6) {
            P.radix[P.npass++] = 16;
            rem -= 4;
        }
        if (rem == 6) {  // 16 * 4 -> 8 * 8 (a radix-4 pass at stride 4 has bank conflicts)
            P.radix[P.npass++] = 8;
            P.radix[P.npass++] = 8;
            rem = 0;
        }
        if (rem == 5) {
            P.radix[P.npass++] = 8;
            rem -= 3;
        }
    }
    if (rem)
        P.radix[P.npass++] = 1 << rem;
    P.radix[P.npass++] = 4;
    return P;
}

// log2 of the stride of the FIRST forward pass (= n / radix[0]); the frequency
// index k of the element stored at position pos after the forward transform
// satisfies  k mod radix[0] == pos >> log2_first_stride.
HD int plan_first_stride_log2(const FftPlan &P)
{
    return (P.npass == 0) ? 0 : P.log2n - ilog2i(P.radix[0]);
}

// frequency index of the element stored at position pos after a forward transform
HD int plan_freq_of_pos(const FftPlan &P, int pos)
{
    int rem = pos, k = 0, mult = 1, l2 = P.log2n;
    for (int p = 0; p < P.npass; ++p) {
        l2 -= ilog2i(P.radix[p]);
        const int j = rem >> l2;
        rem &= (1 << l2) - 1;
        k += j * mult;
        mult *= P.radix[p];
    }
    return k;
}

// ---------------------------------------------------------------------------
// team-scheduled transforms: the CTA is split into teams of `ts` consecutive threads
// (ts = fft_team_size(n, nt)); team t owns the transforms [t*fpt, (t+1)*fpt) for all
// passes, so passes only need a barrier among the team's threads and the teams drift
// apart, overlapping each other's load / compute / store phases.
// ---------------------------------------------------------------------------
HD int fft_team_size(int n, int nt)
{
    int ts = n / 16;
    if (ts < 32)
        ts = 32;
    while (ts < nt && nt / ts > 15 && ts > 32)  // named barriers: at most 15 teams
        ts *= 2;
    if (ts > nt)
        ts = nt;
    return ts;
}

template <int R, int DIR>
HD void fft_pass_team(cplx *S, int f0, int nf, int log2n, int log2s, int lane, int ts,
                      const TwTable &T)
{
    const int log2bpf = log2n - Log2R<R>::value;
    const int nb = nf << log2bpf;
    for (int b = lane; b < nb; b += ts) {
        const int f = f0 + (b >> log2bpf);
        const int u = b & ((1 << log2bpf) - 1);
        fft_pass_butterfly<R, DIR>(S + ((size_t)f << log2n), u, log2s, T);
    }
}

template <int DIR, int MAXR = 16>
HD void fft_pass_team_dispatch(int R, cplx *S, int nfft, int log2n, int log2s, int tid, int nt,
                               int ts, const TwTable &T)
{
    const int nteams = nt / ts;
    const int team = tid / ts, lane = tid - team * ts;
    const int fpt = (nfft + nteams - 1) / nteams;
    const int f0 = team * fpt;
    int nf = nfft - f0;
    if (nf > fpt)
        nf = fpt;
    if (nf <= 0)
        return;
    switch (R) {
    case 16:
        if constexpr (MAXR >= 16)
            fft_pass_team<16, DIR>(S, f0, nf, log2n, log2s, lane, ts, T);
        break;
    case 8:
        if constexpr (MAXR >= 8)
            fft_pass_team<8, DIR>(S, f0, nf, log2n, log2s, lane, ts, T);
        break;
    case 4: fft_pass_team<4, DIR>(S, f0, nf, log2n, log2s, lane, ts, T); break;
    case 2: fft_pass_team<2, DIR>(S, f0, nf, log2n, log2s, lane, ts, T); break;
    default: break;
    }
}

// In-place forward transforms of `nfft` arrays of length P.n (shared memory).
// Must be called by all threads of the block program.  On return the team's own
// transforms are complete (team barrier); callers that read other teams' data must
// BLOCK_SYNC() first.
#define FNFTB_SMEM_FFT_FWD(S, nfft, P, nt, T) FNFTB_SMEM_FFT_FWD_R(S, nfft, P, nt, T, 16)
#define FNFTB_SMEM_FFT_INV(S, nfft, P, nt, T) FNFTB_SMEM_FFT_INV_R(S, nfft, P, nt, T, 16)
#define FNFTB_SMEM_FFT_FWD_R(S, nfft, P, nt, T, MAXR) \
    FNFTB_SMEM_FFT_FWD_SKIP(S, nfft, P, nt, T, MAXR, 0)
#define FNFTB_SMEM_FFT_INV_R(S, nfft, P, nt, T, MAXR) \
    FNFTB_SMEM_FFT_INV_SKIP(S, nfft, P, nt, T, MAXR, 0)
// SKIP = 1 leaves out the stride-1 pass (the last forward / first inverse pass), which the
// caller then performs itself fused with its pointwise work.
#define FNFTB_SMEM_FFT_FWD_SKIP(S, nfft, P, nt, T, MAXR, SKIP)                        \
    do {                                                                              \
        int _l2s = (P).log2n;                                                         \
        const int _ts = fft_team_size((P).n, nt);                                     \
        for (int _p = 0; _p < (P).npass - (SKIP); ++_p) {                                      \
            const int _R = (P).radix[_p];                                             \
            _l2s -= ilog2i(_R);                                                       \
            FOR_THREADS(tid, nt)                                                      \
            {                                                                         \
                fft_pass_team_dispatch<-1, MAXR>(_R, (S), (nfft), (P).log2n, _l2s, tid, nt, _ts, T); \
            }                                                                         \
            TEAM_SYNC(_ts, nt);                                                       \
        }                                                                             \
    } while (0)

// In-place inverse (unnormalised) transforms, consuming the forward's ordering.
#define FNFTB_SMEM_FFT_INV_SKIP(S, nfft, P, nt, T, MAXR, SKIP)                        \
    do {                                                                              \
        int _l2s = (SKIP) ? ilog2i((P).radix[(P).npass - 1]) : 0;                     \
        const int _ts = fft_team_size((P).n, nt);                                     \
        for (int _p = (P).npass - 1 - (SKIP); _p >= 0; --_p) {                                 \
            const int _R = (P).radix[_p];                                             \
            FOR_THREADS(tid, nt)                                                      \
            {                                                                         \
                fft_pass_team_dispatch<+1, MAXR>(_R, (S), (nfft), (P).log2n, _l2s, tid, nt, _ts, T); \
            }                                                                         \
            TEAM_SYNC(_ts, nt);                                                       \
            _l2s += ilog2i(_R);                                                       \
        }                                                                             \
    } while (0)

// ---------------------------------------------------------------------------
// Compile-time specialised transforms (length 2^LOG2N, default radix plan): every stride,
// shift and twiddle step is a constant, so the address arithmetic of a pass folds into
// immediates.  Used by the product-tree kernels for the lengths they actually run
// (32 ... 1024); other lengths use the run-time-plan macros above.
// ---------------------------------------------------------------------------
HD constexpr int plan_log2_prefix(const FftPlan &P, int npasses)
{
    int s = 0;
    for (int i = 0; i < npasses; ++i) {
        int r = P.radix[i], l = 0;
        while ((1 << l) < r)
            ++l;
        s += l;
    }
    return s;
}

template <int R, int DIR, int LOG2N, int LOG2S>
HD void fft_pass_butterfly_ct(cplx *S, int u, const TwTable &T)
{
    constexpr int s = 1 << LOG2S;
    constexpr int LR = Log2R<R>::value;
    const int g = u >> LOG2S;
    const int o = u & (s - 1);
    const int base = (g << (LR + LOG2S)) + o;
    constexpr int log2len = LR + LOG2S;
    cplx v[R];
#pragma unroll
    for (int j = 0; j < R; ++j)
        v[j] = S[swz(base + j * s)];
    if constexpr (DIR > 0 && LOG2S > 0) {
        if (o != 0) {
#pragma unroll
            for (int j = 1; j < R; ++j)
                v[j] = cmul(v[j], tw_lookup<+1>(T, j * o, log2len));
        }
    }
    Dft<R, DIR>::run(v);
    if constexpr (DIR < 0 && LOG2S > 0) {
        if (o != 0) {
#pragma unroll
            for (int j = 1; j < R; ++j)
                v[j] = cmul(v[j], tw_lookup<-1>(T, j * o, log2len));
        }
    }
#pragma unroll
    for (int j = 0; j < R; ++j)
        S[swz(base + j * s)] = v[j];
}

template <int R, int DIR, int LOG2N, int LOG2S>
HD void fft_pass_team_ct(cplx *S, int nfft, int tid, int nt, int ts, const TwTable &T)
{
    constexpr int log2bpf = LOG2N - Log2R<R>::value;
    const int nteams = nt / ts;
    const int team = tid / ts, lane = tid - team * ts;
    const int fpt = (nfft + nteams - 1) / nteams;
    const int f0 = team * fpt;
    int nf = nfft - f0;
    if (nf > fpt)
        nf = fpt;
    if (nf <= 0)
        return;
    const int nb = nf << log2bpf;
    for (int b = lane; b < nb; b += ts) {
        const int f = f0 + (b >> log2bpf);
        const int u = b & ((1 << log2bpf) - 1);
        fft_pass_butterfly_ct<R, DIR, LOG2N, LOG2S>(S + ((size_t)f << LOG2N), u, T);
    }
}

template <int LOG2N, int DIR, int SKIP, int PASS>
struct FftRunnerCt {
    BLK static void run(cplx *S, int nfft, int nt, const TwTable &T)
    {
        constexpr FftPlan P = make_fft_plan(1 << LOG2N, 16);
        constexpr int NP = P.npass - SKIP;
        if constexpr (PASS < NP) {
            constexpr int p = (DIR < 0) ? PASS : (NP - 1 - PASS);
            constexpr int R = P.radix[p];
            constexpr int L2S = LOG2N - plan_log2_prefix(P, p + 1);
            const int ts = fft_team_size(1 << LOG2N, nt);
            FOR_THREADS(tid, nt)
            {
                fft_pass_team_ct<R, DIR, LOG2N, L2S>(S, nfft, tid, nt, ts, T);
            }
            TEAM_SYNC(ts, nt);
            FftRunnerCt<LOG2N, DIR, SKIP, PASS + 1>::run(S, nfft, nt, T);
        }
    }
};

// Dispatch on the run-time length: specialised code for 32..1024, generic otherwise.
// (Same contract as FNFTB_SMEM_FFT_{FWD,INV}_SKIP with the default radix-16 plan.)
#define FNFTB_SMEM_FFT_CT(DIR, S, nfft, P, nt, T, SKIP)                               \
    do {                                                                              \
        switch ((P).log2n) {                                                          \
        case 5: FftRunnerCt<5, DIR, SKIP, 0>::run((S), (nfft), (nt), (T)); break;     \
        case 6: FftRunnerCt<6, DIR, SKIP, 0>::run((S), (nfft), (nt), (T)); break;     \
        case 7: FftRunnerCt<7, DIR, SKIP, 0>::run((S), (nfft), (nt), (T)); break;     \
        case 8: FftRunnerCt<8, DIR, SKIP, 0>::run((S), (nfft), (nt), (T)); break;     \
        case 9: FftRunnerCt<9, DIR, SKIP, 0>::run((S), (nfft), (nt), (T)); break;     \
        case 10: FftRunnerCt<10, DIR, SKIP, 0>::run((S), (nfft), (nt), (T)); break;   \
        default:                                                                      \
            if ((DIR) < 0) {                                                          \
                FNFTB_SMEM_FFT_FWD_SKIP(S, nfft, P, nt, T, 16, SKIP);                 \
            } else {                                                                  \
                FNFTB_SMEM_FFT_INV_SKIP(S, nfft, P, nt, T, 16, SKIP);                 \
            }                                                                         \
            break;                                                                    \
        }                                                                             \
    } while (0)
