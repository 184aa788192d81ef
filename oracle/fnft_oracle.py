"""ORACLE -- numpy restatement of the reference's fast forward NFT hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (fnft_b200/, include/) imports,
calls, links or executes this module; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may, and only as the checker.

Parity status: PINNED.  tests/test_oracle.py checks every function below against
  * the golden vectors of the reference's own unit tests (fmult2x2, chirp-z,
    akns_fscatter per scheme), lifted into tests/golden/golden.npz by
    tests/golden/make_golden.py, and
  * outputs of the unmodified reference library (oracle/_ref, built by
    oracle/Makefile from /root/reference) on seeded inputs, stored in the same file.

Each function cites the reference code it follows (paths relative to the FNFT tree).
The algorithms are restated, not translated line by line: loops over pairs /
frequencies / eigenvalues are numpy-vectorised, the FFT is numpy's (at the same
transform lengths the reference uses, kiss_fft_next_fast_size).
"""
import numpy as np

EPS = np.finfo(np.float64).eps

# values of fnft__akns_discretization_t (include/private/fnft__akns_discretization_t.h:43-72)
AKNS_2SPLIT2_MODAL, AKNS_2SPLIT1A, AKNS_2SPLIT1B, AKNS_2SPLIT2A, AKNS_2SPLIT2B, AKNS_2SPLIT2S = range(6)
AKNS_2SPLIT3A, AKNS_2SPLIT3B, AKNS_2SPLIT3S, AKNS_2SPLIT4A, AKNS_2SPLIT4B = 6, 7, 8, 9, 10
AKNS_2SPLIT5A, AKNS_2SPLIT5B, AKNS_2SPLIT6A, AKNS_2SPLIT6B = 11, 12, 13, 14
AKNS_2SPLIT7A, AKNS_2SPLIT7B, AKNS_2SPLIT8A, AKNS_2SPLIT8B = 15, 16, 17, 18
AKNS_4SPLIT4A = 20
AKNS_4SPLIT4B = 21

# Higher-order splittings as weighted sums of chains (the MATLAB recipes quoted in
# test/fnft__akns_fscatter/fnft__akns_fscatter_test_2split{3A..8B}.c, e.g. _2split6B.c:47-49;
# src/private/fnft__akns_fscatter.c:256-400,435-912 spells the same sums out per coefficient).
# Entry: scheme -> (degree, [(weight, m, leftmost)]) where the chain has m+1 alternating
# factors exp(A.), exp(B.) of sizes 1,2,...,2,1 (in units of eps_t/m), leftmost 'A' or 'B'.
def _richardson(ms):
    return [np.prod([mk * mk / (mk * mk - mj * mj) for mj in ms if mj != mk]) for mk in ms]


def _chain_scheme(order, left, deg):
    ms = list(range(1, order + 1, 2)) if order % 2 else list(range(2, order + 1, 2))
    return deg, [(c, m, left) for c, m in zip(_richardson(ms), ms)]


_CHAINS = {
    AKNS_2SPLIT3A: _chain_scheme(3, 'A', 3), AKNS_2SPLIT3B: _chain_scheme(3, 'B', 3),
    AKNS_2SPLIT3S: (2, [(2 / 3, 2, 'A'), (2 / 3, 2, 'B'), (-1 / 6, 1, 'A'), (-1 / 6, 1, 'B')]),
    AKNS_2SPLIT4A: _chain_scheme(4, 'A', 4), AKNS_4SPLIT4A: _chain_scheme(4, 'A', 4),
    AKNS_2SPLIT5A: _chain_scheme(5, 'A', 15), AKNS_2SPLIT5B: _chain_scheme(5, 'B', 15),
    AKNS_2SPLIT6A: _chain_scheme(6, 'A', 12), AKNS_2SPLIT6B: _chain_scheme(6, 'B', 6),
    AKNS_2SPLIT7A: _chain_scheme(7, 'A', 105), AKNS_2SPLIT7B: _chain_scheme(7, 'B', 105),
    AKNS_2SPLIT8A: _chain_scheme(8, 'A', 24), AKNS_2SPLIT8B: _chain_scheme(8, 'B', 12),
}
# values of fnft_nse_discretization_t (include/fnft_nse_discretization_t.h:104-133)
NSE_2SPLIT2_MODAL, NSE_BO, NSE_2SPLIT1A, NSE_2SPLIT1B, NSE_2SPLIT2A, NSE_2SPLIT2B, NSE_2SPLIT2S = range(7)
NSE_2SPLIT4B = 11
NSE_4SPLIT4B = 21
NSE_CF4_2 = 22
NSE_CF4_3 = 23
NSE_CF5_3 = 24
NSE_CF6_4 = 25
NSE_ES4, NSE_TES4 = 26, 27
_SLOW_UP = {NSE_CF4_2: 2, NSE_CF4_3: 3, NSE_CF5_3: 3, NSE_CF6_4: 4, NSE_ES4: 3, NSE_TES4: 3}   # upsampling factors of the slow schemes
# fnft_kdv_discretization_t (include/fnft_kdv_discretization_t.h:96-122)
KDV_2SPLIT1A, KDV_2SPLIT1B, KDV_2SPLIT2A, KDV_2SPLIT2B, KDV_2SPLIT2S = range(5)
KDV_2SPLIT4B = 9
KDV_4SPLIT4B = 19

# enum -> akns scheme (src/private/fnft__nse_discretization.c:109-202,
# src/private/fnft__kdv_discretization.c:98-193): the polynomial schemes keep their order
_NSE2AKNS = {NSE_2SPLIT2_MODAL: AKNS_2SPLIT2_MODAL}
_NSE2AKNS.update({2 + i: 1 + i for i in range(18)})        # 2SPLIT1A .. 2SPLIT8B
_NSE2AKNS.update({20: AKNS_4SPLIT4A, NSE_4SPLIT4B: AKNS_4SPLIT4B})
_KDV2AKNS = {i: 1 + i for i in range(18)}                  # 2SPLIT1A .. 2SPLIT8B
_KDV2AKNS.update({18: AKNS_4SPLIT4A, KDV_4SPLIT4B: AKNS_4SPLIT4B})


def akns_degree(scheme):
    """src/private/fnft__akns_discretization.c:29-67 (schemes restated here)."""
    if scheme in _CHAINS:
        return _CHAINS[scheme][0]
    return 2 if scheme in (AKNS_2SPLIT4B, AKNS_4SPLIT4B) else 1


def akns_upsampling(scheme):
    """src/private/fnft__akns_discretization.c:114-154."""
    return 2 if scheme in (AKNS_4SPLIT4A, AKNS_4SPLIT4B) else 1


def next_fast_size(n):
    """kiss_fft_next_fast_size, src/3rd_party/kiss_fft/kiss_fft.c:396-408."""
    while True:
        m = n
        for f in (2, 3, 5):
            while m % f == 0:
                m //= f
        if m <= 1:
            return n
        n += 1


def nextpow2(n):
    """misc_nextpowerof2, src/private/fnft__misc.c:316-324."""
    r = 1
    while r < n:
        r *= 2
    return r


def csinc(x):
    """misc_CSINC, src/private/fnft__misc.c:306-314."""
    x = np.asarray(x, dtype=np.complex128)
    small = np.abs(x) < 1.0e-8
    safe = np.where(small, 1.0, x)
    return np.where(small, np.cos(x / np.sqrt(3.0 + 0j)), np.sin(safe) / safe)


def zero_freq_expm(h, q, r):
    """akns_fscatter_zero_freq_scatter_matrix, src/private/fnft__akns_fscatter.c:46-59.
    Returns (M0, M1, M2) = (cos D, q*h*sinc D, r*h*sinc D)."""
    Delta = h * np.sqrt(-q * r + 0j)
    dl = h * csinc(Delta)
    return np.cos(Delta), q * dl, r * dl


def akns_leaves(q, r, eps_t, scheme):
    """Per-sample 2x2 polynomial matrices, src/private/fnft__akns_fscatter.c:116-433.
    Returns p[4, D, deg+1]; matrix k belongs to sample D-1-k (:408), coefficients
    highest power first."""
    q = np.asarray(q, dtype=np.complex128)[::-1]
    r = np.asarray(r, dtype=np.complex128)[::-1]
    D = q.shape[0]
    deg = akns_degree(scheme)
    p = np.zeros((4, D, deg + 1), dtype=np.complex128)
    if scheme == AKNS_2SPLIT2_MODAL:  # :118-147
        scl = 1.0 / np.sqrt(1 - eps_t * q * eps_t * r + 0j)
        p[0, :, 1] = scl
        p[1, :, 0] = scl * eps_t * q
        p[2, :, 1] = scl * eps_t * r
        p[3, :, 0] = scl
    elif scheme == AKNS_2SPLIT1A:  # :149-176
        e0, e1, e2 = zero_freq_expm(eps_t / deg, q, r)
        p[0, :, 1] = e0
        p[1, :, 1] = e1
        p[2, :, 0] = e2
        p[3, :, 0] = e0
    elif scheme in (AKNS_2SPLIT1B, AKNS_2SPLIT2A):  # :178-203
        e0, e1, e2 = zero_freq_expm(eps_t / deg, q, r)
        p[0, :, 1] = e0
        p[1, :, 0] = e1
        p[2, :, 1] = e2
        p[3, :, 0] = e0
    elif scheme == AKNS_2SPLIT2B:  # :204-230
        e0, e1, e2 = zero_freq_expm(0.5 * eps_t / deg, q, r)
        p[0, :, 0] = e1 * e2
        p[0, :, 1] = e0 * e0
        p[1, :, 0] = p[1, :, 1] = e0 * e1
        p[2, :, 0] = p[2, :, 1] = e0 * e2
        p[3, :, 0] = p[0, :, 1]
        p[3, :, 1] = p[0, :, 0]
    elif scheme == AKNS_2SPLIT2S:  # :232-258
        e0, e1, e2 = zero_freq_expm(eps_t / deg, q, r)
        p[0, :, 1] = e0
        p[1, :, 0] = p[1, :, 1] = e1 / 2
        p[2, :, 0] = p[2, :, 1] = e2 / 2
        p[3, :, 0] = e0
    elif scheme in (AKNS_2SPLIT4B, AKNS_4SPLIT4B):  # :402-433
        a0, a1, a2 = zero_freq_expm(0.5 * eps_t / deg, q, r)
        b0, b1, b2 = zero_freq_expm(eps_t / deg, q, r)
        p[0, :, 0] = (4 * b0 * a1 * a2 - b1 * b2) / 3
        p[0, :, 1] = 4 * (b1 * a0 * a2 + b2 * a0 * a1) / 3
        p[0, :, 2] = (4 * b0 * a0 * a0 - b0 * b0) / 3
        p[1, :, 0] = (4 * b0 * a0 * a1 - b0 * b1) / 3
        p[1, :, 1] = 4 * (b1 * a0 * a0 + b2 * a1 * a1) / 3
        p[1, :, 2] = p[1, :, 0]
        p[2, :, 0] = (4 * b0 * a0 * a2 - b0 * b2) / 3
        p[2, :, 1] = 4 * (b2 * a0 * a0 + b1 * a2 * a2) / 3
        p[2, :, 2] = p[2, :, 0]
        p[3, :, 0] = p[0, :, 2]
        p[3, :, 1] = p[0, :, 1]
        p[3, :, 2] = p[0, :, 0]
    elif scheme in _CHAINS:
        for c, m, left in _CHAINS[scheme][1]:
            p += c * _chain_polys(q, r, eps_t, deg, m, left)
    else:
        raise ValueError("scheme not restated in the oracle")
    return p


def _chain_polys(q, r, eps_t, deg, m, left):
    """One chain as a 2x2 matrix of polynomials in z (dense polynomial products, highest power
    first): exp(A f eps_t) = diag(1, z^(f deg)) up to a scalar, exp(B f eps_t) from
    zero_freq_expm (src/private/fnft__akns_fscatter.c:46-59)."""
    D = q.shape[0]
    # P[i][j]: coefficient arrays [D, deg+1] in ASCENDING powers while multiplying
    P = [[np.zeros((D, deg + 1), dtype=np.complex128) for _ in range(2)] for _ in range(2)]
    P[0][0][:, 0] = 1.0
    P[1][1][:, 0] = 1.0
    for i in range(m + 1):
        size = 1 if i in (0, m) else 2
        is_a = (i % 2 == 0) == (left == 'A')
        if is_a:  # right-multiply by diag(1, z^k): shifts column 1
            k = size * deg // m
            assert size * deg % m == 0
            for row in range(2):
                P[row][1] = np.concatenate([np.zeros((D, k), dtype=np.complex128), P[row][1][:, :deg + 1 - k]], axis=1)
        else:
            c, sq, sr = zero_freq_expm(size * eps_t / m, q, r)
            E = [[c, sq], [sr, c]]
            P = [[P[row][0] * E[0][col][:, None] + P[row][1] * E[1][col][:, None] for col in range(2)]
                 for row in range(2)]
    out = np.empty((4, D, deg + 1), dtype=np.complex128)
    for row in range(2):
        for col in range(2):
            out[2 * row + col] = P[row][col][:, ::-1]
    return out


def poly_fmult2x2(p, normalize=True):
    """fnft__poly_fmult2x2, src/private/fnft__poly_fmult.c:381-546.
    p: [4, n, deg+1].  Returns (result[4, deg_out+1], deg_out, W).
    Level loop :460-519; pair product by FFT at length next_fast_size(2*deg+1)
    (:45-48, 239-328); per-pair rescale by 2^-floor(log2(max|c|)) (:330-374)."""
    p = np.array(p, dtype=np.complex128)
    _, n, d1 = p.shape
    deg0 = d1 - 1
    npad = nextpow2(n)
    if npad > n:  # pad with z^deg * I, :404-445
        pad = np.zeros((4, npad - n, d1), dtype=np.complex128)
        pad[0, :, 0] = 1.0
        pad[3, :, 0] = 1.0
        p = np.concatenate([p, pad], axis=1)
    W = 0
    deg = deg0
    while p.shape[1] >= 2:
        L = next_fast_size(2 * deg + 1)
        F = np.fft.fft(p, n=L, axis=2)               # zero-padded forward transforms
        A, Bm = F[:, 0::2, :], F[:, 1::2, :]          # left / right factor of each pair
        C = np.empty((4,) + A.shape[1:], dtype=np.complex128)
        C[0] = A[0] * Bm[0] + A[1] * Bm[2]
        C[1] = A[0] * Bm[1] + A[1] * Bm[3]
        C[2] = A[2] * Bm[0] + A[3] * Bm[2]
        C[3] = A[2] * Bm[1] + A[3] * Bm[3]
        c = np.fft.ifft(C, axis=2)[:, :, :2 * deg + 1]
        if normalize:
            mx = np.abs(c).max(axis=(0, 2))
            a = np.where(mx > 0, np.floor(np.log2(np.where(mx > 0, mx, 1.0))), 0.0)
            c = c * (2.0 ** (-a))[None, :, None]
            W += int(a.sum())
        p = c
        deg *= 2
    res = p[:, 0, :]
    deg_out = deg0 * n
    return res[:, :deg_out + 1].copy(), deg_out, W


def akns_fscatter(q, r, eps_t, scheme, normalize=True):
    """fnft__akns_fscatter, src/private/fnft__akns_fscatter.c:64-925."""
    return poly_fmult2x2(akns_leaves(q, r, eps_t, scheme), normalize)


def nse_fscatter(q, eps_t, kappa, nse_disc, normalize=True):
    """fnft__nse_fscatter, src/private/fnft__nse_fscatter.c:44-91 (r = -kappa*conj(q))."""
    q = np.asarray(q, dtype=np.complex128)
    return akns_fscatter(q, -kappa * np.conj(q), eps_t, _NSE2AKNS[nse_disc], normalize)


def kdv_fscatter(u, eps_t, kdv_disc, normalize=True):
    """fnft__kdv_fscatter, src/private/fnft__kdv_fscatter.c:45-83 (r = -1)."""
    u = np.asarray(u, dtype=np.complex128)
    return akns_fscatter(u, -np.ones_like(u), eps_t, _KDV2AKNS[kdv_disc], normalize)


def poly_chirpz(p, A, W, M):
    """fnft__poly_chirpz, src/private/fnft__poly_chirpz.c:33-105: evaluates p (highest
    power first) at z_m = 1/(A*W^-m), m < M, by Bluestein with three FFTs."""
    p = np.asarray(p, dtype=np.complex128)
    deg = p.shape[0] - 1
    N = deg + 1
    L = next_fast_size(N + M - 1)
    A = complex(A)
    W = complex(W)
    n = np.arange(N, dtype=np.float64)
    y = np.zeros(L, dtype=np.complex128)
    y[:N] = p[::-1] * np.power(A, -n) * np.power(W, 0.5 * n * n)
    v = np.zeros(L, dtype=np.complex128)
    m = np.arange(M, dtype=np.float64)
    v[:M] = np.power(W, -0.5 * m * m)
    k = np.arange(L - N + 1, L, dtype=np.float64)
    v[L - N + 1:] = np.power(W, -0.5 * (L - k) * (L - k))
    g = np.fft.ifft(np.fft.fft(y) * np.fft.fft(v))    # ifft already divides by L (:95)
    return np.power(W, 0.5 * m * m) * g[:M]


def lambda_to_z(lam, eps_t, scheme):
    """fnft__akns_discretization_lambda_to_z, src/private/fnft__akns_discretization.c:204-219."""
    return np.exp(2j * lam * eps_t / (akns_degree(scheme) * akns_upsampling(scheme)))


def resample(q, eps_t, delta):
    """misc_resample, src/private/fnft__misc.c:326-407 (band-limited shift by delta)."""
    q = np.asarray(q, dtype=np.complex128)
    D = q.shape[0]
    i = np.arange(D)
    freq = np.where(i < D // 2, i, i - D) / (D * eps_t)
    return np.fft.ifft(np.fft.fft(q) * np.exp(2j * np.pi * delta * freq))


def preprocess_signal(q, eps_t, kappa, nse_disc):
    """nse_discretization_preprocess_signal without subsampling,
    src/private/fnft__nse_discretization.c:386-656 (:467-473 copy, :474-503 4SPLIT4)."""
    q = np.asarray(q, dtype=np.complex128)
    if nse_disc in (20, NSE_4SPLIT4B, NSE_CF4_2):   # 4SPLIT4A, 4SPLIT4B, CF4_2
        s = np.sqrt(3.0) / 6.0
        q1 = resample(q, eps_t, -eps_t * s)
        q2 = resample(q, eps_t, +eps_t * s)
        w = (0.25 + s, 0.25 - s, 0.25 - s, 0.25 + s)  # fnft__akns_discretization.c:284-298
        out = np.empty(2 * q.shape[0], dtype=np.complex128)
        out[0::2] = w[0] * q1 + w[1] * q2
        out[1::2] = w[2] * q1 + w[3] * q2
        return out
    if nse_disc == NSE_CF4_3:                       # :505-531
        s = np.sqrt(3.0 / 20.0)
        q1 = resample(q, eps_t, -eps_t * s)
        q3 = resample(q, eps_t, +eps_t * s)
        w = cf4_3_weights()
        out = np.empty(3 * q.shape[0], dtype=np.complex128)
        for i in range(3):
            out[i::3] = w[i, 0] * q1 + w[i, 1] * q + w[i, 2] * q3
        return out
    if nse_disc in (NSE_ES4, NSE_TES4):             # :609-631: (q, q', q'') by central differences, zero outside
        D = q.shape[0]
        qz = np.concatenate([[0.0], q, [0.0]])
        out = np.empty(3 * D, dtype=np.complex128)
        out[0::3] = q
        out[1::3] = (qz[2:] - qz[:-2]) / (2 * eps_t)
        out[2::3] = (qz[2:] - 2 * q + qz[:-2]) / (eps_t * eps_t)
        return out
    return q.copy()


def _pauli_exp(a1, a2, a3):
    """exp of a1*sigma1 + a2*sigma2 + a3*sigma3 as the reference writes it
    (src/private/fnft__akns_scatter_matrix.c:464-480): returns (U11, U12, U21, U22), s, c, w."""
    w = np.sqrt(-(a1 * a1) - (a2 * a2) - (a3 * a3) + 0j)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(w != 0, np.sin(w) / np.where(w != 0, w, 1.0), 1.0)
    c = np.cos(w)
    return (c + s * a3, s * (a1 - 1j * a2), s * (a1 + 1j * a2), c - s * a3), s, c, w


def _mm(A, B):
    return (A[0] * B[0] + A[1] * B[2], A[0] * B[1] + A[1] * B[3],
            A[2] * B[0] + A[3] * B[2], A[2] * B[1] + A[3] * B[3])


def _es_step(q3, r3, l, h, tes, with_d=False):
    """One step of ES4 (tes False) / TES4 (tes True) from the samples (q, q', q'') of a grid point, step h
    (negative: the backward sweep, whose pre-computed quantities are the forward ones with -h,
    src/private/fnft__nse_scatter_bound_states.c:124-183, :343-470, :535-630; continuous spectrum:
    src/private/fnft__akns_scatter_matrix.c:259-320, :464-515).  Returns (U, Ud); Ud follows the
    reference's formulas literally (TES4: s_d = sin(w*eps_t)/w with w already proportional to eps_t)."""
    q, qd, qdd = q3
    r, rd, rdd = r3
    h2, h3 = h * h, h * h * h
    if not tes:
        t1 = h3 * (qdd + rdd) / 48.0 + (h * (q + r)) * 0.5
        t2 = (h * (q - r) * 1j) * 0.5 + (h3 * (qdd - rdd) * 1j) / 48.0
        t3 = -h3 * (q * rd - qd * r) / 12.0
        a1 = t1 + h3 * (l * 1j * (qd - rd)) / 12.0
        a2 = t2 - h3 * l * (qd + rd) / 12.0
        a3 = -h * 1j * l + t3
        U, s, c, w = _pauli_exp(a1, a2, a3)
        Ud = None
        if with_d:
            d1 = 1j * h3 * (qd - rd) / 12.0
            d2 = -h3 * (qd + rd) / 12.0
            d3 = -1j * h
            with np.errstate(divide="ignore", invalid="ignore"):
                w_d = -(1 / w) * (a1 * d1 + a2 * d2 + a3 * d3)
                c_d = -np.sin(w) * w_d
                s_d = w_d * (c - s) / w
            Ud = (c_d + s_d * a3 + s * d3, s_d * a1 + s * d1 - 1j * s_d * a2 - 1j * s * d2,
                  s_d * a1 + s * d1 + 1j * s_d * a2 + 1j * s * d2, c_d - s_d * a3 - s * d3)
        return U, Ud
    e1 = (h3 * (qdd + rdd)) / 96.0 - (h2 * (qd + rd)) / 24.0, (h3 * (qdd - rdd) * 1j) / 96.0 + (h2 * (rd - qd) * 1j) / 24.0
    e3 = (h3 * (qdd + rdd)) / 96.0 + (h2 * (qd + rd)) / 24.0, (h3 * (qdd - rdd) * 1j) / 96.0 + (h2 * (qd - rd) * 1j) / 24.0
    E1, _, _, _ = _pauli_exp(e1[0], e1[1], 0.0)
    a1 = (h * (q + r)) * 0.5
    a2 = (h * (q * 1j - r * 1j)) * 0.5
    a3 = -h * l * 1j
    E2, s, c, w = _pauli_exp(a1, a2, a3)
    E3, _, _, _ = _pauli_exp(e3[0], e3[1], 0.0)
    U = _mm(E3, _mm(E2, E1))
    Ud = None
    if with_d:
        with np.errstate(divide="ignore", invalid="ignore"):
            s_d = np.sin(w * h) / w
            c_d = -h * l * s_d
            w_d = l * (h * w * np.cos(w * h) - np.sin(w * h)) / (w * w * w)
        UD = (c_d - 1j * s_d, w_d * q, w_d * r, c_d + 1j * s_d)
        Ud = _mm(E3, _mm(UD, E1))
    return U, Ud


def cf_complex_weights(nse_disc):
    """akns_discretization_method_weights for CF5_3 (3x3) and CF6_4 (4x3),
    src/private/fnft__akns_discretization.c:330-367."""
    if nse_disc == NSE_CF5_3:
        s15 = np.sqrt(15.0)
        w = np.zeros(9, dtype=np.complex128)
        w[0] = ((145.0 + 37.0 * s15) / 900.0) + 1j * ((5.0 + 3.0 * s15) / 300.0)
        w[1] = (-1.0 / 45.0) + 1j * (1.0 / 15.0)
        w[2] = ((145.0 - 37.0 * s15) / 900.0) + 1j * ((5.0 - 3.0 * s15) / 300.0)
        w[3] = (-2.0 / 45.0) + 1j * (-s15 / 50.0)
        w[4] = 22.0 / 45.0
        w[5:9] = np.conj(w[3::-1])
        return w.reshape(3, 3)
    h = np.array([0.245985577298764 + 0.038734389227165j, -0.046806149832549 + 0.012442141491185j,
                  0.010894359342569 - 0.004575808769067j, 0.062868370946917 - 0.048761268117765j,
                  0.269028372054771 - 0.012442141491185j, -0.041970529810473 + 0.014602687659668j])
    return np.concatenate([h, h[::-1]]).reshape(4, 3)


def preprocess_signal_qr(q, eps_t, kappa, nse_disc):
    """(q_preprocessed, r_preprocessed) of nse_discretization_preprocess_signal.  For CF5_3 / CF6_4
    (src/private/fnft__nse_discretization.c:532-604) the complex weights are applied to r = -kappa*conj(q)
    unconjugated, so r_preprocessed is not -kappa*conj(q_preprocessed); for every other scheme it is."""
    q = np.asarray(q, dtype=np.complex128)
    if nse_disc in (NSE_CF5_3, NSE_CF6_4):
        d = eps_t * np.sqrt(15.0) / 10.0
        qs = (resample(q, eps_t, -d), q, resample(q, eps_t, +d))
        rs = tuple(-kappa * np.conj(x) for x in qs)
        w = cf_complex_weights(nse_disc)
        up = w.shape[0]
        qp = np.empty(up * q.shape[0], dtype=np.complex128)
        rp = np.empty(up * q.shape[0], dtype=np.complex128)
        for i in range(up):
            qp[i::up] = w[i, 0] * qs[0] + w[i, 1] * qs[1] + w[i, 2] * qs[2]
            rp[i::up] = w[i, 0] * rs[0] + w[i, 1] * rs[1] + w[i, 2] * rs[2]
        return qp, rp
    qp = preprocess_signal(q, eps_t, kappa, nse_disc)
    return qp, -kappa * np.conj(qp)


def cf4_3_weights():
    """akns_discretization_method_weights for CF4_3, src/private/fnft__akns_discretization.c:299-327:
    Legendre expansion of the coefficient table f at the three Gauss nodes.  Row sums (the weights of
    the spectral parameter, fnft__akns_scatter_matrix.c:101-109): 11/40, 9/20, 11/40."""
    f = np.array([[11.0 / 40.0, 20.0 / 87.0, 7.0 / 50.0],
                  [9.0 / 20.0, 0.0, -7.0 / 25.0],
                  [11.0 / 40.0, -20.0 / 87.0, 7.0 / 50.0]])
    wm = np.array([5.0 / 18.0, 4.0 / 9.0, 5.0 / 18.0])
    xm = np.array([2.0 * np.sqrt(3.0 / 20.0), 0.0, -2.0 * np.sqrt(3.0 / 20.0)])
    P = np.stack([np.ones(3), xm, 0.5 * (3.0 * xm * xm - 1.0)])      # P[n, m]
    w = np.zeros((3, 3))
    for m in range(3):
        for i in range(3):
            w[i, m] = wm[m] * sum((2 * n + 1) * P[n, m] * f[i, n] for n in range(3))
    return w


def slow_lweights(upsampling, nse_disc=None):
    """Weights of the spectral parameter per effective sample (period = upsampling),
    src/private/fnft__akns_scatter_matrix.c:101-160: 1 (BO), 0.5 (CF4_2), row sums of the weight matrix
    (CF4_3 real; CF5_3, CF6_4 complex)."""
    if nse_disc in (NSE_CF5_3, NSE_CF6_4):
        return cf_complex_weights(nse_disc).sum(axis=1)
    if upsampling == 1:
        return np.array([1.0])
    if upsampling == 2:
        return np.array([0.5])
    return cf4_3_weights().sum(axis=1)


def nsev_contspec(q, T, M, XI, kappa=+1, nse_disc=NSE_2SPLIT4B, cstype=0, normalize=True,
                  evaluate=None):
    """Continuous spectrum of fnft_nsev for the fast discretizations:
    src/fnft_nsev.c:133-314 (driver), :458-542 (base), :744-891 (contspec).
    cstype 0: rho[M]; 1: [a | b]; 2: [rho | a | b].
    evaluate(p, xi) -> p(z(xi)), if given, replaces the chirp-z evaluation (used by
    tests/golden/make_golden.py to evaluate the same polynomials in long double)."""
    q = np.asarray(q, dtype=np.complex128)
    D = q.shape[0]
    scheme = _NSE2AKNS[nse_disc]
    eps_t = (T[1] - T[0]) / (D - 1)
    qp = preprocess_signal(q, eps_t, kappa, nse_disc)
    tm, deg, W = akns_fscatter(qp, -kappa * np.conj(qp), eps_t, scheme, normalize)
    eps_xi = (XI[1] - XI[0]) / (M - 1)
    xi = XI[0] + eps_xi * np.arange(M)
    V = lambda_to_z(eps_xi, eps_t, scheme)     # :822-827
    A = lambda_to_z(-XI[0], eps_t, scheme)
    if evaluate is None:
        H11 = poly_chirpz(tm[0], A, V, M)
        H21 = poly_chirpz(tm[2], A, V, M)
    else:
        H11, H21 = evaluate(tm[0], xi), evaluate(tm[2], xi)
    bc = 0.5
    d1 = akns_degree(scheme)
    extra = eps_t / d1 if nse_disc in (NSE_2SPLIT2A, NSE_2SPLIT2_MODAL) else 0.0
    ph_rho = -2.0 * (T[1] + eps_t * bc) + extra                        # nse_discretization.c:240-256
    ph_a = -eps_t * D + (T[1] + eps_t * bc) - (T[0] - eps_t * bc)      # :263-313
    ph_b = -eps_t * D - (T[1] + eps_t * bc) - (T[0] - eps_t * bc) + extra  # :320-379
    out = []
    if cstype in (0, 2):
        out.append(H21 * np.exp(1j * xi * ph_rho) / H11)               # :846-855
    if cstype in (1, 2):
        scale = 2.0 ** W                                               # :863-876
        out.append(H11 * scale * np.exp(1j * xi * ph_a))
        out.append(H21 * scale * np.exp(1j * xi * ph_b))
    return np.concatenate(out)


def nsev_contspec_slow(q, T, M, XI, kappa=+1, nse_disc=NSE_BO, cstype=0):
    """Continuous spectrum of fnft_nsev for the slow discretizations BO, CF4_2 and CF4_3: one product of D
    step matrices per spectral point (fnft__akns_scatter_matrix, src/private/fnft__akns_scatter_matrix.c:
    112-126,206-232, derivative_flag 0), then the epilogue of src/fnft_nsev.c:836-876 with the phase
    factors of the slow branch (src/private/fnft__nse_discretization.c:240-379)."""
    q = np.asarray(q, dtype=np.complex128)
    D = q.shape[0]
    eps_t = (T[1] - T[0]) / (D - 1)
    qp, rp = preprocess_signal_qr(q, eps_t, kappa, nse_disc)
    lws = slow_lweights(_SLOW_UP.get(nse_disc, 1), nse_disc)
    xi = XI[0] + (XI[1] - XI[0]) / (M - 1) * np.arange(M)
    S11, S12 = np.ones(M, dtype=np.complex128), np.zeros(M, dtype=np.complex128)
    S21, S22 = np.zeros(M, dtype=np.complex128), np.ones(M, dtype=np.complex128)
    if nse_disc in (NSE_ES4, NSE_TES4):
        for n in range(0, qp.shape[0], 3):
            (u11, u12, u21, u22), _ = _es_step(qp[n:n + 3], rp[n:n + 3], xi + 0j, eps_t, nse_disc == NSE_TES4)
            S11, S12, S21, S22 = (u11 * S11 + u12 * S21, u11 * S12 + u12 * S22,
                                  u21 * S11 + u22 * S21, u21 * S12 + u22 * S22)
        qp = qp[:0]
    for n in range(qp.shape[0]):
        l = xi * lws[n % lws.size] + 0j
        (u11, u12, u21, u22), _ = _bo_step(qp[n], rp[n], l, eps_t)
        S11, S12, S21, S22 = (u11 * S11 + u12 * S21, u11 * S12 + u12 * S22,
                              u21 * S11 + u22 * S21, u21 * S12 + u22 * S22)
    bc = 0.5
    ph_rho = -2.0 * (T[1] + eps_t * bc)
    ph_a = (T[1] + eps_t * bc) - (T[0] - eps_t * bc)
    ph_b = -(T[1] + eps_t * bc) - (T[0] - eps_t * bc)
    out = []
    if cstype in (0, 2):
        out.append(S21 * np.exp(1j * xi * ph_rho) / S11)
    if cstype in (1, 2):
        out.append(S11 * np.exp(1j * xi * ph_a))
        out.append(S21 * np.exp(1j * xi * ph_b))
    return np.concatenate(out)


def kdvv(u, T, M, XI, kdv_disc=KDV_2SPLIT4B, evaluate=None):
    """fnft_kdvv, src/fnft_kdvv.c:59-209 (no preprocessing; xi grid negated).
    evaluate: see nsev_contspec."""
    u = np.asarray(u, dtype=np.complex128)
    D = u.shape[0]
    scheme = _KDV2AKNS[kdv_disc]
    deg1 = akns_degree(scheme)
    eps_t = (T[1] - T[0]) / (D - 1)
    eps_xi = (XI[1] - XI[0]) / (M - 1)
    tm, deg, _ = akns_fscatter(u, -np.ones_like(u), eps_t, scheme, normalize=False)
    V = np.exp(-2j * eps_xi * eps_t / deg1)      # :169-170
    A = np.exp(2j * XI[0] * eps_t / deg1)
    xi = -XI[0] - np.arange(M) * eps_xi
    if evaluate is None:
        H12 = poly_chirpz(tm[1], A, V, M)
        H22 = poly_chirpz(tm[3], A, V, M)
    else:
        H12, H22 = evaluate(tm[1], xi), evaluate(tm[3], xi)
    if kdv_disc == KDV_2SPLIT2A:                 # :186-195
        H12 = H12 / np.exp(1j * xi * eps_t / deg1)
    return np.exp(2j * xi * (T[1] + 0.5 * eps_t)) * H12 / (2j * xi * H22 - H12)  # :198-203


# ---------------------------------------------------------------------------------
# bound states
# ---------------------------------------------------------------------------------
def _bo_step(q, r, l, h):
    """One Boffetta-Osborne step and its lambda-derivative,
    src/private/fnft__nse_scatter_bound_states.c:297-322 (vectorised over l)."""
    ks = q * r - l * l
    k = np.sqrt(ks + 0j)
    ch = np.cosh(k * h)
    with np.errstate(divide="ignore", invalid="ignore"):
        sh = np.where(ks != 0, np.sinh(k * h) / np.where(ks != 0, k, 1.0), h)
        chi = ch / ks
        u1 = l * sh * 1j
        ud1 = h * l * l * chi * 1j
        ud2 = l * (h * ch - sh) / ks
        U = (ch - u1, q * sh, r * sh, ch + u1)
        Ud = (ud1 - (l * h + 1j + (l * l * 1j) / ks) * sh, -q * ud2, -r * ud2,
              -ud1 - (l * h - 1j - (l * l * 1j) / ks) * sh)
    return U, Ud


def nse_scatter_bound_states(q, T, lam, upsampling=1, r=None, nse_disc=None, ties=None):
    """fnft__nse_scatter_bound_states for BO (upsampling 1), CF4_2 (upsampling 2) and CF4_3 (3),
    src/private/fnft__nse_scatter_bound_states.c:29-667.  q are the effective
    (preprocessed) samples, r = -conj(q).  Returns (a, aprime, b).
    ties: optional list; receives, per spectral point, the b values of ALL sample points whose error metric is
    within 1e-9 (relative) of the minimum -- for a lambda that is not an eigenvalue of a symmetric potential the
    metric has exact two-way ties and the last bit decides which sample point the reference takes."""
    q = np.asarray(q, dtype=np.complex128)
    lam = np.asarray(lam, dtype=np.complex128)
    D = q.shape[0]
    Dg = D // upsampling
    r = -np.conj(q) if r is None else np.asarray(r, dtype=np.complex128)   # CF5_3 / CF6_4: explicit r
    eps_t = (T[1] - T[0]) / (Dg - 1)
    bc = 0.5
    es = nse_disc in (NSE_ES4, NSE_TES4)
    lws = slow_lweights(upsampling, nse_disc)  # sums of the method weights (:214-221, 232-268)
    scl = 1.0 if es else 1.0 / upsampling      # :225, 235, 247; ES4 / TES4: :132, :157
    K = lam.shape[0]
    PHI = np.zeros((Dg + 1, 2, K), dtype=np.complex128)
    tb = T[0] - eps_t * bc
    phi1 = np.exp(-1j * lam * tb)              # :281-284
    phi2 = np.zeros(K, dtype=np.complex128)
    d1 = phi1 * (-1j * tb)
    d2 = np.zeros(K, dtype=np.complex128)
    PHI[0, 0], PHI[0, 1] = phi1, phi2
    ng = 0
    for n in range(D):                          # :289-338 (ES4 / TES4: one step per grid point, :343-470)
        if es and n % 3 != 0:
            continue
        l = lam * lws[n % lws.size]
        U, Ud = _es_step(q[n:n + 3], r[n:n + 3], lam, eps_t, nse_disc == NSE_TES4, True) if es else \
            _bo_step(q[n], r[n], l, eps_t)
        c = Ud[0] * phi1 + Ud[1] * phi2 + U[0] * d1 + U[1] * d2
        d2 = Ud[2] * phi1 + Ud[3] * phi2 + U[2] * d1 + U[3] * d2
        d1 = c
        c = U[2] * phi1 + U[3] * phi2
        phi1 = U[0] * phi1 + U[1] * phi2
        phi2 = c
        if es or (n + 1) % upsampling == 0:
            ng += 1
            PHI[ng, 0], PHI[ng, 1] = phi1, phi2
    te = T[1] + eps_t * bc
    ex = np.exp(1j * lam * te)
    a = PHI[Dg, 0] * ex                         # :639-640
    ap = scl * (d1 * ex + (1j * te) * a)
    # backward sweep :480-530
    PSI = np.zeros((Dg + 1, 2, K), dtype=np.complex128)
    psi1 = np.zeros(K, dtype=np.complex128)
    psi2 = ex.copy()
    PSI[Dg, 0], PSI[Dg, 1] = psi1, psi2
    ng = Dg
    for n in range(D - 1, -1, -1):
        if es and n % 3 != 0:
            continue
        l = lam * lws[n % lws.size]
        U, _ = _es_step(q[n:n + 3], r[n:n + 3], lam, -eps_t, nse_disc == NSE_TES4) if es else \
            _bo_step(q[n], r[n], l, -eps_t)
        c = U[2] * psi1 + U[3] * psi2
        psi1 = U[0] * psi1 + U[1] * psi2
        psi2 = c
        if n % upsampling == 0:
            ng -= 1
            PSI[ng, 0], PSI[ng, 1] = psi1, psi2
    # choice of b :642-654: first strict minimum of the error metric
    with np.errstate(divide="ignore", invalid="ignore"):
        metric = np.abs(0.5 * np.log(np.abs((PHI[:, 1] / PSI[:, 1]) / (PHI[:, 0] / PSI[:, 0]))))
        ratio = PHI[:, 0] / PSI[:, 0]
    b = np.zeros(K, dtype=np.complex128)
    for k in range(K):
        best = np.inf
        for n in range(Dg + 1):
            if metric[n, k] < best:
                best = metric[n, k]
                b[k] = ratio[n, k]
        if ties is not None:
            ties.append(ratio[np.nonzero(metric[:, k] <= best * (1 + 1e-9))[0], k])
    return a, ap, b


def l2norm2(z, a, b):
    """misc_l2norm2, src/private/fnft__misc.c:90-112."""
    z = np.asarray(z)
    N = z.shape[0]
    h = (b - a) / N
    m = np.abs(z) ** 2
    return 0.5 * h * m[0] + h * m[1:-1].sum() + 0.5 * h * m[-1]


def misc_filter(vals, box):
    """misc_filter, src/private/fnft__misc.c:114-157 (order preserving)."""
    return [v for v in vals if (v.real >= box[0]) and (v.real <= box[1]) and (v.imag >= box[2])
            and (v.imag <= box[3])]


def misc_merge(vals, tol):
    """misc_merge, src/private/fnft__misc.c:228-259 (compares with the ORIGINAL list)."""
    vals = list(vals)
    keep = vals[:1]
    for i in range(1, len(vals)):
        dist = -1.0
        for j in range(i):
            dist = abs(vals[j] - vals[i])
            if dist < tol:
                break
        if dist < tol:
            continue
        keep.append(vals[i])
    return keep


def nsev_bound_states_newton(q, T, guesses, nse_disc=NSE_2SPLIT4B, niter=10, bsfilt=2,
                             dstype=0):
    """Discrete spectrum of fnft_nsev with bsloc_NEWTON for the fast discretizations:
    nsev_compute_boundstates src/fnft_nsev.c:595-741, Newton loop :973-1038, norming
    constants / residues :895-970.  Returns (bound_states, normconsts_or_residues)."""
    q = np.asarray(q, dtype=np.complex128)
    D = q.shape[0]
    eps_t = (T[1] - T[0]) / (D - 1)
    up = {NSE_4SPLIT4B: 2, 20: 2, **_SLOW_UP}.get(nse_disc, 1)
    qp, rp = preprocess_signal_qr(q, eps_t, +1, nse_disc)
    wd = nse_disc if nse_disc in (NSE_CF5_3, NSE_CF6_4, NSE_ES4, NSE_TES4) else None
    deg1 = 0 if nse_disc == NSE_BO or nse_disc in _SLOW_UP else akns_degree(_NSE2AKNS[nse_disc])
    map_coeff = 2.0 / deg1 if deg1 else 2.0     # src/fnft_nsev.c:612-616
    if bsfilt == 2:      # FULL :633-653
        re = 0.9 * np.pi / abs(map_coeff * eps_t)
        qg = qp if up == 1 else up * qp[1::up]
        box = [-re, re, 0.0, 1.5 * 0.25 * l2norm2(qg, T[0], T[1])]
    elif bsfilt == 1:    # BASIC :628-632
        box = [-np.inf, np.inf, 0.0, np.inf]
    else:
        box = [-np.inf, np.inf, -np.inf, np.inf]
    out = []
    for lam in np.asarray(guesses, dtype=np.complex128):
        it = 0
        while niter > 0:  # :1007-1034
            a, ap, _ = nse_scatter_bound_states(qp, T, np.array([lam]), up, rp, wd)
            if a[0] == 0:
                break
            err = a[0] / ap[0]
            lam = lam - err
            it += 1
            if lam.imag > box[3] or lam.real > box[1] or lam.real < box[0] or lam.imag < box[2]:
                break
            if not (abs(err) > 100 * EPS and it < niter):
                break
        out.append(lam)
    if bsfilt != 0:      # :717-724
        out = misc_merge(misc_filter(out, box), np.sqrt(EPS))
    bs = np.array(out, dtype=np.complex128)
    if len(bs) == 0:
        return bs, np.zeros(0, dtype=np.complex128)
    a, ap, b = nse_scatter_bound_states(qp, T, bs, up, rp, wd)
    if dstype == 0:
        nc = b
    elif dstype == 1:
        nc = b / ap
    else:
        nc = np.concatenate([b, b / ap])
    return bs, nc


# ---------------------------------------------------------------------------------
# periodic NFT: the polynomials of the grid search and one window of the search itself
# ---------------------------------------------------------------------------------
def nsep_gridsearch_polys(q, T, kappa=+1, nse_disc=NSE_2SPLIT4B):
    """Polynomials whose roots fnft_nsep's grid search locates (src/fnft_nsep.c:222-436, no phase shift):
    p+ and p- = z^(deg/2) (Delta(z) -/+ 2)-type polynomials of the main spectrum (:318-320, :358) and tm12 of the
    auxiliary spectrum (:399).  Returns (p_plus, p_minus, p_aux, eps_t, deg); eps_t = (T1 - T0)/D (:252)."""
    q = np.asarray(q, dtype=np.complex128)
    D = q.shape[0]
    eps_t = (T[1] - T[0]) / D
    qp = preprocess_signal(q, eps_t, kappa, nse_disc)
    tm, deg, W = akns_fscatter(qp, -kappa * np.conj(qp), eps_t, _NSE2AKNS[nse_disc], True)
    pp = tm[0] + np.conj(tm[0][::-1])
    pp[deg // 2] += 2.0 * 2.0 ** (-W)
    pm = pp.copy()
    pm[deg // 2] -= 4.0 * 2.0 ** (-W)
    return pp, pm, tm[1].copy(), eps_t, deg


def fftgridsearch_window(p, PHI, M, i, evaluate):
    """One step i of poly_roots_fftgridsearch (src/private/fnft__poly_roots_fftgridsearch.c:78-148) with the nine
    polynomial values supplied by evaluate(p, z) (e.g. Horner in long double) instead of the three chirp-z calls
    (:68-75), which evaluate p at exp(i(PHI0 + j eps))/(1 + k eps).  Returns the root estimate or None."""
    ld = np.longdouble
    eps = (ld(PHI[1]) - ld(PHI[0])) / (M - 1)
    ang = ld(PHI[0]) + np.arange(i - 1, i + 2).astype(ld) * eps
    z = np.stack([np.exp(1j * ang.astype(np.clongdouble)) / (1 + k * eps) for k in (-1, 0, 1)])   # [k+1][j-i+1]
    y = evaluate(p, z.reshape(-1)).reshape(3, 3)
    y0 = y[1, 1]
    if (np.abs(y) < np.abs(y0)).any():      # minimum modulus test :84-100
        return None
    z0 = np.exp(1j * np.clongdouble(ang[1]))
    c, tmp = np.clongdouble(0), ld(0)
    for jj, j in enumerate(range(i - 1, i + 2)):
        for k in (-1, 0, 1):
            if j == 0 and k == 0:           # (sic) :112-113
                continue
            zi = (1 - k * eps) * np.exp(1j * np.clongdouble(ang[jj]))
            c += np.conj(zi - z0) * (y[k + 1, jj] - y0)
            tmp += np.abs(zi - z0) ** 2
    c /= tmp
    if c == 0:
        return z0 if y0 == 0 else None
    zr = z0 - y0 / c
    return None if np.abs(zr - z0) > eps else zr


def fftgridsearch_windows(p, PHI, M, idx, evaluate):
    """fftgridsearch_window for an array of steps idx at once (one batched call of evaluate); NaN where the step
    yields no root.  Same arithmetic, src/private/fnft__poly_roots_fftgridsearch.c:78-148."""
    ld = np.longdouble
    idx = np.asarray(idx, dtype=np.int64)
    eps = (ld(PHI[1]) - ld(PHI[0])) / (M - 1)
    ang = ld(PHI[0]) + (idx[:, None] + np.arange(-1, 2)[None, :]).astype(ld) * eps          # [n][jj]
    e = np.exp(1j * ang.astype(np.clongdouble))
    rad = np.array([1 / (1 + k * eps) for k in (-1, 0, 1)], dtype=ld)                     # evaluation radii :68-75
    z = e[:, None, :] * rad[None, :, None]                                                # [n][k+1][jj]
    y = evaluate(p, z.reshape(-1)).reshape(z.shape)
    y0 = y[:, 1, 1]
    ok = ~(np.abs(y) < np.abs(y0)[:, None, None]).any(axis=(1, 2))
    z0 = e[:, 1]
    c = np.zeros(len(idx), dtype=np.clongdouble)
    tmp = np.zeros(len(idx), dtype=ld)
    for jj in range(3):
        for k in (-1, 0, 1):
            skip = ((idx + jj - 1) == 0) & (k == 0)                                        # (sic) :112-113
            zi = (1 - k * eps) * e[:, jj]
            c += np.where(skip, 0, np.conj(zi - z0) * (y[:, k + 1, jj] - y0))
            tmp += np.where(skip, 0, np.abs(zi - z0) ** 2)
    c = c / tmp
    with np.errstate(all="ignore"):
        zr = z0 - y0 / c
    ok &= (c != 0) & (np.abs(zr - z0) <= eps)
    return np.where(ok, zr, np.nan + 0j)


def poly_specfact(poly, oversampling_factor, kappa):
    """poly_specfact, src/private/fnft__poly_specfact.c:25-147 (cepstral spectral factorisation, Dumitrescu B.4):
    log-magnitude on an oversampled grid (:76-108), Hilbert transform including the zeroed bin M/2 - 1 (:116-121),
    exponential, truncation (:130-137).  numpy's FFT at the reference's length kiss_fft_next_fast_size((deg+1) *
    oversampling); pinned to the reference to 1e-15 where that length is a power of two -- at other lengths the
    reference's Kiss FFT (radix-3 / radix-5 butterflies) is itself only accurate to ~1e-12, which the logarithm and
    exponential turn into 2e-11 (tests/test_oracle.py::test_specfact_vs_reference_runs)."""
    poly = np.asarray(poly, dtype=np.complex128)
    deg = poly.shape[0] - 1
    M = next_fast_size((deg + 1) * oversampling_factor)
    buf = np.zeros(M, dtype=np.complex128)
    buf[:deg + 1] = poly
    ab = np.abs(np.fft.fft(buf))
    if kappa == 0:
        x = np.log(ab + 0j)
    else:
        x = 0.5 * np.log(1.0 - kappa * ab * ab + 0j)
    X = np.fft.fft(x)
    X[0] = 0.0
    X[1:M // 2 - 1] *= -1j / M
    X[M // 2 - 1] = 0.0
    X[M // 2:] *= 1j / M
    y = np.fft.ifft(X) * M                      # unnormalised inverse transform
    out = np.fft.ifft(np.exp(x - 1j * y) / M) * M
    return np.conj(out[deg::-1])


def nse_finvscatter(tm, eps_t, kappa, nse_disc):
    """nse_finvscatter, src/private/fnft__nse_finvscatter.c:70-366, for 2SPLIT2A / 2SPLIT2_MODAL.  tm: [4][deg+1],
    highest power first.  The reference splits the samples recursively and multiplies polynomial matrices with FFTs;
    the base case (:157-196) reads Q = -kappa conj(T21(0) / T11(0)), q = atan|Q| e^{i arg Q} / eps_t (2SPLIT2A) or
    Q / eps_t (modal), and the inverse of the one-sample matrix is c [[z, -Q z], [kappa conj(Q), 1]], c = (1 + kappa
    |Q|^2)^(-1/2).  Restated here as plain layer peeling (one sample per step, O(D^2)): the same arithmetic in exact
    terms, only practical for small D.  Returns q[deg]."""
    tm = np.asarray(tm, dtype=np.complex128).reshape(4, -1)
    n = tm.shape[1] - 1
    R = [tm[e, ::-1].copy() for e in range(4)]          # by power of z
    q = np.zeros(n, dtype=np.complex128)
    for k in range(n):
        Q = -kappa * np.conj(R[2][0] / R[0][0])
        den = 1.0 + kappa * abs(Q) ** 2
        if den <= 0:
            raise ValueError("A reconstruced sample violates the condition |q[n]|<1.")   # :172-176
        c = 1.0 / np.sqrt(den)
        q[n - 1 - k] = Q / eps_t if nse_disc == NSE_2SPLIT2_MODAL else np.arctan(abs(Q)) * np.exp(1j * np.angle(Q)) / eps_t
        m = n - k
        new = [c * (R[0][:m] - Q * R[2][:m]), c * (R[1][:m] - Q * R[3][:m]),
               c * (kappa * np.conj(Q) * R[0][1:m + 1] + R[2][1:m + 1]),
               c * (kappa * np.conj(Q) * R[1][1:m + 1] + R[3][1:m + 1])]
        R = new
    return q


def nsev_inverse_pure_solitons(bound_states, normconsts, D, T, residues=False):
    """fnft_nsev_inverse without a continuous spectrum, src/fnft_nsev_inverse.c:680-846: sort by descending imaginary
    part (:741-752), residues -> norming constants (:764-789), then per sample the recursion over rho_k (:808-846)."""
    bs = np.array(bound_states, dtype=np.complex128)
    nc = np.array(normconsts, dtype=np.complex128)
    K = bs.shape[0]
    for i in range(K):                                   # the reference's exchange sort
        for j in range(i + 1, K):
            if bs[i].imag < bs[j].imag:
                bs[i], bs[j] = bs[j], bs[i]
                nc[i], nc[j] = nc[j], nc[i]
    if residues:
        for i in range(K):
            tmp = 1.0 + 0j
            for j in range(K):
                if j != i:
                    tmp = tmp * (bs[i] - bs[j]) / (bs[i] - np.conj(bs[j]))
            nc[i] = (nc[i] / (2j * bs[i].imag)) * tmp
    eps_t = (T[1] - T[0]) / (D - 1)
    t = T[0] + eps_t * np.arange(D)
    zc = int(np.argmax(t >= 0.0)) if (t >= 0.0).any() else 0
    right = np.arange(D) >= zc
    rhok = [np.where(right, nc[i] * np.exp(2j * bs[i] * t), (1 / nc[i]) * np.exp(-2j * bs[i] * t)) for i in range(K)]
    qt = np.zeros(D, dtype=np.complex128)
    for i in range(K):
        rho = rhok[i]
        f = (2j * bs[i].imag) / (1 + np.abs(rho) ** 2)
        qt = qt + 2j * np.conj(rho) * f
        for j in range(i + 1, K):
            rhok[j] = ((bs[j] - bs[i]) * rhok[j] + (rhok[j] - rho) * f) / \
                      (bs[j] - np.conj(bs[i]) - (1 + np.conj(rho) * rhok[j]) * f)
    return np.where(right, qt, np.conj(qt))


def misc_rel_err(num, exact):
    """misc_rel_err, src/private/fnft__misc.c:41-51 -- THE parity metric."""
    num = np.asarray(num)
    exact = np.asarray(exact)
    return np.abs(num - exact).sum() / np.abs(exact).sum()
