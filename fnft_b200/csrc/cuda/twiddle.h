// fnft_b200 -- host-side generation of the twiddle table used by the
// shared-memory FFT passes: tw[k] = exp(-2*pi*i*k/n), computed in long double
// with octant symmetry so that the table is correctly rounded and exactly
// symmetric.  (The reference recomputes sin/cos in double for every plan,
// /root/reference/src/3rd_party/kiss_fft/kiss_fft.c:357-363.)
#pragma once
#include <math.h>
#include <stddef.h>

static inline void fnftb_fill_twiddles(double *tw_re_im /* 2*n doubles */, size_t n)
{
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (size_t k = 0; k < n; ++k) {
        // reduce to the first octant: angle = 2*pi*k/n
        size_t kk = k % n;
        long double c, s;
        // use symmetries around multiples of n/8 when n is divisible by 8
        if (n % 8 == 0) {
            const size_t o = n / 8;
            const size_t oct = kk / o;
            const size_t rem = kk % o;
            long double a;
            long double cc, ss;
            if (oct % 2 == 0) {
                a = two_pi * (long double)rem / (long double)n;
                cc = cosl(a);
                ss = sinl(a);
            } else {
                a = two_pi * (long double)(o - rem) / (long double)n;
                cc = sinl(a);
                ss = cosl(a);
            }
            // (cc, ss) = (cos, sin) of the angle within the quadrant pair
            switch (oct / 2) {
            case 0: c = cc; s = ss; break;
            case 1: c = -ss; s = cc; break;
            case 2: c = -cc; s = -ss; break;
            default: c = ss; s = -cc; break;
            }
        } else {
            const long double a = two_pi * (long double)kk / (long double)n;
            c = cosl(a);
            s = sinl(a);
        }
        tw_re_im[2 * k] = (double)c;
        tw_re_im[2 * k + 1] = (double)(-s);
    }
}
