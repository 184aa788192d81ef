// fnft_b200 -- shared device/host helpers for the sm_100a kernels.
//
// Every block-level routine in this library is written as a "block program":
// a function of (arguments, block index, #threads, shared-memory pointer) whose
// body is a sequence of FOR_THREADS(tid) { ... } phases separated by
// BLOCK_SYNC().  Compiled by nvcc for the device the phases run on the CTA's
// threads with __syncthreads() between them; compiled with -DFNFTB_EMUL the same
// source runs on the host with the thread loop made explicit, which is how the
// index arithmetic is unit-tested in the (GPU-less) build container.  The
// emulation build is test tooling only -- the product library never contains it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifdef FNFTB_EMUL
#define HD inline
#define DEV inline
#define BLK inline
#define FOR_THREADS(tid, nt) for (int tid = 0; tid < (nt); ++tid)
#define BLOCK_SYNC() ((void)0)
#define TEAM_SYNC(ts, nt) ((void)0)
#define WARP_MAX(v) (v)
#define LDG(p) (*(p))
struct cplx {
    double x, y;
};
static inline cplx make_cplx(double x, double y)
{
    cplx c;
    c.x = x;
    c.y = y;
    return c;
}
struct blk3 {
    unsigned x, y, z;
};
static inline void emul_sincospi(double a, double *s, double *c)
{
    // exact-ish: reduce a mod 2 first (a is a dyadic rational in all callers)
    double r = fmod(a, 2.0);
    *s = sin(M_PI * r);
    *c = cos(M_PI * r);
}
#define SINCOSPI(a, s, c) emul_sincospi((a), (s), (c))
#define SINCOS(a, s, c) sincos((a), (s), (c))
#else
#include <cuda_runtime.h>
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#define BLK __device__ __forceinline__
// Runs the body exactly once with tid = threadIdx.x.
#define FOR_THREADS(tid, nt) for (int tid = threadIdx.x, _once = 1; _once; _once = 0)
#define BLOCK_SYNC() __syncthreads()
#ifndef FNFTB_NO_NAMED_BARRIERS
#define FNFTB_NO_NAMED_BARRIERS 0
#endif
// Barrier among the `ts` consecutive threads of the caller's team (ts a multiple of 32
// dividing nt).  One warp: __syncwarp; the whole CTA: __syncthreads; otherwise a named
// barrier (ids 1..15, so at most 15 teams).
#define TEAM_SYNC(ts, nt)                                                            \
    do {                                                                             \
        if ((ts) == 32) {                                                            \
            __syncwarp();                                                            \
        } else if ((ts) >= (nt) || FNFTB_NO_NAMED_BARRIERS) {                        \
            __syncthreads();                                                         \
        } else {                                                                     \
            asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)threadIdx.x / (ts)), "r"(ts) : "memory"); \
        }                                                                            \
    } while (0)
// max over the lanes of a (fully active) warp; the emulation build keeps per-lane values
#define WARP_MAX(v) fnftb_warp_max(v)
static __device__ __forceinline__ double fnftb_warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#ifdef __CUDA_ARCH__
#define LDG(p) __ldg(p)
#else
#define LDG(p) (*(p))
#endif
typedef double2 cplx;
#define make_cplx(x, y) make_double2((x), (y))
typedef uint3 blk3;
#define SINCOSPI(a, s, c) sincospi((a), (s), (c))
#define SINCOS(a, s, c) sincos((a), (s), (c))
#endif

// ---------------------------------------------------------------------------
// complex arithmetic on double2
// ---------------------------------------------------------------------------
HD cplx cadd(cplx a, cplx b) { return make_cplx(a.x + b.x, a.y + b.y); }
HD cplx csub(cplx a, cplx b) { return make_cplx(a.x - b.x, a.y - b.y); }
HD cplx cmul(cplx a, cplx b) { return make_cplx(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
HD cplx cmulc(cplx a, cplx b) { return make_cplx(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
HD cplx cscale(cplx a, double s) { return make_cplx(a.x * s, a.y * s); }
HD cplx cconj(cplx a) { return make_cplx(a.x, -a.y); }
HD cplx cneg(cplx a) { return make_cplx(-a.x, -a.y); }
// acc += a*b
HD void cfma(cplx &acc, cplx a, cplx b)
{
    acc.x += a.x * b.x - a.y * b.y;
    acc.y += a.x * b.y + a.y * b.x;
}
HD cplx cmuli(cplx a) { return make_cplx(-a.y, a.x); }   // a * i
HD cplx cmulmi(cplx a) { return make_cplx(a.y, -a.x); }  // a * (-i)
HD double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
HD cplx czero() { return make_cplx(0.0, 0.0); }
// robust complex division a/b (Smith)
HD cplx cdiv(cplx a, cplx b)
{
    if (fabs(b.x) >= fabs(b.y)) {
        const double r = b.y / b.x, den = b.x + b.y * r;
        return make_cplx((a.x + a.y * r) / den, (a.y - a.x * r) / den);
    } else {
        const double r = b.x / b.y, den = b.x * r + b.y;
        return make_cplx((a.x * r + a.y) / den, (a.y * r - a.x) / den);
    }
}

// exp(i*pi*a) with exact argument reduction (a is typically m*2/L)
HD cplx cispi(double a)
{
    double s, c;
    SINCOSPI(a, &s, &c);
    return make_cplx(c, s);
}

// ilog2 of a power of two
HD int ilog2i(unsigned v)
{
    int l = 0;
    while ((1u << l) < v)
        ++l;
    return l;
}

// exact floor(log2(x)) for finite x > 0 (normal or subnormal)
HD int floor_log2(double x)
{
    int e;
    (void)frexp(x, &e);  // x = m * 2^e, m in [0.5,1)
    return e - 1;
}

#define FNFTB_MAX_PASSES 8
// Radix plan of an in-shared-memory FFT of length n = prod radix[i].
struct FftPlan {
    int n;
    int log2n;
    int npass;
    int radix[FNFTB_MAX_PASSES];
};
