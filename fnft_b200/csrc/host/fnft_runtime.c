/*
 * fnft_b200 host library -- per-thread GPU context and runtime controls.
 * Each calling thread owns one device context (one device, one stream, grow-only
 * workspaces); contexts are created lazily on first use.
 */
#include "fnft_internal.h"
#include <stdlib.h>
#include <stdio.h>

static __thread fnftb_ctx *tl_ctx = NULL;
static __thread int tl_device = -1; /* -1: take FNFT_B200_DEVICE or 0 on first use */
static __thread int tl_devptr = 0;
static __thread size_t tl_limit = 0;

fnftb_ctx *fnftb__ctx(void)
{
    if (tl_ctx != NULL)
        return tl_ctx;
    int dev = tl_device;
    if (dev < 0) {
        const char *env = getenv("FNFT_B200_DEVICE");
        dev = (env != NULL && env[0] != '\0') ? atoi(env) : 0;
    }
    if (tl_limit == 0) {
        const char *env = getenv("FNFT_B200_WORKSPACE_MB");
        if (env != NULL && env[0] != '\0')
            tl_limit = (size_t)strtoull(env, NULL, 10) << 20;
    }
    if (fnftb_ctx_create(&tl_ctx, dev) != 0) {
        tl_ctx = NULL;
        (void)E_DEVICE;
        return NULL;
    }
    tl_device = dev;
    return tl_ctx;
}

int fnftb__device_pointers(void) { return tl_devptr; }
size_t fnftb__workspace_limit(void) { return tl_limit; }

/* number of chunks the pipelined batch loops aim for (env FNFT_B200_PIPE, 0 = no
 * pipelining, default 8: with the tapered first / last chunks of fnft_nsev.c larger chunks win,
 * measured 56.3 ms per 4096 signals against 57.9 ms with 16) */
int fnftb__pipe_chunks(void)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("FNFT_B200_PIPE");
        v = (e && e[0]) ? atoi(e) : 8;
        if (v < 0)
            v = 0;
    }
    return v;
}

/* Size of the chunk that starts at signal b0 in a pipelined batch loop.  The copy-in of the
 * first chunk and the copy-out of the last one are the only transfers that nothing overlaps,
 * so the sizes taper at both ends: c/4, c/2, c, ..., c, c/2, c/4 (FNFT_B200_PIPE_TAPER=0: off). */
FNFT_UINT fnftb__pipe_step(FNFT_UINT b0, FNFT_UINT B, FNFT_UINT chunk)
{
    static int taper = -1;
    if (taper < 0) {
        const char *e = getenv("FNFT_B200_PIPE_TAPER");
        taper = (e && e[0]) ? atoi(e) : 1;
    }
    FNFT_UINT step = chunk;
    if (!taper || b0 >= B || chunk < 64)
        return step;
    const FNFT_UINT rem = B - b0, q4 = chunk / 4, h2 = chunk / 2;
    if (b0 == 0)
        step = q4;
    else if (b0 == q4)
        step = h2;
    if (rem <= q4)
        step = rem;
    else if (rem <= q4 + h2)
        step = rem - q4;
    else if (step > rem - q4 - h2)
        step = rem - q4 - h2;
    return step;
}

FNFT_INT fnft_b200_device_count(void) { return (FNFT_INT)fnftb_device_count(); }

FNFT_INT fnft_b200_set_device(FNFT_INT device)
{
    if (device < 0 || device >= fnftb_device_count())
        return E_INVALID_ARGUMENT(device);
    if (tl_ctx != NULL && fnftb_ctx_device(tl_ctx) != device) {
        fnftb_ctx_destroy(tl_ctx);
        tl_ctx = NULL;
    }
    tl_device = device;
    return FNFT_SUCCESS;
}

FNFT_INT fnft_b200_set_device_pointers(FNFT_INT flag)
{
    tl_devptr = (flag != 0);
    return FNFT_SUCCESS;
}

FNFT_INT fnft_b200_synchronize(void)
{
    if (tl_ctx == NULL)
        return FNFT_SUCCESS;
    if (fnftb_ctx_sync(tl_ctx) != 0)
        return E_DEVICE;
    return FNFT_SUCCESS;
}

FNFT_INT fnft_b200_set_workspace_limit(FNFT_UINT bytes)
{
    tl_limit = bytes;
    return FNFT_SUCCESS;
}

void *fnft_b200_stream(void)
{
    fnftb_ctx *c = fnftb__ctx();
    return c ? fnftb_ctx_stream(c) : NULL;
}

unsigned long long fnft_b200_launch_count(void) { return fnftb_launch_count(); }

void fnft_b200_profile_enable(FNFT_INT on) { fnftb_profile_enable(on); }

const char *fnft_b200_profile_report(void) { return fnftb_profile_report(); }

void fnft_b200_release(void)
{
    if (tl_ctx != NULL)
        fnftb_ctx_destroy(tl_ctx);
    tl_ctx = NULL;
}

/*
 * ln|z| and arg z.  The chirp-z kernels evaluate W^(n^2/2) as
 * exp((n^2/2)*ln|W|) * cis((n^2/2)*arg W) with n^2/2 up to ~1e11, so ln|z| must
 * resolve |z|-1 far below double precision: x^2+y^2-1 is accumulated exactly with
 * fused multiply-adds before log1p.  (The reference gets the same information from
 * glibc's clog inside cpow, src/private/fnft__poly_chirpz.c:69-95.)
 */
void fnftb__logpolar(FNFT_COMPLEX z, double *ln_abs, double *arg)
{
    const double x = creal(z), y = cimag(z);
    const double px = x * x, ex = fma(x, x, -px);
    const double py = y * y, ey = fma(y, y, -py);
    const long double m2m1 = (((long double)px - 1.0L) + (long double)py) + ((long double)ex + (long double)ey);
    if (fabsl(m2m1) < 0.5L)
        *ln_abs = (double)(0.5L * log1pl(m2m1));
    else
        *ln_abs = (double)(0.5L * logl((long double)px + (long double)py));
    *arg = (double)atan2l((long double)y, (long double)x);
}
