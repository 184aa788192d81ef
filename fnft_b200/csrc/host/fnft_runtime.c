/*
 * fnft_b200 host library -- per-thread GPU context and runtime controls.
 * Each calling thread owns one device context (one device, one stream, grow-only
 * workspaces); contexts are created lazily on first use.
 */
#include "fnft_internal.h"
#include <stdlib.h>
#include <stdio.h>
#include <pthread.h>

static __thread fnftb_ctx *tl_ctx = NULL;
static __thread int tl_device = -1; /* -1: take FNFT_B200_DEVICE or 0 on first use */
static __thread int tl_devptr = 0;
static __thread size_t tl_limit = 0;

/* A thread that exits without calling fnft_b200_release() must not leak its context (device
 * workspaces, streams): a pthread key whose destructor releases it. */
static pthread_key_t ctx_key;
static pthread_once_t ctx_key_once = PTHREAD_ONCE_INIT;
static void ctx_key_dtor(void *p)
{
    if (p != NULL)
        fnftb_ctx_destroy((fnftb_ctx *)p);
}
static void ctx_key_make(void) { (void)pthread_key_create(&ctx_key, ctx_key_dtor); }
static void ctx_key_set(fnftb_ctx *c)
{
    (void)pthread_once(&ctx_key_once, ctx_key_make);
    (void)pthread_setspecific(ctx_key, c);
}

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
    ctx_key_set(tl_ctx);
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
 * so the sizes taper at both ends over `depth` halvings (FNFT_B200_PIPE_TAPER, 0: off, default 2):
 * c/4, c/2, c, ..., c, c/2, c/4 for depth 2. */
FNFT_UINT fnftb__pipe_step(FNFT_UINT b0, FNFT_UINT B, FNFT_UINT chunk)
{
    static int taper = -1;
    if (taper < 0) {
        const char *e = getenv("FNFT_B200_PIPE_TAPER");
        taper = (e && e[0]) ? atoi(e) : 2;
        if (taper == 1)
            taper = 2; /* the round-1 meaning of "on" */
        if (taper > 5)
            taper = 5;
    }
    if (!taper || b0 >= B || chunk < ((FNFT_UINT)16 << taper))
        return chunk;
    /* ramp: chunk >> taper, ..., chunk >> 1; its total length */
    FNFT_UINT ramp = 0;
    for (int k = taper; k >= 1; k--)
        ramp += chunk >> k;
    if (B < 2 * ramp + chunk)
        return chunk; /* too small a batch for both ramps */
    /* inside the leading ramp? */
    FNFT_UINT pos = 0;
    for (int k = taper; k >= 1; k--) {
        if (b0 == pos)
            return chunk >> k;
        pos += chunk >> k;
    }
    /* trailing ramp: the last `ramp` signals, sizes chunk >> 1, ..., chunk >> taper */
    const FNFT_UINT rem = B - b0;
    if (rem <= ramp) {
        FNFT_UINT left = ramp;
        for (int k = 1; k <= taper; k++) {
            if (rem == left)
                return chunk >> k;
            left -= chunk >> k;
        }
        return rem; /* not on a ramp boundary (cannot happen when the loop only uses these steps) */
    }
    /* middle: full chunks, the last one shortened so that the trailing ramp starts exactly */
    const FNFT_UINT mid = rem - ramp;
    return mid < chunk ? mid : chunk;
}

FNFT_INT fnft_b200_device_count(void) { return (FNFT_INT)fnftb_device_count(); }

FNFT_INT fnft_b200_set_device(FNFT_INT device)
{
    if (device < 0 || device >= fnftb_device_count())
        return E_INVALID_ARGUMENT(device);
    if (tl_ctx != NULL && fnftb_ctx_device(tl_ctx) != device) {
        fnftb_ctx_destroy(tl_ctx);
        tl_ctx = NULL;
        ctx_key_set(NULL);
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
    fnftb__fanout_shutdown();
    if (tl_ctx != NULL)
        fnftb_ctx_destroy(tl_ctx);
    tl_ctx = NULL;
    ctx_key_set(NULL);
}

double fnft_b200_probe_fp64_tflops(void)
{
    fnftb_ctx *c = fnftb__ctx();
    return c ? fnftb_probe_fp64_tflops(c) : 0.0;
}

/*
 * ---- several GPUs behind ONE batched call -------------------------------
 * fnft_b200_set_devices(n, ids) (or the environment variable FNFT_B200_DEVICES = "all" | "0,1,..")
 * gives the calling thread a set of devices.  A *_batch call with host buffers then splits its batch
 * into n contiguous shards; shard i runs on device ids[i] in a persistent worker thread that owns
 * that device's context (so workspaces, pinned status arrays and streams survive between calls) and
 * writes straight into the caller's [B][...] arrays.  Signals are independent: there is no
 * exchange between the devices and no collective.
 */
#define FNFTB_MAX_DEV 16
typedef struct {
    pthread_t th;
    int device, index, started;
    struct fnftb_pool *pool;
} fnftb_worker;
typedef struct fnftb_pool {
    int n;
    fnftb_worker w[FNFTB_MAX_DEV];
    pthread_mutex_t mu;
    pthread_cond_t cv_job, cv_done;
    unsigned long generation; /* incremented per job */
    int pending, quit, nshards;
    fnftb_shard_fn fn;
    void *arg;
    fnft_printf_ptr_t printf_ptr;
    size_t limit;
} fnftb_pool;

static __thread int tl_ndev = -1; /* -1: FNFT_B200_DEVICES not looked at yet */
static __thread int tl_devs[FNFTB_MAX_DEV];
/* 1: tl_devs = {d, d} was filled in automatically (two pipelines on the current device, see fnftb__fanout_shards);
 * the user has not configured any devices and fnft_b200_get_devices keeps reporting none */
static __thread int tl_auto = 0;
static __thread fnftb_pool *tl_pool = NULL;
static __thread int tl_in_worker = 0;

static void *worker_main(void *p)
{
    fnftb_worker *w = (fnftb_worker *)p;
    fnftb_pool *pool = w->pool;
    tl_in_worker = 1;
    tl_ndev = 0;
    tl_device = w->device;
    unsigned long seen = 0;
    for (;;) {
        pthread_mutex_lock(&pool->mu);
        while (!pool->quit && pool->generation == seen)
            pthread_cond_wait(&pool->cv_job, &pool->mu);
        if (pool->quit) {
            pthread_mutex_unlock(&pool->mu);
            break;
        }
        seen = pool->generation;
        const fnftb_shard_fn fn = pool->fn;
        void *arg = pool->arg;
        const int nshards = pool->nshards;
        fnft_errwarn_setprintf(pool->printf_ptr); /* messages go where the caller's go */
        tl_limit = pool->limit;
        pthread_mutex_unlock(&pool->mu);
        if (w->index < nshards)
            fn(arg, w->index, nshards);
        pthread_mutex_lock(&pool->mu);
        if (--pool->pending == 0)
            pthread_cond_signal(&pool->cv_done);
        pthread_mutex_unlock(&pool->mu);
    }
    if (tl_ctx != NULL) {
        fnftb_ctx_destroy(tl_ctx);
        tl_ctx = NULL;
        ctx_key_set(NULL);
    }
    return NULL;
}

void fnftb__fanout_shutdown(void)
{
    fnftb_pool *pool = tl_pool;
    if (pool == NULL)
        return;
    pthread_mutex_lock(&pool->mu);
    pool->quit = 1;
    pthread_cond_broadcast(&pool->cv_job);
    pthread_mutex_unlock(&pool->mu);
    for (int i = 0; i < pool->n; i++)
        if (pool->w[i].started)
            pthread_join(pool->w[i].th, NULL);
    pthread_mutex_destroy(&pool->mu);
    pthread_cond_destroy(&pool->cv_job);
    pthread_cond_destroy(&pool->cv_done);
    free(pool);
    tl_pool = NULL;
}

static void devices_from_env(void)
{
    tl_ndev = 0;
    const char *e = getenv("FNFT_B200_DEVICES");
    if (e == NULL || e[0] == '\0')
        return;
    const int have = fnftb_device_count();
    if (strcmp(e, "all") == 0) {
        for (int i = 0; i < have && i < FNFTB_MAX_DEV; i++)
            tl_devs[tl_ndev++] = i;
        return;
    }
    const char *p = e;
    while (*p != '\0' && tl_ndev < FNFTB_MAX_DEV) {
        char *end = NULL;
        const long v = strtol(p, &end, 10);
        if (end == p)
            break;
        if (v >= 0 && v < have)
            tl_devs[tl_ndev++] = (int)v;
        p = (*end == ',') ? end + 1 : end;
    }
}

FNFT_INT fnft_b200_set_devices(FNFT_INT n, FNFT_INT const *devices)
{
    if (n < 0 || n > FNFTB_MAX_DEV)
        return E_INVALID_ARGUMENT(n);
    if (n > 0 && devices == NULL)
        return E_INVALID_ARGUMENT(devices);
    const int have = fnftb_device_count();
    for (FNFT_INT i = 0; i < n; i++) /* an id may repeat: two contexts on one GPU overlap each other's copies */
        if (devices[i] < 0 || devices[i] >= have)
            return E_INVALID_ARGUMENT(devices);
    fnftb__fanout_shutdown(); /* the workers of the previous set go away with their contexts */
    tl_auto = 0;
    tl_ndev = (int)n;
    for (FNFT_INT i = 0; i < n; i++)
        tl_devs[i] = (int)devices[i];
    return FNFT_SUCCESS;
}

FNFT_INT fnft_b200_get_devices(FNFT_INT *devices, FNFT_INT capacity)
{
    if (tl_ndev < 0)
        devices_from_env();
    if (tl_auto)
        return 0;
    for (int i = 0; i < tl_ndev && i < capacity && devices != NULL; i++)
        devices[i] = tl_devs[i];
    return (FNFT_INT)tl_ndev;
}

/* Two pipelines on ONE device.  A large batch given with host buffers runs as chunks whose copies overlap the kernels
 * of the neighbouring chunks (fnftb_pipeline_*); what stays exposed are the tails of the ~10 kernels of every chunk
 * and the first copy-in / last copy-out.  Two contexts on the same GPU, each with one half of the batch, fill each
 * other's gaps: BASELINE config 2 (4096 signals, D = M = 16384) 46.24 -> 44.85 ms per call, measured
 * (profiles/r02_two_contexts.txt).  The mechanism is the fan-out below with the current device listed twice; it is
 * applied automatically when the caller has not configured devices and the batch has at least
 * FNFT_B200_CTX_MIN_BATCH signals (default 2048: config 4, fnft_kdvv with 2048 signals of 8192 samples, 24.65 -> 23.40 ms; FNFT_B200_CTX_PER_DEVICE=1 turns it off). */
static int auto_contexts(FNFT_UINT B)
{
    static int per_dev = -1;
    static long min_batch = 2048;
    if (per_dev < 0) {
        const char *e = getenv("FNFT_B200_CTX_PER_DEVICE");
        per_dev = (e && e[0]) ? atoi(e) : 2;
        e = getenv("FNFT_B200_CTX_MIN_BATCH");
        if (e && e[0])
            min_batch = atol(e);
    }
    return (per_dev >= 2 && min_batch > 0 && B >= (FNFT_UINT)min_batch) ? 2 : 1;
}

/* Number of shards a batch of B signals given with host buffers is split into (1: no fan-out). */
int fnftb__fanout_shards(FNFT_UINT B)
{
    if (tl_in_worker || tl_devptr)
        return 1;
    if (tl_ndev < 0)
        devices_from_env();
    if (tl_ndev == 0 || tl_auto) { /* no devices configured by the caller */
        if (auto_contexts(B) < 2 || fnftb_device_count() < 1)
            return 1;
        if (tl_device < 0) { /* same rule as fnftb__ctx */
            const char *env = getenv("FNFT_B200_DEVICE");
            const int dev = (env != NULL && env[0] != '\0') ? atoi(env) : 0;
            if (dev < 0 || dev >= fnftb_device_count())
                return 1;
            tl_device = dev;
        }
        if (!tl_auto || tl_devs[0] != tl_device) { /* first use, or the thread moved to another device */
            fnftb__fanout_shutdown();
            tl_devs[0] = tl_devs[1] = tl_device;
            tl_ndev = 2;
            tl_auto = 1;
        }
        return 2;
    }
    if (tl_ndev < 2 || B < 2)
        return 1;
    return (B < (FNFT_UINT)tl_ndev) ? (int)B : tl_ndev;
}

void fnftb__shard_range(FNFT_UINT B, int shard, int nshards, FNFT_UINT *b0, FNFT_UINT *b1)
{
    const FNFT_UINT base = B / (FNFT_UINT)nshards, extra = B % (FNFT_UINT)nshards;
    const FNFT_UINT s = (FNFT_UINT)shard;
    *b0 = s * base + (s < extra ? s : extra);
    *b1 = *b0 + base + (s < extra ? 1 : 0);
}

/* Runs fn(arg, i, nshards) for i < nshards, shard i on the worker of device tl_devs[i]; returns when
 * all have finished.  0 on success. */
int fnftb__fanout_run(int nshards, fnftb_shard_fn fn, void *arg)
{
    if (nshards < 1 || nshards > tl_ndev)
        return -1;
    fnftb_pool *pool = tl_pool;
    if (pool == NULL) {
        pool = calloc(1, sizeof(*pool));
        if (pool == NULL)
            return -1;
        pthread_mutex_init(&pool->mu, NULL);
        pthread_cond_init(&pool->cv_job, NULL);
        pthread_cond_init(&pool->cv_done, NULL);
        pool->n = tl_ndev;
        tl_pool = pool;
        for (int i = 0; i < pool->n; i++) {
            pool->w[i].device = tl_devs[i];
            pool->w[i].index = i;
            pool->w[i].pool = pool;
            if (pthread_create(&pool->w[i].th, NULL, worker_main, &pool->w[i]) != 0) {
                fnftb__fanout_shutdown();
                return -1;
            }
            pool->w[i].started = 1;
        }
    }
    /* every worker wakes up for every job; those with index >= nshards have nothing to do */
    pthread_mutex_lock(&pool->mu);
    pool->fn = fn;
    pool->arg = arg;
    pool->nshards = nshards;
    pool->printf_ptr = fnft_errwarn_getprintf();
    pool->limit = tl_limit;
    pool->pending = pool->n;
    pool->generation++;
    pthread_cond_broadcast(&pool->cv_job);
    while (pool->pending > 0)
        pthread_cond_wait(&pool->cv_done, &pool->mu);
    pthread_mutex_unlock(&pool->mu);
    return 0;
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
