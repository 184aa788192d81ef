// fnft_b200 -- implementation of the thin C-ABI declared in fnftb_device.h:
// context / workspace management and the kernel pipelines.
#include "fnftb_device.h"

#include "bound_kernels.cuh"
#include "bound_warp.cuh"
#include "chirpz_driver.cuh"
#include "chirpz2.cuh"
#include "nsep_kernels.cuh"
#include "poly_roots.cuh"
#include "nsep_refine.cuh"
#include "slow_scatter.cuh"
#include "resample_kernels.cuh"
#include "tree_driver.cuh"
#include "twiddle.h"

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

std::atomic<unsigned long long> g_fnftb_launch_count{0};
// bit 0: CUDA-event timing of every launch (fnftb_profile_enable), bit 1: NVTX ranges around every launch, named like the
// entries of the profile report (FNFT_B200_NVTX=1; read once at load time, header-only NVTX3: nothing happens unless a
// tool is attached).  The launch sites test the word, fnftb_profile_begin / _end look at the bits.
int g_fnftb_profile_on = [] {
    const char *e = getenv("FNFT_B200_NVTX");
    return (e && e[0] == '1') ? 2 : 0;
}();

// ---------------------------------------------------------------------------
// optional per-launch timing (used by bench.py for the roofline numbers)
// ---------------------------------------------------------------------------
#include <map>
struct ProfRec {
    const char *name;
    cudaEvent_t e0, e1;
};
// per host thread, like the contexts (events belong to the thread's device)
static thread_local std::vector<ProfRec> g_prof;
static thread_local std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event()
{
    if (!g_prof_pool.empty()) {
        cudaEvent_t e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void fnftb_profile_begin(const char *name, cudaStream_t st)
{
    if (g_fnftb_profile_on & 2)
        nvtxRangePushA(name);
    if (!(g_fnftb_profile_on & 1))
        return;
    ProfRec r;
    r.name = name;
    r.e0 = prof_event();
    r.e1 = prof_event();
    cudaEventRecord(r.e0, st);
    g_prof.push_back(r);
}
void fnftb_profile_end(cudaStream_t st)
{
    if (g_fnftb_profile_on & 1)
        cudaEventRecord(g_prof.back().e1, st);
    if (g_fnftb_profile_on & 2)
        nvtxRangePop();
}
// NVTX ranges of the host layer (one per public *_batch call and per phase of a chunk); no-ops without FNFT_B200_NVTX=1
extern "C" void fnftb_range_push(const char *name)
{
    if (g_fnftb_profile_on & 2)
        nvtxRangePushA(name);
}
extern "C" void fnftb_range_pop(void)
{
    if (g_fnftb_profile_on & 2)
        nvtxRangePop();
}

static thread_local std::string g_err;

// optional timeline of the pipelined mode (env FNFT_B200_PIPE_TRACE=1): timed events at the
// start / end of every copy-in, compute and copy-out, printed by fnftb_pipeline_end
struct TraceRec {
    const char *what;
    int chunk;
    cudaEvent_t ev;
};
static thread_local std::vector<TraceRec> g_trace;
static int trace_on()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("FNFT_B200_PIPE_TRACE");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}
static void trace_mark(const char *what, int chunk, cudaStream_t st)
{
    if (!trace_on())
        return;
    TraceRec r;
    r.what = what;
    r.chunk = chunk;
    cudaEventCreate(&r.ev);
    cudaEventRecord(r.ev, st);
    g_trace.push_back(r);
}
static void trace_dump()
{
    if (!trace_on() || g_trace.empty())
        return;
    cudaDeviceSynchronize();
    for (size_t i = 0; i < g_trace.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_trace[0].ev, g_trace[i].ev);
        fprintf(stderr, "[pipe] chunk %2d %-10s %8.3f ms\n", g_trace[i].chunk, g_trace[i].what, ms);
    }
    for (auto &r : g_trace)
        cudaEventDestroy(r.ev);
    g_trace.clear();
}
static thread_local int g_trace_chunk = 0;

static int fail(int code, const char *what, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "%s (code %d) at %s:%d", what, code, file, line);
    g_err = buf;
    return code ? code : -1;
}
#define CU(call)                                                                    \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess)                                                      \
            return fail((int)_e, cudaGetErrorString(_e), __FILE__, __LINE__);       \
    } while (0)
#define RC(call)                                                                    \
    do {                                                                            \
        int _rc = (call);                                                           \
        if (_rc != 0) {                                                             \
            const char *_m = (_rc > 0) ? cudaGetErrorString((cudaError_t)_rc)       \
                                       : "internal error";                          \
            return fail(_rc, _m, __FILE__, __LINE__);                               \
        }                                                                           \
    } while (0)

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

struct fnftb_ctx {
    int device = 0;
    cudaStream_t st = nullptr;
    cplx *tw = nullptr;
    int twn = 4096;
    // staged input
    size_t B = 0, D = 0;
    const cplx *q = nullptr, *r = nullptr;
    Buf qbuf, rbuf, qpre, warn;
    const void *qbuf_host = nullptr;  // host batch that qbuf holds (fnftb_signals_staged), with its shape
    size_t qbuf_B = 0, qbuf_D = 0;
    // tree workspace
    Buf lev0, lev1, mx0, mx1, gbuf, colbuf, W, status, tm, tt0, tt1, twmem;
    TwSet tws;
    TreeDeferred deferred = {};  // pending blk_tree_final (see ctx_finalize)
    // result description
    size_t deg = 0;        // degree of the transfer matrices held in tm
    size_t tmB = 0;        // number of matrices held
    size_t tm_entries = 4; // 4 for matrices, 1 for a standalone polynomial
    // chirp-z workspace
    Buf ybuf, vhat, outbuf, pbuf, cztab;
    // bound-state workspace
    Buf box3, lam, kcnt, flag, aout, apout, bout, phi, koff;
    // nsep workspace
    Buf fpoly, vals, roots, nraw, nkept;
    // root-finder workspace
    Buf rt_roots, rt_absc, rt_lg, rt_hull, rt_info, rt_lam, rt_cnt;
    size_t rt_B = 0, rt_n = 0;  // polynomials / degree of the last root-finder call
    // general-length resampling workspace
    Buf rs_a, rs_b;
    // separate homes for the de-rotated (nsep) and the subsampled signals, and saved selections
    Buf qrot, qsub;
    const cplx *saved_q[2] = {nullptr, nullptr}, *saved_r[2] = {nullptr, nullptr};
    size_t saved_D[2] = {0, 0};
    int have_box3 = 0;
    // slow discretizations with several exponentials per step (set by the CF resampling, reset whenever
    // other signals are staged): weight selector of bo_l_at and, for CF5_3 / CF6_4, the explicit r samples
    int slow_wsel = 0;
    const cplx *rpre = nullptr;
    Buf rprebuf;
    // continuous spectrum by segments (signals longer than one product tree): accumulated and current (a, b)
    Buf segacc, segcur;
    // pipelined host transfers (fnftb_pipeline_begin): two slots, copy streams, events
    int pipe_on = 0, slot = 0;
    cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    int comp_valid[2] = {0, 0}, d2h_valid[2] = {0, 0};
    Buf qslot[2], outslot[2], stslot[2];
    int32_t *st_pinned[2] = {nullptr, nullptr};  // pinned host copies of the per-chunk status
    size_t st_pinned_cap[2] = {0, 0};
    // fully asynchronous mode (fnftb_pipeline_begin with total > 0): one pinned status array for the whole
    // batch, filled chunk by chunk; the host never waits inside the loop
    int32_t *st_all = nullptr;
    size_t st_all_cap = 0, st_all_off = 0, st_all_total = 0;
    // workspace of the inverse transform (inverse_api.cu, its own translation unit) and its destructor
    void *inv_ws = nullptr;
    void (*inv_free)(void *) = nullptr;
};

static TwTable ctx_tw(const fnftb_ctx *c)
{
    TwTable T;
    T.tw = c->tw;
    T.twn = c->twn;
    T.log2twn = 0;
    while ((1 << T.log2twn) < c->twn)
        ++T.log2twn;
    return T;
}

static int ensure(Buf &b, size_t bytes)
{
    if (bytes <= b.cap)
        return 0;
    if (b.p)
        CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    // grow with some slack to avoid repeated reallocations
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        want = bytes;
        CU(cudaMalloc(&b.p, want));
    }
    b.cap = want;
    return 0;
}

static void release(Buf &b)
{
    if (b.p)
        cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

extern "C" {

int fnftb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}

const char *fnftb_last_error(void) { return g_err.c_str(); }

unsigned long long fnftb_launch_count(void) { return g_fnftb_launch_count.load(); }

void fnftb_profile_enable(int on) { g_fnftb_profile_on = (g_fnftb_profile_on & 2) | (on ? 1 : 0); }

// Sums the recorded launches per kernel name into a text report
// "name count total_ms\n..." (buffer owned by the library) and clears the records.
const char *fnftb_profile_report(void)
{
    static thread_local std::string rep;
    std::map<std::string, std::pair<long, double>> acc;
    cudaDeviceSynchronize();
    for (ProfRec &r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            auto &a = acc[r.name];
            a.first += 1;
            a.second += ms;
        }
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof.clear();
    rep.clear();
    char line[256];
    for (auto &kv : acc) {
        snprintf(line, sizeof(line), "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        rep += line;
    }
    return rep.c_str();
}

void fnftb_ctx_destroy(fnftb_ctx *c);

static int ctx_init(fnftb_ctx *c)
{
    CU(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    std::vector<double> tw(2 * (size_t)c->twn);
    fnftb_fill_twiddles(tw.data(), (size_t)c->twn);
    CU(cudaMalloc((void **)&c->tw, sizeof(cplx) * c->twn));
    CU(cudaMemcpy(c->tw, tw.data(), sizeof(cplx) * c->twn, cudaMemcpyHostToDevice));
    // pass-major twiddle tables of the spectrum-carry tree kernels (tw_tables.cuh)
    TwSet tmp;
    const size_t n = twset_layout(&tmp);
    RC(ensure(c->twmem, n * sizeof(cplx)));
    RC(twset_build(&c->tws, (cplx *)c->twmem.p, c->st));
    CU(cudaStreamSynchronize(c->st));
    return 0;
}

int fnftb_ctx_create(fnftb_ctx **out, int device)
{
    if (!out)
        return fail(-2, "null argument", __FILE__, __LINE__);
    *out = nullptr;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev <= 0)
        return fail(-3, "no CUDA device available (the fnft_b200 hot path has no CPU fallback)",
                    __FILE__, __LINE__);
    if (device < 0)
        CU(cudaGetDevice(&device));
    if (device >= ndev)
        return fail(-4, "device index out of range", __FILE__, __LINE__);
    CU(cudaSetDevice(device));
    fnftb_ctx *c = new fnftb_ctx();
    c->device = device;
    const int rc = ctx_init(c);
    if (rc != 0) {  // nothing of a half-built context survives
        fnftb_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

void fnftb_ctx_destroy(fnftb_ctx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    if (c->st)
        cudaStreamSynchronize(c->st);
    Buf *all[] = {&c->qbuf, &c->rbuf, &c->lev0, &c->lev1, &c->mx0, &c->mx1, &c->gbuf, &c->colbuf, &c->W,
                  &c->status, &c->tm, &c->tt0, &c->tt1, &c->twmem, &c->ybuf, &c->vhat, &c->outbuf, &c->pbuf, &c->cztab,
                  &c->qpre, &c->warn, &c->box3, &c->lam, &c->kcnt, &c->flag, &c->aout, &c->apout, &c->bout, &c->phi,
                  &c->fpoly, &c->vals, &c->roots, &c->nraw, &c->nkept,
                  &c->rt_roots, &c->rt_absc, &c->rt_lg, &c->rt_hull, &c->rt_info, &c->rt_lam, &c->rt_cnt, &c->rs_a, &c->rs_b, &c->qrot, &c->qsub,
                  &c->qslot[0], &c->qslot[1], &c->outslot[0], &c->outslot[1], &c->stslot[0], &c->stslot[1],
                  &c->rprebuf, &c->koff, &c->segacc, &c->segcur};
    for (Buf *b : all)
        release(*b);
    if (c->inv_ws && c->inv_free)
        c->inv_free(c->inv_ws);
    c->inv_ws = nullptr;
    if (c->st_h2d)
        cudaStreamDestroy(c->st_h2d);
    if (c->st_d2h)
        cudaStreamDestroy(c->st_d2h);
    for (int i = 0; i < 2; ++i) {
        if (c->st_pinned[i])
            cudaFreeHost(c->st_pinned[i]);
        if (i == 0 && c->st_all)
            cudaFreeHost(c->st_all);
        if (c->ev_h2d[i])
            cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_comp[i])
            cudaEventDestroy(c->ev_comp[i]);
        if (c->ev_d2h[i])
            cudaEventDestroy(c->ev_d2h[i]);
    }
    if (c->tw)
        cudaFree(c->tw);
    if (c->st)
        cudaStreamDestroy(c->st);
    delete c;
}

int fnftb_ctx_device(const fnftb_ctx *c) { return c ? c->device : -1; }

int fnftb_ctx_sync(fnftb_ctx *c)
{
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->st));
    return 0;
}

void *fnftb_ctx_stream(fnftb_ctx *c) { return (void *)c->st; }

// ---------------------------------------------------------------------------
// Pipelined host transfers.  Between begin and end, fnftb_set_signals (host q) copies on
// its own stream into one of two input slots and fnftb_contspec (host out) returns
// without waiting: the device->host copy of the results runs on a third stream.  The
// caller alternates slots 0, 1, 0, ... with its chunks and calls fnftb_pipeline_wait(slot)
// before it reads the outputs / status of the chunk that used that slot.
// ---------------------------------------------------------------------------
int fnftb_pipeline_begin(fnftb_ctx *c, size_t total_signals)
{
    if (!c)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    if (total_signals > c->st_all_cap) {
        if (c->st_all)
            CU(cudaFreeHost(c->st_all));
        c->st_all = nullptr;
        c->st_all_cap = 0;
        CU(cudaMallocHost((void **)&c->st_all, total_signals * sizeof(int32_t)));
        c->st_all_cap = total_signals;
    }
    c->st_all_total = total_signals;
    c->st_all_off = 0;
    if (!c->st_h2d) {
        CU(cudaStreamCreateWithFlags(&c->st_h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->st_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
        }
    }
    CU(cudaStreamSynchronize(c->st));
    c->pipe_on = 1;
    c->slot = 0;
    c->comp_valid[0] = c->comp_valid[1] = 0;
    c->d2h_valid[0] = c->d2h_valid[1] = 0;
    return 0;
}

int fnftb_pipeline_wait(fnftb_ctx *c, int slot, const int32_t **status)
{
    if (!c || slot < 0 || slot > 1)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    if (c->d2h_valid[slot])
        CU(cudaEventSynchronize(c->ev_d2h[slot]));
    if (status)
        *status = c->st_pinned[slot];
    return 0;
}

int fnftb_pipeline_status(fnftb_ctx *c, const int32_t **status)
{
    if (!c || !status)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    *status = c->st_all;
    return 0;
}

int fnftb_pipeline_end(fnftb_ctx *c)
{
    if (!c)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    if (!c->pipe_on)
        return 0;
    c->pipe_on = 0;
    trace_dump();
    g_trace_chunk = 0;
    if (c->st_h2d) {
        CU(cudaStreamSynchronize(c->st_h2d));
        CU(cudaStreamSynchronize(c->st));
        CU(cudaStreamSynchronize(c->st_d2h));
    }
    return 0;
}

static size_t per_signal_bytes(size_t D, int deg0, size_t M, int npoly)
{
    const size_t npad = next_pow2_sz(D);
    const size_t dtree = deg0 <= 2 ? (size_t)deg0 : next_pow2_sz((size_t)deg0);  // tree_leaf_degree
    size_t b = 0;
    b += 2 * tree_lev_elems(1, npad, dtree) * sizeof(cplx);
    b += 2 * npad * sizeof(double);
    b += 2 * tree_gbuf_elems(1, npad, dtree) * sizeof(cplx);
    b += 4 * ((size_t)deg0 * D + 1) * sizeof(cplx);
    if (M > 0) {
        const CzGeom g = cz_geometry((int)((size_t)deg0 * D), (int)M);
        b += (size_t)npoly * g.L * sizeof(cplx);
        b += 3 * M * sizeof(cplx);
    }
    b += 2 * D * sizeof(cplx);
    return b;
}

size_t fnftb_max_chunk(const fnftb_ctx *c, size_t D, int deg0, size_t M, int npoly,
                       size_t budget_bytes)
{
    return fnftb_max_chunk_ex(c, D, deg0, M, npoly, 0, budget_bytes);
}

size_t fnftb_max_chunk_ex(const fnftb_ctx *c, size_t D, int deg0, size_t M, int npoly, size_t extra_per_signal,
                          size_t budget_bytes)
{
    if (budget_bytes == 0) {
        size_t free_b = 0, total_b = 0;
        cudaSetDevice(c->device);
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
            free_b = (size_t)8 << 30;
        budget_bytes = free_b / 2;
        const size_t cap = (size_t)24 << 30;
        if (budget_bytes > cap)
            budget_bytes = cap;
    }
    size_t n = budget_bytes / (per_signal_bytes(D, deg0, M, npoly) + extra_per_signal);
    return n < 1 ? 1 : n;
}

int fnftb_set_signals(fnftb_ctx *c, size_t B, size_t D, const void *q, const void *r, int on_device)
{
    if (!c || !q || B == 0 || D == 0)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    c->B = B;
    c->D = D;
    c->have_box3 = 0;
    c->slow_wsel = 0;
    c->rpre = nullptr;
    if (on_device) {
        c->q = (const cplx *)q;
        c->r = (const cplx *)r;
        return 0;
    }
    const size_t bytes = B * D * sizeof(cplx);
    if (c->pipe_on && !r) {
        const int sl = c->slot;
        RC(ensure(c->qslot[sl], bytes));
        if (c->comp_valid[sl])  // the chunk that used this slot before has finished computing
            CU(cudaStreamWaitEvent(c->st_h2d, c->ev_comp[sl], 0));
        trace_mark("h2d_start", g_trace_chunk, c->st_h2d);
        CU(cudaMemcpyAsync(c->qslot[sl].p, q, bytes, cudaMemcpyHostToDevice, c->st_h2d));
        trace_mark("h2d_end", g_trace_chunk, c->st_h2d);
        CU(cudaEventRecord(c->ev_h2d[sl], c->st_h2d));
        CU(cudaStreamWaitEvent(c->st, c->ev_h2d[sl], 0));
        trace_mark("comp_start", g_trace_chunk, c->st);
        c->q = (const cplx *)c->qslot[sl].p;
        c->r = nullptr;
        return 0;
    }
    RC(ensure(c->qbuf, bytes));
    CU(cudaMemcpyAsync(c->qbuf.p, q, bytes, cudaMemcpyHostToDevice, c->st));
    c->q = (const cplx *)c->qbuf.p;
    c->r = nullptr;
    c->qbuf_host = q;
    c->qbuf_B = B;
    c->qbuf_D = D;
    if (r) {
        RC(ensure(c->rbuf, bytes));
        CU(cudaMemcpyAsync(c->rbuf.p, r, bytes, cudaMemcpyHostToDevice, c->st));
        c->r = (const cplx *)c->rbuf.p;
    }
    return 0;
}

// One shot: 1 when the signal buffer still holds the host batch (q, B, D) that the previous fnftb_set_signals call
// uploaded (the sub-sampling / resampling steps write elsewhere); the batch is then staged again without a second copy.
// fnft_nsev uses it between the two passes of SUBSAMPLE_AND_REFINE over the same chunk.
int fnftb_signals_staged(fnftb_ctx *c, size_t B, size_t D, const void *q)
{
    if (!c)
        return 0;
    const int same = (q != nullptr && c->qbuf_host == q && c->qbuf_B == B && c->qbuf_D == D && c->qbuf.p != nullptr);
    c->qbuf_host = nullptr;
    if (!same || cudaSetDevice(c->device) != cudaSuccess)
        return 0;
    c->B = B;
    c->D = D;
    c->have_box3 = 0;
    c->slow_wsel = 0;
    c->rpre = nullptr;
    c->q = (const cplx *)c->qbuf.p;
    c->r = nullptr;
    return 1;
}

static int ensure_tree(fnftb_ctx *c, size_t B, size_t npad, size_t deg0, size_t deg_out)
{
    RC(ensure(c->lev0, tree_lev_elems(B, npad, deg0) * sizeof(cplx)));
    RC(ensure(c->lev1, tree_lev_elems(B, npad, deg0) * sizeof(cplx)));
    RC(ensure(c->mx0, B * npad * sizeof(double)));
    RC(ensure(c->mx1, B * npad * sizeof(double)));
    RC(ensure(c->gbuf, tree_gbuf_elems(B, npad, deg0) * sizeof(cplx)));
    RC(ensure(c->colbuf, tree_gbuf_elems(B, npad, deg0) * sizeof(cplx)));
    RC(ensure(c->W, B * sizeof(int)));
    RC(ensure(c->status, B * sizeof(int)));
    RC(ensure(c->tm, B * 4 * (deg_out + 1) * sizeof(cplx)));
    // tops of the spectrum path: one per 256 samples (low kernels), or per degree-1024 matrix when a
    // coefficient level is converted (tree_convert.cuh)
    const size_t ntops = B * npad / 256 + B * npad * deg0 / 1024 + 1;
    RC(ensure(c->tt0, ntops * 8 * sizeof(cplx)));
    RC(ensure(c->tt1, ntops * 8 * sizeof(cplx)));
    return 0;
}

static TreeWork tree_work(fnftb_ctx *c)
{
    TreeWork w;
    w.lev[0] = (cplx *)c->lev0.p;
    w.lev[1] = (cplx *)c->lev1.p;
    w.mx[0] = (double *)c->mx0.p;
    w.mx[1] = (double *)c->mx1.p;
    w.gbuf = (cplx *)c->gbuf.p;
    w.colbuf = (cplx *)c->colbuf.p;
    w.W = (int *)c->W.p;
    w.status = (int *)c->status.p;
    w.tt[0] = c->tt0.p;
    w.tt[1] = c->tt1.p;
    w.tws = &c->tws;
    return w;
}

// runs the deferred blk_tree_final, for consumers that need the full transfer matrix
static int ctx_finalize(fnftb_ctx *c)
{
    if (!c->deferred.valid)
        return 0;
    RC(tree_finish_cols(c->deferred, c->st));
    const TreeDeferred &f = c->deferred;
    RC(tree_finalize(tree_work(c), f.cur, f.B, f.d_full, f.deg_out, f.normalize, (cplx *)c->tm.p, c->st,
                     f.sym, f.kappa));
    c->deferred.valid = 0;
    return 0;
}

int fnftb_fscatter(fnftb_ctx *c, const fnftb_scatter_desc *d)
{
    if (!c || !d || !c->q)
        return fail(-2, "invalid argument / no signals staged", __FILE__, __LINE__);
    if (d->rmode == FNFTB_RMODE_EXPLICIT && !c->r)
        return fail(-2, "explicit r requested but not staged", __FILE__, __LINE__);
    if (d->deg0 < 1 || d->deg0 > 105)
        return fail(-5, "discretization not implemented on the GPU path", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    const size_t npad = next_pow2_sz(c->D);
    const size_t deg_out = (size_t)d->deg0 * c->D;
    const size_t dtree = (size_t)tree_leaf_degree(d->scheme, d->deg0);
    // longest pair product: cyclic length dtree*npad <= 2^16 rows of <= 64 * 1024 (tree_driver.cuh)
    // longest product: degree 2^18 (operand length 2^18 at the last level of the spectrum-carry path)
    if (dtree * npad > ((size_t)1 << 18))
        return fail(-6, "signal too long for this build", __FILE__, __LINE__);
    RC(ensure_tree(c, c->B, npad, dtree, deg_out));
    const TwTable T = ctx_tw(c);
    static const int knob_defer = tree_knob("FNFT_B200_DEFER_FINAL", 1);
    RC(tree_fscatter(tree_work(c), c->q, c->r, (int)c->B, (int)c->D, d->deg0, d->rmode, d->kappa,
                     d->scheme, d->eps_t, d->normalize, (cplx *)c->tm.p, T, c->st, 1, FNFTB_TREE_SMEM_N,
                     (knob_defer && d->defer_final) ? &c->deferred : nullptr));
    if (!(knob_defer && d->defer_final))
        c->deferred.valid = c->deferred.cols_pending = 0;
    c->deg = deg_out;
    c->tmB = c->B;
    c->tm_entries = 4;
    return 0;
}

int fnftb_fmult2x2(fnftb_ctx *c, size_t deg, size_t n, const void *p_host, int normalize)
{
    if (!c || !p_host || n == 0 || deg == 0)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    const size_t npad = next_pow2_sz(n);
    if (deg * npad > ((size_t)1 << 22))
        return fail(-6, "product too long for this build", __FILE__, __LINE__);
    RC(ensure_tree(c, 1, npad, deg, deg * n));
    const size_t bytes = 4 * n * (deg + 1) * sizeof(cplx);
    RC(ensure(c->pbuf, bytes));
    CU(cudaMemcpyAsync(c->pbuf.p, p_host, bytes, cudaMemcpyHostToDevice, c->st));
    const TwTable T = ctx_tw(c);
    RC(tree_fmult2x2(tree_work(c), (const cplx *)c->pbuf.p, (int)n, (int)deg, normalize,
                     (cplx *)c->tm.p, T, c->st));
    c->deferred.valid = c->deferred.cols_pending = 0;
    c->deg = deg * n;
    c->tmB = 1;
    c->tm_entries = 4;
    return 0;
}

size_t fnftb_result_degree(const fnftb_ctx *c) { return c ? c->deg : 0; }

int fnftb_get_transfer_matrix(fnftb_ctx *c, void *tm_host, int32_t *W_host)
{
    if (!c || c->tmB == 0)
        return fail(-2, "no result held", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ctx_finalize(c));
    if (tm_host)
        CU(cudaMemcpyAsync(tm_host, c->tm.p, c->tmB * c->tm_entries * (c->deg + 1) * sizeof(cplx),
                           cudaMemcpyDeviceToHost, c->st));
    if (W_host)
        CU(cudaMemcpyAsync(W_host, c->W.p, c->tmB * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return 0;
}

int fnftb_get_status(fnftb_ctx *c, int32_t *status_host)
{
    if (!c || !status_host || c->tmB == 0)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(status_host, c->status.p, c->tmB * sizeof(int), cudaMemcpyDeviceToHost,
                       c->st));
    CU(cudaStreamSynchronize(c->st));
    return 0;
}

int fnftb_set_polynomial(fnftb_ctx *c, size_t deg, const void *p_host)
{
    if (!c || !p_host)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ensure(c->tm, (deg + 1) * sizeof(cplx)));
    RC(ensure(c->W, sizeof(int)));
    RC(ensure(c->status, sizeof(int)));
    CU(cudaMemcpyAsync(c->tm.p, p_host, (deg + 1) * sizeof(cplx), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemsetAsync(c->W.p, 0, sizeof(int), c->st));
    CU(cudaMemsetAsync(c->status.p, 0, sizeof(int), c->st));
    c->deferred.valid = c->deferred.cols_pending = 0;
    c->deg = deg;
    c->tmB = 1;
    c->tm_entries = 1;
    return 0;
}

int fnftb_contspec(fnftb_ctx *c, const fnftb_contspec_desc *d, void *out, size_t out_sstride,
                   int on_device, int32_t *status_host)
{
    if (!c || !d || !out || c->tmB == 0 || d->M == 0 || d->npoly < 1 || d->npoly > 2)
        return fail(-2, "invalid argument / no transfer matrix held", __FILE__, __LINE__);
    for (int j = 0; j < d->npoly; ++j)
        if (d->ent[j] < 0 || (size_t)d->ent[j] >= c->tm_entries)
            return fail(-2, "entry index out of range", __FILE__, __LINE__);
    if (c->deg + d->M > ((size_t)1 << 24))
        return fail(-6, "chirp-z too long for this build", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    const size_t B = c->tmB;
    const CzGeom g = cz_geometry((int)c->deg, (int)d->M);
    RC(ensure(c->ybuf, cz_ybuf_elems(g, B, d->npoly) * sizeof(cplx)));
    RC(ensure(c->vhat, (size_t)g.L * sizeof(cplx)));
    RC(ensure(c->cztab, cz_table_elems(g, (int)c->deg, (int)d->M) * sizeof(cplx)));
    cplx *dst = (cplx *)out;
    const bool piped = c->pipe_on && !on_device;
    const int sl = c->slot;
    if (piped) {
        RC(ensure(c->outslot[sl], B * out_sstride * sizeof(cplx)));
        RC(ensure(c->stslot[sl], B * sizeof(int)));
        if (c->st_pinned_cap[sl] < B) {  // a pageable destination would make the copy synchronous
            if (c->st_pinned[sl])
                CU(cudaFreeHost(c->st_pinned[sl]));
            c->st_pinned[sl] = nullptr;
            c->st_pinned_cap[sl] = 0;
            CU(cudaMallocHost((void **)&c->st_pinned[sl], B * sizeof(int32_t)));
            c->st_pinned_cap[sl] = B;
        }
        if (c->d2h_valid[sl])  // results of the previous user of this slot have left the device
            CU(cudaStreamWaitEvent(c->st, c->ev_d2h[sl], 0));
        dst = (cplx *)c->outslot[sl].p;
    } else if (!on_device) {
        RC(ensure(c->outbuf, B * out_sstride * sizeof(cplx)));
        dst = (cplx *)c->outbuf.p;
    }
    CzArgs a;
    memset(&a, 0, sizeof(a));
    a.tm = (const cplx *)c->tm.p;
    a.tm_sstride = c->tm_entries * (c->deg + 1);
    a.ent[0] = d->ent[0];
    a.ent[1] = d->ent[1];
    a.npoly = d->npoly;
    a.deg = (int)c->deg;
    a.B = (int)B;
    a.M = (int)d->M;
    a.lwr = d->lwr;
    a.lwi = d->lwi;
    a.lar = d->lar;
    a.lai = d->lai;
    a.ybuf = (cplx *)c->ybuf.p;
    a.vhat = (cplx *)c->vhat.p;
    a.T = ctx_tw(c);
    a.mode = d->mode;
    a.cstype = d->cstype;
    a.out = dst;
    a.out_sstride = out_sstride;
    a.W = (const int *)c->W.p;
    a.xi0 = d->xi0;
    a.eps_xi = d->eps_xi;
    a.ph_rho = d->ph_rho;
    a.ph_a = d->ph_a;
    a.ph_b = d->ph_b;
    a.kdv_ph = d->kdv_ph;
    a.kdv_sqrtz = d->kdv_sqrtz;
    a.status = (int *)c->status.p;
    {
        static const int knob_cz2 = tree_knob("FNFT_B200_CZ2", 1);
        const bool fast = knob_cz2 && cz2_supported((int)c->deg, (int)d->M);
        Cz2SymSrc src;
        memset(&src, 0, sizeof(src));
        if (c->deferred.valid) {
            const TreeDeferred &f = c->deferred;
            if (fast && f.sym && d->mode == FNFTB_MODE_NSEV && d->npoly == 2 && d->ent[0] == 0 && d->ent[1] == 2) {
                // polynomials straight from the level buffer: H11 = a, H21 = -kappa * b#
                const TreeWork w = tree_work(c);
                src.lev = w.lev[f.cur];
                src.mx = w.mx[f.cur];
                src.W = (int *)c->W.p;
                src.d_full = f.d_full;
                src.kappa = f.kappa;
                src.normalize = f.normalize;
                if (f.cols_pending) {
                    // the last column pass of the tree and the first chirp-z stage in one kernel, when that
                    // combination of radices exists; otherwise the column pass runs now
                    if (cz2_fused_cols_supported(f.cols, (int)c->deg, (int)d->M, f.d_full)) {
                        src.fused = 1;
                        src.up = f.cols;
                        c->deferred.cols_pending = 0;
                    } else {
                        RC(tree_finish_cols(c->deferred, c->st));
                    }
                }
                c->deferred.valid = 0;  // the exponent of the last matrix is added to W by the kernel
            } else {
                RC(ctx_finalize(c));
            }
        }
        if (fast)
            RC(cz2_run(a, (cplx *)c->cztab.p, c->tws, c->st, src.lev ? &src : nullptr));
        else
            RC(cz_run_fastrows(a, (cplx *)c->cztab.p, c->tws, c->st));
    }
    if (piped) {
        CU(cudaMemcpyAsync(c->stslot[sl].p, c->status.p, B * sizeof(int), cudaMemcpyDeviceToDevice, c->st));
        CU(cudaEventRecord(c->ev_comp[sl], c->st));
        trace_mark("comp_end", g_trace_chunk, c->st);
        c->comp_valid[sl] = 1;
        CU(cudaStreamWaitEvent(c->st_d2h, c->ev_comp[sl], 0));
        trace_mark("d2h_start", g_trace_chunk, c->st_d2h);
        CU(cudaMemcpyAsync(out, dst, B * out_sstride * sizeof(cplx), cudaMemcpyDeviceToHost, c->st_d2h));
        if (c->st_all_total > 0 && c->st_all_off + B <= c->st_all_total) {
            CU(cudaMemcpyAsync(c->st_all + c->st_all_off, c->stslot[sl].p, B * sizeof(int), cudaMemcpyDeviceToHost,
                               c->st_d2h));
            c->st_all_off += B;
        } else {
            CU(cudaMemcpyAsync(c->st_pinned[sl], c->stslot[sl].p, B * sizeof(int), cudaMemcpyDeviceToHost,
                               c->st_d2h));
        }
        CU(cudaEventRecord(c->ev_d2h[sl], c->st_d2h));
        trace_mark("d2h_end", g_trace_chunk++, c->st_d2h);
        c->d2h_valid[sl] = 1;
        c->slot ^= 1;
        return 0;
    }
    if (!on_device) {
        CU(cudaMemcpyAsync(out, dst, B * out_sstride * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
    }
    if (status_host)
        CU(cudaMemcpyAsync(status_host, c->status.p, B * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    if (!on_device || status_host)
        CU(cudaStreamSynchronize(c->st));
    return 0;
}


int fnftb_slow_contspec(fnftb_ctx *c, const fnftb_contspec_desc *d, int upsampling, int kappa, double eps_t,
                        void *out, size_t out_sstride, int on_device, int32_t *status_host)
{
    if (!c || !d || !out || !c->q || d->M == 0 || upsampling < 1 || upsampling > 4)
        return fail(-2, "invalid argument / no signals staged", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    const size_t B = c->B;
    RC(ensure(c->status, B * sizeof(int)));
    CU(cudaMemsetAsync(c->status.p, 0, B * sizeof(int), c->st));
    cplx *dst = (cplx *)out;
    if (!on_device) {
        RC(ensure(c->outbuf, B * out_sstride * sizeof(cplx)));
        dst = (cplx *)c->outbuf.p;
    }
    SlowCsArgs a;
    memset(&a, 0, sizeof(a));
    a.q = c->q;
    a.r = c->rpre;
    a.wsel = c->slow_wsel;
    a.B = (int)B;
    a.D = (int)c->D;
    a.upsampling = upsampling;
    a.kappa = kappa;
    a.M = (int)d->M;
    a.cstype = d->cstype;
    a.eps_t = eps_t;
    a.lweight = (upsampling == 2) ? 0.5 : 1.0;  // CF4_3: per-sample weights, see bo_l_at
    a.xi0 = d->xi0;
    a.eps_xi = d->eps_xi;
    a.ph_rho = d->ph_rho;
    a.ph_a = d->ph_a;
    a.ph_b = d->ph_b;
    a.out = dst;
    a.out_sstride = out_sstride;
    a.status = (int *)c->status.p;
    const size_t n = B * (size_t)d->M;
    if (g_fnftb_profile_on)
        fnftb_profile_begin("slow_contspec", c->st);
    k_slow_contspec<<<(unsigned)((n + 3) / 4), 128, 0, c->st>>>(a);
    if (g_fnftb_profile_on)
        fnftb_profile_end(c->st);
    ++g_fnftb_launch_count;
    CU(cudaGetLastError());
    if (!on_device)
        CU(cudaMemcpyAsync(out, dst, B * out_sstride * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
    if (status_host)
        CU(cudaMemcpyAsync(status_host, c->status.p, B * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    if (!on_device || status_host)
        CU(cudaStreamSynchronize(c->st));
    return 0;
}

// ---------------------------------------------------------------------------
// periodic NFT: grid search
// ---------------------------------------------------------------------------
int fnftb_nsep_derotate(fnftb_ctx *c, double lam_shift, double T0, double eps_t)
{
    if (!c || !c->q)
        return fail(-2, "no signals staged", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ensure(c->qrot, c->B * c->D * sizeof(cplx)));  // (not qpre: the resampling step writes there)
    DerotArgs da;
    da.q = c->q;
    da.out = (cplx *)c->qrot.p;
    da.B = (int)c->B;
    da.D = (int)c->D;
    da.lam_shift = lam_shift;
    da.T0 = T0;
    da.eps_t = eps_t;
    const long long tot = (long long)c->B * c->D;
    RC((launch_blocks<DerotArgs, blk_nsep_derotate>(da, (unsigned)((tot + 255) / 256), 256, 0, c->st,
                                                     "nsep_derotate")));
    c->q = (const cplx *)c->qrot.p;
    return 0;
}

int fnftb_signals_save(fnftb_ctx *c, int slot)
{
    if (!c || !c->q || slot < 0 || slot > 1)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    c->saved_q[slot] = c->q;
    c->saved_r[slot] = c->r;
    c->saved_D[slot] = c->D;
    return 0;
}

int fnftb_signals_restore(fnftb_ctx *c, int slot)
{
    if (!c || slot < 0 || slot > 1 || !c->saved_q[slot])
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    c->q = c->saved_q[slot];
    c->r = c->saved_r[slot];
    c->D = c->saved_D[slot];
    c->have_box3 = 0;
    return 0;
}

static int roots_run(fnftb_ctx *c, const cplx *coef, long long cstride, size_t B, size_t n, void *roots_host,
                     int32_t *info_host);

int fnftb_nsep_floquet_roots(fnftb_ctx *c, double rhs, void *roots_host, int32_t *info_host)
{
    if (!c || c->tmB == 0 || c->deg < 2 || c->tm_entries != 4)
        return fail(-2, "invalid argument / no transfer matrix held", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ctx_finalize(c));
    const size_t B = c->tmB, d1 = c->deg + 1;
    RC(ensure(c->fpoly, B * 2 * d1 * sizeof(cplx)));
    FloquetRhsArgs fa;
    fa.tm = (const cplx *)c->tm.p;
    fa.W = (const int *)c->W.p;
    fa.P = (cplx *)c->fpoly.p;
    fa.B = (int)B;
    fa.deg = (int)c->deg;
    fa.rhs = rhs;
    const long long tot = (long long)B * (long long)d1;
    k_nsep_floquet_rhs<<<(unsigned)((tot + 255) / 256), 256, 0, c->st>>>(fa);
    ++g_fnftb_launch_count;
    CU(cudaGetLastError());
    return roots_run(c, (const cplx *)c->fpoly.p, (long long)d1, B, c->deg, roots_host, info_host);
}

int fnftb_nsep_refine(fnftb_ctx *c, const fnftb_refine_desc *d, const int32_t *K_host, void *lam_host,
                      int32_t *flag_host)
{
    if (!c || !d || !c->q || !K_host || !lam_host || d->Kstride < 1 || (d->upsampling != 1 && d->upsampling != 2))
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    const size_t n = c->B * (size_t)d->Kstride;
    RC(ensure(c->lam, n * sizeof(cplx)));
    RC(ensure(c->kcnt, c->B * sizeof(int)));
    RC(ensure(c->flag, n * sizeof(int)));
    CU(cudaMemcpyAsync(c->lam.p, lam_host, n * sizeof(cplx), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->kcnt.p, K_host, c->B * sizeof(int), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemsetAsync(c->flag.p, 0, n * sizeof(int), c->st));
    NsepRefineArgs a;
    memset(&a, 0, sizeof(a));
    a.q = c->q;
    a.B = (int)c->B;
    a.D = (int)c->D;
    a.upsampling = d->upsampling;
    a.kappa = d->kappa;
    a.Kstride = d->Kstride;
    a.K = (const int *)c->kcnt.p;
    a.lam = (cplx *)c->lam.p;
    a.flag = (int *)c->flag.p;
    a.eps_t = d->eps_t;
    a.lweight = (d->upsampling == 2) ? 0.5 : 1.0;  // sum of the CF4_2 weights of a node
    a.scl = (d->upsampling == 2) ? 0.5 : 1.0;      // fnft__akns_scatter_matrix.c:113,121
    a.rhs = d->rhs;
    a.tol = d->tol;
    a.max_evals = d->max_evals;
    a.mode = d->mode;
    if (g_fnftb_profile_on)
        fnftb_profile_begin("nsep_refine", c->st);
    k_nsep_refine<<<(unsigned)((n + 3) / 4), 128, 0, c->st>>>(a);
    if (g_fnftb_profile_on)
        fnftb_profile_end(c->st);
    ++g_fnftb_launch_count;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(lam_host, c->lam.p, n * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
    if (flag_host)
        CU(cudaMemcpyAsync(flag_host, c->flag.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return 0;
}

// ln|z| and arg z of a rounded complex double, resolving |z|-1 exactly (same
// construction as fnftb__logpolar in the host library)
static void logpolar_ld(double x, double y, double *ln_abs, double *arg)
{
    const double px = x * x, ex = fma(x, x, -px);
    const double py = y * y, ey = fma(y, y, -py);
    const long double m2m1 = (((long double)px - 1.0L) + (long double)py) + ((long double)ex + (long double)ey);
    if (fabsl(m2m1) < 0.5L)
        *ln_abs = (double)(0.5L * log1pl(m2m1));
    else
        *ln_abs = (double)(0.5L * logl((long double)px + (long double)py));
    *arg = (double)atan2l((long double)y, (long double)x);
}

// bytes of device workspace the grid search needs per signal
static size_t nsep_bytes_per_signal(size_t deg, size_t Mpts)
{
    const CzGeom g = cz_geometry((int)deg, (int)Mpts);
    return (2 * (size_t)g.L + 2 * 3 * Mpts + 2 * (deg + 1) + 3 * (deg + 1)) * sizeof(cplx);
}

size_t fnftb_nsep_chunk(const fnftb_ctx *c, size_t D_eff, int deg0, size_t budget_bytes)
{
    if (budget_bytes == 0) {
        size_t free_b = 0, total_b = 0;
        cudaSetDevice(c->device);
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
            free_b = (size_t)8 << 30;
        budget_bytes = free_b / 2;
        const size_t cap = (size_t)16 << 30;
        if (budget_bytes > cap)
            budget_bytes = cap;
    }
    const size_t deg = (size_t)deg0 * D_eff;
    const size_t per = nsep_bytes_per_signal(deg, 32 * deg) + per_signal_bytes(D_eff, deg0, 0, 0);
    size_t n = budget_bytes / per;
    return n < 1 ? 1 : n;
}

// Requires fnftb_fscatter to have run on the staged signals.  For every signal:
// roots of p+ and p- (main spectrum, concatenated, at most Kmax) and of tm12
// (auxiliary spectrum, at most Mmax), mapped to lambda, filtered, in scan order.
// status_host[b]: 0 ok, 1 = more roots than the polynomial degree (src/fnft_nsep.c:329-332),
// bit 4 set = main spectrum truncated, bit 5 set = auxiliary spectrum truncated.
int fnftb_nsep_gridsearch(fnftb_ctx *c, const fnftb_nsep_desc *d, uint64_t *K_host, void *main_host,
                          uint64_t *M_host, void *aux_host, int32_t *status_host)
{
    if (!c || !d || c->tmB == 0 || c->tm_entries != 4)
        return fail(-2, "invalid argument / no transfer matrix held", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ctx_finalize(c));
    const size_t B = c->tmB, deg = c->deg, d1 = deg + 1;
    if (deg < 2)
        return fail(-2, "degree too small", __FILE__, __LINE__);
    const size_t Mpts = 32 * deg;  // oversampling_factor, src/fnft_nsep.c:43
    if (deg + Mpts > ((size_t)1 << 24))
        return fail(-6, "grid search too long for this build", __FILE__, __LINE__);
    const CzGeom g = cz_geometry((int)deg, (int)Mpts);
    const int want_main = (main_host != nullptr), want_aux = (aux_host != nullptr);
    RC(ensure(c->fpoly, B * 2 * d1 * sizeof(cplx)));
    RC(ensure(c->ybuf, B * 2 * (size_t)g.L * sizeof(cplx)));
    RC(ensure(c->vhat, (size_t)g.L * sizeof(cplx)));
    RC(ensure(c->vals, B * 2 * 3 * Mpts * sizeof(cplx)));
    RC(ensure(c->cztab, cz_table_elems(g, (int)deg, (int)Mpts) * sizeof(cplx)));
    const size_t cap = deg;  // more than deg roots is an error anyway
    RC(ensure(c->roots, B * 2 * cap * sizeof(cplx)));
    RC(ensure(c->nraw, B * 2 * sizeof(int)));
    RC(ensure(c->nkept, B * 2 * sizeof(int)));
    std::vector<int> nraw(B * 2), nkept(B * 2);
    std::vector<cplx> roots;
    for (size_t b = 0; b < B; ++b)
        status_host[b] = 0;

    // FNFT_B200_NSEP_TIMING=1: host wall time of the phases (each closed by a stream synchronisation)
    static const int knob_timing = tree_knob("FNFT_B200_NSEP_TIMING", 0);
    struct timespec t_last;
    clock_gettime(CLOCK_MONOTONIC, &t_last);
    auto tick = [&](const char *label) {
        if (!knob_timing)
            return;
        cudaStreamSynchronize(c->st);
        struct timespec now;
        clock_gettime(CLOCK_MONOTONIC, &now);
        fprintf(stderr, "[nsep timing]     %-24s %8.3f ms\n", label,
                (now.tv_sec - t_last.tv_sec) * 1e3 + (now.tv_nsec - t_last.tv_nsec) * 1e-6);
        t_last = now;
    };
    const double eps = (d->PHI1 - d->PHI0) / (double)(Mpts - 1);
    for (int pass = 0; pass < 2; ++pass) {  // 0: main (p+, p-), 1: aux (tm12)
        if ((pass == 0 && !want_main) || (pass == 1 && !want_aux))
            continue;
        const int npoly = (pass == 0) ? 2 : 1;
        if (pass == 0) {
            FloquetArgs fa;
            fa.tm = (const cplx *)c->tm.p;
            fa.W = (const int *)c->W.p;
            fa.P = (cplx *)c->fpoly.p;
            fa.B = (int)B;
            fa.deg = (int)deg;
            const long long tot = (long long)B * d1;
            RC((launch_blocks<FloquetArgs, blk_nsep_polys>(fa, (unsigned)((tot + 255) / 256), 256, 0,
                                                          c->st, "nsep_polys")));
        }
        for (int k = -1; k <= 1; ++k) {  // three rings, fftgridsearch.c:70-77
            CzArgs a;
            memset(&a, 0, sizeof(a));
            if (pass == 0) {
                a.tm = (const cplx *)c->fpoly.p;
                a.tm_sstride = 2 * d1;
                a.ent[0] = 0;
                a.ent[1] = 1;
            } else {
                a.tm = (const cplx *)c->tm.p;
                a.tm_sstride = 4 * d1;
                a.ent[0] = 1;  // tm12
            }
            a.npoly = npoly;
            a.deg = (int)deg;
            a.B = (int)B;
            a.M = (int)Mpts;
            // W = CEXP(I*eps), A = (1 + k*eps)*CEXP(-I*PHI0) formed as rounded complex
            // doubles like the reference (fftgridsearch.c:68-73), then taken apart
            logpolar_ld(cos(eps), sin(eps), &a.lwr, &a.lwi);
            {
                const double rad = 1.0 + k * eps;
                logpolar_ld(rad * cos(-d->PHI0), rad * sin(-d->PHI0), &a.lar, &a.lai);
            }
            a.ybuf = (cplx *)c->ybuf.p;
            a.vhat = (cplx *)c->vhat.p;
            a.T = ctx_tw(c);
            a.mode = FNFTB_CZ_RAW;
            a.out = (cplx *)c->vals.p + (size_t)(k + 1) * Mpts;
            a.out_sstride = (size_t)npoly * 3 * Mpts;
            a.out_jstride = 3 * Mpts;
            a.status = (int *)c->status.p;
            RC(cz_run_fastrows(a, (cplx *)c->cztab.p, c->tws, c->st));
        }
        tick("polys + 3 chirp-z rings");
        ScanArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.vals = (const cplx *)c->vals.p;
        sa.B = (int)B;
        sa.npoly = npoly;
        sa.M = (int)Mpts;
        sa.PHI0 = d->PHI0;
        sa.eps = eps;
        sa.lam_den = d->lam_den;
        sa.filtering = d->filtering;
        for (int i = 0; i < 4; ++i)
            sa.box[i] = d->box[i];
        sa.lam_shift = d->lam_shift;
        sa.cap = (int)cap;
        sa.out = (cplx *)c->roots.p;
        sa.n_raw = (int *)c->nraw.p;
        sa.n_kept = (int *)c->nkept.p;
        static const int knob_scan = tree_knob("FNFT_B200_NSEP_SCAN_TILED", 1);  // 0: blk_nsep_scan (round 1)
        if (knob_scan) {
            if (g_fnftb_profile_on)
                fnftb_profile_begin("nsep_scan", c->st);
            k_nsep_scan_tiled<512><<<(unsigned)(B * npoly), 512, 0, c->st>>>(sa);
            if (g_fnftb_profile_on)
                fnftb_profile_end(c->st);
            ++g_fnftb_launch_count;
            CU(cudaGetLastError());
        } else {
            const int nt = 256;
            RC((launch_blocks<ScanArgs, blk_nsep_scan>(sa, (unsigned)(B * npoly), nt, 2 * nt * sizeof(int),
                                                       c->st, "nsep_scan")));
        }
        CU(cudaMemcpyAsync(nraw.data(), c->nraw.p, B * npoly * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CU(cudaMemcpyAsync(nkept.data(), c->nkept.p, B * npoly * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        tick("scan + counts");
        // copy only as many values per row as the fullest row holds (typically tens of the `cap` slots)
        size_t maxn = 0;
        for (size_t i = 0; i < B * npoly; ++i) {
            const size_t n = (size_t)nkept[i] > cap ? cap : (size_t)nkept[i];
            if (n > maxn)
                maxn = n;
        }
        roots.resize(B * npoly * maxn + 1);  // compact host copy: row pitch maxn
        if (maxn > 0) {
            CU(cudaMemcpy2DAsync(roots.data(), maxn * sizeof(cplx), c->roots.p, cap * sizeof(cplx),
                                 maxn * sizeof(cplx), B * npoly, cudaMemcpyDeviceToHost, c->st));
            CU(cudaStreamSynchronize(c->st));
        }
        tick("copy of the kept values");
        for (size_t b = 0; b < B; ++b) {
            if (pass == 0) {
                cplx *dst = (cplx *)main_host + b * d->Kmax;
                size_t K = 0;
                for (int j = 0; j < 2; ++j) {
                    if ((size_t)nraw[b * 2 + j] > deg) {
                        status_host[b] = 1;
                        break;
                    }
                    size_t n = (size_t)nkept[b * 2 + j];
                    if (n > cap)
                        n = cap;
                    if (K + n > d->Kmax) {
                        status_host[b] |= 16;
                        n = (K < d->Kmax) ? d->Kmax - K : 0;
                    }
                    memcpy(dst + K, roots.data() + (b * 2 + j) * maxn, n * sizeof(cplx));
                    K += n;
                }
                K_host[b] = (status_host[b] == 1) ? 0 : K;
            } else {
                cplx *dst = (cplx *)aux_host + b * d->Mmax;
                size_t n = (size_t)nkept[b];
                if (n > cap)
                    n = cap;
                if (n > d->Mmax) {
                    status_host[b] |= 32;
                    n = d->Mmax;
                }
                memcpy(dst, roots.data() + b * maxn, n * sizeof(cplx));
                M_host[b] = n;
            }
        }
        tick("assembly of the user arrays");
    }
    return 0;
}

// ---------------------------------------------------------------------------
// 4SPLIT4 preprocessing
// ---------------------------------------------------------------------------
int fnftb_resample_4split4(fnftb_ctx *c, double eps_t, int32_t *warn_host)
{
    return fnftb_resample_4split4_sub(c, eps_t, 1, c ? c->D : 0, warn_host);
}

int fnftb_subsample(fnftb_ctx *c, size_t nskip, size_t Dsub)
{
    if (!c || !c->q)
        return fail(-2, "no signals staged", __FILE__, __LINE__);
    if (nskip < 1 || Dsub < 1 || (Dsub - 1) * nskip >= c->D)
        return fail(-2, "invalid subsampling", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ensure(c->qsub, c->B * Dsub * sizeof(cplx)));
    SubsampleArgs sa;
    sa.q = c->q;
    sa.out = (cplx *)c->qsub.p;
    sa.B = (int)c->B;
    sa.D = (int)c->D;
    sa.nskip = (int)nskip;
    sa.Dsub = (int)Dsub;
    const long long total = (long long)c->B * (long long)Dsub;
    RC((launch_blocks<SubsampleArgs, blk_subsample>(sa, (unsigned)((total + 255) / 256), 256, 0, c->st,
                                                    "subsample")));
    c->q = (const cplx *)c->qsub.p;
    c->r = nullptr;
    c->D = Dsub;
    c->have_box3 = 0;
    c->slow_wsel = 0;
    c->rpre = nullptr;
    return 0;
}

// 4SPLIT4 preprocessing for any number of samples: length-D DFTs as chirp-z transforms
// (resample_kernels.cuh, second half)
static int resample_general(fnftb_ctx *c, double eps_t, size_t nskip, size_t Dsub, int32_t *warn_host, int up = 2,
                            int wsel = 0, int kappa = 1)
{
    const size_t B = c->B, D = c->D;
    if (2 * D > ((size_t)1 << 24))
        return fail(-6, "signal too long for the GPU resampling step", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ensure(c->rs_a, B * 2 * D * sizeof(cplx)));
    RC(ensure(c->rs_b, B * 2 * D * sizeof(cplx)));
    Buf &dstbuf = (nskip > 1) ? c->qsub : c->qpre;  // a subsampled copy never replaces the full one
    RC(ensure(dstbuf, B * (size_t)up * Dsub * sizeof(cplx)));
    RC(ensure(c->warn, B * sizeof(int)));
    RC(ensure(c->status, B * sizeof(int)));
    const CzGeom g = cz_geometry((int)D - 1, (int)D);
    RC(ensure(c->ybuf, cz_ybuf_elems(g, B, 2) * sizeof(cplx)));
    RC(ensure(c->vhat, (size_t)g.L * sizeof(cplx)));
    RC(ensure(c->cztab, cz_table_elems(g, (int)D - 1, (int)D) * sizeof(cplx)));
    RsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.B = (int)B;
    ra.D = (int)D;
    ra.nskip = (int)nskip;
    ra.Dsub = (int)Dsub;
    ra.eps_t = eps_t;
    ra.warn = (int *)c->warn.p;
    ra.up = up;
    ra.q0 = c->q;
    ra.wsel = wsel;
    ra.kappa = kappa;
    if (wsel >= 2) {
        RC(ensure(c->rprebuf, B * (size_t)up * Dsub * sizeof(cplx)));
        ra.rout = (cplx *)c->rprebuf.p;
    }
    if (up == 3) {
        // CF4_3 weights (/root/reference/src/private/fnft__akns_discretization.c:299-327): Legendre
        // expansion of the coefficient table f at the three Gauss nodes
        const double f[3][3] = {{11.0 / 40.0, 20.0 / 87.0, 7.0 / 50.0},
                                {9.0 / 20.0, 0.0, -7.0 / 25.0},
                                {11.0 / 40.0, -20.0 / 87.0, 7.0 / 50.0}};
        const double wm[3] = {5.0 / 18.0, 4.0 / 9.0, 5.0 / 18.0};
        const double xm[3] = {2.0 * sqrt(3.0 / 20.0), 0.0, -2.0 * sqrt(3.0 / 20.0)};
        for (int m = 0; m < 3; ++m) {
            const double P[3] = {1.0, xm[m], 0.5 * (3.0 * xm[m] * xm[m] - 1.0)};
            for (int i = 0; i < 3; ++i) {
                double w = 0.0;
                for (int n = 0; n < 3; ++n)
                    w += (2 * n + 1) * P[n] * f[i][n];
                ra.w3[i * 3 + m] = w * wm[m];
            }
        }
    }
    const long long tot = (long long)B * (long long)D;
    // 1. q reversed -> rs_a[B][D]
    ra.in = c->q;
    ra.out = (cplx *)c->rs_a.p;
    RC((launch_blocks<RsArgs, blk_rs_reverse>(ra, (unsigned)((tot + 255) / 256), 256, 0, c->st, "resample_reverse")));
    // 2. X = DFT_D(q) -> rs_b[B][D]
    CzArgs a;
    memset(&a, 0, sizeof(a));
    a.tm = (const cplx *)c->rs_a.p;
    a.tm_sstride = D;
    a.ent[0] = 0;
    a.ent[1] = 0;
    a.npoly = 1;
    a.deg = (int)D - 1;
    a.B = (int)B;
    a.M = (int)D;
    a.lwr = 0.0;
    a.lwi = -2.0 * 3.14159265358979323846 / (double)D;  // W = exp(-2 pi i / D), A = 1
    a.dft_n = (int)D;
    a.ybuf = (cplx *)c->ybuf.p;
    a.vhat = (cplx *)c->vhat.p;
    a.T = ctx_tw(c);
    a.mode = FNFTB_CZ_RAW;
    a.out = (cplx *)c->rs_b.p;
    a.out_sstride = D;
    a.status = (int *)c->status.p;
    RC(cz_run_fastrows(a, (cplx *)c->cztab.p, c->tws, c->st));
    // 3. band-limit check, the two phase ramps -> rs_a[B][2][D] (reversed)
    ra.in = (const cplx *)c->rs_b.p;
    ra.out = (cplx *)c->rs_a.p;
    RC((launch_blocks<RsArgs, blk_rs_shift>(ra, (unsigned)B, 256, 3 * 256 * sizeof(double), c->st, "resample_shift")));
    // 4. the two inverse DFTs (unnormalised) -> rs_b[B][2][D]
    a.tm = (const cplx *)c->rs_a.p;
    a.tm_sstride = 2 * D;
    a.ent[0] = 0;
    a.ent[1] = 1;
    a.npoly = 2;
    a.lwi = 2.0 * 3.14159265358979323846 / (double)D;
    a.out = (cplx *)c->rs_b.p;
    a.out_sstride = 2 * D;
    a.out_jstride = D;
    RC(cz_run_fastrows(a, (cplx *)c->cztab.p, c->tws, c->st));
    // 5. weights + subsampling -> qpre[B][2*Dsub]
    ra.in = (const cplx *)c->rs_b.p;
    ra.out = (cplx *)dstbuf.p;
    const long long tot2 = (long long)B * (long long)Dsub;
    RC((launch_blocks<RsArgs, blk_rs_weights>(ra, (unsigned)((tot2 + 255) / 256), 256, 0, c->st, "resample_weights")));
    c->q = (const cplx *)dstbuf.p;
    c->r = nullptr;
    c->D = (size_t)up * Dsub;
    c->have_box3 = 0;
    c->slow_wsel = wsel;
    c->rpre = (wsel >= 2) ? (const cplx *)c->rprebuf.p : nullptr;
    if (warn_host) {
        CU(cudaMemcpyAsync(warn_host, c->warn.p, B * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    return 0;
}

// CF4_3 preprocessing (/root/reference/src/private/fnft__nse_discretization.c:505-531): always
// through the general (chirp-z DFT) path; the staged signals become [B][3*Dsub]
int fnftb_resample_cf4_3_sub(fnftb_ctx *c, double eps_t, size_t nskip, size_t Dsub, int32_t *warn_host)
{
    if (!c || !c->q)
        return fail(-2, "no signals staged", __FILE__, __LINE__);
    if (nskip < 1 || Dsub < 1 || (Dsub - 1) * nskip >= c->D)
        return fail(-2, "invalid subsampling", __FILE__, __LINE__);
    if (c->D < 4)
        return fail(-6, "resampling needs at least 4 samples", __FILE__, __LINE__);
    return resample_general(c, eps_t, nskip, Dsub, warn_host, 3, 1, 1);
}

// CF5_3 (wsel 2) and CF6_4 (wsel 3) preprocessing (:532-604): complex weights, explicit r samples
int fnftb_resample_cf_sub(fnftb_ctx *c, int wsel, int kappa, double eps_t, size_t nskip, size_t Dsub,
                          int32_t *warn_host)
{
    if (wsel == 1)
        return fnftb_resample_cf4_3_sub(c, eps_t, nskip, Dsub, warn_host);
    if (!c || !c->q || (wsel != 2 && wsel != 3) || (kappa != 1 && kappa != -1))
        return fail(-2, "invalid argument / no signals staged", __FILE__, __LINE__);
    if (nskip < 1 || Dsub < 1 || (Dsub - 1) * nskip >= c->D)
        return fail(-2, "invalid subsampling", __FILE__, __LINE__);
    if (c->D < 4)
        return fail(-6, "resampling needs at least 4 samples", __FILE__, __LINE__);
    return resample_general(c, eps_t, nskip, Dsub, warn_host, wsel == 2 ? 3 : 4, wsel, kappa);
}

// ES4 (wsel 4) / TES4 (wsel 5) preprocessing: (q, q', q'') per sub-sampled grid point by finite differences
// (/root/reference/src/private/fnft__nse_discretization.c:609-631); the staged signals become [B][3*Dsub]
int fnftb_preprocess_es4(fnftb_ctx *c, int wsel, double eps_t, size_t nskip, size_t Dsub)
{
    if (!c || !c->q || (wsel != 4 && wsel != 5))
        return fail(-2, "invalid argument / no signals staged", __FILE__, __LINE__);
    if (nskip < 1 || Dsub < 2 || (Dsub - 1) * nskip >= c->D)
        return fail(-2, "invalid subsampling", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    Buf &dstbuf = (nskip > 1) ? c->qsub : c->qpre;  // a subsampled copy never replaces the full one
    RC(ensure(dstbuf, c->B * 3 * Dsub * sizeof(cplx)));
    Es4Args ea;
    ea.q = c->q;
    ea.out = (cplx *)dstbuf.p;
    ea.B = (int)c->B;
    ea.D = (int)c->D;
    ea.nskip = (int)nskip;
    ea.Dsub = (int)Dsub;
    ea.eps_sub = eps_t * (double)nskip;
    const long long total = (long long)c->B * (long long)Dsub;
    RC((launch_blocks<Es4Args, blk_es4_preprocess>(ea, (unsigned)((total + 255) / 256), 256, 0, c->st,
                                                   "es4_preprocess")));
    c->q = (const cplx *)dstbuf.p;
    c->r = nullptr;
    c->D = 3 * Dsub;
    c->have_box3 = 0;
    c->slow_wsel = wsel;
    c->rpre = nullptr;
    return 0;
}

// The staged signals are preprocessed samples supplied by the caller (private API): CF4_3 (wsel 1), ES4 (4), TES4 (5)
int fnftb_set_slow_weights(fnftb_ctx *c, int wsel)
{
    if (!c || (wsel != 0 && wsel != 1 && wsel != 4 && wsel != 5))
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    c->slow_wsel = wsel;
    c->rpre = nullptr;
    return 0;
}

int fnftb_resample_4split4_sub(fnftb_ctx *c, double eps_t, size_t nskip, size_t Dsub, int32_t *warn_host)
{
    if (!c || !c->q)
        return fail(-2, "no signals staged", __FILE__, __LINE__);
    if (nskip < 1 || Dsub < 1 || (Dsub - 1) * nskip >= c->D)
        return fail(-2, "invalid subsampling", __FILE__, __LINE__);
    const size_t D = c->D;
    if (D < 4)
        return fail(-6, "resampling needs at least 4 samples", __FILE__, __LINE__);
    if ((D & (D - 1)) != 0 || D > 4096)
        return resample_general(c, eps_t, nskip, Dsub, warn_host);
    CU(cudaSetDevice(c->device));
    Buf &dstbuf = (nskip > 1) ? c->qsub : c->qpre;  // a subsampled copy never replaces the full one
    RC(ensure(dstbuf, c->B * 2 * Dsub * sizeof(cplx)));
    RC(ensure(c->warn, c->B * sizeof(int)));
    ResampleArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.q = c->q;
    ra.out = (cplx *)dstbuf.p;
    ra.warn = (int *)c->warn.p;
    ra.B = (int)c->B;
    ra.D = (int)D;
    ra.nskip = (int)nskip;
    ra.Dsub = (int)Dsub;
    ra.eps_t = eps_t;
    ra.plan = make_fft_plan((int)D);
    ra.T = ctx_tw(c);
    const int nt = 256;
    RC((launch_blocks<ResampleArgs, blk_resample_4split4>(ra, (unsigned)c->B, nt,
                                                          resample_smem_bytes((int)D, nt), c->st, "resample_4split4")));
    c->q = (const cplx *)dstbuf.p;
    c->r = nullptr;
    c->D = 2 * Dsub;
    c->have_box3 = 0;
    c->slow_wsel = 0;
    c->rpre = nullptr;
    if (warn_host) {
        CU(cudaMemcpyAsync(warn_host, c->warn.p, c->B * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    return 0;
}

// ---------------------------------------------------------------------------
// polynomial roots (poly_roots.cuh)
// ---------------------------------------------------------------------------
static int roots_run(fnftb_ctx *c, const cplx *coef, long long cstride, size_t B, size_t n, void *roots_host,
                     int32_t *info_host)
{
    RC(ensure(c->rt_roots, B * n * sizeof(cplx)));
    RC(ensure(c->rt_absc, B * (n + 1) * sizeof(double)));
    RC(ensure(c->rt_lg, B * (n + 1) * sizeof(double)));
    RC(ensure(c->rt_hull, B * (n + 2) * sizeof(int)));
    RC(ensure(c->rt_info, B * 4 * sizeof(int)));
    RootsArgs ra;
    ra.coef = coef;
    ra.cstride = cstride;
    ra.n = (int)n;
    ra.roots = (cplx *)c->rt_roots.p;
    ra.absc = (double *)c->rt_absc.p;
    ra.lg = (double *)c->rt_lg.p;
    ra.hull = (int *)c->rt_hull.p;
    ra.info = (int *)c->rt_info.p;
    // Sweeps after which a root that still moves is given up (returned as NaN, dropped by the callers' filters).  Roots that
    // converge do so within ~20 sweeps; a few ill-conditioned ones per polynomial never meet the residual test and kept every
    // CTA sweeping with almost all threads idle: with 200 sweeps config 7 ran at 6.1 k signals/s, with 64 at 7.7 k, with the
    // same bound states on all 1024 signals and the whole reference suite green (scripts/roots_maxit.sh)
    static const int knob_maxit = tree_knob("FNFT_B200_ROOTS_MAXIT", 64);
    ra.maxit = knob_maxit;
    ra.in_global = 0;
    ra.stats = 0;
    if (g_fnftb_profile_on)
        fnftb_profile_begin("poly_roots", c->st);
    const int rc = roots_launch(ra, (int)B, c->st);
    if (g_fnftb_profile_on)
        fnftb_profile_end(c->st);
    if (rc == -6)
        return fail(-6, "polynomial degree too large for the GPU root finder (max 32768)", __FILE__, __LINE__);
    if (rc)
        return fail(rc, "root finder launch failed", __FILE__, __LINE__);
    c->rt_B = B;
    c->rt_n = n;
    if (roots_host)
        CU(cudaMemcpyAsync(roots_host, c->rt_roots.p, B * n * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
    if (info_host)
        CU(cudaMemcpyAsync(info_host, c->rt_info.p, B * 4 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    if (roots_host || info_host)
        CU(cudaStreamSynchronize(c->st));
    return 0;
}

int fnftb_roots_lambda(fnftb_ctx *c, double lam_den, const double *box, int use_box3, void *lam_host,
                       size_t stride, int32_t *count_host)
{
    if (!c || !lam_host || !count_host || c->rt_B == 0 || stride == 0)
        return fail(-2, "invalid argument / no roots held", __FILE__, __LINE__);
    if (use_box3 && (!c->have_box3 || c->B != c->rt_B))
        return fail(-2, "no per-signal bound available", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    const size_t B = c->rt_B, n = c->rt_n;
    RC(ensure(c->rt_lam, B * n * sizeof(cplx)));
    RC(ensure(c->rt_cnt, B * sizeof(int)));
    RootsLamArgs a;
    a.roots = (const cplx *)c->rt_roots.p;
    a.info = (const int *)c->rt_info.p;
    a.lam = (cplx *)c->rt_lam.p;
    a.count = (int *)c->rt_cnt.p;
    a.n = (int)n;
    a.lam_den = lam_den;
    a.filtering = box ? 1 : 0;
    for (int i = 0; i < 4; ++i)
        a.box[i] = box ? box[i] : 0.0;
    a.box3 = use_box3 ? (const double *)c->box3.p : nullptr;
    k_roots_lambda<<<(unsigned)B, 256, 0, c->st>>>(a);
    ++g_fnftb_launch_count;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(count_host, c->rt_cnt.p, B * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    size_t maxn = 0;
    for (size_t b = 0; b < B; ++b) {
        const size_t k = (size_t)count_host[b] > stride ? stride : (size_t)count_host[b];
        if (k > maxn)
            maxn = k;
    }
    if (maxn > 0) {
        CU(cudaMemcpy2DAsync(lam_host, stride * sizeof(cplx), c->rt_lam.p, n * sizeof(cplx), maxn * sizeof(cplx), B,
                             cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    return 0;
}

int fnftb_poly_roots(fnftb_ctx *c, int ent, void *roots_host, int32_t *info_host)
{
    if (!c || c->tmB == 0 || c->deg < 1 || ent < 0 || ent >= (int)c->tm_entries)
        return fail(-2, "invalid argument / no transfer matrix held", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ctx_finalize(c));
    const size_t n = c->deg;
    return roots_run(c, (const cplx *)c->tm.p + (size_t)ent * (n + 1), (long long)(c->tm_entries * (n + 1)),
                     c->tmB, n, roots_host, info_host);
}

// ---------------------------------------------------------------------------
// bound states
// ---------------------------------------------------------------------------
static BoundArgs bound_args(fnftb_ctx *c, const fnftb_bound_desc *d)
{
    BoundArgs a;
    memset(&a, 0, sizeof(a));
    a.q = c->q;
    a.B = (int)c->B;
    a.D = (int)c->D;
    a.upsampling = d->upsampling;
    a.wsel = c->slow_wsel;
    a.r = c->rpre;
    a.Kmax = d->Kmax;
    a.K = (const int *)c->kcnt.p;
    a.lam = (cplx *)c->lam.p;
    a.T0 = d->T0;
    a.T1 = d->T1;
    a.eps_t = d->eps_t;
    a.bc = d->bc;
    a.lweight = d->lweight;
    a.scl = d->scl;
    a.niter = d->niter;
    a.box0 = d->box0;
    a.box1 = d->box1;
    a.box2 = d->box2;
    a.box3 = (d->use_box3 && c->have_box3) ? (const double *)c->box3.p : nullptr;
    a.flag = (int *)c->flag.p;
    a.a_out = (cplx *)c->aout.p;
    a.ap_out = (cplx *)c->apout.p;
    a.b_out = (cplx *)c->bout.p;
    a.phi = (cplx *)c->phi.p;
    return a;
}

int fnftb_imbound(fnftb_ctx *c, int upsampling, double T0, double T1, double *box3_host)
{
    if (!c || !c->q)
        return fail(-2, "no signals staged", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ensure(c->box3, c->B * sizeof(double)));
    NormArgs na;
    na.q = c->q;
    na.B = (int)c->B;
    na.D = (int)c->D;
    na.upsampling = upsampling;
    na.T0 = T0;
    na.T1 = T1;
    na.out = (double *)c->box3.p;
    RC((launch_blocks<NormArgs, blk_imbound>(na, (unsigned)c->B, 256, 256 * sizeof(double), c->st, "bound_imbound")));
    c->have_box3 = 1;
    if (box3_host) {
        CU(cudaMemcpyAsync(box3_host, c->box3.p, c->B * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    return 0;
}

static int stage_eigs(fnftb_ctx *c, const fnftb_bound_desc *d, const int32_t *K_host, const void *lam_host)
{
    const size_t n = c->B * (size_t)d->Kmax;
    RC(ensure(c->lam, n * sizeof(cplx)));
    RC(ensure(c->kcnt, c->B * sizeof(int)));
    RC(ensure(c->flag, n * sizeof(int)));
    CU(cudaMemcpyAsync(c->lam.p, lam_host, n * sizeof(cplx), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->kcnt.p, K_host, c->B * sizeof(int), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemsetAsync(c->flag.p, 0, n * sizeof(int), c->st));
    return 0;
}

int fnftb_newton(fnftb_ctx *c, const fnftb_bound_desc *d, const int32_t *K_host, void *lam_host,
                 int32_t *flag_host)
{
    if (!c || !d || !c->q || !K_host || !lam_host || d->Kmax < 1)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(stage_eigs(c, d, K_host, lam_host));
    const size_t n = c->B * (size_t)d->Kmax;
    BoundArgs a = bound_args(c, d);
    static const int knob_warp = tree_knob("FNFT_B200_BOUND_WARP", 1);
    if (knob_warp || c->slow_wsel >= 4) {  // (ES4 / TES4 exist in the warp kernels only)
        // one eigenvalue per warp (bound_warp.cuh)
        if (g_fnftb_profile_on)
            fnftb_profile_begin("bound_newton_warp", c->st);
        k_newton_warp<<<(unsigned)((n + 3) / 4), 128, 0, c->st>>>(a);
        if (g_fnftb_profile_on)
            fnftb_profile_end(c->st);
        ++g_fnftb_launch_count;
        CU(cudaGetLastError());
    } else {
        RC((launch_blocks<BoundArgs, blk_newton, 128>(a, (unsigned)((n + 63) / 64), 64, 0, c->st, "bound_newton")));
    }
    CU(cudaMemcpyAsync(lam_host, c->lam.p, n * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
    if (flag_host)
        CU(cudaMemcpyAsync(flag_host, c->flag.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return 0;
}

int fnftb_normconsts(fnftb_ctx *c, const fnftb_bound_desc *d, const int32_t *K_host,
                     const void *lam_host, void *a_host, void *ap_host, void *b_host)
{
    if (!c || !d || !c->q || !K_host || !lam_host || d->Kmax < 1)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(stage_eigs(c, d, K_host, lam_host));
    const size_t n = c->B * (size_t)d->Kmax;
    RC(ensure(c->aout, n * sizeof(cplx)));
    RC(ensure(c->apout, n * sizeof(cplx)));
    RC(ensure(c->bout, n * sizeof(cplx)));
    RC(ensure(c->koff, c->B * sizeof(int)));
    const size_t Dg = c->D / (size_t)d->upsampling;
    CU(cudaMemsetAsync(c->aout.p, 0, n * sizeof(cplx), c->st));
    CU(cudaMemsetAsync(c->apout.p, 0, n * sizeof(cplx), c->st));
    CU(cudaMemsetAsync(c->bout.p, 0, n * sizeof(cplx), c->st));
    // The PHI scratch has one slot of (D_given + 1) vectors per eigenvalue that is actually present (not per
    // Kmax: callers following the reference convention pass Kmax = fnft_nsev_max_K(D)).  Signals are
    // processed in groups whose slots fit the scratch budget.
    static const size_t phi_budget = (size_t)tree_knob("FNFT_B200_PHI_MB", 4096) << 20;
    const size_t slot_bytes = (Dg + 1) * 2 * sizeof(cplx);
    std::vector<int> koff(c->B);
    static const int knob_warp = tree_knob("FNFT_B200_BOUND_WARP", 1);
    for (size_t b0 = 0; b0 < c->B;) {
        size_t b1 = b0, ktot = 0;
        while (b1 < c->B && (b1 == b0 || (ktot + (size_t)K_host[b1]) * slot_bytes <= phi_budget)) {
            koff[b1] = (int)ktot;
            ktot += (size_t)K_host[b1];
            ++b1;
        }
        if (ktot > 0) {
            if (ktot > (size_t)0x7fffffff)
                return fail(-6, "too many eigenvalues in one signal group", __FILE__, __LINE__);
            RC(ensure(c->phi, ktot * slot_bytes));
            CU(cudaMemcpyAsync((int *)c->koff.p + b0, koff.data() + b0, (b1 - b0) * sizeof(int),
                               cudaMemcpyHostToDevice, c->st));
            BoundArgs a = bound_args(c, d);
            const size_t e0 = b0 * (size_t)d->Kmax;
            a.q += b0 * c->D;
            if (a.r)
                a.r += b0 * c->D;
            a.B = (int)(b1 - b0);
            a.K += b0;
            a.lam += e0;
            a.flag += e0;
            a.a_out += e0;
            a.ap_out += e0;
            a.b_out += e0;
            if (a.box3)
                a.box3 += b0;
            a.koff = (const int *)c->koff.p + b0;
            a.ktot = (int)ktot;
            const size_t ng = (size_t)a.B * (size_t)d->Kmax;
            if (knob_warp || c->slow_wsel >= 4) {
                if (g_fnftb_profile_on)
                    fnftb_profile_begin("bound_normconsts_warp", c->st);
                k_normconsts_warp<<<(unsigned)((ng + 3) / 4), 128, 0, c->st>>>(a);
                if (g_fnftb_profile_on)
                    fnftb_profile_end(c->st);
                ++g_fnftb_launch_count;
                CU(cudaGetLastError());
            } else {
                RC((launch_blocks<BoundArgs, blk_normconsts, 128>(a, (unsigned)((ng + 63) / 64), 64, 0, c->st,
                                                                  "bound_normconsts")));
            }
            // koff (host vector) is reused by the next group only after this copy has been consumed
            CU(cudaStreamSynchronize(c->st));
        }
        b0 = b1;
    }
    if (a_host)
        CU(cudaMemcpyAsync(a_host, c->aout.p, n * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
    if (ap_host)
        CU(cudaMemcpyAsync(ap_host, c->apout.p, n * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
    if (b_host)
        CU(cudaMemcpyAsync(b_host, c->bout.p, n * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Continuous spectrum by SEGMENTS, for signals whose transfer matrix is longer than one product tree can hold
// (degree > 2^18).  The reference has no such limit (src/private/fnft__poly_fmult.c:404-445 multiplies whatever
// it is given); here the host (fnft_nsev.c: nsev_pass_segmented) cuts the signal into pieces the tree can
// multiply, gets the scattering coefficients (a_s, b_s)(xi) of every piece on its own time window through the
// normal path (tree + chirp-z, contspec type AB), and chains them on the real xi grid:
//     a <- a_s a - kappa conj(b_s) b,     b <- b_s a + conj(a_s) b        (pieces in order of increasing time)
// which is the product of the pieces' transfer matrices [a_s, -kappa b_s*; b_s, a_s*] evaluated point by point
// instead of coefficient by coefficient.  The epilogue of src/fnft_nsev.c:846-876 follows on the chained values.
// ---------------------------------------------------------------------------------------
__global__ void k_seg_compose(cplx *acc, const cplx *cur, size_t n, int M, double kap, int first)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const size_t s = i / (size_t)M, m = i % (size_t)M;
    const cplx as = cur[s * 2 * M + m], bs = cur[s * 2 * M + M + m];
    if (first) {
        acc[s * 2 * M + m] = as;
        acc[s * 2 * M + M + m] = bs;
        return;
    }
    const cplx a = acc[s * 2 * M + m], b = acc[s * 2 * M + M + m];
    cplx an = cmul(as, a);
    cfmac(an, cscale(b, -kap), bs);  // - kappa * b * conj(b_s)
    cplx bn = cmul(bs, a);
    cfmac(bn, b, as);                // + b * conj(a_s)
    acc[s * 2 * M + m] = an;
    acc[s * 2 * M + M + m] = bn;
}

__global__ void k_seg_finish(const cplx *acc, cplx *out, size_t out_sstride, size_t n, int M, int cstype, int *status)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const size_t s = i / (size_t)M, m = i % (size_t)M;
    const cplx a = acc[s * 2 * M + m], b = acc[s * 2 * M + M + m];
    cplx *o = out + s * out_sstride;
    size_t off = 0;
    if (cstype == 0 || cstype == 2) {
        if (a.x == 0.0 && a.y == 0.0) {
            status[s] = 3;  // FNFT_EC_DIV_BY_ZERO, src/fnft_nsev.c:850-852
            o[m] = make_cplx(NAN, NAN);
        } else {
            o[m] = cdiv(b, a);
        }
        off = (size_t)M;
    }
    if (cstype == 1 || cstype == 2) {
        o[off + m] = a;
        o[off + M + m] = b;
    }
}

// general 2x2 chaining (KdV): cur = [H12 | H22 | H11 | H21] of the piece at the same points z_m, acc = second column
// (v1, v2) of the product of the pieces so far; later pieces multiply from the left
__global__ void k_seg_compose_general(cplx *acc, const cplx *cur, size_t n, int M, int first)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const size_t s = i / (size_t)M, m = i % (size_t)M;
    const cplx *c = cur + s * 4 * M;
    cplx *a = acc + s * 2 * M;
    if (first) {
        a[m] = c[m];
        a[M + m] = c[M + m];
        return;
    }
    const cplx v1 = a[m], v2 = a[M + m];
    cplx n1 = cmul(c[2 * (size_t)M + m], v1);
    cfma(n1, c[m], v2);
    cplx n2 = cmul(c[3 * (size_t)M + m], v1);
    cfma(n2, c[M + m], v2);
    a[m] = n1;
    a[M + m] = n2;
}

// src/fnft_kdvv.c:186-203 on the chained (H12, H22); same formulas as the epilogue of blk_cz_cols_inv
__global__ void k_seg_finish_kdv(const cplx *acc, cplx *out, size_t out_sstride, size_t n, int M, double xi0, double eps_xi,
                                 double kdv_ph, double kdv_sqrtz)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const size_t s = i / (size_t)M, m = i % (size_t)M;
    const double xi = -xi0 - (double)m * eps_xi;
    cplx h12 = acc[s * 2 * M + m];
    const cplx h22 = acc[s * 2 * M + M + m];
    double sn, cs;
    if (kdv_sqrtz != 0.0) {
        sincos(xi * kdv_sqrtz, &sn, &cs);
        h12 = cdiv(h12, make_cplx(cs, sn));
    }
    sincos(2.0 * xi * kdv_ph, &sn, &cs);
    const cplx num = cmul(make_cplx(cs, sn), h12);
    const cplx den = make_cplx(-2.0 * xi * h22.y - h12.x, 2.0 * xi * h22.x - h12.y);
    out[s * out_sstride + m] = cdiv(num, den);
}

extern "C" {

size_t fnftb_tree_max_samples(int scheme, int deg0)
{
    if (deg0 < 1)
        return 0;
    const size_t dtree = (size_t)tree_leaf_degree(scheme, deg0);
    size_t n = 1;
    while (2 * n * dtree <= ((size_t)1 << 18))  // the same bound as fnftb_fscatter
        n *= 2;
    // FNFT_B200_TREE_MAX_SAMPLES: a smaller limit (tests exercise the segmented path at sizes the direct path
    // also handles)
    static const int knob = tree_knob("FNFT_B200_TREE_MAX_SAMPLES", 0);
    if (knob >= 2 && (size_t)knob < n)
        n = (size_t)knob;
    return n;
}

// stage B pieces of Dseg samples each, taken with a row stride of `stride` samples from q (host or device)
int fnftb_set_signals_strided(fnftb_ctx *c, size_t B, size_t Dseg, const void *q, size_t stride, int on_device)
{
    if (!c || !q || B == 0 || Dseg == 0 || stride < Dseg)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    c->B = B;
    c->D = Dseg;
    c->have_box3 = 0;
    c->slow_wsel = 0;
    c->rpre = nullptr;
    RC(ensure(c->qbuf, B * Dseg * sizeof(cplx)));
    c->qbuf_host = nullptr;
    CU(cudaMemcpy2DAsync(c->qbuf.p, Dseg * sizeof(cplx), q, stride * sizeof(cplx), Dseg * sizeof(cplx), B,
                         on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->st));
    c->q = (const cplx *)c->qbuf.p;
    c->r = nullptr;
    return 0;
}

// device buffer [B][2M] that receives the (a, b) of the current piece (fnftb_contspec with on_device = 1)
void *fnftb_seg_buffer(fnftb_ctx *c, size_t B, size_t M, int nent)
{
    if (!c || nent < 2 || nent > 4 || cudaSetDevice(c->device) != cudaSuccess)
        return nullptr;
    if (ensure(c->segcur, B * (size_t)nent * M * sizeof(cplx)) != 0 || ensure(c->segacc, B * 2 * M * sizeof(cplx)) != 0)
        return nullptr;
    return c->segcur.p;
}

int fnftb_seg_compose_general(fnftb_ctx *c, size_t B, size_t M, int first)
{
    if (!c || B == 0 || M == 0 || !c->segcur.p || !c->segacc.p)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    const size_t n = B * M;
    k_seg_compose_general<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>((cplx *)c->segacc.p, (const cplx *)c->segcur.p, n,
                                                                        (int)M, first);
    ++g_fnftb_launch_count;
    CU(cudaGetLastError());
    return 0;
}

int fnftb_seg_finish_kdv(fnftb_ctx *c, size_t B, size_t M, double xi0, double eps_xi, double kdv_ph, double kdv_sqrtz,
                         void *out, size_t out_sstride, int on_device)
{
    if (!c || !out || B == 0 || M == 0 || !c->segacc.p)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    cplx *dst = (cplx *)out;
    if (!on_device) {
        RC(ensure(c->outbuf, B * out_sstride * sizeof(cplx)));
        dst = (cplx *)c->outbuf.p;
    }
    const size_t n = B * M;
    k_seg_finish_kdv<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>((const cplx *)c->segacc.p, dst, out_sstride, n, (int)M, xi0,
                                                                   eps_xi, kdv_ph, kdv_sqrtz);
    ++g_fnftb_launch_count;
    CU(cudaGetLastError());
    if (!on_device) {
        CU(cudaMemcpyAsync(out, dst, B * out_sstride * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    return 0;
}

int fnftb_seg_compose(fnftb_ctx *c, size_t B, size_t M, int kappa, int first)
{
    if (!c || B == 0 || M == 0 || !c->segcur.p || !c->segacc.p)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    const size_t n = B * M;
    k_seg_compose<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>((cplx *)c->segacc.p, (const cplx *)c->segcur.p, n, (int)M,
                                                                (double)kappa, first);
    ++g_fnftb_launch_count;
    CU(cudaGetLastError());
    return 0;
}

// epilogue on the chained values; out: [B][out_sstride] (host, or device if on_device); status_host may be NULL
int fnftb_seg_finish(fnftb_ctx *c, size_t B, size_t M, int cstype, void *out, size_t out_sstride, int on_device,
                     int32_t *status_host)
{
    if (!c || !out || B == 0 || M == 0 || !c->segacc.p)
        return fail(-2, "invalid argument", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ensure(c->status, B * sizeof(int)));
    CU(cudaMemsetAsync(c->status.p, 0, B * sizeof(int), c->st));
    cplx *dst = (cplx *)out;
    if (!on_device) {
        RC(ensure(c->outbuf, B * out_sstride * sizeof(cplx)));
        dst = (cplx *)c->outbuf.p;
    }
    const size_t n = B * M;
    k_seg_finish<<<(unsigned)((n + 255) / 256), 256, 0, c->st>>>((const cplx *)c->segacc.p, dst, out_sstride, n, (int)M, cstype,
                                                               (int *)c->status.p);
    ++g_fnftb_launch_count;
    CU(cudaGetLastError());
    if (!on_device) {
        CU(cudaMemcpyAsync(out, dst, B * out_sstride * sizeof(cplx), cudaMemcpyDeviceToHost, c->st));
        if (status_host)
            CU(cudaMemcpyAsync(status_host, c->status.p, B * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// hooks for inverse_api.cu (ctx_hooks.h): the inverse transform lives in its own translation
// unit and borrows the stream, the general 2x2 pair product and the any-length DFT from here
// ---------------------------------------------------------------------------------------
#include "ctx_hooks.h"
cudaStream_t fnftb__stream(fnftb_ctx *c) { return c->st; }
int fnftb__activate(fnftb_ctx *c)
{
    CU(cudaSetDevice(c->device));
    return 0;
}
void **fnftb__inv_slot(fnftb_ctx *c, void (***dtor)(void *))
{
    *dtor = &c->inv_free;
    return &c->inv_ws;
}
int fnftb__fail(int code, const char *what, const char *file, int line) { return fail(code, what, file, line); }

// B products of two 2x2 polynomial matrices of degree d (a power of two; general coefficient path of
// tree_driver.cuh, no normalisation).  prepare: *lev0 = operand buffer [B][2][4][d+1] to be filled on the
// context's stream; run: *res = result [B][4][2d+1] = matrix 0 * matrix 1 (valid until the next tree call).
int fnftb__pair2x2_prepare(fnftb_ctx *c, size_t B, size_t d, cplx **lev0)
{
    if (!c || B == 0 || d == 0 || (d & (d - 1)) != 0 || d > ((size_t)1 << 15))
        return fail(-6, "pair product: degree must be a power of two <= 32768", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    RC(ensure_tree(c, B, 2, d, 2 * d));
    c->deferred.valid = c->deferred.cols_pending = 0;
    c->tmB = 0;
    *lev0 = (cplx *)c->lev0.p;
    return 0;
}
int fnftb__pair2x2_run(fnftb_ctx *c, size_t B, size_t d, const cplx **res)
{
    const TwTable T = ctx_tw(c);
    int cur = 0;
    const TreeWork w = tree_work(c);
    RC(tree_levels(w, (int)B, 2, (int)d, 0, T, c->st, &cur));
    *res = w.lev[cur];
    return 0;
}

// B discrete Fourier transforms of any length n (sign -1: forward, +1: inverse, both unnormalised) as
// chirp-z transforms (W = exp(sign 2 pi i / n), A = 1) through the general four-step path, like the
// resampling step.  in_rev holds the inputs REVERSED (x[n-1-k], polynomial coefficients in descending order).
int fnftb__dft(fnftb_ctx *c, size_t B, size_t n, const cplx *in_rev, cplx *out, int sign)
{
    if (!c || B == 0 || n < 2 || 2 * n > ((size_t)1 << 24))
        return fail(-6, "DFT length not supported", __FILE__, __LINE__);
    CU(cudaSetDevice(c->device));
    const CzGeom g = cz_geometry((int)n - 1, (int)n);
    RC(ensure(c->ybuf, cz_ybuf_elems(g, B, 1) * sizeof(cplx)));
    RC(ensure(c->vhat, (size_t)g.L * sizeof(cplx)));
    RC(ensure(c->cztab, cz_table_elems(g, (int)n - 1, (int)n) * sizeof(cplx)));
    RC(ensure(c->status, B * sizeof(int)));
    CzArgs a;
    memset(&a, 0, sizeof(a));
    a.tm = in_rev;
    a.tm_sstride = n;
    a.ent[0] = 0;
    a.ent[1] = 0;
    a.npoly = 1;
    a.deg = (int)n - 1;
    a.B = (int)B;
    a.M = (int)n;
    a.lwr = 0.0;
    a.lwi = (double)sign * 2.0 * 3.14159265358979323846 / (double)n;
    a.dft_n = (int)n;
    a.ybuf = (cplx *)c->ybuf.p;
    a.vhat = (cplx *)c->vhat.p;
    a.T = ctx_tw(c);
    a.mode = FNFTB_CZ_RAW;
    a.out = out;
    a.out_sstride = n;
    a.status = (int *)c->status.p;
    // the first-generation row kernels (table twiddles, ~1e-15) rather than the register-derived twiddles of the
    // fast row kernel (~1e-12): the spectral factorisation takes logarithms and exponentials of these transforms
    RC(cz_run_exact(a, (cplx *)c->cztab.p, c->st));
    return 0;
}

