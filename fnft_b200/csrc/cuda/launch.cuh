// fnft_b200 -- kernel launch shim.  A "block program" F(args, blockIdx, nthreads,
// smem) becomes a __global__ kernel on the device build and a pair of host loops
// on the -DFNFTB_EMUL build (test tooling, see common.cuh).
#pragma once
#include "common.cuh"

#ifdef FNFTB_EMUL
typedef void *fnftb_stream_t;
template <class Args, void (*F)(const Args &, blk3, int, void *), int MAXT = 256, int MINB = 1>
static inline int launch_blocks(const Args &a, unsigned grid, int nt, size_t smem_bytes,
                                fnftb_stream_t /*st*/, const char * /*name*/ = "")
{
    void *smem = smem_bytes ? malloc(smem_bytes) : NULL;
    for (unsigned b = 0; b < grid; ++b) {
        blk3 bid;
        bid.x = b;
        bid.y = 0;
        bid.z = 0;
        F(a, bid, nt, smem);
    }
    free(smem);
    return 0;
}
#else
typedef cudaStream_t fnftb_stream_t;
template <class Args, void (*F)(const Args &, blk3, int, void *), int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) fnftb_kernel(const Args a)
{
    extern __shared__ double2 fnftb_smem[];
    F(a, blockIdx, (int)blockDim.x, (void *)fnftb_smem);
}

// number of kernel launches issued by this library (reported by bench.py)
#include <atomic>
extern std::atomic<unsigned long long> g_fnftb_launch_count;  // contexts are per host thread
// optional per-launch timing with CUDA events (device_api.cu); name is a literal
void fnftb_profile_begin(const char *name, cudaStream_t st);
void fnftb_profile_end(cudaStream_t st);
extern int g_fnftb_profile_on;

// Opt-in to more than 48 KiB of dynamic shared memory, ONCE per (device, kernel) and to the device maximum.
// Setting the attribute to the size of the current launch on every launch is a race when several host threads (the
// per-device workers of fnft_b200_set_devices, possibly several on one GPU) launch the same kernel with different
// sizes: a smaller value set by one thread between another thread's set and its launch makes that launch fail with
// cudaErrorInvalidValue.  The maximum costs nothing: occupancy follows the size given at launch.
#include <mutex>
#include <set>
#include <utility>
static inline int fnftb_smem_optin(const void *kernel, size_t smem_bytes)
{
    if (smem_bytes <= 48 * 1024)
        return 0;
    static std::mutex mu;
    static std::set<std::pair<int, const void *>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return (int)e;
    std::lock_guard<std::mutex> lk(mu);
    if (done.count(std::make_pair(dev, kernel)))
        return 0;
    int maxs = 0;
    e = cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess)
        return (int)e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs);
    if (e != cudaSuccess)
        return (int)e;
    done.insert(std::make_pair(dev, kernel));
    return 0;
}

template <class Args, void (*F)(const Args &, blk3, int, void *), int MAXT = 256, int MINB = 1>
static inline int launch_blocks(const Args &a, unsigned grid, int nt, size_t smem_bytes,
                                fnftb_stream_t st, const char *name = "")
{
    if (grid == 0)
        return 0;
    if (nt > MAXT)
        return -77;
    {
        const int e = fnftb_smem_optin((const void *)fnftb_kernel<Args, F, MAXT, MINB>, smem_bytes);
        if (e != 0)
            return e;
    }
    if (g_fnftb_profile_on)
        fnftb_profile_begin(name, st);
    fnftb_kernel<Args, F, MAXT, MINB><<<grid, nt, smem_bytes, st>>>(a);
    if (g_fnftb_profile_on)
        fnftb_profile_end(st);
    ++g_fnftb_launch_count;
    return (int)cudaGetLastError();
}
#endif
