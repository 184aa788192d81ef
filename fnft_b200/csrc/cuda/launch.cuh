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

template <class Args, void (*F)(const Args &, blk3, int, void *), int MAXT = 256, int MINB = 1>
static inline int launch_blocks(const Args &a, unsigned grid, int nt, size_t smem_bytes,
                                fnftb_stream_t st, const char *name = "")
{
    if (grid == 0)
        return 0;
    if (nt > MAXT)
        return -77;
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fnftb_kernel<Args, F, MAXT, MINB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem_bytes);
        if (e != cudaSuccess)
            return (int)e;
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
