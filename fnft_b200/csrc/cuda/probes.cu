// fnft_b200 -- hardware probes used for the roofline denominators (bench.py, profiles/):
// the FP64 pipe's DFMA throughput measured on the device the context runs on.  Diagnostics only,
// nothing on the transform path calls them.
#include "fnftb_device.h"
#include <cuda_runtime.h>

// 8 independent dependent-FMA chains per thread; 2 flops per FMA
__global__ void __launch_bounds__(256) k_probe_dfma(double *out, int iters, double x, double y)
{
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        a[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                a[k] = fma(a[k], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        s += a[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" double fnftb_probe_fp64_tflops(fnftb_ctx *c)
{
    if (!c || cudaSetDevice(fnftb_ctx_device(c)) != cudaSuccess)
        return 0.0;
    cudaStream_t st = (cudaStream_t)fnftb_ctx_stream(c);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, fnftb_ctx_device(c));
    const int grid = sms * 8, nt = 256, iters = 4096;
    double *out = nullptr;
    if (cudaMalloc((void **)&out, (size_t)grid * nt * sizeof(double)) != cudaSuccess)
        return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {  // first one warms up
        cudaEventRecord(e0, st);
        k_probe_dfma<<<grid, nt, 0, st>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8 * 4 * (double)iters * grid * nt / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best)
            best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}
