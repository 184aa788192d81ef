// Micro-benchmark (diagnostics, not product): how fast can CTAs shaped like the X stage of k_up_smem read
// 4 operand streams spaced `stride` elements apart and write 1 output stream, as a function of the stream
// spacing, the CTA shape and the CTAs resident per SM?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_probe stream_probe.cu && ./stream_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef double2 cplx;

template <int G, bool STORE, bool NOALLOC>
__global__ void k_probe(const cplx *in, cplx *out, int N, long long stride, int nblk_per_item)
{
    extern __shared__ double2 sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const long long item = blockIdx.x;
    const cplx *a0 = in + item * 4 * stride;  // 4 streams: a0, a0+stride, ...
    cplx *o = out + item * (long long)N;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    cplx acc = make_double2(0.0, 0.0);
    for (int cb = warp * 32 * G; cb < N; cb += nwarps * 32 * G) {
        cplx x[G], y[G], z[G], w[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int pos = cb + 32 * i + lane;
            if (NOALLOC) {
                asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(x[i].x), "=d"(x[i].y) : "l"(a0 + pos));
                asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(y[i].x), "=d"(y[i].y) : "l"(a0 + stride + pos));
                asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(z[i].x), "=d"(z[i].y) : "l"(a0 + 2 * stride + pos));
                asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(w[i].x), "=d"(w[i].y) : "l"(a0 + 3 * stride + pos));
            } else {
                x[i] = __ldg(a0 + pos);
                y[i] = __ldg(a0 + stride + pos);
                z[i] = __ldg(a0 + 2 * stride + pos);
                w[i] = __ldg(a0 + 3 * stride + pos);
            }
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            cplx r = make_double2(x[i].x * z[i].x - y[i].y * w[i].y, x[i].x * z[i].y + y[i].x * w[i].x);
            if (STORE)
                o[cb + 32 * i + lane] = r;
            else {
                acc.x += r.x;
                acc.y += r.y;
            }
            sm[(cb + 32 * i + lane) & 1023] = r;
        }
    }
    if (!STORE && acc.x == 1.2345e300)
        o[tid] = acc;
}

template <int G, bool STORE, bool NOALLOC>
static double run(const cplx *in, cplx *out, int N, long long stride, int nt, int smem, long long items)
{
    cudaFuncSetAttribute(k_probe<G, STORE, NOALLOC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_probe<G, STORE, NOALLOC><<<(unsigned)items, nt, smem>>>(in, out, N, stride, 1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best)
            best = ms;
    }
    if (cudaGetLastError() != cudaSuccess)
        return -1.0;
    const double bytes = (double)items * N * 16.0 * (4 + (STORE ? 1 : 0));
    return bytes / (best * 1e-3) / 1e12;
}

int main()
{
    const size_t total = (size_t)6 << 30;  // 6 GiB of input
    cplx *in, *out;
    cudaMalloc(&in, total + (64 << 20));
    cudaMalloc(&out, total / 4 + (64 << 20));
    cudaMemset(in, 0, total);
    printf("%-8s %-10s %-5s %-8s %-6s | read-only TB/s (ldg / noalloc) | read+store TB/s (ldg / noalloc)\n", "N", "stride", "nt", "smemKB", "G");
    const int Ns[3] = {2048, 4096, 8192};
    for (int in_ = 0; in_ < 3; ++in_) {
        const int N = Ns[in_];
        const int nt = N / 32;
        const int smemKB = N * 16 / 1024;
        for (int pad = 0; pad < 3; ++pad) {
            const long long stride = N + (pad == 0 ? 0 : (pad == 1 ? 16 : 80));
            const long long items = (long long)(total / 16) / (4 * stride);
            const double r0 = run<4, false, false>(in, out, N, stride, nt, smemKB * 1024, items);
            const double r1 = run<4, false, true>(in, out, N, stride, nt, smemKB * 1024, items);
            const double s0 = run<4, true, false>(in, out, N, stride, nt, smemKB * 1024, items);
            const double s1 = run<4, true, true>(in, out, N, stride, nt, smemKB * 1024, items);
            printf("%-8d %-10lld %-5d %-8d %-6d | %6.2f %6.2f | %6.2f %6.2f\n", N, stride, nt, smemKB, 4, r0, r1, s0, s1);
        }
        // same shape with 2x the threads, and with small shared memory (many CTAs per SM)
        const long long items = (long long)(total / 16) / (4 * (long long)N);
        printf("%-8d %-10d %-5d %-8d %-6d | %6.2f %6.2f | %6.2f %6.2f   (2x threads)\n", N, N, 2 * nt, smemKB, 4,
               run<4, false, false>(in, out, N, N, 2 * nt, smemKB * 1024, items), run<4, false, true>(in, out, N, N, 2 * nt, smemKB * 1024, items),
               run<4, true, false>(in, out, N, N, 2 * nt, smemKB * 1024, items), run<4, true, true>(in, out, N, N, 2 * nt, smemKB * 1024, items));
        printf("%-8d %-10d %-5d %-8d %-6d | %6.2f %6.2f | %6.2f %6.2f   (16 KB smem)\n", N, N, nt, 16, 4,
               run<4, false, false>(in, out, N, N, nt, 16 * 1024, items), run<4, false, true>(in, out, N, N, nt, 16 * 1024, items),
               run<4, true, false>(in, out, N, N, nt, 16 * 1024, items), run<4, true, true>(in, out, N, N, nt, 16 * 1024, items));
        printf("%-8d %-10d %-5d %-8d %-6d | %6.2f %6.2f | %6.2f %6.2f   (G = 8)\n", N, N, nt, smemKB, 8,
               run<8, false, false>(in, out, N, N, nt, smemKB * 1024, items), run<8, false, true>(in, out, N, N, nt, smemKB * 1024, items),
               run<8, true, false>(in, out, N, N, nt, smemKB * 1024, items), run<8, true, true>(in, out, N, N, nt, smemKB * 1024, items));
    }
    return 0;
}
