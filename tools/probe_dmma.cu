// Dev probe: are the FP64 tensor path (DMMA) and the FP64 vector pipe (DFMA) separate
// pipes that can overlap, and what is the dependent-issue latency of DMMA.8x8x4?
//   (1) one warp alone on an SM sub-partition: cycles per DMMA for 1 / 2 / 4 independent accumulator chains;
//   (2) whole GPU, 4 warps per sub-partition: DFMA alone, DMMA alone, both in the SAME warp (interleaved), both in
//       DIFFERENT warps of a sub-partition.  If the pipes were separate the mixed kernels would take max(t_dfma,
//       t_dmma); if DMMA runs on the FP64 units they take t_dfma + t_dmma.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/probe_dmma tools/probe_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void dmma16(double (&c)[4], double a0, double a1, double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}

template <int CHAINS>
__global__ void lat16(double *out, long long *cyc, double a, double b, int iters)
{
    double c[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
#pragma unroll
            for (int m = 0; m < CHAINS; m++) dmma16(c[m], a, a + 1.0, b);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = c[0][0] + c[1][1] + c[0][2] + c[1][3];
}

template <int CHAINS>
__global__ void lat(double *out, long long *cyc, double a, double b, int iters)
{
    double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
#pragma unroll
            for (int m = 0; m < CHAINS; m++) dmma(c[m][0], c[m][1], a, b);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = c[0][0] + c[1][1] + c[2][0] + c[3][1];
}

// MODE 0: DFMA only (ND per iteration per thread)   1: DMMA only (NM per iteration per warp)
//      2: both in every warp                         3: even warps DFMA, odd warps DMMA (twice the per-warp amount)
template <int MODE>
__global__ void __launch_bounds__(512, 1) mix(double *out, double a, double b, int iters)
{
    double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = 1.0 + threadIdx.x * 1e-6 + k;
    const bool odd = (threadIdx.x >> 5) & 1;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (MODE == 0 || MODE == 2 || (MODE == 3 && !odd)) {
#pragma unroll
            for (int r = 0; r < (MODE == 3 ? 16 : 8); r++)
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = fma(v[k], a, b);          // 64 (128) DFMA = 128 (256) issue cycles of the FP64 unit
        }
        if (MODE == 1 || MODE == 2 || (MODE == 3 && odd)) {
#pragma unroll
            for (int r = 0; r < (MODE == 3 ? 4 : 2); r++)
#pragma unroll
                for (int m = 0; m < 4; m++) dmma(c[m][0], c[m][1], a, b);     // 8 (16) DMMA.8x8x4 = 2048 (4096) lane-FMAs
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + c[0][0] + c[1][1] + c[2][0] + c[3][1];
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 8 * 512 * 256); cudaMalloc(&cyc, 8);
    long long h;
    const int iters = 256;
#define LAT(C) lat<C><<<1, 32>>>(out, cyc, 0.999, 1e-3, iters); lat<C><<<1, 32>>>(out, cyc, 0.999, 1e-3, iters); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("one warp, %d independent DMMA chain(s): %.1f cycles per DMMA.8x8x4\n", C, (double)h / (iters * 16 * C));
    LAT(1) LAT(2) LAT(4)
#define LAT16(C) lat16<C><<<1, 32>>>(out, cyc, 0.999, 1e-3, iters); lat16<C><<<1, 32>>>(out, cyc, 0.999, 1e-3, iters); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("one warp, %d independent m16n8k4 chain(s): %.1f cycles per mma (= two DMMA.8x8x4)\n", C, (double)h / (iters * 16 * C));
    LAT16(1) LAT16(2)
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int it2 = 20000;
    float t[4];
#define MIX(M) mix<M><<<sms, 512>>>(out, 0.999, 1e-3, it2); cudaEventRecord(e0); mix<M><<<sms, 512>>>(out, 0.999, 1e-3, it2); cudaEventRecord(e1); \
    cudaEventSynchronize(e1); cudaEventElapsedTime(&t[M], e0, e1);
    MIX(0) MIX(1) MIX(2) MIX(3)
    const double fl_dfma = 2.0 * 64 * 512.0 * sms * it2, fl_dmma = 2.0 * 256 * 8 * 16.0 * sms * it2;
    printf("DFMA only          : %.3f ms  %.2f TFLOP/s\n", t[0], fl_dfma / t[0] * 1e-9);
    printf("DMMA only          : %.3f ms  %.2f TFLOP/s\n", t[1], fl_dmma / t[1] * 1e-9);
    printf("both, same warp    : %.3f ms  %.2f TFLOP/s   (separate pipes would give %.3f ms, a shared unit %.3f ms)\n", t[2],
           (fl_dfma + fl_dmma) / t[2] * 1e-9, t[0] > t[1] ? t[0] : t[1], t[0] + t[1]);
    printf("both, other warps  : %.3f ms  %.2f TFLOP/s   (same totals as the line above)\n", t[3], (fl_dfma + fl_dmma) / t[3] * 1e-9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
