// Dev probe: dependent-issue latency of FP64 / FP32 instructions for one warp alone on an SM sub-partition.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/probe_lat tools/probe_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void lat(double *out, long long *cyc, double a, double b, int iters)
{
    double v = 1.0 + threadIdx.x * 1e-3, w = v + 1.0;   // non-zero: a zero dividend sends the whole warp through the division slow path
    float f = threadIdx.x * 1e-3f;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 64; k++) {
            if (KIND == 0) v = fma(v, a, b);                       // dependent DFMA
            if (KIND == 1) { v = fma(v, a, b); w = fma(w, a, b); } // two independent chains
            if (KIND == 2) f = fmaf(f, (float)a, (float)b);        // dependent FFMA
            if (KIND == 3) v = v / a;                              // dependent DDIV (fast path)
            if (KIND == 6) v = (v * 0.0) / a;                     // DDIV with a zero dividend in some lanes: slow path
            if (KIND == 4) v = exp(v * 1e-3) - 0.5;                      // dependent exp
            if (KIND == 5) v = log(v + 2.0);                       // dependent log
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = v + w + f;
}
int main()
{
    double *out; long long *cyc; cudaMalloc(&out, 8 * 32); cudaMalloc(&cyc, 8);
    const char *names[] = {"DFMA dependent", "DFMA 2 chains (per pair)", "FFMA dependent", "DDIV dependent", "exp dependent", "log dependent", "DDIV slow path (0 / a)"};
    long long h;
    const int iters = 64;
#define RUN(K) lat<K><<<1, 32>>>(out, cyc, 0.999, 1e-3, iters); lat<K><<<1, 32>>>(out, cyc, 0.999, 1e-3, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-28s %.1f cycles per op\n", names[K], (double)h / (iters * 64));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    return 0;
}
