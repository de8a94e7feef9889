// Dev probe: throughput of the register-resident LV RK4 step (30 FP64-pipe instructions)
// as a function of resident warps per SM and per-thread ILP, plus a pure DFMA ceiling.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/probe_rk4 tools/probe_rk4.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../yagre_mcmc_b200/csrc/lv_model.cuh"
void yg_set_error(const char *, ...) {}

template <int ILP>
__global__ void rk4_probe(double *out, int nsteps, double b0, double d0)
{
    LvRates r[ILP];
    double x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        r[i] = lv_rates(0.8, 0.4, 10.0, 512, b0 + 1e-3 * (threadIdx.x + i), d0 + 1e-3 * i);
        x[i] = 1.0 + 1e-3 * threadIdx.x; y[i] = 0.8 + 1e-3 * i;
    }
#pragma unroll 1
    for (int s = 0; s < nsteps; s += 2) {
#pragma unroll
        for (int i = 0; i < ILP; i++) lv_rk4_step(r[i], x[i], y[i]);
#pragma unroll
        for (int i = 0; i < ILP; i++) lv_rk4_step(r[i], x[i], y[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + y[i];
    if (s == 1.2345) out[0] = s;
}

template <int ILP>
double run(int sms, int blocks_per_sm, int threads, int nsteps, double *sink)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        rk4_probe<ILP><<<sms * blocks_per_sm, threads>>>(sink, nsteps, 0.4, 0.6);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1);
        if (rep > 0 && t < best) best = t;
    }
    const double instr = 30.0 * nsteps * ILP * (double)sms * blocks_per_sm * threads;
    return instr * 2.0 / (best * 1e-3) / 1e12;     // "FMA-equivalent" TFLOP/s: 2 flop per pipe instruction
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink; cudaMalloc(&sink, 8);
    printf("SMs %d\n", sms);
    printf("%8s %6s %6s %12s\n", "warps/SM", "ILP", "thr", "pipeTFLOPs");
    const int cfg[][2] = {{1, 32}, {1, 64}, {1, 128}, {1, 256}, {1, 384}, {1, 512}, {1, 768}, {1, 1024}, {2, 1024}};
    for (auto &c : cfg) {
        const int steps = 4096;
        printf("%8d %6d %6d %12.2f\n", c[0] * c[1] / 32, 1, c[1], run<1>(sms, c[0], c[1], steps, sink));
        printf("%8d %6d %6d %12.2f\n", c[0] * c[1] / 32, 2, c[1], run<2>(sms, c[0], c[1], steps, sink));
        if (c[0] * c[1] <= 1024) printf("%8d %6d %6d %12.2f\n", c[0] * c[1] / 32, 4, c[1], run<4>(sms, c[0], c[1], steps, sink));
    }
    return 0;
}
