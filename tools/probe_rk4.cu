// Dev probe: throughput of the register-resident LV RK4 step (30 FP64-pipe instructions) as a
// function of resident threads per SM, items interleaved per thread (ILP) and of where the two
// chain-independent rate constants live (constant bank / uniform registers vs per-lane registers).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/probe_rk4 tools/probe_rk4.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../yagre_mcmc_b200/csrc/lv_model.cuh"
void yg_set_error(const char *, ...) {}

template <int ILP, bool REG_RATES>
__global__ void __launch_bounds__(1024, 1) rk4_probe(double *out, int nsteps, double b0, double d0, double ha, double hg)
{
    LvRates r[ILP];
    double x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        r[i].ha = ha; r[i].hg = hg;
        if (REG_RATES) { r[i].ha += 1e-9 * threadIdx.x; r[i].hg += 1e-9 * (threadIdx.x + i); }
        r[i].hb = b0 + 1e-6 * (threadIdx.x + i);
        r[i].hd = d0 + 1e-6 * i;
        x[i] = 1.0 + 1e-3 * threadIdx.x; y[i] = 0.8 + 1e-3 * i;
    }
    const LvConsts k = lv_consts();
#pragma unroll 1
    for (int s = 0; s < nsteps; s += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) lv_rk4_step(r[i], k, x[i], y[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + y[i];
    if (s == 1.2345) out[0] = s;
}

template <int ILP, bool REG>
double run(int sms, int threads, int nsteps, double *sink)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        rk4_probe<ILP, REG><<<sms, threads>>>(sink, nsteps, 0.4 * 10 / 512, 0.6 * 10 / 512, 0.8 * 10 / 512, 0.4 * 10 / 512);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1);
        if (rep > 0 && t < best) best = t;
    }
    const double instr = 30.0 * nsteps * ILP * (double)sms * threads;
    return instr * 2.0 / (best * 1e-3) / 1e12;     // "FMA-equivalent" TFLOP/s: 2 flop per pipe instruction
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink; cudaMalloc(&sink, 8);
    printf("SMs %d; pipe TFLOP/s (2 flop per FP64 instruction)\n", sms);
    printf("%6s %4s %10s %10s\n", "thr/SM", "ILP", "const", "regs");
    for (int thr : {128, 256, 512, 768, 1024}) {
        const int steps = 8192;
        printf("%6d %4d %10.2f %10.2f\n", thr, 1, run<1, false>(sms, thr, steps, sink), run<1, true>(sms, thr, steps, sink));
        printf("%6d %4d %10.2f %10.2f\n", thr, 2, run<2, false>(sms, thr, steps, sink), run<2, true>(sms, thr, steps, sink));
        if (thr <= 512) printf("%6d %4d %10.2f %10.2f\n", thr, 4, run<4, false>(sms, thr, steps, sink), run<4, true>(sms, thr, steps, sink));
    }
    return 0;
}
