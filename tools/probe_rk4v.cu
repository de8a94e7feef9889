// Dev probe: instruction-count variants of the LV RK4 step on the FP64 pipe.
//   V30: the step of lv_model.cuh before this probe (k-form, 30 FP64 instructions per step)
//   V20: stage-point form (20 FP64 instructions per step), see lv_model.cuh
// Prints RK4 steps/s per variant and the deviation of both from a long-double textbook RK4.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/probe_rk4v tools/probe_rk4v.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

struct C30 { double ha, hg; };
struct C20 { double ha6, ha2, ha1, ha6p, mhg6, mhg2, mhg1, mhg6p; };

__device__ __forceinline__ void step30(double ha, double hg, double hb, double hd, double third, double sixth, double &x,
                                       double &y)
{
    double kx = x * fma(-hb, y, ha);
    double ky = y * fma(hd, x, -hg);
    double xs = fma(0.5, kx, x), ys = fma(0.5, ky, y);
    double ax = fma(sixth, kx, x), ay = fma(sixth, ky, y);
    kx = xs * fma(-hb, ys, ha);
    ky = ys * fma(hd, xs, -hg);
    xs = fma(0.5, kx, x); ys = fma(0.5, ky, y);
    ax = fma(third, kx, ax); ay = fma(third, ky, ay);
    kx = xs * fma(-hb, ys, ha);
    ky = ys * fma(hd, xs, -hg);
    xs = x + kx; ys = y + ky;
    ax = fma(third, kx, ax); ay = fma(third, ky, ay);
    kx = xs * fma(-hb, ys, ha);
    ky = ys * fma(hd, xs, -hg);
    x = fma(sixth, kx, ax);
    y = fma(sixth, ky, ay);
}

// mb* = -h*beta*{1/6,1/2,1}, hd* = h*delta*{1/6,1/2,1}: per-chain registers; c.*: constant bank
__device__ __forceinline__ void step20(const C20 &c, double mb6, double mb2, double mb1, double hd6, double hd2, double hd1,
                                       double &x, double &y)
{
    const double tx = x * fma(mb6, y, c.ha6);          // k1x / 6
    const double ty = y * fma(hd6, x, c.mhg6);
    const double x2 = fma(3.0, tx, x);                 // x + k1x / 2
    const double y2 = fma(3.0, ty, y);
    const double x3 = fma(x2, fma(mb2, y2, c.ha2), x); // x + k2x / 2
    const double y3 = fma(y2, fma(hd2, x2, c.mhg2), y);
    const double x4 = fma(x3, fma(mb1, y3, c.ha1), x); // x + k3x
    const double y4 = fma(y3, fma(hd1, x3, c.mhg1), y);
    const double wx = fma(mb6, y4, c.ha6p);            // 1/3 + (ha - hb y4) / 6
    const double wy = fma(hd6, x4, c.mhg6p);
    const double sx = fma(2.0 / 3.0, x3, tx);
    const double sy = fma(2.0 / 3.0, y3, ty);
    x = fma(x4, wx, sx);                               // (x2 + 2 x3 + x4 - x) / 3 + k4x / 6
    y = fma(y4, wy, sy);
}

template <int V>
__global__ void __launch_bounds__(1024, 1) probe(double *out, int nsteps, double hb0, double hd0, C30 c30, C20 c20, int dump)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const double hb = hb0 * (1.0 + 1e-3 * (threadIdx.x & 63)), hd = hd0 * (1.0 + 2e-3 * (threadIdx.x & 31));
    double x = 0.5 + (threadIdx.x & 15) / 16.0, y = 1.5 - (threadIdx.x & 7) / 8.0;
    if (V == 30) {
        double third = 1.0 / 3.0, sixth = 1.0 / 6.0;
        asm volatile("" : "+d"(third), "+d"(sixth));
#pragma unroll 1
        for (int s = 0; s < nsteps; s += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) step30(c30.ha, c30.hg, hb, hd, third, sixth, x, y);
        }
    } else {
        const double mb6 = -hb / 6.0, mb2 = -0.5 * hb, mb1 = -hb, hd6 = hd / 6.0, hd2 = 0.5 * hd;
#pragma unroll 1
        for (int s = 0; s < nsteps; s += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) step20(c20, mb6, mb2, mb1, hd6, hd2, hd, x, y);
        }
    }
    if (dump) {
        if (g < dump) { out[2 * g] = x; out[2 * g + 1] = y; }
    } else if (x + y == 1.2345) out[0] = x;
}

static void host_ref(int t, int nsteps, double h, double alpha, double gamma, double beta0, double delta0, long double &X,
                     long double &Y)
{
    const double hb = (h * beta0) * (1.0 + 1e-3 * (t & 63)), hd = (h * delta0) * (1.0 + 2e-3 * (t & 31));
    long double b = (long double)hb / h, d = (long double)hd / h, a = alpha, g = gamma, hh = h;
    long double x = 0.5 + (t & 15) / 16.0, y = 1.5 - (t & 7) / 8.0;
    auto fx = [&](long double x, long double y) { return a * x - b * x * y; };
    auto fy = [&](long double x, long double y) { return d * x * y - g * y; };
    for (int s = 0; s < nsteps; s++) {
        long double k1x = fx(x, y), k1y = fy(x, y);
        long double k2x = fx(x + hh / 2 * k1x, y + hh / 2 * k1y), k2y = fy(x + hh / 2 * k1x, y + hh / 2 * k1y);
        long double k3x = fx(x + hh / 2 * k2x, y + hh / 2 * k2y), k3y = fy(x + hh / 2 * k2x, y + hh / 2 * k2y);
        long double k4x = fx(x + hh * k3x, y + hh * k3y), k4y = fy(x + hh * k3x, y + hh * k3y);
        x += hh / 6 * (k1x + 2 * k2x + 2 * k3x + k4x);
        y += hh / 6 * (k1y + 2 * k2y + 2 * k3y + k4y);
    }
    X = x; Y = y;
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink; cudaMalloc(&sink, 8 * 4096);
    const double T = 10.0, alpha = 0.8, gamma = 0.4, beta0 = 0.4, delta0 = 0.6;
    for (int N : {512, 64}) {
        const double h = T / N;
        C30 c30{h * alpha, h * gamma};
        C20 c20{h * alpha / 6, h * alpha / 2, h * alpha, h * alpha / 6 + 1.0 / 3.0,
                -h * gamma / 6, -h * gamma / 2, -h * gamma, 1.0 / 3.0 - h * gamma / 6};
        // numerics: N steps (T = 10), 256 lanes
        double r30[512], r20[512];
        probe<30><<<1, 256>>>(sink, N, h * beta0, h * delta0, c30, c20, 256);
        cudaMemcpy(r30, sink, sizeof(r30), cudaMemcpyDeviceToHost);
        probe<20><<<1, 256>>>(sink, N, h * beta0, h * delta0, c30, c20, 256);
        cudaMemcpy(r20, sink, sizeof(r20), cudaMemcpyDeviceToHost);
        double e30 = 0, e20 = 0;
        for (int t = 0; t < 256; t++) {
            long double X, Y;
            host_ref(t, N, h, alpha, gamma, beta0, delta0, X, Y);
            e30 = fmax(e30, fmax(fabs((double)((r30[2 * t] - X) / X)), fabs((double)((r30[2 * t + 1] - Y) / Y))));
            e20 = fmax(e20, fmax(fabs((double)((r20[2 * t] - X) / X)), fabs((double)((r20[2 * t + 1] - Y) / Y))));
        }
        printf("N=%d: max rel deviation from long-double textbook RK4: V30 %.3e  V20 %.3e   (x[3]=%.17g / %.17g)\n", N, e30,
               e20, r30[6], r20[6]);
    }
    const double h = T / 512;
    C30 c30{h * alpha, h * gamma};
    C20 c20{h * alpha / 6, h * alpha / 2, h * alpha, h * alpha / 6 + 1.0 / 3.0,
            -h * gamma / 6, -h * gamma / 2, -h * gamma, 1.0 / 3.0 - h * gamma / 6};
    printf("SMs %d\n%6s %14s %14s %8s | pipe TFLOP/s (2 x instr) V30 V20\n", sms, "thr/SM", "V30 steps/s", "V20 steps/s", "ratio");
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int thr : {256, 512, 768, 1024}) {
        const int steps = 16384;
        double rate[2];
        for (int v = 0; v < 2; v++) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                if (v == 0) probe<30><<<sms, thr>>>(sink, steps, h * beta0, h * delta0, c30, c20, 0);
                else probe<20><<<sms, thr>>>(sink, steps, h * beta0, h * delta0, c30, c20, 0);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float t; cudaEventElapsedTime(&t, e0, e1);
                if (rep > 0 && t < best) best = t;
            }
            rate[v] = (double)steps * sms * thr / (best * 1e-3);
        }
        printf("%6d %14.4e %14.4e %8.3f | %8.2f %8.2f\n", thr, rate[0], rate[1], rate[1] / rate[0], rate[0] * 60e-12,
               rate[1] * 40e-12);
    }
    return 0;
}
