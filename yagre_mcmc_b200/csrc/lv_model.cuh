// lv_model.cuh -- Lotka-Volterra forward model by fixed-step classical RK4.
//
// Replaces LotkaVolterraSolver.invoke (reference yagremcmc/test/testSetup.py:109-141,
// flow__ :96-99) with the fixed-step RK4 north_star prescribes; the arithmetic
// specification is oracle/ref_harness.py RK4LotkaVolterraSolver.
//
// FP64-pipe formulation: the step size is folded into the four rate constants
// (ha = h*alpha, hb = h*beta, hd = h*delta, hg = h*gamma) and each stage uses the
// factored flow  h*fx = x*(ha - hb*y),  h*fy = y*(hd*x - hg), so one RK4 step
// of the 2-state system is exactly 30 FP64-pipe instructions (DFMA/DMUL/DADD)
// for the 58 algorithmic flop of the unfused textbook form.  The
// result differs from the oracle's unfused order by a few ulp per step; the
// parity tests bound the accumulated effect on the log-posterior (1e-10 rel).
#pragma once
#include "common.cuh"

struct LvRates {
    double ha, hb, hd, hg;
};

YG_DEVFN LvRates lv_rates(double alpha, double gamma, double T, int N, double beta, double delta)
{
    const double h = T / (double)N;
    LvRates r;
    r.ha = h * alpha;
    r.hb = h * beta;
    r.hd = h * delta;
    r.hg = h * gamma;
    return r;
}

struct LvConsts {
    double third, sixth;
};

// 1/3 and 1/6 pinned in vector registers: left to itself ptxas rematerialises them with
// four UMOVs inside the integration loop, and on the FP64-bound path every non-FP64
// instruction costs an issue slot (a DFMA holds the dispatch port for two cycles).
YG_DEVFN LvConsts lv_consts()
{
    LvConsts c;
    c.third = 1.0 / 3.0;
    c.sixth = 1.0 / 6.0;
    asm volatile("" : "+d"(c.third), "+d"(c.sixth));
    return c;
}

YG_DEVFN void lv_rk4_step(const LvRates &r, const LvConsts &k, double &x, double &y)
{
    // stage 1
    double kx = x * fma(-r.hb, y, r.ha);
    double ky = y * fma(r.hd, x, -r.hg);
    double xs = fma(0.5, kx, x), ys = fma(0.5, ky, y);
    double ax = fma(k.sixth, kx, x), ay = fma(k.sixth, ky, y);
    // stage 2
    kx = xs * fma(-r.hb, ys, r.ha);
    ky = ys * fma(r.hd, xs, -r.hg);
    xs = fma(0.5, kx, x);  ys = fma(0.5, ky, y);
    ax = fma(k.third, kx, ax); ay = fma(k.third, ky, ay);
    // stage 3
    kx = xs * fma(-r.hb, ys, r.ha);
    ky = ys * fma(r.hd, xs, -r.hg);
    xs = x + kx;  ys = y + ky;
    ax = fma(k.third, kx, ax); ay = fma(k.third, ky, ay);
    // stage 4
    kx = xs * fma(-r.hb, ys, r.ha);
    ky = ys * fma(r.hd, xs, -r.hg);
    x = fma(k.sixth, kx, ax);
    y = fma(k.sixth, ky, ay);
}

// Integrates N steps.  The caller maps non-finite END states to +inf (=> logL = -inf =>
// rejected by the unchanged acceptance rule), the policy of the oracle plugin.
YG_DEVFN void lv_integrate(const LvRates &r, int N, double &x, double &y)
{
    const LvConsts k = lv_consts();
    int i = 0;
#pragma unroll 1
    for (; i + 8 <= N; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) lv_rk4_step(r, k, x, y);
    }
#pragma unroll 1
    for (; i < N; i++) lv_rk4_step(r, k, x, y);
}

YG_DEVFN void lv_finite_or_inf(double &x, double &y)
{
    x = isfinite(x) ? x : CUDART_INF;
    y = isfinite(y) ? y : CUDART_INF;
}
