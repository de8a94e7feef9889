// lv_model.cuh -- Lotka-Volterra forward model by fixed-step classical RK4.
//
// Replaces LotkaVolterraSolver.invoke (reference yagremcmc/test/testSetup.py:109-141,
// flow__ :96-99) with the fixed-step RK4 north_star prescribes; the arithmetic
// specification is oracle/ref_harness.py RK4LotkaVolterraSolver.
//
// FP64-pipe formulation ("stage-point form"): with the step size folded into the rates
// (ha = h*alpha, hb = h*beta, hd = h*delta, hg = h*gamma) the flow factors as
//     h*fx = x*(ha - hb*y),   h*fy = y*(hd*x - hg),
// so every RK4 stage POINT is one nested FMA pair,  x_next = fma(x_s, fma(-c*hb, y_s, c*ha), x),
// and the increments are never formed.  With x2 = x + k1/2, x3 = x + k2/2, x4 = x + k3:
//     x' = x + (k1 + 2 k2 + 2 k3 + k4)/6 = (x2 - x)/3 + (2/3) x3 + x4*(1/3 + (ha - hb*y4)/6)
// where (x2 - x)/3 = k1/6 =: t is the product the first stage computes anyway.  One RK4 step
// of the 2-state system is 20 FP64-pipe instructions (18 DFMA + 2 DMUL) for the 58 flop of the
// unfused textbook form; the k-form used before needed 30.  Six of the 20 have three
// vector-register sources (1.5 issue slots each on B200, tools/probe_rk4v.cu), the other 14 take
// the chain-independent rates as constant-bank operands: 23 issue slots against 30, measured
// 1.31x the RK4 steps/s of the k-form.  No term cancels (x' is a sum of positive parts), the
// result differs from the oracle's unfused order by a few ulp per step (3e-14 relative after 512
// steps, tools/probe_rk4v.cu); the parity tests bound the effect on the log-posterior (1e-10 rel).
#pragma once
#include "common.cuh"

// chain independent: constant-bank operands of the integration loop
struct LvStepConsts {
    double ha6, ha2, ha1, ha6p;        // h*alpha * {1/6, 1/2, 1},  h*alpha/6 + 1/3
    double mhg6, mhg2, mhg1, mhg6p;    // -h*gamma * {1/6, 1/2, 1},  1/3 - h*gamma/6
};

__host__ __device__ inline LvStepConsts lv_step_consts(double alpha, double gamma, double h)
{
    const double ha = h * alpha, hg = h * gamma;
    LvStepConsts c;
    c.ha1 = ha;  c.ha2 = 0.5 * ha;  c.ha6 = ha / 6.0;  c.ha6p = c.ha6 + 1.0 / 3.0;
    c.mhg1 = -hg; c.mhg2 = -0.5 * hg; c.mhg6 = -(hg / 6.0); c.mhg6p = 1.0 / 3.0 + c.mhg6;
    return c;
}

// per chain: (beta, delta) = exp(theta)
struct LvRates {
    double mb6, mb2, mb1;              // -h*beta * {1/6, 1/2, 1}
    double hd6, hd2, hd1;              //  h*delta * {1/6, 1/2, 1}
};

YG_DEVFN LvRates lv_rates(double h, double beta, double delta)
{
    const double hb = h * beta, hd = h * delta;
    LvRates r;
    r.mb1 = -hb; r.mb2 = -0.5 * hb; r.mb6 = -(hb * (1.0 / 6.0));
    r.hd1 = hd;  r.hd2 = 0.5 * hd;  r.hd6 = hd * (1.0 / 6.0);
    return r;
}

YG_DEVFN void lv_rk4_step(const LvStepConsts &c, const LvRates &r, double &x, double &y)
{
    const double tx = x * fma(r.mb6, y, c.ha6);               // k1x / 6
    const double ty = y * fma(r.hd6, x, c.mhg6);
    const double x2 = fma(3.0, tx, x);                        // x + k1x / 2
    const double y2 = fma(3.0, ty, y);
    const double x3 = fma(x2, fma(r.mb2, y2, c.ha2), x);      // x + k2x / 2
    const double y3 = fma(y2, fma(r.hd2, x2, c.mhg2), y);
    const double x4 = fma(x3, fma(r.mb1, y3, c.ha1), x);      // x + k3x
    const double y4 = fma(y3, fma(r.hd1, x3, c.mhg1), y);
    const double wx = fma(r.mb6, y4, c.ha6p);                 // 1/3 + (ha - hb*y4) / 6
    const double wy = fma(r.hd6, x4, c.mhg6p);
    const double sx = fma(2.0 / 3.0, x3, tx);
    const double sy = fma(2.0 / 3.0, y3, ty);
    x = fma(x4, wx, sx);
    y = fma(y4, wy, sy);
}

// Integrates N steps.  The caller maps non-finite END states to +inf (=> logL = -inf =>
// rejected by the unchanged acceptance rule), the policy of the oracle plugin.
YG_DEVFN void lv_integrate(const LvStepConsts &c, const LvRates &r, int N, double &x, double &y)
{
    int i = 0;
#pragma unroll 1
    for (; i + 8 <= N; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) lv_rk4_step(c, r, x, y);
    }
#pragma unroll 1
    for (; i < N; i++) lv_rk4_step(c, r, x, y);
}

YG_DEVFN void lv_finite_or_inf(double &x, double &y)
{
    x = isfinite(x) ? x : CUDART_INF;
    y = isfinite(y) ? y : CUDART_INF;
}
