// probe_gemm_warp_tile.cuh -- tile products of the warp-specialised tensor-path experiment (commit 285f523), kept for
// tools/probe_gemm_warp.cu: the weighted misfit of 8 (or 16) chains as a swap-AB m16n8k4 GEMM out of shared memory.
#pragma once
#include <stdint.h>
#ifndef YG_DEVFN
#define YG_DEVFN __device__ __forceinline__
#endif

YG_DEVFN void dmma_m16n8k4(double &c0, double &c1, double &c2, double &c3, double a0, double a1, double b0)
{
    asm volatile(
        "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3)
        : "d"(a0), "d"(a1), "d"(b0));
}

// Shared-memory load of a GEMM operand that ptxas may not move across its neighbours.  Left to itself ptxas gathers the
// DMMAs of one accumulator into one dependent chain (whatever the order of the PTX), and a DMMA that waits for its
// predecessor issues every 26 cycles instead of every 16.  Volatile loads keep their program order, so the operands of
// k-step i + 1 of ALL accumulator chains are fetched before the DMMAs of k-step i: gathering a chain would mean
// keeping every other chain's operands alive in registers, and the scheduler keeps the interleaved order instead
// (SASS: DMMA R28 / R32 / R36 / R24 round robin; profiles/r02_linear_dmma.md).
YG_DEVFN double lds_ordered(const double *p)
{
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

struct SmemLevel {
    const double *G;        // [np][ks]  sqrt(w_row) G_row, w = n_data * noise precision
    const double *bd;       // [np]      sqrt(w_row) (b - mean over the data rows)   (likelihood.py:74-75 broadcasts F against the rows)
    const double *pmean;    // [kp]
    const double *pprec;    // [kp]   (zero beyond dim)
    double q_const;         // sum_col prec_col * sum_rows (d_row,col - mean_col)^2
    int np;                 // rows of the operand rounded up to a multiple of 32 (zero rows)
};


// Weighted misfit of the 8 chains of a tile against one level: the GEMM warp's whole job.  b[i] = theta[chain g][4 i + t]
// is the B operand (the layout the chain warp wrote its proposal in).  On return qa / qb hold this lane's partial sums
// (rows nb + g and nb + 8 + g of every block) of chains 2t and 2t + 1.
//   sum_rows ||F - d_row||^2_P = sum_col w_col (F_col - mean_col)^2 + q_const  (exact identity; no cancellation: the row
// scatter is a precomputed constant).  Accumulator layout of m16n8k4 with G as the A operand:
// c0, c1 = F[row nb + g][chains 2t, 2t + 1], c2, c3 = F[row nb + 8 + g][the same chains].
// ONE warp has to keep the FP64 pipe of its sub-partition full (16 cycles per DMMA.8x8x4, in-order issue):
//   * a pass covers 32 rows = two m16n8k4 per k-step = four independent DMMA.8x8x4 accumulator chains;
//   * the A operands of a GROUP of two k-steps (8 loads) are fetched in one burst a whole group (8 DMMAs, 128 pipe
//     cycles) ahead of their use, into the register buffer the group before last was read from, across the pass boundary
//     too.  The __syncwarp() after the burst is there for ptxas: it may not move a shared-memory load across it.
//     Without it ptxas sinks every load to the slot right behind the DMMA that read the same registers and packs the
//     rotation into three register pairs -- three DMMA slots between a load and its use -- and a lone warp then issues
//     a DMMA every 21.7 cycles (tools/probe_gemm_warp.cu);
//   * the accumulators alternate between two register sets and the epilogue of a pass (8 DADD + 8 DFMA) is placed
//     after the first group of the following pass: nothing reads an accumulator while the pipe could stall on it.
template <int KQ>
YG_DEVFN void misfit_tile(const SmemLevel &L, const int ks, const double (&a)[KQ], const int g, const int t, double &qa, double &qb)
{
    double sa = 0.0, sb = 0.0;
    auto epilogue = [&](const int nb, const double (&c)[2][4]) {
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const double b0 = L.bd[nb + 16 * m + g], b1 = L.bd[nb + 16 * m + 8 + g];
            const double e0 = c[m][0] + b0, e1 = c[m][1] + b0, e2 = c[m][2] + b1, e3 = c[m][3] + b1;     // sqrt(w) (A @ theta + b - mean(data))
            sa = fma(e2, e2, fma(e0, e0, sa));
            sb = fma(e3, e3, fma(e1, e1, sb));
        }
    };
    const int n_pass = L.np >> 5;                  // np is a multiple of 32 (zero rows add exactly 0.0)
    const size_t stride = (size_t)32 * ks;
    const double *Gl = L.G + (size_t)g * ks + t;
    const double *Gend = Gl + (size_t)(n_pass - 1) * stride;
    constexpr int GK = 2, NGRP = KQ / GK;          // k-steps per group, groups per pass (even)
    double A[2][GK][2][2];
    auto load_group = [&](double (&X)[GK][2][2], const double *Gb, const int grp) {
#pragma unroll
        for (int k = 0; k < GK; k++)
#pragma unroll
            for (int m = 0; m < 2; m++) {
                X[k][m][0] = lds_ordered(Gb + (size_t)(16 * m) * ks + 4 * (GK * grp + k));
                X[k][m][1] = lds_ordered(Gb + (size_t)(16 * m + 8) * ks + 4 * (GK * grp + k));
            }
        __syncwarp();
    };
    load_group(A[0], Gl, 0);
    // one pass over the 32 rows at Gb into c; `mid` runs after the first group
    auto pass = [&](double (&c)[2][4], const double *Gb, auto &&mid) {
        const double *Gn = Gb == Gend ? Gb : Gb + stride;       // the last pass prefetches its own rows again (never used)
#pragma unroll
        for (int m = 0; m < 2; m++) c[m][0] = c[m][1] = c[m][2] = c[m][3] = 0.0;
#pragma unroll
        for (int grp = 0; grp < NGRP; grp++) {
            if (grp + 1 < NGRP) load_group(A[(grp + 1) & 1], Gb, grp + 1);
            else load_group(A[(grp + 1) & 1], Gn, 0);
#pragma unroll
            for (int k = 0; k < GK; k++)
#pragma unroll
                for (int m = 0; m < 2; m++)
                    dmma_m16n8k4(c[m][0], c[m][1], c[m][2], c[m][3], A[grp & 1][k][m][0], A[grp & 1][k][m][1], a[GK * grp + k]);
            if (grp == 0) mid();
        }
    };
    double ce[2][4], co[2][4];
    int p = 0;
    const double *Gb = Gl;
    while (true) {
        pass(ce, Gb, [&]() { if (p > 0) epilogue(32 * (p - 1), co); });
        Gb += stride;
        if (++p == n_pass) { epilogue(32 * (p - 1), ce); break; }
        pass(co, Gb, [&]() { epilogue(32 * (p - 1), ce); });
        Gb += stride;
        if (++p == n_pass) { epilogue(32 * (p - 1), co); break; }
    }
    qa = sa;
    qb = sb;
}


// Two tiles (16 chains) at once: every A operand feeds two m16n8k4 (one per tile), the B operands come from the tiles in
// shared memory k-step by k-step instead of living in registers.  Eight DMMA.8x8x4 accumulator chains, three loads per
// four DMMAs instead of four.
template <int KQ>
YG_DEVFN void misfit_pair(const SmemLevel &L, const int ks, const double *P0, const double *P1, const int g, const int t,
                          double &qa0, double &qb0, double &qa1, double &qb1)
{
    double s[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const int n_pass = L.np >> 5;
    const double *Gb = L.G + (size_t)g * ks + t;
    const double *B0 = P0 + g * ks + t, *B1 = P1 + g * ks + t;
    for (int p = 0; p < n_pass; p++, Gb += (size_t)32 * ks) {
        double c[2][2][4];
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int m = 0; m < 2; m++) c[q][m][0] = c[q][m][1] = c[q][m][2] = c[q][m][3] = 0.0;
        double A[2][2], An[2][2], b0, b1, b0n, b1n;
#pragma unroll
        for (int m = 0; m < 2; m++) {
            A[m][0] = lds_ordered(Gb + (size_t)(16 * m) * ks);
            A[m][1] = lds_ordered(Gb + (size_t)(16 * m + 8) * ks);
        }
        b0 = lds_ordered(B0);
        b1 = lds_ordered(B1);
#pragma unroll
        for (int i = 0; i < KQ; i++) {
            if (i + 1 < KQ) {
#pragma unroll
                for (int m = 0; m < 2; m++) {
                    An[m][0] = lds_ordered(Gb + (size_t)(16 * m) * ks + 4 * (i + 1));
                    An[m][1] = lds_ordered(Gb + (size_t)(16 * m + 8) * ks + 4 * (i + 1));
                }
                b0n = lds_ordered(B0 + 4 * (i + 1));
                b1n = lds_ordered(B1 + 4 * (i + 1));
            }
#pragma unroll
            for (int m = 0; m < 2; m++) {
                dmma_m16n8k4(c[0][m][0], c[0][m][1], c[0][m][2], c[0][m][3], A[m][0], A[m][1], b0);
                dmma_m16n8k4(c[1][m][0], c[1][m][1], c[1][m][2], c[1][m][3], A[m][0], A[m][1], b1);
            }
#pragma unroll
            for (int m = 0; m < 2; m++) { A[m][0] = An[m][0]; A[m][1] = An[m][1]; }
            b0 = b0n; b1 = b1n;
        }
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const int nb = 32 * p + 16 * m;
                const double d0 = L.bd[nb + g], d1 = L.bd[nb + 8 + g];
                const double e0 = c[q][m][0] + d0, e1 = c[q][m][1] + d0, e2 = c[q][m][2] + d1, e3 = c[q][m][3] + d1;
                s[q][0] = fma(e2, e2, fma(e0, e0, s[q][0]));
                s[q][1] = fma(e3, e3, fma(e1, e1, s[q][1]));
            }
    }
    qa0 = s[0][0]; qb0 = s[0][1]; qa1 = s[1][0]; qb1 = s[1][1];
}
