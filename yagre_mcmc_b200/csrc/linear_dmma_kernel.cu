// linear_dmma_kernel.cu -- Metropolis-Hastings over an ensemble of chains on the LINEAR model
// F = G theta + b when the parameter / data dimensions are too large for one chain per thread
// (d up to 64, data_dim up to 256): the forward model of a tile of chains is a dense FP64 GEMM
// [chains x d] . [d x data_dim], issued on the FP64 tensor path (DMMA, mma.sync m16n8k4.f64),
// with the Gaussian-misfit log-likelihood fused into the accumulator epilogue.
//
// Reference semantics restated (rkutri/yagre-mcmc):
//   linear forward       exampleSetup.py:42-52 (A @ theta + b, broadcast against the data rows)
//   likelihood / prior   statistics/likelihood.py:33-39,74-84, statistics/gaussian.py:19-24,
//                        statistics/covariance.py:19-22,54-55 (diagonal precisions)
//   proposal             statistics/gaussian.py:61-66 with a diagonal factor (covariance.py:51-52)
//   MRW / MLDA ratios    chain/method/mrw.py:51-57, chain/method/mlda.py:100-110,146-154
//   step loop            chain/metropolisHastings.py:55-120
//   Welford diagnostics  chain/diagnostics.py:91-94, statistics/estimation.py:36-53 (diagonal M2)
//
// B200 mapping
//   * persistent CTAs (one per SM), 16 warps; a warp owns a tile of 8 chains (the M extent of
//     m8n8k4 = the native DMMA.8x8x4) for ALL n_steps: 8-row tiles halve the registers per thread
//     (d = 64: 32 per operand copy), so 16 warps are resident and the non-GEMM work of a warp
//     (noise, accept, Welford) overlaps the GEMMs of three others on its sub-partition.  Proposals live in registers in A-fragment layout: lane
//     (g = lane / 4, t = lane % 4) holds p[row g][k = 4 i + t], so a proposal is already
//     the A operand of the GEMM; the current state of the tile sits in a per-warp shared-memory
//     tile in the same lane-private pattern (d = 64 needs 64 registers per operand copy);
//   * G of every level is staged once per CTA into shared memory (row stride = 4 mod 16 doubles:
//     the B-fragment loads of a warp are bank-conflict free) and shared by the 16 warps: one 8-byte
//     load per lane and DMMA, half of the shared-memory bandwidth at full DMMA rate;
//   * epilogue per 8x8 accumulator tile: + b, - data row, * noise precision, squared, summed
//     per chain; a 4-lane butterfly finishes the row sums, so the four lanes of a chain hold
//     bit-identical log-posteriors and take the same accept decision without further traffic;
//   * Philox noise is keyed like the one-chain-per-thread kernels (seed, global chain id, step,
//     sub-step, pair); the Box-Muller transform of this kernel runs in FP32 (see
//     philox_normal_pair_f32): d normals per chain-step make the transform, not the GEMM, the
//     limiter otherwise.
#include "ensemble.h"
#include "big_linear.h"
#include <math_constants.h>

namespace {

YG_DEVFN void dmma_m8n8k4(double &c0, double &c1, double a0, double b0)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a0), "d"(b0));
}

YG_DEVFN void dmma_m16n8k4(double &c0, double &c1, double &c2, double &c3, double a0, double a1, double b0)
{
    asm volatile(
        "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3)
        : "d"(a0), "d"(a1), "d"(b0));
}

// Box-Muller on one Philox block with the TRANSFORM in FP32: the two uniforms keep their 53 Philox bits, but
// log / sqrt / sincospi run on the FP32 pipe (a few dozen FMA-pipe instructions instead of ~150 FP64 ones per
// pair).  The normals are exact N(0,1) draws up to a relative perturbation of ~1e-7 -- immaterial for a proposal
// distribution -- and a recorded stream replays bit-exactly.  This kernel draws d normals per chain and step
// (64 at d = 64) against 2 in the LV kernel, where the FP64 transform stays.
// Not inlined on purpose: inlined copies (Philox rounds + logf + sincospif) per proposal made the step loop
// larger than the instruction cache (stall reason no_instruction 2.9 warps per issue in the ncu capture).
__device__ __noinline__ void philox_normal_pair_f32(uint64_t seed, uint64_t chain, uint64_t step, uint32_t sub, uint32_t b,
                                                    double &z0, double &z1)
{
    const uint4 w = philox_block(seed, chain, step, sub, b);
    const float u1 = (float)(u53(w.x, w.y) + 0x1.0p-53);   // (0,1]
    const float u2 = (float)u53(w.z, w.w);
    const float R = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    z0 = (double)(R * cs);
    z1 = (double)(R * sn);
}

YG_DEVFN double quad_sum(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

struct SmemLevel {
    const double *G;        // [np][ks]
    const double *bd;       // [np]   b - mean over the data rows (likelihood.py:74-75 broadcasts F against the rows)
    const double *nw;       // [np]   n_data * noise precision (zero beyond data_dim)
    const double *pmean;    // [kp]
    const double *pprec;    // [kp]   (zero beyond dim)
    double q_const;         // sum_col prec_col * sum_rows (d_row,col - mean_col)^2
    int np;
};

// log-posterior of the chain (row g of the warp's 8 x d tile) whose parameters are spread over the quad:
// a[i] = theta[4 i + t].  Every lane of a quad returns the same value.
template <int KQ>
YG_DEVFN double logpost_tile(const SmemLevel &L, const int ks, const double (&a)[KQ], const int g, const int t)
{
    // sum_rows ||F - d_row||^2_P = sum_col n P_col (F_col - mean_col)^2 + q_const  (exact identity;
    // no cancellation: the row scatter is a precomputed constant)
    double q = 0.0;
    auto epilogue = [&](const int nb, const double c0, const double c1) {
        const int col = nb + 2 * t;                                      // even: one 16-byte load per operand pair
        const double2 bd = *reinterpret_cast<const double2 *>(L.bd + col);
        const double2 nw = *reinterpret_cast<const double2 *>(L.nw + col);
        const double e0 = c0 + bd.x, e1 = c1 + bd.y;                    // A @ theta + b - mean(data)
        q = fma(nw.y * e1, e1, fma(nw.x * e0, e0, q));
    };
    // four independent accumulator tiles per pass: a chain of dependent DMMAs alone cannot fill the pipe
    int nb = 0;
    for (; nb + 32 <= L.np; nb += 32) {
        double c[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        const double *Gb = L.G + (size_t)(nb + g) * ks + t;
#pragma unroll
        for (int i = 0; i < KQ; i++) {
#pragma unroll
            for (int m = 0; m < 4; m++) dmma_m8n8k4(c[m][0], c[m][1], a[i], Gb[(size_t)(8 * m) * ks + 4 * i]);
        }
#pragma unroll
        for (int m = 0; m < 4; m++) epilogue(nb + 8 * m, c[m][0], c[m][1]);
    }
    for (; nb < L.np; nb += 8) {
        double c0 = 0.0, c1 = 0.0;
        const double *Gb = L.G + (size_t)(nb + g) * ks + t;
#pragma unroll
        for (int i = 0; i < KQ; i++) dmma_m8n8k4(c0, c1, a[i], Gb[4 * i]);
        epilogue(nb, c0, c1);
    }
    double pr = 0.0;
#pragma unroll
    for (int i = 0; i < KQ; i++) {
        const double x = a[i] - L.pmean[4 * i + t];
        pr = fma(L.pprec[4 * i + t] * x, x, pr);
    }
    return -0.5 * (quad_sum(q) + L.q_const) + (-0.5 * quad_sum(pr));
}

constexpr int BIG_WARPS = 16;      // warps per CTA: 8 chains each, <= 128 registers per thread

// FREE_NOISE = true: the production instance (Philox noise only).  The injected / recorded noise paths of the
// parity tests live in the FREE_NOISE = false instance, which keeps the step loop of the production one small
// enough for the instruction cache (stall reason no_instruction in profiles/r01_linear_dmma.md).
template <int KQ, bool TWO_LEVEL, bool FREE_NOISE>
__global__ void __launch_bounds__(BIG_WARPS * 32, 1) linear_dmma_mh_kernel(const RunArgs a, const DevBigHeader *gh)
{
    const int noise_mode = FREE_NOISE ? (int)YG_NOISE_PHILOX : a.noise_mode;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3, odd = t & 1;
    // ---- stage the problem into shared memory -------------------------------------------------
    const DevBigHeader H = *gh;
    const double *gtail = reinterpret_cast<const double *>(gh + 1);
    for (int i = tid; i < H.tail_len; i += blockDim.x) smem[i] = gtail[i];
    __syncthreads();
    SmemLevel Lv[2];
#pragma unroll
    for (int l = 0; l < 2; l++) {
        Lv[l].G = smem + H.lvl[l].G_off;
        Lv[l].bd = smem + H.lvl[l].bd_off;
        Lv[l].nw = smem + H.lvl[l].nw_off;
        Lv[l].pmean = smem + H.lvl[l].pmean_off;
        Lv[l].pprec = smem + H.lvl[l].pprec_off;
        Lv[l].q_const = H.lvl[l].q_const;
        Lv[l].np = H.lvl[l].np;
    }
    const double *propL = smem + H.propL_off;      // [kp] diagonal proposal factor (zero beyond dim)
    // current state of the warp's 8 chains: [8][ks] doubles after the problem blob; lane (g, t) only ever
    // touches its own slots (row g, columns 4 i + t), so no synchronisation is needed, and the row stride
    // (4 mod 16 doubles) makes the accesses bank-conflict free
    const int d = H.dim, ks = H.ks, J = TWO_LEVEL ? H.J : 1, n_lvl = TWO_LEVEL ? 2 : 1;
    double *ths = smem + ((H.tail_len + 1) & ~1) + (size_t)warp * 8 * ks;
#define TH(i) ths[g * ks + 4 * (i) + t]
    const int64_t N = a.n_chains;
    unsigned long long cnt_acc = 0ull, cnt_ev0 = 0ull, cnt_ev1 = 0ull, cnt_tr = 0ull;

    const int64_t n_tiles = (N + 7) / 8;
    for (int64_t tile = (int64_t)blockIdx.x * BIG_WARPS + warp; tile < n_tiles; tile += (int64_t)gridDim.x * BIG_WARPS) {
        // this lane's chain (a row beyond n_chains is computed but never stored)
        const int64_t gr = tile * 8 + g;
        const bool live = gr < N;
        const int64_t gg = live ? gr : 0;
        const uint64_t gid = (uint64_t)(a.chain_offset + gg);
#pragma unroll 1
        for (int i = 0; i < KQ; i++) {
            const int k = 4 * i + t;
            TH(i) = (k < d) ? a.theta[(int64_t)k * N + gg] : 0.0;
        }
        double lp0 = a.logpost[gg], lp1 = TWO_LEVEL ? a.logpost[N + gg] : 0.0;
        unsigned long long nacc = a.n_accept[gg];

        // Welford (estimation.py:36-53, diagonal M2) in run-length form: a chain that stays at x for m
        // consecutive steps contributes  n' = n + m, mean' = mean + (x - mean) m / n',
        // M2' = M2 + (x - mean)^2 n m / n'  -- algebraically the m sequential updates.  With the
        // accumulators in (L2-resident) global memory this touches them once per accepted move
        // instead of once per step; the registers stay with the GEMM operands.  Loads of a batch of
        // columns are issued before any store (a store may alias the next load for the compiler).
        double run = 0.0, wn = (double)a.welford_n0;
        auto welford_flush = [&](const bool f) {
            const bool fl = f && live && run != 0.0;
            if (!__any_sync(0xffffffffu, fl)) return;
            const double n1 = wn + run, c1 = run / n1, c2 = wn * c1;
            constexpr int B = KQ < 8 ? KQ : 8;
#pragma unroll 1
            for (int i0 = 0; i0 < KQ; i0 += B) {
                double m0[B], v0[B];
#pragma unroll
                for (int i = 0; i < B; i++) {
                    const int k = 4 * (i0 + i) + t;
                    const bool on = fl && k < d;
                    m0[i] = on ? a.w_mean[(int64_t)k * N + gr] : 0.0;
                    v0[i] = on ? a.w_m2[(int64_t)big_w2_index(k, d) * N + gr] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < B; i++) {
                    const int k = 4 * (i0 + i) + t;
                    if (fl && k < d) {
                        const double dl = TH(i0 + i) - m0[i];
                        a.w_mean[(int64_t)k * N + gr] = fma(dl, c1, m0[i]);
                        a.w_m2[(int64_t)big_w2_index(k, d) * N + gr] = fma(dl * dl, c2, v0[i]);
                    }
                }
            }
            if (fl) { wn += run; run = 0.0; }
        };

        // p = s + L z with a diagonal L, unfused like numpy; z keyed like the per-thread kernels: pair b of
        // sub-step j gives z[2b], z[2b+1].  Columns 4i+t and 4i+(t^1) of a lane pair are the two halves of
        // pair b = (4i + (t & ~1)) / 2: the even lane draws the pairs of even i, the odd lane those of odd
        // i, and they swap the halves they do not own -- one Philox block + one Box-Muller per lane and
        // 4 parameters.
        auto propose = [&](auto &&src, int64_t n, int j, double (&p)[KQ]) {
            bool same = true;
#pragma unroll
            for (int i2 = 0; i2 < KQ; i2 += 2) {
                double zv[2] = {0.0, 0.0};                 // z of columns 4 i2 + t and 4 (i2 + 1) + t
                if (noise_mode == YG_NOISE_INJECT) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int k = 4 * (i2 + e) + t;
                        if (k < d && i2 + e < KQ) zv[e] = a.z[((n * J + j) * d + k) * N + gg];
                    }
                } else {
                    const int i_mine = i2 + odd;           // the i whose pair this lane draws
                    double z0, z1;
                    philox_normal_pair_f32(a.seed, gid, (uint64_t)(a.step0 + n), (uint32_t)j,
                                           (uint32_t)((4 * i_mine + (t & ~1)) >> 1), z0, z1);
                    // even lane keeps z0 of its pair (column 4 i2 + t), needs z0 of the odd lane's pair (i2 + 1);
                    // odd lane keeps z1 of its pair (column 4 (i2+1) + t), needs z1 of the even lane's pair (i2)
                    const double recv = __shfl_xor_sync(0xffffffffu, odd ? z0 : z1, 1);
                    zv[0] = odd ? recv : z0;
                    zv[1] = odd ? z1 : recv;
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int k = 4 * (i2 + e) + t;
                        if (k >= d || i2 + e >= KQ) zv[e] = 0.0;
                        else if (noise_mode == YG_NOISE_RECORD && live) a.z[((n * J + j) * d + k) * N + gr] = zv[e];
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    if (i2 + e < KQ) {
                        const double sv = src(i2 + e);
                        p[i2 + e] = __dadd_rn(sv, __dmul_rn(propL[4 * (i2 + e) + t], zv[e]));
                        same = same && (p[i2 + e] == sv);
                    }
                }
            }
            // parameter/vector.py:37-45: equal iff every coordinate is equal (all four lanes agree)
            const unsigned m = __ballot_sync(0xffffffffu, same);
            return ((m >> (4 * g)) & 0xFu) == 0xFu;
        };

        int64_t thin_left = a.thin, thin_out = -1;
        for (int64_t n = 0; n < a.n_steps; n++) {
            const uint64_t step = (uint64_t)(a.step0 + n);
            const bool store_now = (--thin_left == 0);     // (n + 1) % thin == 0 without a 64-bit division per step
            if (store_now) { thin_left = a.thin; thin_out++; }
            // FullDiagnostics: Welford of the pre-transition state (diagnostics.py:91-94): the state of
            // this step is seen once more; the accumulators are touched when the state changes.
            run += 1.0;
            bool accepted = false;
            if (!TWO_LEVEL) {
                double p[KQ];
                const bool eq = propose([&](int i) { return TH(i); }, n, 0, p);
                const double lpp = logpost_tile<KQ>(Lv[0], ks, p, g, t);
                if (!eq) {                                              // metropolisHastings.py:60-61
                    if (t == 0 && live) cnt_ev0++;
                    double u;
                    if (noise_mode == YG_NOISE_INJECT) u = a.u_f[n * N + gg];
                    else {
                        u = philox_uniform(a.seed, gid, step, YG_SUB_FINE);
                        if (noise_mode == YG_NOISE_RECORD && live && t == 0) a.u_f[n * N + gg] = u;
                    }
                    accepted = accept_rule(lpp - lp0, u);
                }
                welford_flush(accepted);
                if (accepted) {
#pragma unroll
                    for (int i = 0; i < KQ; i++) TH(i) = p[i];
                    lp0 = lpp;
                }
            } else {
                double s[KQ], p[KQ], lps = lp0;
#pragma unroll
                for (int i = 0; i < KQ; i++) s[i] = TH(i);
                for (int j = 0; j < J; j++) {                           // coarse sub-chain, mlda.py:100-110
                    const bool eq = propose([&](int i) { return s[i]; }, n, j, p);
                    const double lpp = logpost_tile<KQ>(Lv[0], ks, p, g, t);
                    if (eq) continue;
                    if (t == 0 && live) cnt_ev0++;
                    const int64_t ui = (n * J + j) * N + gg;
                    double u;
                    if (noise_mode == YG_NOISE_INJECT) u = a.u_c[ui];
                    else {
                        u = philox_uniform(a.seed, gid, step, (uint32_t)j);
                        if (noise_mode == YG_NOISE_RECORD && live && t == 0) a.u_c[ui] = u;
                    }
                    if (accept_rule(lpp - lps, u)) {
#pragma unroll
                        for (int i = 0; i < KQ; i++) s[i] = p[i];
                        lps = lpp;
                    }
                }
                // the sub-chain's end point is the proposal; no fine evaluation for a chain that did
                // not move (metropolisHastings.py:60-61).  The GEMM is warp wide: it runs when ANY of
                // the 8 chains moved, and only the chains that moved use (and count) its result.
                bool same = true;
#pragma unroll
                for (int i = 0; i < KQ; i++) same = same && (s[i] == TH(i));
                const unsigned m = __ballot_sync(0xffffffffu, same);
                const bool moved = (((m >> (4 * g)) & 0xFu) != 0xFu) && live;
                double lpf = 0.0;
                if (__any_sync(0xffffffffu, moved)) {
                    lpf = logpost_tile<KQ>(Lv[1], ks, s, g, t);
                    if (moved) {
                        if (t == 0) cnt_ev1++;
                        double u;
                        if (noise_mode == YG_NOISE_INJECT) u = a.u_f[n * N + gg];
                        else {
                            u = philox_uniform(a.seed, gid, step, YG_SUB_FINE);
                            if (noise_mode == YG_NOISE_RECORD && t == 0) a.u_f[n * N + gg] = u;
                        }
                        const double delta = lpf + lp0 - lps - lp1;     // mlda.py:148-152, this order
                        accepted = accept_rule(delta, u);
                    }
                }
                welford_flush(accepted);
                if (accepted) {
#pragma unroll
                    for (int i = 0; i < KQ; i++) TH(i) = s[i];
                    lp0 = lps;
                    lp1 = lpf;
                }
            }
            if (live) {
                if (accepted) { nacc++; if (t == 0) cnt_acc++; }
                if (t == 0) {
                    cnt_tr++;
                    if (a.accepted) a.accepted[n * N + gr] = accepted ? 1 : 0;
                }
                if (store_now) {
                    const int64_t o = thin_out;
                    if (a.samples) {
#pragma unroll 1
                        for (int i = 0; i < KQ; i++)
                            if (4 * i + t < d) a.samples[(o * d + 4 * i + t) * N + gr] = TH(i);
                    }
                    if (a.lp_out && t == 0) {
                        a.lp_out[(o * n_lvl) * N + gr] = lp0;
                        if (TWO_LEVEL) a.lp_out[(o * n_lvl + 1) * N + gr] = lp1;
                    }
                }
            }
        }
        // ---- store chain state ------------------------------------------------------------------
        welford_flush(true);
        if (live) {
#pragma unroll 1
            for (int i = 0; i < KQ; i++) {
                const int k = 4 * i + t;
                if (k < d) a.theta[(int64_t)k * N + gr] = TH(i);
            }
            if (t == 0) {
                a.logpost[gr] = lp0;
                if (TWO_LEVEL) a.logpost[N + gr] = lp1;
                a.n_accept[gr] = nacc;
            }
        }
    }
#undef TH
    // ---- counters: warp-shuffle reduction, one atomic per warp ----------------------------------
    unsigned long long v[4] = {cnt_tr, cnt_acc, cnt_ev0, cnt_ev1};
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0 && v[k]) atomicAdd(&a.counters[k], v[k]);
    }
}

// log-posterior of arbitrary points, one chain per thread, plain loops over global memory
// (yg_set_state / yg_logpost: outside the step loop).
__global__ void big_logpost_kernel(const DevBigHeader *gh, int lvl, const double *theta, int64_t n, double *out)
{
    const DevBigHeader &H = *gh;
    const double *tail = reinterpret_cast<const double *>(gh + 1);
    const BigLevel &L = H.lvl[lvl];
    const double *G = tail + L.G_off, *bd = tail + L.bd_off, *nw = tail + L.nw_off;
    const double *pm = tail + L.pmean_off, *pw = tail + L.pprec_off;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int col = 0; col < L.data_dim; col++) {
            double f = 0.0;
            for (int k = 0; k < H.dim; k++) f = fma(G[(size_t)col * H.ks + k], theta[(int64_t)k * n + c], f);
            const double e = f + bd[col];
            s = fma(nw[col] * e, e, s);
        }
        s += L.q_const;
        double p = 0.0;
        for (int k = 0; k < H.dim; k++) {
            const double x = theta[(int64_t)k * n + c] - pm[k];
            p = fma(pw[k] * x, x, p);
        }
        out[c] = -0.5 * s + (-0.5 * p);
    }
}

// FP64 tensor-path micro-benchmark: independent m16n8k4 accumulator chains, 8 warps per SM.
__global__ void __launch_bounds__(256, 1) dmma_peak_kernel(double *sink, int iters, double x)
{
    double c[8][4];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0.0;
    const double a0 = x + threadIdx.x * 1e-9, a1 = x - threadIdx.x * 1e-9, b0 = 1.0 + 1e-9 * threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) dmma_m16n8k4(c[k][0], c[k][1], c[k][2], c[k][3], a0, a1, b0);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    if (s == 123.456) sink[0] = s;
}

template <int KQ>
int launch_t(yg_ensemble *e, const RunArgs &a, cudaStream_t st)
{
    const DevBigHeader *hh = reinterpret_cast<const DevBigHeader *>(e->h_problem.data());
    const size_t smem = sizeof(double) * ((((size_t)hh->tail_len + 1) & ~size_t(1)) + (size_t)BIG_WARPS * 8 * hh->ks);
    const bool free_noise = a.noise_mode == YG_NOISE_PHILOX;
    auto kern = e->cfg.n_levels == 2
                    ? (free_noise ? linear_dmma_mh_kernel<KQ, true, true> : linear_dmma_mh_kernel<KQ, true, false>)
                    : (free_noise ? linear_dmma_mh_kernel<KQ, false, true> : linear_dmma_mh_kernel<KQ, false, false>);
    YG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (a.n_chains + 7) / 8;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((tiles + BIG_WARPS - 1) / BIG_WARPS, e->sm_count));
    kern<<<grid, BIG_WARPS * 32, smem, st>>>(a, reinterpret_cast<const DevBigHeader *>(e->d_problem));
    YG_CUDA_CHECK(cudaGetLastError());
    e->last_grid = grid;
    e->last_block = BIG_WARPS * 32;
    e->last_smem = (int)smem;
    e->launches += 1;
    return YG_OK;
}

}  // namespace

int yg_launch_linear_big(yg_ensemble *e, const RunArgs &a, cudaStream_t st)
{
    const DevBigHeader *hh = reinterpret_cast<const DevBigHeader *>(e->h_problem.data());
    switch (hh->kp / 4) {
    case 1: case 2: case 3: case 4: return launch_t<4>(e, a, st);
    case 5: case 6: case 7: case 8: return launch_t<8>(e, a, st);
    default: return launch_t<16>(e, a, st);
    }
}

int yg_launch_logpost_big(yg_ensemble *e, int level, const double *theta, int64_t n, double *out, cudaStream_t st)
{
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, (int64_t)e->sm_count * 16));
    big_logpost_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<const DevBigHeader *>(e->d_problem), level, theta, n, out);
    YG_CUDA_CHECK(cudaGetLastError());
    e->launches += 1;
    return YG_OK;
}

int yg_dmma_peak(int device, double ms, double *tflops_out)
{
    YG_CUDA_CHECK(cudaSetDevice(device));
    int sms = 148;
    YG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double *sink = nullptr;
    YG_CUDA_CHECK(cudaMalloc((void **)&sink, 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 2000;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        dmma_peak_kernel<<<sms, 256>>>(sink, iters, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        const double flop = 2.0 * 16 * 8 * 4 * 8.0 * iters * 8.0 * sms;       // per mma x 8 chains x 8 warps x SMs
        if (rep > 0) best = std::max(best, flop / (t * 1e-3) / 1e12);
        if (t < ms / 6.0) iters *= 2;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    YG_CUDA_CHECK(cudaGetLastError());
    *tflops_out = best;
    return YG_OK;
}
