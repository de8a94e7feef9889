// abi.cu -- the extern "C" surface of libyagre_b200.so (include/yagre_b200.h).
// Plain pointers and sizes only; no torch types, no exceptions across the boundary.
#include "ensemble.h"
#include "big_linear.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

int yg_pooled_impl(const double *w_mean, const double *w_m2, const unsigned long long *n_accept, int d, int64_t nc,
                   int64_t welford_n, double *partials, int n_part, double *out_dev, cudaStream_t st);

static thread_local char g_err[512] = "";

void yg_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

constexpr int POOL_PARTS = 64;

int check_handle(const yg_ensemble *e, bool need_problem, bool need_state)
{
    if (!e) {
        yg_set_error("null ensemble handle");
        return YG_ERR_INVALID;
    }
    if (need_problem && !e->problem_set) {
        yg_set_error("yg_set_problem has not been called");
        return YG_ERR_STATE;
    }
    if (need_state && !e->state_set) {
        yg_set_error("yg_set_state / yg_load_state has not been called");
        return YG_ERR_STATE;
    }
    return YG_OK;
}

template <typename T>
int dev_alloc(T **p, size_t n)
{
    YG_CUDA_CHECK(cudaMalloc((void **)p, sizeof(T) * std::max<size_t>(n, 1)));
    YG_CUDA_CHECK(cudaMemset(*p, 0, sizeof(T) * std::max<size_t>(n, 1)));
    return YG_OK;
}

void copy_mat(double *dst, const double *src, int rows, int cols)
{
    if (src) memcpy(dst, src, sizeof(double) * rows * cols);
}

RunArgs make_args(yg_ensemble *e)
{
    RunArgs a;
    memset(&a, 0, sizeof(a));
    a.problem = e->d_problem;
    a.problem_bytes = (uint32_t)e->h_problem.size();
    a.thin = 1;
    a.n_chains = e->cfg.n_chains;
    a.chain_offset = e->cfg.chain_offset;
    a.seed = e->cfg.seed;
    a.step0 = e->step_index;
    a.welford_n0 = e->welford_n;
    a.am_t0 = e->am_steps;
    a.theta = e->theta;
    a.logpost = e->logpost;
    a.n_accept = e->n_accept;
    a.w_mean = e->w_mean;
    a.w_m2 = e->w_m2;
    a.am_mean = e->am_mean;
    a.am_m2 = e->am_m2;
    a.prop_L = e->prop_L;
    a.adaptive = e->cfg.adaptive;
    a.am_refresh = std::max(1, e->cfg.am_refresh);
    a.am_idle = e->cfg.am_idle_steps;
    a.am_collect = e->cfg.am_collection_steps;
    a.am_eps = e->cfg.am_eps;
    a.am_scale = e->cfg.am_scale > 0.0 ? e->cfg.am_scale : 2.4 * 2.4 / (double)e->cfg.dim;
    a.aem = e->cfg.aem;
    a.aem_min_data = e->cfg.aem_min_data;
    a.aem_heuristic = e->cfg.aem_heuristic;
    a.welford = e->cfg.acceptance_only ? 0 : 1;
    a.aem_n = e->aem_n;
    a.aem_mean = e->aem_mean;
    a.aem_m2 = e->aem_m2;
    a.aem_cache = e->aem_cache;
    a.counters = e->counters;
    return a;
}

__global__ void broadcast_L_kernel(const DevProblemHeader *pb, double *prop_L, int64_t n)
{
    const int d = pb->dim;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n; g += (int64_t)gridDim.x * blockDim.x)
        for (int k = 0; k < d * d; k++) prop_L[(int64_t)k * n + g] = pb->prop_L[k];
}

bool is_diagonal(const double *M, int n)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (i != j && M[i * n + j] != 0.0) return false;
    return true;
}

// Large linear model: builds the DevBigHeader blob (big_linear.h) for linear_dmma_kernel.cu.
int set_problem_big(yg_ensemble *e, const yg_problem *pb)
{
    const int d = e->cfg.dim, nl = e->cfg.n_levels;
    if (e->cfg.adaptive || e->cfg.eq_mode != YG_EQ_EXACT || e->cfg.aem || nl > 2) {
        yg_set_error("large linear model (dim > %d or data_dim > %d): MRW / pCN / two-level delayed acceptance with exact "
                     "equality only (no adaptive proposal, adaptive error model or third level)", YG_MAX_DIM, YG_MAX_DATA_DIM);
        return YG_ERR_UNSUPPORTED;
    }
    if (pb->proposal != YG_PROPOSAL_MRW && pb->proposal != YG_PROPOSAL_PCN) {
        yg_set_error("unknown proposal kind %d", pb->proposal);
        return YG_ERR_INVALID;
    }
    if (pb->proposal == YG_PROPOSAL_PCN) {
        if (!(pb->pcn_step > 0.0 && pb->pcn_step <= 0.5)) {       // pcn.py:42
            yg_set_error("pCN step size must lie in (0, 0.5], got %g", pb->pcn_step);
            return YG_ERR_INVALID;
        }
        if (nl != 1) {
            yg_set_error("pCN is a single-level method (chain/method/pcn.py)");
            return YG_ERR_UNSUPPORTED;
        }
    }
    for (int l = 0; l < nl; l++)
        if (pb->level[l].tempered) {
            yg_set_error("large linear model: tempered levels are not implemented");
            return YG_ERR_UNSUPPORTED;
        }
    const bool dense_L = !is_diagonal(pb->prop_L, d);
    for (int i = 0; i < d; i++)
        for (int j = i + 1; j < d; j++)
            if (pb->prop_L[(size_t)i * d + j] != 0.0) {
                yg_set_error("prop_L must be lower triangular");
                return YG_ERR_INVALID;
            }
    const int kp = d <= 16 ? 16 : (d <= 32 ? 32 : 64), ks = kp + 4;
    size_t tail_len = 0;
    bool dense_prior[2] = {false, false};
    std::vector<double> Rp[2];            // chol(prior precision)' of the levels with a dense prior
    for (int l = 0; l < nl; l++) {
        const yg_level &L = pb->level[l];
        if (!L.data || !L.noise_prec || !L.prior_mean || !L.prior_prec || !L.G || !L.b || L.n_data < 1 ||
            L.data_dim < 1) {
            yg_set_error("level %d: linear level needs data, noise_prec, prior, G and b", l);
            return YG_ERR_INVALID;
        }
        if (L.data_dim > YG_BIG_MAX_DATA_DIM) {
            yg_set_error("level %d: large linear model is limited to data_dim <= %d", l, YG_BIG_MAX_DATA_DIM);
            return YG_ERR_UNSUPPORTED;
        }
        if (!is_diagonal(L.noise_prec, L.data_dim)) {
            yg_set_error("level %d: large linear model needs diagonal measurement noise", l);
            return YG_ERR_UNSUPPORTED;
        }
        for (int r = 0; r < L.data_dim; r++)
            if (!(L.noise_prec[(size_t)r * L.data_dim + r] >= 0.0)) {
                yg_set_error("level %d: negative noise precision", l);
                return YG_ERR_INVALID;
            }
        dense_prior[l] = !is_diagonal(L.prior_prec, d);
        if (dense_prior[l]) {
            // P = Lp Lp' (lower Cholesky); R = Lp' so that ||R x||^2 = x' P x.  DenseCovarianceMatrix priors
            // (statistics/covariance.py:69-94) arrive here as their dense precision.
            Rp[l].assign((size_t)d * d, 0.0);
            std::vector<double> Lp((size_t)d * d, 0.0);
            for (int j = 0; j < d; j++) {
                double sdiag = L.prior_prec[(size_t)j * d + j];
                for (int k = 0; k < j; k++) sdiag -= Lp[(size_t)j * d + k] * Lp[(size_t)j * d + k];
                if (!(sdiag > 0.0)) {
                    yg_set_error("level %d: the dense prior precision is not positive definite", l);
                    return YG_ERR_INVALID;
                }
                Lp[(size_t)j * d + j] = sqrt(sdiag);
                for (int i = j + 1; i < d; i++) {
                    double v = 0.5 * (L.prior_prec[(size_t)i * d + j] + L.prior_prec[(size_t)j * d + i]);
                    for (int k = 0; k < j; k++) v -= Lp[(size_t)i * d + k] * Lp[(size_t)j * d + k];
                    Lp[(size_t)i * d + j] = v / Lp[(size_t)j * d + j];
                }
            }
            for (int i = 0; i < d; i++)
                for (int j = 0; j < d; j++) Rp[l][(size_t)i * d + j] = Lp[(size_t)j * d + i];
        }
        const size_t rows = (size_t)L.data_dim + (dense_prior[l] ? d : 0);
        const size_t np = (rows + 15) & ~size_t(15);
        tail_len += np * ks + np + 2 * (size_t)kp;
    }
    tail_len += 2 * (size_t)kp + (dense_L ? (size_t)kp * ks : 0);
    tail_len = (tail_len + 1) & ~size_t(1);
    {   // the blob and the tiles of at least four warps (state tile, noise tile, + a scratch tile for a dense factor)
        const size_t per_warp = sizeof(double) * 8 * (size_t)ks * (dense_L ? 2 : 1) + sizeof(float) * 8 * (size_t)(kp + 4);
        if (sizeof(double) * tail_len + 4 * per_warp > 226 * 1024) {
            yg_set_error("large linear model: %zu bytes of G / data do not fit the shared memory of one SM", sizeof(double) * tail_len);
            return YG_ERR_UNSUPPORTED;
        }
    }
    std::vector<char> blob(sizeof(DevBigHeader) + sizeof(double) * tail_len, 0);
    DevBigHeader *h = reinterpret_cast<DevBigHeader *>(blob.data());
    double *tail = reinterpret_cast<double *>(blob.data() + sizeof(DevBigHeader));
    h->dim = d; h->kp = kp; h->ks = ks; h->n_levels = nl;
    h->J = nl == 2 ? e->cfg.sub_chain_length : 1;
    h->tail_len = (int32_t)tail_len;
    h->proposal = pb->proposal;
    h->dense_L = dense_L ? 1 : 0;
    if (pb->proposal == YG_PROPOSAL_PCN) {
        const double t = 2.0 * pb->pcn_step;                      // pcn.py:30
        h->pcn_a = sqrt(1.0 - t);
        h->pcn_b = sqrt(t);
    }
    size_t off = 0;
    for (int l = 0; l < nl; l++) {
        const yg_level &L = pb->level[l];
        BigLevel &B = h->lvl[l];
        const int rows = L.data_dim + (dense_prior[l] ? d : 0);
        const int np = (rows + 15) & ~15;
        B.n_data = L.n_data; B.data_dim = L.data_dim; B.np = np; B.n_rows = rows;
        B.G_off = (int32_t)off;
        B.bd_off = (int32_t)(off + (size_t)np * ks);
        // sum_rows ||F - d_row||^2_P = sum_col n P_col (F_col - mean_col)^2 + sum_col P_col sum_rows (d - mean)^2;
        // sqrt(n P_col) is folded into row `col` of G and into b - mean
        double q_const = 0.0;
        for (int r = 0; r < L.data_dim; r++) {
            double mean = 0.0;
            for (int i = 0; i < L.n_data; i++) mean += L.data[(size_t)i * L.data_dim + r];
            mean /= (double)L.n_data;
            double scatter = 0.0;
            for (int i = 0; i < L.n_data; i++) {
                const double dd = L.data[(size_t)i * L.data_dim + r] - mean;
                scatter += dd * dd;
            }
            const double prec = L.noise_prec[(size_t)r * L.data_dim + r];
            const double sw = sqrt((double)L.n_data * prec);
            for (int k = 0; k < d; k++) tail[off + (size_t)r * ks + k] = sw * L.G[(size_t)r * d + k];
            tail[B.bd_off + r] = sw * (L.b[r] - mean);
            q_const += prec * scatter;
        }
        if (dense_prior[l]) {                    // d more rows: R and -R m
            for (int r = 0; r < d; r++) {
                double rm = 0.0;
                for (int k = 0; k < d; k++) {
                    tail[off + (size_t)(L.data_dim + r) * ks + k] = Rp[l][(size_t)r * d + k];
                    rm += Rp[l][(size_t)r * d + k] * L.prior_mean[k];
                }
                tail[B.bd_off + L.data_dim + r] = -rm;
            }
        }
        B.q_const = q_const;
        off += (size_t)np * ks + np;
        B.pmean_off = (int32_t)off;
        for (int k = 0; k < d; k++) tail[off + k] = L.prior_mean[k];
        off += kp;
        B.pprec_off = (int32_t)off;
        if (!dense_prior[l])
            for (int k = 0; k < d; k++) tail[off + k] = L.prior_prec[(size_t)k * d + k];
        off += kp;
    }
    if (nl == 1) h->lvl[1] = h->lvl[0];
    h->propL_off = (int32_t)off;
    for (int k = 0; k < d; k++) tail[off + k] = pb->prop_L[(size_t)k * d + k];
    off += kp;
    h->pcn_mean_off = (int32_t)off;
    if (pb->proposal == YG_PROPOSAL_PCN && pb->pcn_mean)
        for (int k = 0; k < d; k++) tail[off + k] = pb->pcn_mean[k];
    off += kp;
    h->Ld_off = (int32_t)off;
    if (dense_L)
        for (int r = 0; r < d; r++)
            for (int k = 0; k <= r; k++) tail[off + (size_t)r * ks + k] = pb->prop_L[(size_t)r * d + k];
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    e->problem_set = false;
    if (e->d_problem) cudaFree(e->d_problem);
    e->d_problem = nullptr;
    YG_CUDA_CHECK(cudaMalloc((void **)&e->d_problem, blob.size()));
    YG_CUDA_CHECK(cudaMemcpy(e->d_problem, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    if (!e->big_done) {     // per-tile progress counters of the balanced (tile, step-range) schedule
        int rc = dev_alloc(&e->big_done, ((size_t)e->cfg.n_chains + 7) / 8);
        if (rc) return rc;
    }
    e->h_problem.swap(blob);
    e->problem_set = true;
    e->big = true;
    return YG_OK;
}

}  // namespace

extern "C" const char *yg_last_error(void) { return g_err; }
extern "C" uint32_t yg_abi_version(void) { return YG_ABI_VERSION; }

extern "C" int yg_create(const yg_config *cfg, yg_ensemble **out)
{
    if (!cfg || !out) {
        yg_set_error("yg_create: null argument");
        return YG_ERR_INVALID;
    }
    *out = nullptr;
    if (cfg->abi_version != YG_ABI_VERSION) {
        yg_set_error("abi_version %u != library %u", cfg->abi_version, YG_ABI_VERSION);
        return YG_ERR_ABI;
    }
    const int max_dim = cfg->model == YG_MODEL_LINEAR ? YG_BIG_MAX_DIM : YG_MAX_DIM;
    if (cfg->n_chains < 1 || cfg->dim < 1 || cfg->dim > max_dim) {
        yg_set_error("n_chains=%lld dim=%d out of range (dim 1..%d)", (long long)cfg->n_chains, cfg->dim, max_dim);
        return cfg->dim > max_dim ? YG_ERR_UNSUPPORTED : YG_ERR_INVALID;
    }
    if (cfg->dim > YG_MAX_DIM && cfg->adaptive) {
        yg_set_error("adaptive Metropolis is limited to dim <= %d", YG_MAX_DIM);
        return YG_ERR_UNSUPPORTED;
    }
    if (cfg->n_levels < 1 || cfg->n_levels > YG_MAX_LEVELS) {
        yg_set_error("n_levels=%d: MRW (1), two-level delayed acceptance (2) and MLDA with two surrogates (3) "
                     "exist on the device", cfg->n_levels);
        return YG_ERR_UNSUPPORTED;
    }
    if (cfg->n_levels == 3 && (cfg->model == YG_MODEL_LV_RK4 || cfg->aem || cfg->dim > YG_MAX_DIM)) {
        yg_set_error("three levels: Gaussian targets and the linear model (dim <= %d) only, no adaptive error model",
                     YG_MAX_DIM);
        return YG_ERR_UNSUPPORTED;
    }
    if (cfg->n_levels >= 2 && cfg->sub_chain_length < 1) {
        yg_set_error("sub_chain_length must be >= 1");
        return YG_ERR_INVALID;
    }
    if (cfg->model != YG_MODEL_GAUSS && cfg->model != YG_MODEL_LINEAR && cfg->model != YG_MODEL_LV_RK4) {
        yg_set_error("unknown model %d", cfg->model);
        return YG_ERR_UNSUPPORTED;
    }
    if (cfg->model == YG_MODEL_LV_RK4 && cfg->dim != 2) {
        yg_set_error("the Lotka-Volterra model has two parameters (dim=%d)", cfg->dim);
        return YG_ERR_INVALID;
    }
    if (cfg->adaptive && cfg->dim < 2) {
        yg_set_error("Adaptivity not implemented for scalar chains.");      // chain/adaptive.py:41-43
        return YG_ERR_UNSUPPORTED;
    }
    if (cfg->adaptive && (cfg->am_idle_steps < 0 || cfg->am_collection_steps < 0 || cfg->am_refresh < 0)) {
        yg_set_error("adaptive Metropolis: idle / collection steps and the refresh interval must be >= 0");
        return YG_ERR_INVALID;
    }
    if (cfg->aem) {
        if (cfg->n_levels != 2 || cfg->model != YG_MODEL_LINEAR || cfg->adaptive || cfg->dim > YG_MAX_DIM) {
            yg_set_error("adaptive error model: two levels, linear model (dim <= %d), no adaptive proposal", YG_MAX_DIM);
            return YG_ERR_UNSUPPORTED;   // aem.py:72-75: only defined for a hierarchy of Bayesian models
        }
        if (cfg->aem_min_data < 2) {
            yg_set_error("Smallest senisible data size for AEM is 2.");      // likelihood.py:99-100
            return YG_ERR_INVALID;
        }
    }
    if (cfg->eq_mode == YG_EQ_ISCLOSE && cfg->dim != 1) {
        yg_set_error("isclose equality is the ScalarParameter rule (dim must be 1)");
        return YG_ERR_INVALID;
    }
    int ndev = 0;
    YG_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) {
        yg_set_error("device %d not present (%d visible)", cfg->device, ndev);
        return YG_ERR_INVALID;
    }
    YG_CUDA_CHECK(cudaSetDevice(cfg->device));
    yg_ensemble *e = new (std::nothrow) yg_ensemble();
    if (!e) {
        yg_set_error("out of host memory");
        return YG_ERR_INVALID;
    }
    e->cfg = *cfg;
    if (cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess) {
        yg_set_error("cudaDeviceGetAttribute(multiProcessorCount) failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete e;
        return YG_ERR_CUDA;
    }
    const size_t n = (size_t)cfg->n_chains, d = (size_t)cfg->dim;
    int rc = YG_OK;
    if ((rc = dev_alloc(&e->theta, d * n)) || (rc = dev_alloc(&e->logpost, YG_MAX_LEVELS * n)) ||
        (rc = dev_alloc(&e->w_mean, d * n)) || (rc = dev_alloc(&e->w_m2, (cfg->dim > YG_MAX_DIM ? d : d * d) * n)) ||
        (rc = dev_alloc(&e->n_accept, n)) || (rc = dev_alloc(&e->counters, 8)) ||
        (rc = dev_alloc(&e->pool_partials, (size_t)POOL_PARTS * (size_t)yg_pooled_len(std::min(cfg->dim, YG_MAX_DIM))))) {
        yg_destroy(e);
        return rc;
    }
    if (cfg->adaptive) {
        if ((rc = dev_alloc(&e->am_mean, d * n)) || (rc = dev_alloc(&e->am_m2, d * d * n)) ||
            (rc = dev_alloc(&e->prop_L, d * d * n))) {
            yg_destroy(e);
            return rc;
        }
    }
    *out = e;
    return YG_OK;
}

extern "C" int yg_destroy(yg_ensemble *e)
{
    if (!e) return YG_OK;
    cudaSetDevice(e->cfg.device);
    cudaFree(e->d_problem);
    cudaFree(e->theta);
    cudaFree(e->logpost);
    cudaFree(e->w_mean);
    cudaFree(e->w_m2);
    cudaFree(e->am_mean);
    cudaFree(e->am_m2);
    cudaFree(e->prop_L);
    cudaFree(e->aem_n);
    cudaFree(e->aem_mean);
    cudaFree(e->aem_m2);
    cudaFree(e->aem_cache);
    cudaFree(e->n_accept);
    cudaFree(e->counters);
    cudaFree(e->pool_partials);
    cudaFree(e->big_done);
    delete e;
    return YG_OK;
}

extern "C" int yg_set_problem(yg_ensemble *e, const yg_problem *pb)
{
    int rc = check_handle(e, false, false);
    if (rc) return rc;
    if (!pb || !pb->prop_L) {
        yg_set_error("yg_set_problem: problem / prop_L is null");
        return YG_ERR_INVALID;
    }
    const int d = e->cfg.dim, nl = e->cfg.n_levels, model = e->cfg.model;
    if (model == YG_MODEL_LINEAR) {
        bool big = d > YG_MAX_DIM;
        for (int l = 0; l < nl; l++) big = big || pb->level[l].data_dim > YG_MAX_DATA_DIM;
        if (big) return set_problem_big(e, pb);
    }
    e->big = false;
    if (pb->proposal != YG_PROPOSAL_MRW && pb->proposal != YG_PROPOSAL_PCN) {
        yg_set_error("unknown proposal kind %d", pb->proposal);
        return YG_ERR_INVALID;
    }
    if (pb->proposal == YG_PROPOSAL_PCN) {
        if (!(pb->pcn_step > 0.0 && pb->pcn_step <= 0.5)) {       // pcn.py:42
            yg_set_error("pCN step size must lie in (0, 0.5], got %g", pb->pcn_step);
            return YG_ERR_INVALID;
        }
        if (nl != 1 || e->cfg.adaptive) {
            yg_set_error("pCN is a single-level, non-adaptive method (chain/method/pcn.py)");
            return YG_ERR_UNSUPPORTED;
        }
    }
    // tail: per level data + design
    size_t tail_len = 0;
    for (int l = 0; l < nl; l++) {
        const yg_level &L = pb->level[l];
        if (model == YG_MODEL_GAUSS) {
            if (!L.g_mean || !L.g_prec) {
                yg_set_error("level %d: Gaussian target needs g_mean and g_prec", l);
                return YG_ERR_INVALID;
            }
            continue;
        }
        if (!L.data || !L.noise_prec || !L.prior_mean || !L.prior_prec || L.n_data < 1 || L.data_dim < 1 ||
            L.data_dim > YG_MAX_DATA_DIM) {
            yg_set_error("level %d: regression level needs data[n_data>=1, data_dim<=%d], noise_prec, prior", l,
                         YG_MAX_DATA_DIM);
            return YG_ERR_INVALID;
        }
        if (model == YG_MODEL_LINEAR && (!L.G || !L.b)) {
            yg_set_error("level %d: linear model needs G and b", l);
            return YG_ERR_INVALID;
        }
        if (model == YG_MODEL_LV_RK4) {
            if (!L.design || L.data_dim != 2 || L.rk4_steps < 1 || !(L.T > 0.0)) {
                yg_set_error("level %d: LV model needs design[n_data,2], data_dim=2, rk4_steps>=1, T>0", l);
                return YG_ERR_INVALID;
            }
        }
        tail_len += (size_t)L.n_data * L.data_dim;
        if (model == YG_MODEL_LV_RK4) tail_len += (size_t)L.n_data * 2;
    }
    tail_len = (tail_len + 1) & ~size_t(1);        // keep the blob a multiple of 16 bytes
    std::vector<char> blob(sizeof(DevProblemHeader) + sizeof(double) * tail_len, 0);
    DevProblemHeader *h = reinterpret_cast<DevProblemHeader *>(blob.data());
    double *tail = reinterpret_cast<double *>(blob.data() + sizeof(DevProblemHeader));
    h->model = model;
    h->dim = d;
    h->n_levels = nl;
    h->J = nl >= 2 ? e->cfg.sub_chain_length : 1;
    h->eq_mode = e->cfg.eq_mode;
    h->tail_len = (int32_t)tail_len;
    h->proposal = pb->proposal;
    if (pb->proposal == YG_PROPOSAL_PCN) {
        const double t = 2.0 * pb->pcn_step;                      // pcn.py:30
        h->pcn_a = sqrt(1.0 - t);
        h->pcn_b = sqrt(t);
        if (pb->pcn_mean) copy_mat(h->pcn_mean, pb->pcn_mean, 1, d);
    }
    copy_mat(h->prop_L, pb->prop_L, d, d);
    for (int i = 0; i < d; i++)
        for (int j = i + 1; j < d; j++)
            if (h->prop_L[i * d + j] != 0.0) {
                yg_set_error("prop_L must be lower triangular");
                return YG_ERR_INVALID;
            }
    size_t off = 0;
    for (int l = 0; l < nl; l++) {
        const yg_level &L = pb->level[l];
        DevLevel &D = h->lvl[l];
        if (model == YG_MODEL_GAUSS) {
            copy_mat(D.g_mean, L.g_mean, 1, d);
            copy_mat(D.g_prec, L.g_prec, d, d);
            D.g_logconst = L.g_logconst;
            continue;
        }
        D.n_data = L.n_data;
        D.data_dim = L.data_dim;
        D.tempering = 1.0;
        if (L.tempered) {
            if (!(L.tempering >= 0.0 && L.tempering <= 1.0)) {       // tmlda.py:24-29
                yg_set_error("Invalid tempering parameter at index %d: %g (must be in [0, 1]).", l, L.tempering);
                return YG_ERR_INVALID;
            }
            D.tempering = L.tempering;
        }
        copy_mat(D.noise_prec, L.noise_prec, L.data_dim, L.data_dim);
        copy_mat(D.prior_mean, L.prior_mean, 1, d);
        copy_mat(D.prior_prec, L.prior_prec, d, d);
        D.data_off = (int32_t)off;
        memcpy(tail + off, L.data, sizeof(double) * L.n_data * L.data_dim);
        off += (size_t)L.n_data * L.data_dim;
        if (model == YG_MODEL_LINEAR) {
            copy_mat(D.G, L.G, L.data_dim, d);
            copy_mat(D.b, L.b, 1, L.data_dim);
        } else {
            D.alpha = L.alpha;
            D.gamma = L.gamma;
            D.T = L.T;
            D.rk4_steps = L.rk4_steps;
            D.design_off = (int32_t)off;
            memcpy(tail + off, L.design, sizeof(double) * L.n_data * 2);
            off += (size_t)L.n_data * 2;
        }
    }
    if (e->cfg.aem) {
        const yg_level &Lc = pb->level[0], &Lf = pb->level[1];
        if (Lc.data_dim != Lf.data_dim || pb->proposal != YG_PROPOSAL_MRW) {
            yg_set_error("adaptive error model: both levels must share the data dimension (MRW proposal)");
            return YG_ERR_INVALID;
        }
        for (int i = 0; i < Lc.data_dim; i++)
            for (int j = 0; j < Lc.data_dim; j++)
                if (i != j && Lc.noise_prec[i * Lc.data_dim + j] != 0.0) {
                    // noise.py:29-33
                    yg_set_error("Currently, AEM is only implemented for independent measurement noise.");
                    return YG_ERR_UNSUPPORTED;
                }
    }
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    e->problem_set = false;                     // stays false if anything below fails
    if (e->d_problem) cudaFree(e->d_problem);
    e->d_problem = nullptr;
    YG_CUDA_CHECK(cudaMalloc((void **)&e->d_problem, blob.size()));
    YG_CUDA_CHECK(cudaMemcpy(e->d_problem, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    e->h_problem.swap(blob);
    if (e->cfg.aem) {
        const size_t n = (size_t)e->cfg.n_chains, dd = (size_t)pb->level[0].data_dim;
        cudaFree(e->aem_n); cudaFree(e->aem_mean); cudaFree(e->aem_m2); cudaFree(e->aem_cache);
        e->aem_n = nullptr; e->aem_mean = e->aem_m2 = e->aem_cache = nullptr;
        e->aem_data_dim = (int)dd;
        if ((rc = dev_alloc(&e->aem_n, n)) || (rc = dev_alloc(&e->aem_mean, dd * n)) ||
            (rc = dev_alloc(&e->aem_m2, dd * n)) || (rc = dev_alloc(&e->aem_cache, (size_t)(3 * (d + 1) + 1) * n)))
            return rc;
    }
    e->problem_set = true;
    return YG_OK;
}

extern "C" int yg_set_state(yg_ensemble *e, const double *theta0_dev, int32_t flags, void *stream)
{
    int rc = check_handle(e, true, false);
    if (rc) return rc;
    if (!theta0_dev) {
        yg_set_error("theta0_dev is null");
        return YG_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)e->cfg.n_chains, d = (size_t)e->cfg.dim;
    // a handle that never held a state has nothing to keep
    const bool keep_diag = (flags & YG_KEEP_DIAGNOSTICS) && e->state_set;
    const bool keep_adapt = (flags & YG_KEEP_ADAPTATION) && e->state_set;
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    YG_CUDA_CHECK(cudaMemcpyAsync(e->theta, theta0_dev, sizeof(double) * d * n, cudaMemcpyDeviceToDevice, st));
    if (!keep_diag) {
        YG_CUDA_CHECK(cudaMemsetAsync(e->w_mean, 0, sizeof(double) * d * n, st));
        YG_CUDA_CHECK(cudaMemsetAsync(e->w_m2, 0, sizeof(double) * (d > YG_MAX_DIM ? d : d * d) * n, st));
        YG_CUDA_CHECK(cudaMemsetAsync(e->n_accept, 0, sizeof(unsigned long long) * n, st));
        YG_CUDA_CHECK(cudaMemsetAsync(e->counters, 0, sizeof(unsigned long long) * 8, st));
        e->welford_n = 0;
    }
    if (e->cfg.adaptive && !keep_adapt) {
        YG_CUDA_CHECK(cudaMemsetAsync(e->am_mean, 0, sizeof(double) * d * n, st));
        YG_CUDA_CHECK(cudaMemsetAsync(e->am_m2, 0, sizeof(double) * d * d * n, st));
        broadcast_L_kernel<<<(int)std::min<size_t>((n + 127) / 128, 2048), 128, 0, st>>>(e->d_problem, e->prop_L,
                                                                                      (int64_t)n);
        YG_CUDA_CHECK(cudaGetLastError());
        e->am_steps = 0;
    }
    if (e->cfg.aem && !keep_adapt) {       // empty cache, no error realisations yet
        YG_CUDA_CHECK(cudaMemsetAsync(e->aem_n, 0, sizeof(unsigned long long) * n, st));
        YG_CUDA_CHECK(cudaMemsetAsync(e->aem_mean, 0, sizeof(double) * (size_t)e->aem_data_dim * n, st));
        YG_CUDA_CHECK(cudaMemsetAsync(e->aem_m2, 0, sizeof(double) * (size_t)e->aem_data_dim * n, st));
        YG_CUDA_CHECK(cudaMemsetAsync(e->aem_cache, 0, sizeof(double) * (3 * (d + 1) + 1) * n, st));
    }
    for (int l = 0; l < e->cfg.n_levels; l++) {
        rc = e->big ? yg_launch_logpost_big(e, l, e->theta, (int64_t)n, e->logpost + (size_t)l * n, st)
                    : yg_launch_logpost(e, l, e->theta, (int64_t)n, e->logpost + (size_t)l * n, st);
        if (rc) return rc;
    }
    // e->step_index (the Philox stream position) is deliberately left alone: the reference's numpy
    // generator keeps advancing across run() calls, and so does this stream (yg_seek repositions it)
    e->state_set = true;
    return YG_OK;
}

extern "C" int yg_seek(yg_ensemble *e, int64_t step_index)
{
    int rc = check_handle(e, false, false);
    if (rc) return rc;
    if (step_index < 0) {
        yg_set_error("yg_seek: step_index must be >= 0");
        return YG_ERR_INVALID;
    }
    e->step_index = step_index;
    return YG_OK;
}

// Replaces the proposal factor between launches (pooled proposal covariance across GPUs, north_star;
// the reference's only hook for this is AdaptiveMRWProposal swapping its covariance, chain/adaptive.py:55-60).
extern "C" int yg_set_proposal_factor(yg_ensemble *e, const double *L_host, void *stream)
{
    int rc = check_handle(e, true, false);
    if (rc) return rc;
    const int d = e->cfg.dim;
    if (!L_host) {
        yg_set_error("yg_set_proposal_factor: L is null");
        return YG_ERR_INVALID;
    }
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            const double v = L_host[(size_t)i * d + j];
            if (!std::isfinite(v) || (j > i && v != 0.0) || (i == j && !(v > 0.0))) {
                yg_set_error("proposal factor must be finite, lower triangular with a positive diagonal");
                return YG_ERR_INVALID;
            }
        }
    cudaStream_t st = (cudaStream_t)stream;
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    if (e->big) {
        DevBigHeader *h = reinterpret_cast<DevBigHeader *>(e->h_problem.data());
        if (!h->dense_L && !is_diagonal(L_host, d)) {
            yg_set_error("large linear model: this handle was built for a diagonal proposal factor (its kernels and shared-"
                         "memory plan differ for a dense one): pass a dense factor to yg_set_problem to switch");
            return YG_ERR_UNSUPPORTED;
        }
        if (h->proposal != YG_PROPOSAL_MRW) {
            yg_set_error("the proposal factor of a pCN chain is the prior's (pcn.py:23-35) and cannot be replaced");
            return YG_ERR_UNSUPPORTED;
        }
        double *tail = reinterpret_cast<double *>(e->h_problem.data() + sizeof(DevBigHeader));
        for (int k = 0; k < d; k++) tail[h->propL_off + k] = L_host[(size_t)k * d + k];
        char *base = reinterpret_cast<char *>(e->d_problem) + sizeof(DevBigHeader);
        YG_CUDA_CHECK(cudaMemcpyAsync(base + sizeof(double) * h->propL_off, tail + h->propL_off, sizeof(double) * d,
                                      cudaMemcpyHostToDevice, st));
        if (h->dense_L) {
            for (int r = 0; r < d; r++)
                for (int k = 0; k < d; k++) tail[h->Ld_off + (size_t)r * h->ks + k] = (k <= r) ? L_host[(size_t)r * d + k] : 0.0;
            YG_CUDA_CHECK(cudaMemcpyAsync(base + sizeof(double) * h->Ld_off, tail + h->Ld_off,
                                          sizeof(double) * (size_t)h->kp * h->ks, cudaMemcpyHostToDevice, st));
        }
        return YG_OK;
    }
    DevProblemHeader *h = reinterpret_cast<DevProblemHeader *>(e->h_problem.data());
    if (h->proposal != YG_PROPOSAL_MRW) {
        yg_set_error("the proposal factor of a pCN chain is the prior's (pcn.py:23-35) and cannot be replaced");
        return YG_ERR_UNSUPPORTED;
    }
    copy_mat(h->prop_L, L_host, d, d);
    YG_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<char *>(e->d_problem) + offsetof(DevProblemHeader, prop_L), h->prop_L,
                                  sizeof(h->prop_L), cudaMemcpyHostToDevice, st));
    if (e->cfg.adaptive && e->prop_L) {
        const size_t n = (size_t)e->cfg.n_chains;
        broadcast_L_kernel<<<(int)std::min<size_t>((n + 127) / 128, 2048), 128, 0, st>>>(e->d_problem, e->prop_L, (int64_t)n);
        YG_CUDA_CHECK(cudaGetLastError());
    }
    return YG_OK;
}

extern "C" int yg_run(yg_ensemble *e, int64_t n_steps, int32_t thin, const yg_outputs *out, const yg_noise *noise,
                      void *stream)
{
    int rc = check_handle(e, true, true);
    if (rc) return rc;
    if (n_steps < 0 || thin < 1) {
        yg_set_error("n_steps=%lld thin=%d invalid", (long long)n_steps, thin);
        return YG_ERR_INVALID;
    }
    if (n_steps == 0) return YG_OK;
    RunArgs a = make_args(e);
    a.n_steps = n_steps;
    a.thin = thin;
    if (out) {
        a.samples = out->samples_dev;
        a.accepted = out->accepted_dev;
        a.lp_out = out->logpost_dev;
        if ((a.samples || a.lp_out) && n_steps % thin) {
            yg_set_error("n_steps (%lld) must be a multiple of thin (%d) when samples are stored",
                         (long long)n_steps, thin);
            return YG_ERR_INVALID;
        }
    }
    if (noise && noise->mode != YG_NOISE_PHILOX) {
        if (noise->mode != YG_NOISE_INJECT && noise->mode != YG_NOISE_RECORD) {
            yg_set_error("unknown noise mode %d", noise->mode);
            return YG_ERR_INVALID;
        }
        if (!noise->z_dev || !noise->u_f_dev || (e->cfg.n_levels >= 2 && !noise->u_c_dev)) {
            yg_set_error("noise mode %d needs z_dev, u_f_dev%s", noise->mode,
                         e->cfg.n_levels >= 2 ? " and u_c_dev" : "");
            return YG_ERR_INVALID;
        }
        a.noise_mode = noise->mode;
        a.z = noise->z_dev;
        a.u_c = noise->u_c_dev;
        a.u_f = noise->u_f_dev;
    }
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    rc = e->big ? yg_launch_linear_big(e, a, (cudaStream_t)stream)
         : (e->cfg.model == YG_MODEL_LV_RK4) ? yg_launch_lv(e, a, false, (cudaStream_t)stream)
                                             : yg_launch_generic(e, a, false, (cudaStream_t)stream);
    if (rc) return rc;
    e->step_index += n_steps;
    e->welford_n += n_steps;
    if (e->cfg.adaptive) e->am_steps += n_steps;
    return YG_OK;
}

extern "C" int yg_get_state(yg_ensemble *e, const yg_state *dst, void *stream)
{
    int rc = check_handle(e, true, true);
    if (rc) return rc;
    if (!dst) return YG_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)e->cfg.n_chains, d = (size_t)e->cfg.dim, nl = (size_t)e->cfg.n_levels;
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
    if (dst->theta_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->theta_dev, e->theta, 8 * d * n, k, st));
    if (dst->logpost_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->logpost_dev, e->logpost, 8 * nl * n, k, st));
    if (dst->n_accept_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->n_accept_dev, e->n_accept, 8 * n, k, st));
    if (dst->w_mean_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->w_mean_dev, e->w_mean, 8 * d * n, k, st));
    if (dst->w_m2_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->w_m2_dev, e->w_m2, 8 * (d > YG_MAX_DIM ? d : d * d) * n, k, st));
    if (dst->prop_L_dev) {
        if (!e->cfg.adaptive) {
            yg_set_error("prop_L_dev is per-chain state of adaptive ensembles only");
            return YG_ERR_INVALID;
        }
        YG_CUDA_CHECK(cudaMemcpyAsync(dst->prop_L_dev, e->prop_L, 8 * d * d * n, k, st));
    }
    if (dst->am_mean_dev || dst->am_m2_dev) {
        if (!e->cfg.adaptive) {
            yg_set_error("am_mean_dev / am_m2_dev are per-chain state of adaptive ensembles only");
            return YG_ERR_INVALID;
        }
        if (dst->am_mean_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->am_mean_dev, e->am_mean, 8 * d * n, k, st));
        if (dst->am_m2_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->am_m2_dev, e->am_m2, 8 * d * d * n, k, st));
    }
    if (dst->aem_n_dev || dst->aem_mean_dev || dst->aem_m2_dev || dst->aem_cache_dev) {
        if (!e->cfg.aem) {
            yg_set_error("aem_*_dev are per-chain state of adaptive-error-model ensembles only");
            return YG_ERR_INVALID;
        }
        const size_t dd = (size_t)e->aem_data_dim;
        if (dst->aem_n_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->aem_n_dev, e->aem_n, 8 * n, k, st));
        if (dst->aem_mean_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->aem_mean_dev, e->aem_mean, 8 * dd * n, k, st));
        if (dst->aem_m2_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->aem_m2_dev, e->aem_m2, 8 * dd * n, k, st));
        if (dst->aem_cache_dev) YG_CUDA_CHECK(cudaMemcpyAsync(dst->aem_cache_dev, e->aem_cache, 8 * (3 * (d + 1) + 1) * n, k, st));
    }
    return YG_OK;
}

extern "C" int yg_load_state(yg_ensemble *e, const yg_state *src, int64_t step_index, int64_t welford_n,
                             int64_t am_steps, void *stream)
{
    int rc = check_handle(e, true, false);
    if (rc) return rc;
    if (!src || !src->theta_dev || !src->logpost_dev) {
        yg_set_error("yg_load_state needs at least theta_dev and logpost_dev");
        return YG_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)e->cfg.n_chains, d = (size_t)e->cfg.dim, nl = (size_t)e->cfg.n_levels;
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
    YG_CUDA_CHECK(cudaMemcpyAsync(e->theta, src->theta_dev, 8 * d * n, k, st));
    YG_CUDA_CHECK(cudaMemcpyAsync(e->logpost, src->logpost_dev, 8 * nl * n, k, st));
    if (src->n_accept_dev) YG_CUDA_CHECK(cudaMemcpyAsync(e->n_accept, src->n_accept_dev, 8 * n, k, st));
    if (src->w_mean_dev) YG_CUDA_CHECK(cudaMemcpyAsync(e->w_mean, src->w_mean_dev, 8 * d * n, k, st));
    if (src->w_m2_dev) YG_CUDA_CHECK(cudaMemcpyAsync(e->w_m2, src->w_m2_dev, 8 * (d > YG_MAX_DIM ? d : d * d) * n, k, st));
    if (src->prop_L_dev && e->cfg.adaptive)
        YG_CUDA_CHECK(cudaMemcpyAsync(e->prop_L, src->prop_L_dev, 8 * d * d * n, k, st));
    if (src->am_mean_dev && e->cfg.adaptive)
        YG_CUDA_CHECK(cudaMemcpyAsync(e->am_mean, src->am_mean_dev, 8 * d * n, k, st));
    if (src->am_m2_dev && e->cfg.adaptive)
        YG_CUDA_CHECK(cudaMemcpyAsync(e->am_m2, src->am_m2_dev, 8 * d * d * n, k, st));
    if (e->cfg.aem) {
        const size_t dd = (size_t)e->aem_data_dim;
        if (src->aem_n_dev) YG_CUDA_CHECK(cudaMemcpyAsync(e->aem_n, src->aem_n_dev, 8 * n, k, st));
        if (src->aem_mean_dev) YG_CUDA_CHECK(cudaMemcpyAsync(e->aem_mean, src->aem_mean_dev, 8 * dd * n, k, st));
        if (src->aem_m2_dev) YG_CUDA_CHECK(cudaMemcpyAsync(e->aem_m2, src->aem_m2_dev, 8 * dd * n, k, st));
        if (src->aem_cache_dev) YG_CUDA_CHECK(cudaMemcpyAsync(e->aem_cache, src->aem_cache_dev, 8 * (3 * (d + 1) + 1) * n, k, st));
    }
    e->step_index = step_index;
    e->welford_n = welford_n;
    e->am_steps = am_steps;
    e->state_set = true;
    return YG_OK;
}

extern "C" int yg_get_counters(yg_ensemble *e, int64_t *out_host, void *stream)
{
    int rc = check_handle(e, false, false);
    if (rc) return rc;
    if (!out_host) return YG_ERR_INVALID;
    unsigned long long c[8];
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    YG_CUDA_CHECK(cudaMemcpyAsync(c, e->counters, sizeof(c), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    YG_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    out_host[0] = e->step_index;
    out_host[1] = (int64_t)c[0];
    out_host[2] = (int64_t)c[1];
    out_host[3] = (int64_t)c[2];
    out_host[4] = (int64_t)c[3];
    out_host[5] = e->welford_n;
    out_host[6] = e->am_steps;
    out_host[7] = (int64_t)c[4];
    out_host[8] = (int64_t)c[5];
    return YG_OK;
}

extern "C" int yg_logpost(yg_ensemble *e, int32_t level, const double *theta_dev, int64_t n, double *out_dev,
                          void *stream)
{
    int rc = check_handle(e, true, false);
    if (rc) return rc;
    if (level < 0 || level >= e->cfg.n_levels || !theta_dev || !out_dev || n < 1) {
        yg_set_error("yg_logpost: invalid arguments");
        return YG_ERR_INVALID;
    }
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    return e->big ? yg_launch_logpost_big(e, level, theta_dev, n, out_dev, (cudaStream_t)stream)
                  : yg_launch_logpost(e, level, theta_dev, n, out_dev, (cudaStream_t)stream);
}

extern "C" int yg_pooled_stats(yg_ensemble *e, double *out_dev, void *stream)
{
    int rc = check_handle(e, true, true);
    if (rc) return rc;
    if (!out_dev) return YG_ERR_INVALID;
    if (e->cfg.dim > YG_MAX_DIM) {
        yg_set_error("pooled statistics need the full second-moment matrix (dim <= %d)", YG_MAX_DIM);
        return YG_ERR_UNSUPPORTED;
    }
    YG_CUDA_CHECK(cudaSetDevice(e->cfg.device));
    return yg_pooled_impl(e->w_mean, e->w_m2, e->n_accept, e->cfg.dim, e->cfg.n_chains, e->welford_n,
                          e->pool_partials, POOL_PARTS, out_dev, (cudaStream_t)stream);
}

extern "C" int yg_last_launch(yg_ensemble *e, int32_t *grid, int32_t *block, int32_t *smem_bytes, int64_t *launches)
{
    if (!e) return YG_ERR_INVALID;
    if (grid) *grid = e->last_grid;
    if (block) *block = e->last_block;
    if (smem_bytes) *smem_bytes = e->last_smem;
    if (launches) *launches = e->launches;
    return YG_OK;
}

extern "C" int yg_fp64_tensor_peak(int32_t device, double ms, double *tflops_out)
{
    if (!tflops_out) return YG_ERR_INVALID;
    return yg_dmma_peak(device, ms, tflops_out);
}

#ifdef YG_BOUNDS_CHECK
// Bounds-checked development build only (not part of include/yagre_b200.h): proves that YG_CHK is live in THIS
// library by violating it on purpose.  Returns 1 when the device assert fired (the CUDA context is unusable
// afterwards: call it last, in a process of its own), 0 when the launch completed, i.e. the checks are compiled out.
namespace {
__global__ void bounds_selftest_kernel(int limit) { YG_CHK(threadIdx.x + 1, limit); }
}  // namespace
extern "C" int yg_bounds_check_selftest(void)
{
    bounds_selftest_kernel<<<1, 32>>>(1);
    return cudaDeviceSynchronize() == cudaErrorAssert ? 1 : 0;
}
#endif
