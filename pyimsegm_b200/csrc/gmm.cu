// gmm.cu -- the class model of the pipeline on the device: StandardScaler + full-covariance Gaussian mixture (EM).
//
// Replaces the host round trip of imsegm/graph_cuts.py:73-163 (estim_class_model, default 'GMM'):
//   sklearn Pipeline[StandardScaler, GaussianMixture(n_components=K, covariance_type='full',
//                                                    n_init=int(sqrt(max_iter)), max_iter=max_iter)]
//   followed by predict_proba (imsegm/pipelines.py:95-96).
// The EM restates sklearn.mixture.GaussianMixture (tol 1e-3 on the mean log-likelihood, reg_covar 1e-6,
// weights = (sum resp + 10 eps) / N, centred covariance, precision Cholesky, best of n_init by lower bound).
// Each of the n_init restarts runs in its own CTA, all restarts concurrently; the initial hard assignment is
// either supplied (init_labels: makes the fit deterministic and comparable with sklearn from the same start) or
// k-means++ / Lloyd with a counter-based RNG (the reference leaves the model unseeded, so only the algorithm,
// not a label-for-label result, can be matched; tests compare against sklearn from a shared initialisation).
// N is tiny (superpixels, not pixels): this stage is latency bound, it exists to remove the host sync.
#include "common.cuh"
#include <float.h>

namespace {

constexpr int GT = 512;      // threads per restart CTA
constexpr int DMAX = 16;     // feature dimensions handled on the device
constexpr int KMAX = 8;      // mixture components handled on the device

struct GmmWs {
    double* xs;       // [N, D] standardised features
    double* scale;    // [2 D] mean, scale
    double* resp;     // [n_init, N, K]
    int* lab;         // [n_init, N]
    double* par;      // [n_init, PSTRIDE]: weights K | means K D | cov K D D | prec_chol K D D | lower_bound | n_iter | converged | ok
    double* red;      // [n_init, GT] scratch
};

__host__ __device__ inline int pstride(int K, int D) { return K + K * D + 2 * K * D * D + 4; }

struct Rng {
    unsigned long long s;
    __device__ explicit Rng(unsigned long long seed) : s(seed) {}
    __device__ unsigned long long next()
    {
        unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    __device__ double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

__device__ double block_sum_d(double v, double* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int i = 0; i < GT / 32; ++i) t += s_red[i];
    return t;
}

// StandardScaler: mean / population std per feature (zero variance -> scale 1), single CTA
__global__ void __launch_bounds__(GT) k_gmm_scale(const double* __restrict__ feat, int N_in, const int* n_dev, int D, int ld, int use_scaler,
                                                  GmmWs w)
{
    __shared__ double s_red[GT / 32];
    const int N = n_dev ? min(*n_dev, N_in) : N_in;
    for (int d = 0; d < D; ++d) {
        double s = 0;
        for (int n = threadIdx.x; n < N; n += GT) s += feat[(size_t)n * ld + d];
        double mean = block_sum_d(s, s_red) / N;
        double v = 0;
        for (int n = threadIdx.x; n < N; n += GT) { double t = feat[(size_t)n * ld + d] - mean; v += t * t; }
        double var = block_sum_d(v, s_red) / N;
        double sc = sqrt(var);
        // sklearn _is_constant_feature: var <= N eps var + (N mean eps)^2  -> scale 1
        const double ub = N * DBL_EPSILON * var + (N * mean * DBL_EPSILON) * (N * mean * DBL_EPSILON);
        if (var <= ub) sc = 1.0;
        if (!use_scaler) { mean = 0.0; sc = 1.0; }
        if (threadIdx.x == 0) { w.scale[d] = mean; w.scale[D + d] = sc; }
        for (int n = threadIdx.x; n < N; n += GT) w.xs[(size_t)n * D + d] = (feat[(size_t)n * ld + d] - mean) / sc;
        __syncthreads();
    }
}

// log N(x | mu_k, prec_chol_k) for all k; returns log-sum-exp of (log w_k + log N_k), fills lw[k] with the terms
__device__ __forceinline__ double log_prob_all(const double* x, int D, int K, const double* wts, const double* mu, const double* pc,
                                               const double* logdet, double* lw)
{
    double mx = -DBL_MAX;
    for (int k = 0; k < K; ++k) {
        const double* m = mu + k * D;
        const double* U = pc + (size_t)k * D * D; // upper triangular: y_j = sum_{i<=j} (x_i - m_i) U[i][j]
        double q = 0;
        for (int j = 0; j < D; ++j) {
            double y = 0;
            for (int i = 0; i <= j; ++i) y += (x[i] - m[i]) * U[i * D + j];
            q += y * y;
        }
        double lp = -0.5 * (D * 1.8378770664093453 + q) + logdet[k]; // logdet[k] already holds log|prec_chol_k| + log w_k
        lw[k] = lp;
        mx = fmax(mx, lp);
    }
    double s = 0;
    for (int k = 0; k < K; ++k) s += exp(lw[k] - mx);
    return mx + log(s);
}

// parameters from responsibilities (sklearn _estimate_gaussian_parameters + _compute_precision_cholesky)
// returns false when a covariance is not positive definite
__device__ bool m_step(const double* __restrict__ xs, const double* __restrict__ resp, int N, int D, int K, double reg, double* par,
                       double* s_part, double* s_red)
{
    double* wts = par; double* mu = par + K; double* cov = mu + K * D; double* pc = cov + (size_t)K * D * D;
    // pass 1: nk and means, quantity-parallel over sample slices
    const int Q1 = K * (1 + D);
    for (int q0 = 0; q0 < Q1; q0 += GT) {
        const int nq = min(GT, Q1 - q0);
        const int S = max(1, GT / nq);
        const int q = q0 + (threadIdx.x % nq), sl = threadIdx.x / nq;
        double acc = 0;
        if (sl < S) {
            const int k = q / (1 + D), j = q % (1 + D);
            for (int n = sl; n < N; n += S) {
                double r = resp[(size_t)n * K + k];
                acc += j == 0 ? r : r * xs[(size_t)n * D + j - 1];
            }
        }
        s_part[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < nq) {
            double t = 0;
            for (int s2 = 0; s2 < S; ++s2) t += s_part[s2 * nq + threadIdx.x];
            const int k = q / (1 + D), j = q % (1 + D);
            if (j == 0) wts[k] = t + 10 * DBL_EPSILON; // nk (divided by N at the end)
            else mu[k * D + j - 1] = t;                 // sum r x (divided by nk below)
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < K * D; i += GT) mu[i] = mu[i] / wts[i / D];
    __syncthreads();
    // pass 2: centred covariances (upper triangle), quantity-parallel
    const int T = D * (D + 1) / 2, Q2 = K * T;
    for (int q0 = 0; q0 < Q2; q0 += GT) {
        const int nq = min(GT, Q2 - q0);
        const int S = max(1, GT / nq);
        const int q = q0 + (threadIdx.x % nq), sl = threadIdx.x / nq;
        int k = q / T, t = q % T, a = 0;
        while (t >= D - a) { t -= D - a; ++a; }
        const int b = a + t; // (a, b), a <= b
        double acc = 0;
        if (sl < S) {
            const double ma = mu[k * D + a], mb = mu[k * D + b];
            for (int n = sl; n < N; n += S)
                acc += resp[(size_t)n * K + k] * (xs[(size_t)n * D + a] - ma) * (xs[(size_t)n * D + b] - mb);
        }
        s_part[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x < nq) {
            double tt = 0;
            for (int s2 = 0; s2 < S; ++s2) tt += s_part[s2 * nq + threadIdx.x];
            double c = tt / wts[k] + (a == b ? reg : 0.0);
            cov[(size_t)k * D * D + a * D + b] = c;
            cov[(size_t)k * D * D + b * D + a] = c;
        }
        __syncthreads();
    }
    // Cholesky cov = L L^T, prec_chol = (L^-1)^T, one thread per component
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (threadIdx.x < K) {
        const int k = threadIdx.x;
        const double* Cm = cov + (size_t)k * D * D;
        double* U = pc + (size_t)k * D * D;
        double L[DMAX * DMAX];
        bool ok = true;
        for (int i = 0; i < D && ok; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = Cm[i * D + j];
                for (int p = 0; p < j; ++p) s -= L[i * D + p] * L[j * D + p];
                if (i == j) { if (!(s > 0)) { ok = false; break; } L[i * D + i] = sqrt(s); }
                else L[i * D + j] = s / L[j * D + j];
            }
        if (!ok) s_bad = 1;
        else {
            // Z = L^-1 (lower);  U = Z^T
            for (int c = 0; c < D; ++c)
                for (int r = 0; r < D; ++r) {
                    if (r < c) { U[c * D + r] = 0.0; continue; }
                    double s = (r == c) ? 1.0 : 0.0;
                    for (int p = c; p < r; ++p) s -= L[r * D + p] * U[c * D + p]; // U[c][p] holds Z[p][c]
                    U[c * D + r] = s / L[r * D + r];
                }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += GT) wts[k] = wts[k] / N;
    __syncthreads();
    return s_bad == 0;
}

// one CTA per restart
__global__ void __launch_bounds__(GT) k_gmm_fit(int N_in, const int* n_dev, int D, int K, int max_iter, double tol, double reg,
                                               unsigned long long seed, const int* __restrict__ init_labels, GmmWs w)
{
    __shared__ double s_part[GT];
    __shared__ double s_red[GT / 32];
    __shared__ double s_logdet[KMAX];
    __shared__ double s_cent[KMAX * DMAX];
    __shared__ int s_pick;
    const int N = n_dev ? min(*n_dev, N_in) : N_in;
    const int init = blockIdx.x;
    const double* xs = w.xs;
    double* resp = w.resp + (size_t)init * N_in * K;
    int* lab = w.lab + (size_t)init * N_in;
    double* par = w.par + (size_t)init * pstride(K, D);
    double* wts = par; double* mu = par + K; double* pc = mu + K * D + (size_t)K * D * D;
    double* tail = par + K + K * D + 2 * (size_t)K * D * D; // lower_bound, n_iter, converged, ok

    // ---- initial hard assignment ----
    if (init_labels) {
        for (int n = threadIdx.x; n < N; n += GT) lab[n] = init_labels[(size_t)init * N_in + n];
        __syncthreads();
    } else {
        // k-means++ seeding (one D^2-weighted draw per centre), then Lloyd
        Rng rng(seed * 0x100000001B3ull + 1469598103934665603ull * (unsigned long long)(init + 1));
        double* d2 = w.red + (size_t)init * N_in; // closest squared distance per sample
        int first = (int)(rng.uniform() * N); if (first >= N) first = N - 1;
        for (int d = threadIdx.x; d < D; d += GT) s_cent[d] = xs[(size_t)first * D + d];
        __syncthreads();
        for (int c = 1; c <= K; ++c) {
            // update closest distances with centre c-1
            double loc = 0;
            for (int n = threadIdx.x; n < N; n += GT) {
                double s = 0;
                for (int d = 0; d < D; ++d) { double t = xs[(size_t)n * D + d] - s_cent[(c - 1) * D + d]; s += t * t; }
                double cur = (c == 1) ? s : fmin(d2[n], s);
                d2[n] = cur;
                loc += cur;
            }
            double total = block_sum_d(loc, s_red);
            if (c == K) break;
            const double thr = rng.uniform() * total; // same on every thread (same rng state)
            // contiguous chunks -> prefix over chunk sums -> the chunk holding thr scans itself
            const int chunk = (N + GT - 1) / GT, beg = threadIdx.x * chunk, end = min(beg + chunk, N);
            double cs = 0;
            for (int n = beg; n < end; ++n) cs += d2[n];
            s_part[threadIdx.x] = cs;
            if (threadIdx.x == 0) s_pick = N - 1;
            __syncthreads();
            if (threadIdx.x == 0) {
                double run = 0; int t = 0;
                for (; t < GT; ++t) { if (run + s_part[t] >= thr) break; run += s_part[t]; }
                if (t < GT) {
                    int b2 = t * chunk, e2 = min(b2 + chunk, N), n = b2;
                    for (; n < e2; ++n) { run += d2[n]; if (run >= thr) break; }
                    s_pick = min(n, N - 1);
                }
            }
            __syncthreads();
            for (int d = threadIdx.x; d < D; d += GT) s_cent[c * D + d] = xs[(size_t)s_pick * D + d];
            __syncthreads();
        }
        // Lloyd iterations (sklearn KMeans: max_iter 300, tol 1e-4 * mean feature variance; X is standardised)
        for (int n = threadIdx.x; n < N; n += GT) lab[n] = -1;
        for (int it = 0; it < 300; ++it) {
            int changed = 0;
            for (int n = threadIdx.x; n < N; n += GT) {
                double best = DBL_MAX; int bk = 0;
                for (int k = 0; k < K; ++k) {
                    double s = 0;
                    for (int d = 0; d < D; ++d) { double t = xs[(size_t)n * D + d] - s_cent[k * D + d]; s += t * t; }
                    if (s < best) { best = s; bk = k; }
                }
                if (lab[n] != bk) { lab[n] = bk; changed = 1; }
            }
            changed = __syncthreads_or(changed);
            // new centres: count and coordinate sums of every cluster in ONE quantity-parallel pass over sample slices
            // (quantity q = (k, j): j == 0 the count, j >= 1 the sum of coordinate j-1)
            const int Q = K * (1 + D);                 // <= 8 * 17 = 136 <= GT
            const int S = max(1, GT / Q);
            {
                const int q = threadIdx.x % Q, sl = threadIdx.x / Q;
                double a = 0;
                if (sl < S) {
                    const int k = q / (1 + D), j = q % (1 + D);
                    for (int n = sl; n < N; n += S)
                        if (lab[n] == k) a += j == 0 ? 1.0 : xs[(size_t)n * D + j - 1];
                }
                s_part[threadIdx.x] = a;
            }
            __syncthreads();
            __shared__ double s_tot[KMAX * (1 + DMAX)];
            if (threadIdx.x < Q) {
                double t = 0;
                for (int s2 = 0; s2 < S; ++s2) t += s_part[s2 * Q + threadIdx.x];
                s_tot[threadIdx.x] = t;
            }
            __syncthreads();
            double shift = 0;
            for (int k = 0; k < K; ++k) {
                const double cnt = s_tot[k * (1 + D)];
                if (cnt > 0)
                    for (int d = 0; d < D; ++d) { const double t = s_tot[k * (1 + D) + 1 + d] / cnt - s_cent[k * D + d]; shift += t * t; }
            }
            __syncthreads();
            if (threadIdx.x < K * D) {
                const int k = threadIdx.x / D, d = threadIdx.x % D;
                const double cnt = s_tot[k * (1 + D)];
                if (cnt > 0) s_cent[threadIdx.x] = s_tot[k * (1 + D) + 1 + d] / cnt;
            }
            __syncthreads();
            if (!changed || shift <= 1e-4) break;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < N * K; i += GT) resp[i] = (lab[i / K] == i % K) ? 1.0 : 0.0;
    __syncthreads();

    // ---- EM ----
    bool ok = m_step(xs, resp, N, D, K, reg, par, s_part, s_red);
    double lower = -DBL_MAX;
    int it = 0, conv = 0;
    if (ok) {
        for (it = 1; it <= max_iter; ++it) {
            const double prev = lower;
            if (threadIdx.x < K) {
                double ld = 0;
                for (int j = 0; j < D; ++j) ld += log(pc[(size_t)threadIdx.x * D * D + j * D + j]);
                s_logdet[threadIdx.x] = ld + log(wts[threadIdx.x]);
            }
            __syncthreads();
            double acc = 0;
            for (int n = threadIdx.x; n < N; n += GT) {
                double lw[KMAX];
                double lse = log_prob_all(xs + (size_t)n * D, D, K, wts, mu, pc, s_logdet, lw);
                for (int k = 0; k < K; ++k) resp[(size_t)n * K + k] = exp(lw[k] - lse);
                acc += lse;
            }
            lower = block_sum_d(acc, s_red) / N;
            ok = m_step(xs, resp, N, D, K, reg, par, s_part, s_red);
            if (!ok) break;
            if (it > 1 && fabs(lower - prev) < tol) { conv = 1; break; }
        }
        if (it > max_iter) it = max_iter;
    }
    if (threadIdx.x == 0) { tail[0] = ok ? lower : -DBL_MAX; tail[1] = (double)it; tail[2] = (double)conv; tail[3] = ok ? 1.0 : 0.0; }
}

// select the best restart, evaluate predict_proba for every sample, export the parameters
__global__ void __launch_bounds__(256) k_gmm_predict(int N_in, const int* n_dev, int D, int K, int n_init, GmmWs w, double* proba,
                                                    double* params_out)
{
    const int N = n_dev ? min(*n_dev, N_in) : N_in;
    const int ps = pstride(K, D);
    int best = -1;
    double bl = 0;
    for (int i = 0; i < n_init; ++i) {
        const double* tail = w.par + (size_t)i * ps + K + K * D + 2 * (size_t)K * D * D;
        if (tail[3] != 0.0 && (best < 0 || tail[0] > bl)) { best = i; bl = tail[0]; }
    }
    __shared__ double s_logdet[KMAX];
    if (best < 0) { // every restart hit a singular covariance: NaN probabilities make the failure visible
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K; i += gridDim.x * blockDim.x) proba[i] = nan("");
        if (blockIdx.x == 0 && threadIdx.x == 0 && params_out) params_out[2 * D + ps - 1] = 0.0;
        return;
    }
    const double* par = w.par + (size_t)best * ps;
    const double* wts = par; const double* mu = par + K; const double* pc = mu + K * D + (size_t)K * D * D;
    if (threadIdx.x < K) {
        double ld = 0;
        for (int j = 0; j < D; ++j) ld += log(pc[(size_t)threadIdx.x * D * D + j * D + j]);
        s_logdet[threadIdx.x] = ld + log(wts[threadIdx.x]);
    }
    __syncthreads();
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        double lw[KMAX];
        double lse = log_prob_all(w.xs + (size_t)n * D, D, K, wts, mu, pc, s_logdet, lw);
        for (int k = 0; k < K; ++k) proba[(size_t)n * K + k] = exp(lw[k] - lse);
    }
    if (blockIdx.x == 0 && params_out) {
        for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) params_out[i] = w.scale[i];
        for (int i = threadIdx.x; i < ps; i += blockDim.x) params_out[2 * D + i] = par[i];
        if (threadIdx.x == 0) params_out[2 * D + ps] = (double)best;
    }
}

static size_t carve_gmm(GmmWs& w, void* ws, size_t bytes, int N, int D, int K, int n_init)
{
    WsCarver c(ws, bytes);
    w.xs = c.take<double>((size_t)N * D);
    w.scale = c.take<double>(2 * (size_t)D);
    w.resp = c.take<double>((size_t)n_init * N * K);
    w.lab = c.take<int>((size_t)n_init * N);
    w.par = c.take<double>((size_t)n_init * pstride(K, D));
    w.red = c.take<double>((size_t)n_init * (N > GT ? N : GT));
    return isb_align(c.off);
}

} // namespace

extern "C" size_t isb_gmm_workspace_bytes(int N, int D, int K, int n_init)
{
    GmmWs w;
    return carve_gmm(w, nullptr, 0, N, D, K, n_init);
}

extern "C" int isb_gmm_params_len(int D, int K) { return 2 * D + pstride(K, D) + 1; }

extern "C" int isb_gmm_fit_predict(const double* feat, int N, int D, int ld, const int32_t* n_dev, int K, int n_init, int max_iter,
                                   double tol, double reg_covar, int use_scaler, unsigned long long seed, const int32_t* init_labels,
                                   double* proba, double* params_out, void* ws, size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(feat && proba && ws, "null pointer");
    ISB_REQUIRE(N > 0 && D > 0 && ld >= D && K > 0 && n_init > 0 && max_iter > 0, "bad sizes");
    if (D > DMAX || K > KMAX) { isb_set_error("device GMM handles D <= %d and K <= %d (got D=%d K=%d)", DMAX, KMAX, D, K); return ISB_ERR_UNSUPPORTED; }
    GmmWs w;
    size_t need = carve_gmm(w, ws, ws_bytes, N, D, K, n_init);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_GMM, st);
    k_gmm_scale<<<1, GT, 0, st>>>(feat, N, n_dev, D, ld, use_scaler, w);
    ISB_LAUNCH_CHECK();
    k_gmm_fit<<<n_init, GT, 0, st>>>(N, n_dev, D, K, max_iter, tol, reg_covar, seed, init_labels, w);
    ISB_LAUNCH_CHECK();
    int blocks = (N + 255) / 256;
    if (blocks > 148) blocks = 148;
    k_gmm_predict<<<blocks, 256, 0, st>>>(N, n_dev, D, K, n_init, w, proba, params_out);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
