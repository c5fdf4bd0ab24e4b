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
constexpr int DMAX = 16;     // feature dimensions of the single-kernel path (one CTA per restart, everything on chip)
constexpr int DBIG = 256;    // feature dimensions of the large-D path (batched GEMMs, e.g. colour + Leung-Malik = 189)
constexpr int KMAX = 8;      // mixture components handled on the device

struct GmmWs {
    double* xs;       // [N, D] standardised features
    double* scale;    // [2 D] mean, scale
    double* resp;     // [n_init, N, K]
    int* lab;         // [n_init, N]
    double* par;      // [n_init, PSTRIDE]: weights K | means K D | cov K D D | prec_chol K D D | lower_bound | n_iter | converged | ok
    double* red;      // [n_init, GT] scratch
    // large-D path only (D > DMAX)
    double* big;      // [n_init, K, N, D]  Y = X U (E-step) / sqrt(r) (X - mu) (M-step); never live at the same time
    double* bvec;     // [n_init, K, D]     mu U
    double* ldw;      // [n_init, K]        log|prec_chol| + log weight
    double* lowpart;  // [n_init, ceil(N / 8)] per-block sums of the log-likelihood
    double* cent;     // [n_init, K, D]     k-means centres
    double* tot;      // [n_init, K, 1 + D] k-means counts / coordinate sums
    double* state;    // [n_init, 4]        lower bound of the previous E-step, done, -, failed
    int* flag;        // [1]                restarts still running
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


// ---------------------------------------------------------------------------------------------------------------------
// Large-D path (DMAX < D <= DBIG), e.g. colour + full Leung-Malik statistics: D = 189 (BASELINE config 3).
// The same EM, restructured around its two contractions, all restarts and components in one launch each:
//   E-step:  Y[r,k] = X U[r,k]                (N x D x D per (restart, component))   -> k_dgemm_batched<false>
//   M-step:  C[r,k] = Xw[r,k]^T Xw[r,k],  Xw = sqrt(resp) (X - mu)   (D x D x N)     -> k_dgemm_batched<true>
// with a CTA-parallel Cholesky / triangular inverse per (r,k) between them.  The host loop runs the restarts in lock step
// and reads one int per iteration (restarts still running); a converged restart is frozen (its CTAs exit at once).
// FP64 throughout (the reference's scikit-learn model is float64); explicit fma() because the file is built with -fmad=false.
// ---------------------------------------------------------------------------------------------------------------------

constexpr int TM = 64, TN = 64, TK = 16;

// C[b] (M x Nn) = op(A[b]) B[b], row-major; TRANS_A: A[b] is stored Kd x M.  The sample count may come from the device (n_dev).
// Batch b = (restart r, component k) = (b / per, b % per); operand X of the batch starts at X + r * strideXr + k * strideXk.
// A batch whose restart is done is skipped.
struct BatchStride { size_t ar, ak, br, bk, cr, ck; };
template <bool TRANS_A>
__global__ void __launch_bounds__(256) k_dgemm_batched(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                                       double* __restrict__ C, int ldc, BatchStride bs, int M_in, int Nn, int Kd_in,
                                                       const int* n_dev, int n_is_m, const double* state, int per)
{
    __shared__ double As[TK][TM + 4];
    __shared__ double Bs[TK][TN + 4];
    const int b = blockIdx.z, br = b / per, bk = b % per;
    if (state[(size_t)br * 4 + 1] != 0.0) return;
    const int nlim = n_dev ? *n_dev : 0x7fffffff;
    const int M = n_is_m ? min(M_in, nlim) : M_in;       // the sample count is M (E-step) or Kd (M-step)
    const int Kd = n_is_m ? Kd_in : min(Kd_in, nlim);
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    if (m0 >= M) return;
    A += br * bs.ar + bk * bs.ak; B += br * bs.br + bk * bs.bk; C += br * bs.cr + bk * bs.ck;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < Kd; k0 += TK) {
        for (int i = threadIdx.x; i < TK * TM; i += 256) {
            int kk, mm;
            if (TRANS_A) { kk = i / TM; mm = i % TM; } else { mm = i / TK; kk = i % TK; }
            const int gm = m0 + mm, gk = k0 + kk;
            double v = 0.0;
            if (gm < M && gk < Kd) v = TRANS_A ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[kk][mm] = v;
        }
        for (int i = threadIdx.x; i < TK * TN; i += 256) {
            const int kk = i / TN, nn = i % TN;
            const int gk = k0 + kk, gn = n0 + nn;
            Bs[kk][nn] = (gk < Kd && gn < Nn) ? B[(size_t)gk * ldb + gn] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            double a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; bb[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn < Nn) C[(size_t)gm * ldc + gn] = acc[i][j];
        }
    }
}

// initial hard assignment of every restart: supplied labels, or k-means++ / Lloyd (as k_gmm_fit, centres in global memory)
__global__ void __launch_bounds__(GT) k_big_init(int N_in, const int* n_dev, int D, int K, unsigned long long seed,
                                                const int* __restrict__ init_labels, GmmWs w)
{
    __shared__ double s_part[GT];
    __shared__ double s_red[GT / 32];
    __shared__ int s_pick;
    __shared__ double s_shift;
    const int N = n_dev ? min(*n_dev, N_in) : N_in;
    const int init = blockIdx.x;
    const double* xs = w.xs;
    double* resp = w.resp + (size_t)init * N_in * K;
    int* lab = w.lab + (size_t)init * N_in;
    double* cent = w.cent + (size_t)init * K * D;
    double* tot = w.tot + (size_t)init * K * (1 + D);
    if (threadIdx.x == 0) { w.state[init * 4] = -DBL_MAX; w.state[init * 4 + 1] = 0.0; w.state[init * 4 + 2] = 0.0; w.state[init * 4 + 3] = 0.0; }
    if (init_labels) {
        for (int n = threadIdx.x; n < N; n += GT) lab[n] = init_labels[(size_t)init * N_in + n];
        __syncthreads();
    } else {
        Rng rng(seed * 0x100000001B3ull + 1469598103934665603ull * (unsigned long long)(init + 1));
        double* d2 = w.red + (size_t)init * (N_in > GT ? N_in : GT);
        int first = (int)(rng.uniform() * N); if (first >= N) first = N - 1;
        for (int d = threadIdx.x; d < D; d += GT) cent[d] = xs[(size_t)first * D + d];
        __syncthreads();
        for (int c = 1; c <= K; ++c) {
            double loc = 0;
            for (int n = threadIdx.x; n < N; n += GT) {
                double s = 0;
                for (int d = 0; d < D; ++d) { double t = xs[(size_t)n * D + d] - cent[(c - 1) * D + d]; s += t * t; }
                double cur = (c == 1) ? s : fmin(d2[n], s);
                d2[n] = cur;
                loc += cur;
            }
            double total = block_sum_d(loc, s_red);
            if (c == K) break;
            const double thr = rng.uniform() * total;
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
            for (int d = threadIdx.x; d < D; d += GT) cent[c * D + d] = xs[(size_t)s_pick * D + d];
            __syncthreads();
        }
        for (int n = threadIdx.x; n < N; n += GT) lab[n] = -1;
        const int Q = K * (1 + D);
        for (int it = 0; it < 300; ++it) {
            int changed = 0;
            for (int n = threadIdx.x; n < N; n += GT) {
                double best = DBL_MAX; int bk = 0;
                for (int k = 0; k < K; ++k) {
                    double s = 0;
                    for (int d = 0; d < D; ++d) { double t = xs[(size_t)n * D + d] - cent[k * D + d]; s += t * t; }
                    if (s < best) { best = s; bk = k; }
                }
                if (lab[n] != bk) { lab[n] = bk; changed = 1; }
            }
            changed = __syncthreads_or(changed);
            for (int q = threadIdx.x; q < Q; q += GT) {
                const int k = q / (1 + D), j = q % (1 + D);
                double a = 0;
                for (int n = 0; n < N; ++n)
                    if (lab[n] == k) a += j == 0 ? 1.0 : xs[(size_t)n * D + j - 1];
                tot[q] = a;
            }
            __syncthreads();
            double sh = 0;
            for (int i = threadIdx.x; i < K * D; i += GT) {
                const int k = i / D, d = i % D;
                const double cnt = tot[k * (1 + D)];
                if (cnt > 0) { const double t = tot[k * (1 + D) + 1 + d] / cnt - cent[i]; sh += t * t; }
            }
            sh = block_sum_d(sh, s_red);
            if (threadIdx.x == 0) s_shift = sh;
            __syncthreads();
            for (int i = threadIdx.x; i < K * D; i += GT) {
                const int k = i / D, d = i % D;
                const double cnt = tot[k * (1 + D)];
                if (cnt > 0) cent[i] = tot[k * (1 + D) + 1 + d] / cnt;
            }
            __syncthreads();
            if (!changed || s_shift <= 1e-4) break;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < N * K; i += GT) resp[i] = (lab[i / K] == i % K) ? 1.0 : 0.0;
}

// nk and means of every (restart, component): grid (K, n_init); thread per feature, samples in sequence
__global__ void __launch_bounds__(256) k_big_means(int N_in, const int* n_dev, int D, int K, GmmWs w)
{
    const int k = blockIdx.x, init = blockIdx.y;
    if (w.state[init * 4 + 1] != 0.0) return;
    const int N = n_dev ? min(*n_dev, N_in) : N_in;
    const double* resp = w.resp + (size_t)init * N_in * K;
    double* par = w.par + (size_t)init * pstride(K, D);
    double* wts = par; double* mu = par + K;
    __shared__ double s_nk;
    if (threadIdx.x == 0) {
        double t = 0;
        for (int n = 0; n < N; ++n) t += resp[(size_t)n * K + k];
        s_nk = t + 10 * DBL_EPSILON;
        wts[k] = s_nk;   // nk; divided by N once the covariance has used it (k_big_chol)
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += 256) {
        double a = 0;
        for (int n = 0; n < N; ++n) a = fma(resp[(size_t)n * K + k], w.xs[(size_t)n * D + d], a);
        mu[k * D + d] = a / s_nk;
    }
}

// Xw[r,k][n][d] = sqrt(resp) (x - mu): grid (ceil(N*D/256), K, n_init)
__global__ void __launch_bounds__(256) k_big_weighted(int N_in, const int* n_dev, int D, int K, GmmWs w)
{
    const int k = blockIdx.y, init = blockIdx.z;
    if (w.state[init * 4 + 1] != 0.0) return;
    const int N = n_dev ? min(*n_dev, N_in) : N_in;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)N * D) return;
    const int n = (int)(i / D), d = (int)(i % D);
    const double* mu = w.par + (size_t)init * pstride(K, D) + K;
    const double r = w.resp[((size_t)init * N_in + n) * K + k];
    w.big[(((size_t)init * K + k) * N_in) * D + i] = sqrt(r) * (w.xs[i] - mu[k * D + d]);
}

// per (restart, component): covariance from the raw Gram matrix, Cholesky, prec_chol = (L^-1)^T, log-determinant + log weight,
// b = mu U.  One CTA of 1024 threads; the matrix lives in global memory (L2), __syncthreads orders the accesses.
__global__ void __launch_bounds__(1024) k_big_chol(int N_in, const int* n_dev, int D, int K, double reg, GmmWs w)
{
    const int k = blockIdx.x, init = blockIdx.y;
    if (w.state[init * 4 + 1] != 0.0) return;
    const int N = n_dev ? min(*n_dev, N_in) : N_in;
    double* par = w.par + (size_t)init * pstride(K, D);
    double* wts = par; double* mu = par + K + (size_t)k * D;
    double* Cm = par + K + K * D + (size_t)k * D * D;                       // in: raw Gram; out: covariance
    double* U = par + K + K * D + (size_t)K * D * D + (size_t)k * D * D;   // out: prec_chol (upper); scratch: L (lower)
    const int T = blockDim.x, tid = threadIdx.x;
    const double nk = wts[k];
    __shared__ int s_bad;
    __shared__ double s_piv;
    if (tid == 0) s_bad = 0;
    // covariance (symmetrised from the upper triangle of the Gram matrix) and a working copy L
    for (int i = tid; i < D * D; i += T) {
        const int a = i / D, b = i % D;
        const double g = a <= b ? Cm[a * D + b] : Cm[b * D + a];
        U[i] = g / nk + (a == b ? reg : 0.0);
    }
    __syncthreads();
    for (int i = tid; i < D * D; i += T) Cm[i] = U[i];
    __syncthreads();
    // right-looking Cholesky in U's storage (lower triangle holds L)
    for (int j = 0; j < D; ++j) {
        if (tid == 0) {
            const double d = U[j * D + j];
            if (!(d > 0)) s_bad = 1;
            s_piv = sqrt(d);
            U[j * D + j] = s_piv;
        }
        __syncthreads();
        if (s_bad) break;
        const double piv = s_piv;
        for (int i = j + 1 + tid; i < D; i += T) U[i * D + j] = U[i * D + j] / piv;
        __syncthreads();
        const int m = D - j - 1;
        for (int idx = tid; idx < m * m; idx += T) {
            const int i = j + 1 + idx / m, c = j + 1 + idx % m;
            if (c <= i) U[i * D + c] = fma(-U[i * D + j], U[c * D + j], U[i * D + c]);
        }
        __syncthreads();
    }
    if (s_bad) {
        if (tid == 0) { w.state[init * 4 + 1] = 1.0; w.state[init * 4 + 3] = 1.0; } // done, failed
        return;
    }
    // Z = L^-1 (lower) by forward substitution, one warp per column c:
    //     Z[c][c] = 1 / L[c][c],   Z[r][c] = -(sum_{p=c}^{r-1} L[r][p] Z[p][c]) / L[r][r]   (r > c)
    // Z[p][c] is written to the UPPER triangle at U[c][p] -- that is prec_chol = Z^T, exactly where it has to end up; the strict
    // lower triangle keeps L until every column is done, the diagonal of L moves to shared memory first.
    __shared__ double s_diag[DBIG];
    for (int j = tid; j < D; j += T) s_diag[j] = U[j * D + j];
    __syncthreads();
    const int lane = tid & 31, wid = tid >> 5, nw = T >> 5;
    for (int c = wid; c < D; c += nw) {
        if (lane == 0) U[c * D + c] = 1.0 / s_diag[c];
        __syncwarp();
        for (int r = c + 1; r < D; ++r) {
            double sum = 0;
            for (int p = c + lane; p < r; p += 32) sum = fma(U[r * D + p], U[c * D + p], sum);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) U[c * D + r] = -sum / s_diag[r];
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = tid; i < D * D; i += T) { const int a = i / D, b = i % D; if (a > b) U[i] = 0.0; }
    __syncthreads();
    // log|prec_chol| + log weight, b = mu U, weight
    if (tid == 0) {
        double ld = 0;
        for (int j = 0; j < D; ++j) ld -= log(s_diag[j]);
        w.ldw[init * K + k] = ld + log(nk / N);
    }
    for (int j = tid; j < D; j += T) {
        double a = 0;
        for (int i = 0; i <= j; ++i) a = fma(mu[i], U[i * D + j], a);
        w.bvec[((size_t)init * K + k) * D + j] = a;
    }
    __syncthreads();
    if (tid == 0) wts[k] = nk / N;
}

// E-step after the GEMM: warp per sample; log N_k from |Y[r,k][n] - b[r,k]|^2, responsibilities, per-block log-likelihood sums
__global__ void __launch_bounds__(256) k_big_estep(int N_in, const int* n_dev, int D, int K, GmmWs w)
{
    const int init = blockIdx.y;
    if (w.state[init * 4 + 1] != 0.0) return;
    const int N = n_dev ? min(*n_dev, N_in) : N_in;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int n = blockIdx.x * 8 + wl;
    __shared__ double s_lse[8];
    double lse = 0.0;
    if (n < N) {
        double lw[KMAX];
        double mx = -DBL_MAX;
        for (int k = 0; k < K; ++k) {
            const double* y = w.big + (((size_t)init * K + k) * N_in + n) * D;
            const double* b = w.bvec + ((size_t)init * K + k) * D;
            double q = 0;
            for (int j = lane; j < D; j += 32) { const double t = y[j] - b[j]; q = fma(t, t, q); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            lw[k] = -0.5 * (D * 1.8378770664093453 + q) + w.ldw[init * K + k];
            mx = fmax(mx, lw[k]);
        }
        double s = 0;
        for (int k = 0; k < K; ++k) s += exp(lw[k] - mx);
        lse = mx + log(s);
        if (lane == 0)
            for (int k = 0; k < K; ++k) w.resp[((size_t)init * N_in + n) * K + k] = exp(lw[k] - lse);
    }
    if (lane == 0) s_lse[wl] = lse;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < 8; ++i) t += s_lse[i];
        w.lowpart[(size_t)init * ((N_in + 7) / 8) + blockIdx.x] = t;
    }
}

// after E + M: the lower bound of this iteration, convergence, bookkeeping (sklearn: the M-step runs before the test)
__global__ void k_big_converge(int N_in, const int* n_dev, int D, int K, int n_init, int it, int max_iter, double tol, GmmWs w)
{
    const int N = n_dev ? min(*n_dev, N_in) : N_in;
    const int init = threadIdx.x;
    const int nblk = (N + 7) / 8, stride = (N_in + 7) / 8;
    __shared__ int s_running;
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    if (init < n_init) {
        double* st = w.state + init * 4;
        double* tail = w.par + (size_t)init * pstride(K, D) + K + K * D + 2 * (size_t)K * D * D;
        if (st[1] == 0.0) {
            double t = 0;
            for (int i = 0; i < nblk; ++i) t += w.lowpart[(size_t)init * stride + i];
            const double lower = t / N;
            const bool conv = it > 1 && fabs(lower - st[0]) < tol;
            st[0] = lower;
            tail[0] = lower; tail[1] = (double)it; tail[2] = conv ? 1.0 : 0.0; tail[3] = 1.0;
            if (conv || it >= max_iter) st[1] = 1.0; else atomicAdd(&s_running, 1);
        } else if (st[3] != 0.0) {
            tail[0] = -DBL_MAX; tail[3] = 0.0; // a covariance was not positive definite
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *w.flag = s_running;
}

static int fit_big(int N, const int* n_dev, int D, int K, int n_init, int max_iter, double tol, double reg, unsigned long long seed,
                   const int* init_labels, GmmWs& w, cudaStream_t st)
{
    ISB_REQUIRE(n_init <= 1024, "too many restarts");
    const int RK = n_init * K;
    const size_t sND = (size_t)N * D, sDD = (size_t)D * D;
    const int ps = pstride(K, D);
    auto m_step = [&]() -> int {
        k_big_means<<<dim3(K, n_init), 256, 0, st>>>(N, n_dev, D, K, w);
        ISB_LAUNCH_CHECK();
        k_big_weighted<<<dim3((unsigned)((sND + 255) / 256), K, n_init), 256, 0, st>>>(N, n_dev, D, K, w);
        ISB_LAUNCH_CHECK();
        // Gram matrices straight into the covariance slots of the parameter blocks (par + r * ps + K + K D + k D D)
        {
            const BatchStride bs = { (size_t)K * sND, sND, (size_t)K * sND, sND, (size_t)ps, sDD };
            k_dgemm_batched<true><<<dim3((D + TN - 1) / TN, (D + TM - 1) / TM, RK), 256, 0, st>>>(
                w.big, D, w.big, D, w.par + K + K * D, D, bs, D, D, N, n_dev, 0, w.state, K);
            ISB_LAUNCH_CHECK();
        }
        k_big_chol<<<dim3(K, n_init), 1024, 0, st>>>(N, n_dev, D, K, reg, w);
        ISB_LAUNCH_CHECK();
        return ISB_OK;
    };
    k_big_init<<<n_init, GT, 0, st>>>(N, n_dev, D, K, seed, init_labels, w);
    ISB_LAUNCH_CHECK();
    if (int rc = m_step()) return rc;
    for (int it = 1; it <= max_iter; ++it) {
        {
            const BatchStride bs = { 0, 0, (size_t)ps, sDD, (size_t)K * sND, sND };
            k_dgemm_batched<false><<<dim3((D + TN - 1) / TN, (N + TM - 1) / TM, RK), 256, 0, st>>>(
                w.xs, D, w.par + K + K * D + (size_t)K * sDD, D, w.big, D, bs, N, D, D, n_dev, 1, w.state, K);
            ISB_LAUNCH_CHECK();
        }
        k_big_estep<<<dim3((N + 7) / 8, n_init), 256, 0, st>>>(N, n_dev, D, K, w);
        ISB_LAUNCH_CHECK();
        if (int rc = m_step()) return rc;
        k_big_converge<<<1, 1024, 0, st>>>(N, n_dev, D, K, n_init, it, max_iter, tol, w);
        ISB_LAUNCH_CHECK();
        int running = 0;
        ISB_CUDA_CHECK(cudaMemcpyAsync(&running, w.flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        ISB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (running == 0) break;
    }
    return ISB_OK;
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
    if (D > DMAX) {
        w.big = c.take<double>((size_t)n_init * K * N * D);
        w.bvec = c.take<double>((size_t)n_init * K * D);
        w.ldw = c.take<double>((size_t)n_init * K);
        w.lowpart = c.take<double>((size_t)n_init * ((N + 7) / 8));
        w.cent = c.take<double>((size_t)n_init * K * D);
        w.tot = c.take<double>((size_t)n_init * K * (1 + D));
        w.state = c.take<double>((size_t)n_init * 4);
        w.flag = c.take<int>(1);
    }
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
    if (D > DBIG || K > KMAX) { isb_set_error("device GMM handles D <= %d and K <= %d (got D=%d K=%d)", DBIG, KMAX, D, K); return ISB_ERR_UNSUPPORTED; }
    GmmWs w;
    size_t need = carve_gmm(w, ws, ws_bytes, N, D, K, n_init);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_GMM, st);
    k_gmm_scale<<<1, GT, 0, st>>>(feat, N, n_dev, D, ld, use_scaler, w);
    ISB_LAUNCH_CHECK();
    if (D > DMAX) {
        if (int rc = fit_big(N, n_dev, D, K, n_init, max_iter, tol, reg_covar, seed, init_labels, w, st)) return rc;
    } else {
        k_gmm_fit<<<n_init, GT, 0, st>>>(N, n_dev, D, K, max_iter, tol, reg_covar, seed, init_labels, w);
        ISB_LAUNCH_CHECK();
    }
    int blocks = (N + 255) / 256;
    if (blocks > 148) blocks = 148;
    k_gmm_predict<<<blocks, 256, 0, st>>>(N, n_dev, D, K, n_init, w, proba, params_out);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
