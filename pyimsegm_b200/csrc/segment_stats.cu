// segment_stats.cu -- per-superpixel colour statistics + centroids.
//
// Replaces the reference's native module imsegm/features_cython.pyx:
//   computeColorImage2dMean :81, computeColorImage2dEnergy :101, computeColorImage2dVariance :122,
//   normColorFeatures :59 (count + divide), and regionprops centroids of imsegm/superpixels.py:205-224.
// The reference makes 4 passes over the labels and 3 strided passes over the image PER statistic; here one
// pass produces sum, sum of squares, count and coordinate sums, a second pass the squared deviations from the
// f32 mean (the reference's two-pass variance, descriptors.py:291-295).  Pixels are converted to f32 and the
// products are formed in f32 exactly as the Cython code does (float val; val * val), accumulation is f64.
//
// Mapping: a thread owns one image column inside a strip of SROWS rows, so warp loads are coalesced along x and
// label runs along y (~ one superpixel height) are accumulated in registers; one flush of atomics per run.
// Algorithmic HBM bytes: pass 1 = image bytes + 4 B/px labels, pass 2 the same.
#include "common.cuh"

namespace {

constexpr int SROWS = 16;

struct StatWs {
    double* acc;      // [nb][6]  sum c0..c2, sumsq c0..c2
    double* var;      // [nb][3]
    long long* iacc;  // [nb][3]  count, sum row, sum col
    float* meanf;     // [nb][3]
};

__device__ __forceinline__ float clean(float v) { return isnan(v) ? 0.0f : v; } // np.nan_to_num (descriptors.py:824)

__global__ void __launch_bounds__(256) k_stats_pass1(const void* __restrict__ img, int dtype, const int* __restrict__ seg, int H, int W,
                                                     int y_off, StatWs ws)
{
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const int y0 = blockIdx.y * SROWS, y1 = min(y0 + SROWS, H);
    int cur = -1;
    double s0 = 0, s1 = 0, s2 = 0, e0 = 0, e1 = 0, e2 = 0;
    long long cnt = 0, sy = 0;
    for (int y = y0; y <= y1; ++y) {
        int l = -1;
        float v0 = 0, v1 = 0, v2 = 0;
        if (y < y1) {
            size_t p = (size_t)y * W + x;
            l = seg[p];
            if (img) {
                v0 = clean(load_as_f32(img, dtype, 3 * p));
                v1 = clean(load_as_f32(img, dtype, 3 * p + 1));
                v2 = clean(load_as_f32(img, dtype, 3 * p + 2));
            }
        }
        if (l != cur) {
            if (cur >= 0) {
                double* a = ws.acc + 6 * (size_t)cur;
                atomicAdd(a, s0); atomicAdd(a + 1, s1); atomicAdd(a + 2, s2);
                atomicAdd(a + 3, e0); atomicAdd(a + 4, e1); atomicAdd(a + 5, e2);
                unsigned long long* ia = (unsigned long long*)(ws.iacc + 3 * (size_t)cur);
                atomicAdd(ia, (unsigned long long)cnt);
                atomicAdd(ia + 1, (unsigned long long)sy);
                atomicAdd(ia + 2, (unsigned long long)(cnt * x));
            }
            cur = l;
            s0 = s1 = s2 = e0 = e1 = e2 = 0;
            cnt = 0; sy = 0;
        }
        if (y < y1) {
            s0 += (double)v0; s1 += (double)v1; s2 += (double)v2;
            e0 += (double)__fmul_rn(v0, v0); e1 += (double)__fmul_rn(v1, v1); e2 += (double)__fmul_rn(v2, v2);
            cnt += 1; sy += y + y_off;
        }
    }
}

__global__ void k_stats_means(int nb, StatWs ws)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nb) return;
    long long c = ws.iacc[3 * (size_t)k];
    for (int z = 0; z < 3; ++z) {
        double m = ws.acc[6 * (size_t)k + z];
        if (c > 0) m = m / (double)c;
        ws.meanf[3 * (size_t)k + z] = (float)m; // np.array(means, dtype=np.float32), descriptors.py:293
    }
}

__global__ void __launch_bounds__(256) k_stats_pass2(const void* __restrict__ img, int dtype, const int* __restrict__ seg, int H, int W,
                                                     StatWs ws)
{
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const int y0 = blockIdx.y * SROWS, y1 = min(y0 + SROWS, H);
    int cur = -1;
    double a0 = 0, a1 = 0, a2 = 0;
    float m0 = 0, m1 = 0, m2 = 0;
    for (int y = y0; y <= y1; ++y) {
        int l = -1;
        size_t p = (size_t)y * W + x;
        if (y < y1) l = seg[p];
        if (l != cur) {
            if (cur >= 0) {
                double* a = ws.var + 3 * (size_t)cur;
                atomicAdd(a, a0); atomicAdd(a + 1, a1); atomicAdd(a + 2, a2);
            }
            cur = l;
            a0 = a1 = a2 = 0;
            if (l >= 0) { m0 = ws.meanf[3 * (size_t)l]; m1 = ws.meanf[3 * (size_t)l + 1]; m2 = ws.meanf[3 * (size_t)l + 2]; }
        }
        if (y < y1) {
            float d0 = __fsub_rn(clean(load_as_f32(img, dtype, 3 * p)), m0);
            float d1 = __fsub_rn(clean(load_as_f32(img, dtype, 3 * p + 1)), m1);
            float d2 = __fsub_rn(clean(load_as_f32(img, dtype, 3 * p + 2)), m2);
            a0 += (double)__fmul_rn(d0, d0); a1 += (double)__fmul_rn(d1, d1); a2 += (double)__fmul_rn(d2, d2);
        }
    }
}

__device__ __forceinline__ double tidy(double v)
{
    if (isnan(v)) return 0.0;          // np.nan_to_num (descriptors.py:857)
    if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return v == 0.0 ? 0.0 : v;         // features[features == 0] = 0  (-0 -> +0)
}

__global__ void k_stats_finalize(int nb, int flags, StatWs ws, double* feat, int ld, int col0, double* centres, int* counts)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nb) return;
    long long c = ws.iacc[3 * (size_t)k];
    double dn = (double)c;
    int col = col0;
    double* row = feat ? feat + (size_t)k * ld : nullptr;
    if (row) {
        if (flags & 1) { for (int z = 0; z < 3; ++z) { double v = ws.acc[6 * (size_t)k + z]; if (c > 0) v = v / dn; row[col++] = tidy(v); } }
        if (flags & 2) { for (int z = 0; z < 3; ++z) { double v = ws.var[3 * (size_t)k + z]; if (c > 0) v = v / dn; row[col++] = tidy(sqrt(v)); } }
        if (flags & 4) { for (int z = 0; z < 3; ++z) { double v = ws.acc[6 * (size_t)k + 3 + z]; if (c > 0) v = v / dn; row[col++] = tidy(v); } }
    }
    if (centres) {
        if (c > 0) { centres[2 * (size_t)k] = (double)ws.iacc[3 * (size_t)k + 1] / dn; centres[2 * (size_t)k + 1] = (double)ws.iacc[3 * (size_t)k + 2] / dn; }
        else { centres[2 * (size_t)k] = -1.0; centres[2 * (size_t)k + 1] = -1.0; }
    }
    if (counts) counts[k] = (int)c;
}

static size_t carve(StatWs& w, void* ws, size_t bytes, int nb)
{
    WsCarver c(ws, bytes);
    w.acc = c.take<double>(6 * (size_t)nb);
    w.var = c.take<double>(3 * (size_t)nb);
    w.iacc = c.take<long long>(3 * (size_t)nb);
    w.meanf = c.take<float>(3 * (size_t)nb);
    return isb_align(c.off);
}

} // namespace

extern "C" size_t isb_segment_stats_workspace_bytes(int nb)
{
    StatWs w;
    return carve(w, nullptr, 0, nb);
}

extern "C" int isb_segment_stats_2d(const void* img, int dtype, const int32_t* seg, int H, int W, int nb, int flags, double* feat,
                                    int ld, int col0, double* centres, int32_t* counts, void* ws, size_t ws_bytes,
                                    isb_stream_t stream)
{
    ISB_REQUIRE(seg && ws && (img || (flags & 7) == 0), "null pointer");
    ISB_REQUIRE(H > 0 && W > 0 && nb > 0, "bad sizes");
    ISB_REQUIRE(dtype >= ISB_U8 && dtype <= ISB_F64, "bad dtype");
    StatWs w;
    size_t need = carve(w, ws, ws_bytes, nb);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_STATS, st);
    ISB_CUDA_CHECK(cudaMemsetAsync(ws, 0, need, st));
    dim3 grid((W + 255) / 256, (H + SROWS - 1) / SROWS);
    k_stats_pass1<<<grid, 256, 0, st>>>(img, dtype, seg, H, W, 0, w);
    ISB_LAUNCH_CHECK();
    if (flags & 2) {
        k_stats_means<<<(nb + 255) / 256, 256, 0, st>>>(nb, w);
        ISB_LAUNCH_CHECK();
        k_stats_pass2<<<grid, 256, 0, st>>>(img, dtype, seg, H, W, w);
        ISB_LAUNCH_CHECK();
    }
    k_stats_finalize<<<(nb + 255) / 256, 256, 0, st>>>(nb, flags, w, feat, ld, col0, centres, counts);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

// ---- caller-owned accumulators (row bands of one image merged by a collective between the calls) -------------------------

extern "C" int isb_segment_stats_accumulate(const void* img, int dtype, const int32_t* seg, int H, int W, int y_off, int nb, double* acc,
                                            int64_t* iacc, isb_stream_t stream)
{
    ISB_REQUIRE(seg && acc && iacc, "null pointer");
    ISB_REQUIRE(H > 0 && W > 0 && nb > 0 && y_off >= 0, "bad sizes");
    ISB_REQUIRE(dtype >= ISB_U8 && dtype <= ISB_F64, "bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_STATS, st);
    StatWs w;
    w.acc = acc; w.iacc = (long long*)iacc; w.var = nullptr; w.meanf = nullptr;
    dim3 grid((W + 255) / 256, (H + SROWS - 1) / SROWS);
    k_stats_pass1<<<grid, 256, 0, st>>>(img, dtype, seg, H, W, y_off, w);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" int isb_segment_stats_deviation(const void* img, int dtype, const int32_t* seg, int H, int W, int nb, const double* acc,
                                           const int64_t* iacc, float* meanf_scratch, double* var, isb_stream_t stream)
{
    ISB_REQUIRE(img && seg && acc && iacc && meanf_scratch && var, "null pointer");
    ISB_REQUIRE(H > 0 && W > 0 && nb > 0, "bad sizes");
    ISB_REQUIRE(dtype >= ISB_U8 && dtype <= ISB_F64, "bad dtype");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_STATS, st);
    StatWs w;
    w.acc = (double*)acc; w.iacc = (long long*)iacc; w.var = var; w.meanf = meanf_scratch;
    k_stats_means<<<(nb + 255) / 256, 256, 0, st>>>(nb, w);
    ISB_LAUNCH_CHECK();
    dim3 grid((W + 255) / 256, (H + SROWS - 1) / SROWS);
    k_stats_pass2<<<grid, 256, 0, st>>>(img, dtype, seg, H, W, w);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" int isb_segment_stats_finish(int nb, int flags, const double* acc, const double* var, const int64_t* iacc, double* feat, int ld,
                                        int col0, double* centres, int32_t* counts, isb_stream_t stream)
{
    ISB_REQUIRE(acc && iacc && (var || !(flags & 2)), "null pointer");
    ISB_REQUIRE(nb > 0, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    StatWs w;
    w.acc = (double*)acc; w.iacc = (long long*)iacc; w.var = (double*)var; w.meanf = nullptr;
    k_stats_finalize<<<(nb + 255) / 256, 256, 0, st>>>(nb, flags, w, feat, ld, col0, centres, counts);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
