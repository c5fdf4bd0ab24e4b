// native_misc.cu -- the remaining functions of the reference's native module imsegm/features_cython.pyx:
//   computeGrayImage3dMean :144, ...Energy :169, ...Variance :194   (f32 voxels, f64 accumulators; one channel, any rank)
//   computeLabelHistogram2d :222                                     (labels under a structuring element)
//   computeRayFeaturesBinary2d :239                                  (distance to the first boundary along rays, f32 marching)
// They sit beside the hot path (3-D gray pipeline, RG2SP / centre detection: SURVEY.md section 8f) and are provided so that
// the whole native surface of the reference has a device entry point.
#include "common.cuh"

namespace {

constexpr int GSTRIP = 16;

__device__ __forceinline__ float clean(float v) { return isnan(v) ? 0.0f : v; }

// pass 0: sum, sum of squares (f32 product), count.  pass 1: sum (v - mean_f32)^2.
__global__ void __launch_bounds__(256) k_gray_stats(const void* __restrict__ img, int dtype, const int* __restrict__ seg, long long n, int pass,
                                                    double* acc /* [nb][3]: sum, sumsq, var */, long long* cnt, const float* meanf)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long beg = t * GSTRIP, end = min(beg + GSTRIP, n);
    if (beg >= n) return;
    int cur = -1;
    double s = 0, e = 0;
    long long c = 0;
    float m = 0.f;
    for (long long i = beg; i <= end; ++i) {
        const int l = i < end ? seg[i] : -1;
        if (l != cur) {
            if (cur >= 0) {
                if (pass == 0) { atomicAdd(&acc[3 * (size_t)cur], s); atomicAdd(&acc[3 * (size_t)cur + 1], e); atomicAdd((unsigned long long*)&cnt[cur], (unsigned long long)c); }
                else atomicAdd(&acc[3 * (size_t)cur + 2], s);
            }
            cur = l; s = 0; e = 0; c = 0;
            if (pass == 1 && l >= 0) m = meanf[l];
        }
        if (i < end) {
            const float v = clean(load_as_f32(img, dtype, (size_t)i));
            if (pass == 0) { s += (double)v; e += (double)__fmul_rn(v, v); c += 1; }
            else { const float d = __fsub_rn(v, m); s += (double)__fmul_rn(d, d); }
        }
    }
}

__global__ void k_gray_means(int nb, const double* acc, const long long* cnt, float* meanf)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nb) return;
    double m = acc[3 * (size_t)k];
    if (cnt[k] > 0) m = m / (double)cnt[k];
    meanf[k] = (float)m;
}

__global__ void k_gray_finalize(int nb, int flags, const double* acc, const long long* cnt, double* feat, int ld, int col0)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nb) return;
    const long long c = cnt[k];
    int col = col0;
    double* row = feat + (size_t)k * ld;
    for (int st = 0; st < 3; ++st) {
        if (!(flags & (1 << st))) continue;
        double v = acc[3 * (size_t)k + (st == 0 ? 0 : (st == 1 ? 2 : 1))];
        if (c > 0) v = v / (double)c;
        if (st == 1) v = sqrt(v);
        row[col++] = isnan(v) ? 0.0 : (v == 0.0 ? 0.0 : v);
    }
}

__global__ void k_label_hist(const short* __restrict__ segm, const short* __restrict__ selem, long long n, int nb_labels, unsigned* hist)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = segm[i];
    if (l >= 0 && l < nb_labels && selem[i] == 1) atomicAdd(&hist[l], 1u);
}

// one thread per (position, angle)
__global__ void k_ray_features(const signed char* __restrict__ seg, int H, int W, const int* __restrict__ pos, int n_pos, const float* __restrict__ sin_a,
                               const float* __restrict__ cos_a, int n_ang, int edge, float* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pos * n_ang) return;
    const int p = t / n_ang, i = t - p * n_ang;
    const int pr = pos[2 * p], pc = pos[2 * p + 1];
    const signed char start = seg[(size_t)pr * W + pc];
    if (start && edge == 1) { out[t] = 0.0f; return; } // the position sits inside the border label
    float dist = -1.0f;
    const int diag = (int)sqrt((double)W * W + (double)H * H);
    float pos0 = (float)pr, pos1 = (float)pc;
    float g0 = sin_a[i], g1 = cos_a[i];
    const float gmax = fmaxf(fabsf(g0), fabsf(g1));
    g0 = __fdiv_rn(g0, gmax); g1 = __fdiv_rn(g1, gmax);
    signed char last = start;
    for (int s = 0; s < diag; ++s) {
        pos0 = __fadd_rn(pos0, g0); pos1 = __fadd_rn(pos1, g1);
        if (pos0 < 0 || roundf(pos0) >= H || pos1 < 0 || roundf(pos1) >= W) break;
        const signed char actual = seg[(size_t)((int)roundf(pos0)) * W + (int)roundf(pos1)];
        if ((edge == 1 && actual) || (edge == -1 && last && !actual)) {
            const float dx = __fsub_rn(pos0, (float)pr), dy = __fsub_rn(pos1, (float)pc);
            dist = (float)sqrt((double)__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            break;
        }
        last = actual;
    }
    out[t] = dist;
}

// joint histogram of two label maps: hist[a][b] = #{p : slic[p] == a, annot[p] == b}; a thread walks a short column strip so
// that runs of equal (a, b) cost one atomic
__global__ void __launch_bounds__(256) k_region_label_hist(const int* __restrict__ slic, const int* __restrict__ annot, int H, int W, int nb_annot,
                                                           unsigned* hist)
{
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const int y0 = blockIdx.y * GSTRIP, y1 = min(y0 + GSTRIP, H);
    int ca = -1, cb = -1;
    unsigned run = 0;
    for (int y = y0; y <= y1; ++y) {
        int a = -1, b = -1;
        if (y < y1) { a = slic[(size_t)y * W + x]; b = annot[(size_t)y * W + x]; }
        if (a != ca || b != cb) {
            if (run) atomicAdd(&hist[(size_t)ca * nb_annot + cb], run);
            ca = a; cb = b; run = 0;
        }
        if (y < y1) ++run;
    }
}

} // namespace

extern "C" int isb_region_label_hist(const int32_t* slic, const int32_t* annot, int H, int W, int nb_slic, int nb_annot, uint32_t* hist,
                                     isb_stream_t stream)
{
    ISB_REQUIRE(slic && annot && hist && H > 0 && W > 0 && nb_slic > 0 && nb_annot > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    ISB_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)nb_slic * nb_annot, st));
    k_region_label_hist<<<dim3((W + 255) / 256, (H + GSTRIP - 1) / GSTRIP), 256, 0, st>>>(slic, annot, H, W, nb_annot, hist);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" size_t isb_gray_stats_workspace_bytes(int nb) { return isb_align(sizeof(double) * 3 * (size_t)nb) + isb_align(sizeof(long long) * (size_t)nb) + isb_align(sizeof(float) * (size_t)nb) + 1024; }

extern "C" int isb_gray_stats(const void* img, int dtype, const int32_t* seg, long long n, int nb, int flags, double* feat, int ld, int col0,
                              void* ws, size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(img && seg && feat && ws, "null pointer");
    ISB_REQUIRE(n > 0 && nb > 0 && dtype >= ISB_U8 && dtype <= ISB_F64, "bad arguments");
    ISB_REQUIRE(ws_bytes >= isb_gray_stats_workspace_bytes(nb), "workspace too small");
    WsCarver c(ws, ws_bytes);
    double* acc = c.take<double>(3 * (size_t)nb);
    long long* cnt = c.take<long long>(nb);
    float* meanf = c.take<float>(nb);
    cudaStream_t st = (cudaStream_t)stream;
    ISB_CUDA_CHECK(cudaMemsetAsync(ws, 0, isb_align(c.off), st));
    const unsigned blocks = (unsigned)(((n + GSTRIP - 1) / GSTRIP + 255) / 256);
    k_gray_stats<<<blocks, 256, 0, st>>>(img, dtype, seg, n, 0, acc, cnt, meanf);
    ISB_LAUNCH_CHECK();
    if (flags & 2) {
        k_gray_means<<<(nb + 255) / 256, 256, 0, st>>>(nb, acc, cnt, meanf);
        ISB_LAUNCH_CHECK();
        k_gray_stats<<<blocks, 256, 0, st>>>(img, dtype, seg, n, 1, acc, cnt, meanf);
        ISB_LAUNCH_CHECK();
    }
    k_gray_finalize<<<(nb + 255) / 256, 256, 0, st>>>(nb, flags, acc, cnt, feat, ld, col0);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" int isb_label_hist_2d(const int16_t* segm_select, const int16_t* struc_elem, int H, int W, int nb_labels, uint32_t* hist,
                                 isb_stream_t stream)
{
    ISB_REQUIRE(segm_select && struc_elem && hist && H > 0 && W > 0 && nb_labels > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    ISB_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)nb_labels, st));
    const long long n = (long long)H * W;
    k_label_hist<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(segm_select, struc_elem, n, nb_labels, hist);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" int isb_ray_features_2d(const int8_t* seg_binary, int H, int W, const int32_t* positions, int n_pos, const float* sin_a,
                                   const float* cos_a, int n_ang, int edge, float* out, isb_stream_t stream)
{
    ISB_REQUIRE(seg_binary && positions && sin_a && cos_a && out, "null pointer");
    ISB_REQUIRE(H > 0 && W > 0 && n_pos > 0 && n_ang > 0 && (edge == 1 || edge == -1), "bad arguments");
    const int n = n_pos * n_ang;
    k_ray_features<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seg_binary, H, W, positions, n_pos, sin_a, cos_a, n_ang, edge, out);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
