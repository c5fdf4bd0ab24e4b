// lm_texture.cu -- Leung-Malik texture descriptors: background subtraction + filter-bank contraction + statistics.
//
// Replaces imsegm/descriptors.py:1041-1106 compute_texture_desc_lm_img2d_clr:
//   :1078      img - gaussian_filter(img.astype(float), 150)      (sigma 150 on ALL THREE axes, mode reflect, 1201 taps)
//   :1085-1098 per battery: ndimage.convolve of every channel with every 33x33 kernel, max over orientations (:951-966),
//              clip at 1e6 (:1088), scale by log(1 + ||r||) / 0.03 / ||r|| over the whole [3,H,W] response (:1090-1094),
//              per-superpixel mean / std / energy (compute_image2d_color_statistic -> features_cython.pyx)
//
// (1) background: separable FP64 correlation; one kernel blurs along the image rows axis with coalesced column access
//     (a thread owns a column and 16 consecutive output rows, the 1201 weights slide through registers), the other axis
//     reuses it on a transposed copy; the channel axis (length 3, reflected 1201-tap kernel) folds into a 3x3 mix.
// (2) contraction: implicit GEMM  response[pixel, filter] = sum_taps patch[pixel, tap] * K[tap, filter]  on the TENSOR
//     cores: mma.sync.m16n8k8 TF32 with the 3xTF32 split (a_hi b_hi + a_hi b_lo + a_lo b_hi, FP32 accumulate), which
//     keeps the result at f32 accuracy -- the reference rounds every response to f32 before its statistics
//     (descriptors.py:233).  A fragments are read straight from the shared-memory image tile (shifted windows, no im2col
//     copy); the pre-split weights stream through a cp.async double buffer, one filter row (33 -> 40 taps) per stage.
//     FLOPs: 3 channels x 76 kernels x 33^2 x 2 = 496 584 per pixel (x3 for the split).
// (3) epilogue, fused: max over the orientations of a battery (quad shuffles), clip, then sum r and sum r^2 per
//     (superpixel, battery, channel) and globally per battery -- the responses are never written to memory.  The
//     log-norm scale is applied to the sums afterwards (mean and std scale with a, energy with a^2).
// Tolerance against the float64 oracle: 2e-4 relative on the features (stated in tests/test_gpu_texture.py).
#include "common.cuh"
#include <cuda_pipeline.h>

namespace {

// ---------------------------------------------------------------- (1) background ----------------------------------------------------

constexpr int VB_R = 16;      // output rows per thread
constexpr int VB_T = 128;     // threads (columns) per CTA

__device__ __forceinline__ int reflect_idx(int i, int n)
{
    if (n == 1) return 0;
    int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return (i < n) ? i : (p - 1 - i);
}

__global__ void __launch_bounds__(256) k_lm_to_planar(const void* __restrict__ img, int dtype, size_t npx, double* __restrict__ out)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c * npx + p] = load_as_f64(img, dtype, 3 * p + c);
}

// blur along axis 0 of [planes][n0][n1] (n1 contiguous).  wfull: 2*radius+1 weights (symmetric), in global memory.
__global__ void __launch_bounds__(VB_T) k_lm_vblur(const double* __restrict__ in, int n0, int n1, const double* __restrict__ wfull, int radius,
                                                   double* __restrict__ out)
{
    extern __shared__ double s_w[]; // padded weights: index d + radius + VB_R for d in [-(radius+VB_R), radius+VB_R]
    const int wn = 2 * (radius + VB_R) + 1;
    for (int i = threadIdx.x; i < wn; i += VB_T) {
        int d = i - (radius + VB_R);
        s_w[i] = (d >= -radius && d <= radius) ? wfull[d + radius] : 0.0;
    }
    __syncthreads();
    const int x = blockIdx.x * VB_T + threadIdx.x;
    const int y0 = blockIdx.y * VB_R;
    const size_t plane = (size_t)blockIdx.z * n0 * n1;
    if (x >= n1) return;
    double acc[VB_R];
#pragma unroll
    for (int r = 0; r < VB_R; ++r) acc[r] = 0.0;
    // input rows i = y0 - radius .. y0 + VB_R - 1 + radius; output row y0 + r uses weight w[i - (y0 + r)]
    const int i_beg = y0 - radius, i_end = y0 + VB_R - 1 + radius;
    const double* wc = s_w + (radius + VB_R); // wc[d]
    for (int i = i_beg; i <= i_end; ++i) {
        const double v = in[plane + (size_t)reflect_idx(i, n0) * n1 + x];
        const int d0 = i - y0;
#pragma unroll
        for (int r = 0; r < VB_R; ++r) acc[r] = fma(wc[d0 - r], v, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < VB_R; ++r)
        if (y0 + r < n0) out[plane + (size_t)(y0 + r) * n1 + x] = acc[r];
}

// [planes][n0][n1] -> [planes][n1][n0]
__global__ void __launch_bounds__(256) k_lm_transpose(const double* __restrict__ in, int n0, int n1, double* __restrict__ out)
{
    __shared__ double tile[32][33];
    const size_t plane = (size_t)blockIdx.z * n0 * n1;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8)
        if (by + j < n0 && bx + tx < n1) tile[j][tx] = in[plane + (size_t)(by + j) * n1 + bx + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (bx + j < n1 && by + tx < n0) out[plane + (size_t)(bx + j) * n0 + by + tx] = tile[tx][j];
}

// out[c][y][x] = (float)(in[c][y][x] - sum_c' mix[c][c'] * blurT[c'][x][y])   (blurT is [3][W][H])
struct Mix3 { double m[9]; };
__global__ void __launch_bounds__(256) k_lm_mix_sub(const double* __restrict__ in, const double* __restrict__ blurT, int H, int W, Mix3 mix,
                                                    float* __restrict__ out)
{
    __shared__ double tile[3][32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t npx = (size_t)H * W;
    for (int c = 0; c < 3; ++c)
        for (int j = ty; j < 32; j += 8)
            if (bx + j < W && by + tx < H) tile[c][j][tx] = blurT[c * npx + (size_t)(bx + j) * H + by + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int y = by + j, x = bx + tx;
        if (y >= H || x >= W) continue;
        const double b0 = tile[0][tx][j], b1 = tile[1][tx][j], b2 = tile[2][tx][j];
        for (int c = 0; c < 3; ++c) {
            const double bl = mix.m[3 * c] * b0 + mix.m[3 * c + 1] * b1 + mix.m[3 * c + 2] * b2;
            out[c * npx + (size_t)y * W + x] = (float)(in[c * npx + (size_t)y * W + x] - bl);
        }
    }
}

// ---------------------------------------------------------------- (2)+(3) contraction + statistics ---------------------------------

constexpr int KW = 33;             // kernel edge
constexpr int KWP = 40;            // taps of one kernel row, padded to a multiple of 8
constexpr int KRAD = 16;
constexpr int TM_X = 16;           // an M tile = 16 consecutive pixels of one image row
constexpr int TM_ROWS = 16;        // image rows per CTA: 8 warps x 2 M tiles
constexpr int TILE_W = TM_X + KWP; // 56 floats per staged image row
constexpr int TILE_H = TM_ROWS + KW - 1; // 48
constexpr int NPMAX = 80;          // padded filter count (10 n-tiles)
constexpr int NSTRIDE = 88;        // smem row stride of the weight stage (bank-conflict free for the B fragments)
constexpr int LSLOTS = 8;          // distinct superpixels a 16x16 tile may touch before the slow path

struct LmArgs {
    const float* img;   // [3][H][W] background-subtracted
    const int* seg;     // [H][W]
    int H, W, nb;
    const float* w_hi;  // [KW][KWP][NP] tf32-rounded weights (kernel already flipped: correlation form)
    const float* w_lo;  // [KW][KWP][NP] tf32-rounded remainders
    int NP, n_tiles, orient, n_orient_tiles, n_batt;
    double* S1;         // [nb][n_batt*3] sum r
    double* S2;         // [nb][n_batt*3] sum r^2
    double* G2;         // [n_batt] global sum r^2 over all pixels and channels
};

__device__ __forceinline__ unsigned f2tf32(float x)
{
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// battery index of an oriented n-tile half / of a single column
__device__ __forceinline__ int batt_of_oriented(int tile, int half, int orient)
{
    // orient 8: tile 2s = edge of sigma s, tile 2s+1 = bar;  orient 4: tile s holds edge (cols 0-3) and bar (cols 4-7)
    return orient == 8 ? 5 * (tile >> 1) + (tile & 1) : 5 * tile + half;
}
__device__ __forceinline__ int batt_of_single(int j) { return 5 * (j / 3) + 2 + j % 3; }

template <int NT8> // number of n-tiles (10 for the full bank, 5 for the short one)
__global__ void __launch_bounds__(256, 2) k_lm_conv(LmArgs a)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    float* s_img = (float*)smraw;                       // [TILE_H][TILE_W]
    float* s_w = s_img + TILE_H * TILE_W;               // [2 stages][hi|lo][KWP][NSTRIDE]
    double* s_acc = (double*)(s_w + 2 * 2 * KWP * NSTRIDE); // [LSLOTS][n_batt][2]
    __shared__ int s_lab[LSLOTS];
    __shared__ int s_nlab;
    __shared__ double s_g2[20];
    __shared__ int s_plab[TM_ROWS * TM_X];          // label of every pixel of the tile (-1 outside the image)
    __shared__ signed char s_pslot[TM_ROWS * TM_X]; // its slot in s_lab (-1: more than LSLOTS labels in this tile -> global atomics)

    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * TM_X, y0 = blockIdx.y * TM_ROWS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const size_t npx = (size_t)a.H * a.W;
    const float* img = a.img + (size_t)ch * npx;
    const int NP = a.NP;

    // stage the image tile (reflect at the borders: ndimage.convolve mode='reflect')
    for (int i = threadIdx.x; i < TILE_H * TILE_W; i += 256) {
        const int ty = i / TILE_W, tx = i - ty * TILE_W;
        const int gy = reflect_idx(y0 + ty - KRAD, a.H), gx = reflect_idx(x0 + tx - KRAD, a.W);
        s_img[i] = img[(size_t)gy * a.W + gx];
    }
    {
        const int py = threadIdx.x / TM_X, px = threadIdx.x % TM_X; // 256 threads = 16 x 16 pixels
        const int y = y0 + py, x = x0 + px;
        s_plab[threadIdx.x] = (y < a.H && x < a.W) ? a.seg[(size_t)y * a.W + x] : -1;
    }
    if (threadIdx.x < 20) s_g2[threadIdx.x] = 0.0;
    for (int i = threadIdx.x; i < LSLOTS * a.n_batt * 2; i += 256) s_acc[i] = 0.0;

    __syncthreads();
    if (threadIdx.x == 0) { // distinct labels of the tile (a 16x16 tile touches a handful of superpixels)
        int n = 0;
        for (int i = 0; i < TM_ROWS * TM_X; ++i) {
            const int lb = s_plab[i];
            int slot = -1;
            if (lb >= 0) {
                for (int k = 0; k < n; ++k) if (s_lab[k] == lb) { slot = k; break; }
                if (slot < 0 && n < LSLOTS) { s_lab[n] = lb; slot = n++; }
            }
            s_pslot[i] = (signed char)slot;
        }
        s_nlab = n;
    }
    __syncthreads();

    auto load_stage = [&](int dy, int buf) {
        // [KWP][NP] hi and lo of kernel row dy -> smem rows of stride NSTRIDE, 16-byte cp.async
        const float* src_hi = a.w_hi + (size_t)dy * KWP * NP;
        const float* src_lo = a.w_lo + (size_t)dy * KWP * NP;
        float* dst = s_w + (size_t)buf * 2 * KWP * NSTRIDE;
        const int vec_per_row = NP / 4;
        for (int i = threadIdx.x; i < 2 * KWP * vec_per_row; i += 256) {
            const int half = i / (KWP * vec_per_row), rem = i - half * KWP * vec_per_row;
            const int k = rem / vec_per_row, v = rem - k * vec_per_row;
            const float* s = (half ? src_lo : src_hi) + (size_t)k * NP + 4 * v;
            __pipeline_memcpy_async(dst + (size_t)half * KWP * NSTRIDE + k * NSTRIDE + 4 * v, s, 16);
        }
        __pipeline_commit();
    };

    float acc[2][NT8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NT8; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;

    load_stage(0, 0);
    for (int dy = 0; dy < KW; ++dy) {
        const int buf = dy & 1;
        if (dy + 1 < KW) load_stage(dy + 1, buf ^ 1);
        if (dy + 1 < KW) __pipeline_wait_prior(1); else __pipeline_wait_prior(0);
        __syncthreads();
        const float* wh = s_w + (size_t)buf * 2 * KWP * NSTRIDE;
        const float* wl = wh + KWP * NSTRIDE;
#pragma unroll
        for (int q = 0; q < KWP / 8; ++q) {
            // A fragments of the two M tiles of this warp (image rows warp and warp + 8 of the CTA tile)
            unsigned ah[2][4], al[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const float* row = s_img + (warp + 8 * m + dy) * TILE_W + 8 * q + t;
                const float v0 = row[g], v1 = row[g + 8], v2 = row[g + 4], v3 = row[g + 12];
                // fragment order: (row g, col t), (row g+8, col t), (row g, col t+4), (row g+8, col t+4); pixel rx = g or g+8, tap dx = 8q+t(+4)
                ah[m][0] = f2tf32(v0); al[m][0] = f2tf32(v0 - __uint_as_float(ah[m][0]));
                ah[m][1] = f2tf32(v1); al[m][1] = f2tf32(v1 - __uint_as_float(ah[m][1]));
                ah[m][2] = f2tf32(v2); al[m][2] = f2tf32(v2 - __uint_as_float(ah[m][2]));
                ah[m][3] = f2tf32(v3); al[m][3] = f2tf32(v3 - __uint_as_float(ah[m][3]));
            }
#pragma unroll
            for (int j = 0; j < NT8; ++j) {
                const unsigned bh0 = __float_as_uint(wh[(8 * q + t) * NSTRIDE + 8 * j + g]);
                const unsigned bh1 = __float_as_uint(wh[(8 * q + t + 4) * NSTRIDE + 8 * j + g]);
                const unsigned bl0 = __float_as_uint(wl[(8 * q + t) * NSTRIDE + 8 * j + g]);
                const unsigned bl1 = __float_as_uint(wl[(8 * q + t + 4) * NSTRIDE + 8 * j + g]);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    mma_tf32(acc[m][j], al[m], bh0, bh1);
                    mma_tf32(acc[m][j], ah[m], bl0, bl1);
                    mma_tf32(acc[m][j], ah[m], bh0, bh1);
                }
            }
        }
        __syncthreads();
    }

    // ---- epilogue: battery responses of the 2 x 16 pixels of this warp, statistics ----
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int prow = (warp + 8 * m) * TM_X;
        const int lab0 = s_plab[prow + g], lab1 = s_plab[prow + g + 8];   // pixels (g) and (g + 8) of this M tile
        const int slot0 = s_pslot[prow + g], slot1 = s_pslot[prow + g + 8];
        auto add = [&](int lb, int slot, int batt, float r) {
            if (lb < 0) return;
            if (r > 1.e6f) r = 1.e6f; // MAX_SIGNAL_RESPONSE
            const double rd = (double)r, r2 = rd * rd;
            if (slot >= 0) {
                atomicAdd(&s_acc[(slot * a.n_batt + batt) * 2], rd);
                atomicAdd(&s_acc[(slot * a.n_batt + batt) * 2 + 1], r2);
            } else {
                atomicAdd(&a.S1[(size_t)lb * a.n_batt * 3 + batt * 3 + ch], rd);
                atomicAdd(&a.S2[(size_t)lb * a.n_batt * 3 + batt * 3 + ch], r2);
                atomicAdd(&a.G2[batt], r2);
            }
        };
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
            if (j < a.n_orient_tiles) {
                float m0 = fmaxf(acc[m][j][0], acc[m][j][1]), m1 = fmaxf(acc[m][j][2], acc[m][j][3]);
                m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
                if (a.orient == 8) {
                    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
                    if (t == 0) { const int b = batt_of_oriented(j, 0, 8); add(lab0, slot0, b, m0); add(lab1, slot1, b, m1); }
                } else if ((t & 1) == 0) {
                    const int b = batt_of_oriented(j, t >> 1, 4); add(lab0, slot0, b, m0); add(lab1, slot1, b, m1);
                }
            } else {
                // single-filter batteries: this thread holds columns 2t and 2t+1 of the tile
                const int c0 = 8 * (j - a.n_orient_tiles) + 2 * t;
                const int nsingle = a.n_batt - 2 * (a.n_batt / 5);
                if (c0 < nsingle) { const int b = batt_of_single(c0); add(lab0, slot0, b, acc[m][j][0]); add(lab1, slot1, b, acc[m][j][2]); }
                if (c0 + 1 < nsingle) { const int b = batt_of_single(c0 + 1); add(lab0, slot0, b, acc[m][j][1]); add(lab1, slot1, b, acc[m][j][3]); }
            }
        }
    }
    __syncthreads();
    const int nl = min(s_nlab, LSLOTS);
    for (int i = threadIdx.x; i < nl * a.n_batt; i += 256) {
        const int slot = i / a.n_batt, b = i - slot * a.n_batt;
        const int lb = s_lab[slot];
        const double v1 = s_acc[(slot * a.n_batt + b) * 2], v2 = s_acc[(slot * a.n_batt + b) * 2 + 1];
        if (v2 != 0.0 || v1 != 0.0) {
            atomicAdd(&a.S1[(size_t)lb * a.n_batt * 3 + b * 3 + ch], v1);
            atomicAdd(&a.S2[(size_t)lb * a.n_batt * 3 + b * 3 + ch], v2);
            atomicAdd(&s_g2[b], v2);
        }
    }
    __syncthreads();
    if (threadIdx.x < a.n_batt && s_g2[threadIdx.x] != 0.0) atomicAdd(&a.G2[threadIdx.x], s_g2[threadIdx.x]);
}

__device__ __forceinline__ double tidy(double v)
{
    if (isnan(v)) return 0.0;
    return v == 0.0 ? 0.0 : v;
}

// features = statistics of the scaled responses: a = log(1 + ||r||) / 0.03 / ||r|| per battery
__global__ void k_lm_finalize(int nb, int n_batt, int flags, const double* __restrict__ S1, const double* __restrict__ S2,
                              const double* __restrict__ G2, const int* __restrict__ counts, double* feat, int ld, int col0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb * n_batt) return;
    const int k = i / n_batt, b = i - k * n_batt;
    const double norm = sqrt(G2[b]);
    const double al = (norm == 0.0 || isinf(norm)) ? 0.0 : (log(1.0 + norm) / 0.03) / norm;
    const int nfl = ((flags & 1) ? 1 : 0) + ((flags & 2) ? 1 : 0) + ((flags & 4) ? 1 : 0);
    const double n = (double)counts[k];
    double* row = feat + (size_t)k * ld + col0 + (size_t)b * 3 * nfl;
    for (int c = 0; c < 3; ++c) {
        const double s1 = S1[(size_t)k * n_batt * 3 + b * 3 + c], s2 = S2[(size_t)k * n_batt * 3 + b * 3 + c];
        double mean = 0, en = 0, sd = 0;
        if (n > 0) {
            mean = al * s1 / n;
            en = al * al * s2 / n;
            double var = en - mean * mean;
            sd = var > 0 ? sqrt(var) : 0.0;
        }
        int col = 0;
        if (flags & 1) { row[col * 3 + c] = tidy(mean); ++col; }
        if (flags & 2) { row[col * 3 + c] = tidy(sd); ++col; }
        if (flags & 4) { row[col * 3 + c] = tidy(en); ++col; }
    }
}

__global__ void k_lm_counts(const int* __restrict__ seg, size_t npx, int* counts)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npx) atomicAdd(&counts[seg[p]], 1);
}

struct LmWs { double* p0; double* p1; double* p2; float* imgf; double* S1; double* S2; double* G2; int* counts; };

static size_t carve_lm(LmWs& w, void* ws, size_t bytes, int H, int W, int nb, int n_batt)
{
    WsCarver c(ws, bytes);
    const size_t n = 3 * (size_t)H * W;
    w.p0 = c.take<double>(n); w.p1 = c.take<double>(n); w.p2 = c.take<double>(n);
    w.imgf = c.take<float>(n);
    w.S1 = c.take<double>((size_t)nb * n_batt * 3); w.S2 = c.take<double>((size_t)nb * n_batt * 3);
    w.G2 = c.take<double>(32);
    w.counts = c.take<int>(nb);
    return isb_align(c.off);
}

} // namespace

extern "C" size_t isb_lm_workspace_bytes(int H, int W, int nb, int n_batt)
{
    LmWs w;
    return carve_lm(w, nullptr, 0, H, W, nb, n_batt);
}

extern "C" int isb_lm_texture(const void* img, int dtype, const int32_t* seg, int H, int W, int nb, const double* bg_weights, int bg_radius,
                              const double* chmix_host, const float* w_hi, const float* w_lo, int NP, int orient, int n_batt, int flags,
                              double* feat, int ld, int col0, void* ws, size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(img && seg && w_hi && w_lo && feat && ws && chmix_host, "null pointer");
    ISB_REQUIRE(H > 0 && W > 0 && nb > 0, "bad sizes");
    ISB_REQUIRE(dtype >= ISB_U8 && dtype <= ISB_F64, "bad dtype");
    ISB_REQUIRE((orient == 8 && NP == 80 && n_batt == 20) || (orient == 4 && NP == 40 && n_batt == 15),
                "filter bank layout must be the full (8 orientations, 80 padded filters, 20 batteries) or the short one (4, 40, 15)");
    ISB_REQUIRE(bg_radius >= 0 && (bg_radius == 0 || bg_weights), "background weights missing");
    LmWs w;
    size_t need = carve_lm(w, ws, ws_bytes, H, W, nb, n_batt);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_LM, st);
    const size_t npx = (size_t)H * W;
    k_lm_to_planar<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(img, dtype, npx, w.p0);
    ISB_LAUNCH_CHECK();
    Mix3 mix;
    for (int i = 0; i < 9; ++i) mix.m[i] = chmix_host[i];
    if (bg_radius > 0) {
        const size_t smem = sizeof(double) * (2 * (size_t)(bg_radius + VB_R) + 1);
        ISB_REQUIRE(smem <= 200 * 1024, "background radius too large");
        ISB_CUDA_CHECK(cudaFuncSetAttribute(k_lm_vblur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // axis 0 (rows): p0 [3][H][W] -> p1
        k_lm_vblur<<<dim3((W + VB_T - 1) / VB_T, (H + VB_R - 1) / VB_R, 3), VB_T, smem, st>>>(w.p0, H, W, bg_weights, bg_radius, w.p1);
        ISB_LAUNCH_CHECK();
        // axis 1 (cols): transpose, blur along the (new) rows axis; the result stays transposed [3][W][H]
        k_lm_transpose<<<dim3((W + 31) / 32, (H + 31) / 32, 3), 256, 0, st>>>(w.p1, H, W, w.p2);
        ISB_LAUNCH_CHECK();
        k_lm_vblur<<<dim3((H + VB_T - 1) / VB_T, (W + VB_R - 1) / VB_R, 3), VB_T, smem, st>>>(w.p2, W, H, bg_weights, bg_radius, w.p1);
        ISB_LAUNCH_CHECK();
    } else {
        ISB_CUDA_CHECK(cudaMemsetAsync(w.p1, 0, sizeof(double) * 3 * npx, st));
    }
    // axis 2 (channels) folded into the 3x3 mix; subtract; to f32 planar
    k_lm_mix_sub<<<dim3((W + 31) / 32, (H + 31) / 32), 256, 0, st>>>(w.p0, w.p1, H, W, mix, w.imgf);
    ISB_LAUNCH_CHECK();
    ISB_CUDA_CHECK(cudaMemsetAsync(w.S1, 0, sizeof(double) * (size_t)nb * n_batt * 3, st));
    ISB_CUDA_CHECK(cudaMemsetAsync(w.S2, 0, sizeof(double) * (size_t)nb * n_batt * 3, st));
    ISB_CUDA_CHECK(cudaMemsetAsync(w.G2, 0, sizeof(double) * 32, st));
    ISB_CUDA_CHECK(cudaMemsetAsync(w.counts, 0, sizeof(int) * (size_t)nb, st));
    k_lm_counts<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(seg, npx, w.counts);
    ISB_LAUNCH_CHECK();
    LmArgs a;
    a.img = w.imgf; a.seg = seg; a.H = H; a.W = W; a.nb = nb; a.w_hi = w_hi; a.w_lo = w_lo; a.NP = NP; a.n_tiles = NP / 8;
    a.orient = orient; a.n_orient_tiles = orient == 8 ? 8 : 3; a.n_batt = n_batt; a.S1 = w.S1; a.S2 = w.S2; a.G2 = w.G2;
    const size_t smem = sizeof(float) * (TILE_H * TILE_W + 2 * 2 * KWP * NSTRIDE) + sizeof(double) * LSLOTS * n_batt * 2;
    dim3 grid((W + TM_X - 1) / TM_X, (H + TM_ROWS - 1) / TM_ROWS, 3);
    if (orient == 8) {
        ISB_CUDA_CHECK(cudaFuncSetAttribute(k_lm_conv<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_lm_conv<10><<<grid, 256, smem, st>>>(a);
    } else {
        ISB_CUDA_CHECK(cudaFuncSetAttribute(k_lm_conv<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_lm_conv<5><<<grid, 256, smem, st>>>(a);
    }
    ISB_LAUNCH_CHECK();
    k_lm_finalize<<<(nb * n_batt + 255) / 256, 256, 0, st>>>(nb, n_batt, flags, w.S1, w.S2, w.G2, w.counts, feat, ld, col0);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
