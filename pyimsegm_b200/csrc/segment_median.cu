// segment_median.cu -- per-segment, per-channel MEDIAN of an image, and binary morphology with a disc.
//
// Median: replaces imsegm/descriptors.py:420-455 numpy_img2d_color_median and :651-676 numpy_img3d_gray_median -- pure-Python
// loops in the reference (a list append per pixel and channel, then np.median per label).  Here:
//   1. counting sort of the pixel indices by label (histogram, single-CTA scan, scatter; the order inside a label is irrelevant),
//   2. one CTA per (label, channel): 8-bit MSB-first radix SELECT on the order-preserving 64-bit image of the doubles -- eight
//      passes over the label's pixels find the lower middle value exactly, one more pass finds the upper middle one
//      (np.median averages the two for an even count).
// Morphology: skimage.morphology.opening(mask, disk(r)) as imsegm/descriptors.py:1873-1876 applies it to the boundary mask of the
// Ray features = grey erosion then grey dilation with a disc footprint, borders reflected (scipy.ndimage default mode).
#include "common.cuh"

namespace {

constexpr int MT = 128;   // threads of a select CTA

__global__ void __launch_bounds__(256) k_med_count(const int* __restrict__ seg, size_t n, int nb, int* __restrict__ counts)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int lb = seg[p];
    if (lb >= 0 && lb < nb) atomicAdd(&counts[lb], 1);
}

// exclusive scan counts[0..nb) -> start[0..nb], cursor = start (single CTA; nb is a superpixel count)
__global__ void __launch_bounds__(1024) k_med_scan(const int* __restrict__ counts, int nb, int* __restrict__ start, int* __restrict__ cursor)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_w[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int t = s_w[lane];
            int ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, ti, o); if (lane >= o) ti += u; }
            s_w[lane] = ti - t;
        }
        __syncthreads();
        const int excl = s_carry + s_w[wid] + incl - v;
        if (i < nb) { start[i] = excl; cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[nb] = s_carry;
}

__global__ void __launch_bounds__(256) k_med_scatter(const int* __restrict__ seg, size_t n, int nb, int* __restrict__ cursor, unsigned* __restrict__ order)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int lb = seg[p];
    if (lb >= 0 && lb < nb) order[atomicAdd(&cursor[lb], 1)] = (unsigned)p;
}

// blockIdx.x = label, blockIdx.y = channel
__global__ void __launch_bounds__(MT) k_med_select(const void* __restrict__ img, int dtype, int C, const int* __restrict__ start,
                                                   const unsigned* __restrict__ order, double* __restrict__ out)
{
    __shared__ int s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_k;
    __shared__ unsigned long long s_next;   // smallest key above the selected one
    __shared__ int s_le;                    // how many keys are <= the selected one
    const int lb = blockIdx.x, c = blockIdx.y;
    const int beg = start[lb], n = start[lb + 1] - beg;
    if (n <= 0) { if (threadIdx.x == 0) out[(size_t)lb * C + c] = nan(""); return; }
    const unsigned* ord = order + beg;
    const int k_lo = (n - 1) / 2, k_hi = n / 2;
    if (threadIdx.x == 0) { s_prefix = 0ull; s_k = k_lo; }
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += MT) s_hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int i = threadIdx.x; i < n; i += MT) {
            const unsigned long long key = f64_ordered(load_as_f64(img, dtype, (size_t)ord[i] * C + c));
            if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(int)((key >> shift) & 255ull)], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int k = s_k, d = 0;
            for (; d < 255; ++d) { if (k < s_hist[d]) break; k -= s_hist[d]; }
            s_k = k;
            s_prefix = prefix | ((unsigned long long)d << shift);
        }
        __syncthreads();
    }
    const unsigned long long sel = s_prefix;    // key of the element of rank k_lo
    double hi = f64_unordered(sel);
    if (k_hi != k_lo) {
        if (threadIdx.x == 0) { s_next = ~0ull; s_le = 0; }
        __syncthreads();
        int le = 0;
        unsigned long long nx = ~0ull;
        for (int i = threadIdx.x; i < n; i += MT) {
            const unsigned long long key = f64_ordered(load_as_f64(img, dtype, (size_t)ord[i] * C + c));
            if (key <= sel) ++le; else if (key < nx) nx = key;
        }
        atomicAdd(&s_le, le);
        atomicMin(&s_next, nx);
        __syncthreads();
        if (s_le < k_hi + 1) hi = f64_unordered(s_next);   // the upper middle element is the next larger value
    }
    if (threadIdx.x == 0) out[(size_t)lb * C + c] = 0.5 * (f64_unordered(sel) + hi);
}

struct MedWs { int* counts; int* start; int* cursor; unsigned* order; };
static size_t carve_med(MedWs& w, void* ws, size_t bytes, size_t n, int nb)
{
    WsCarver c(ws, bytes);
    w.counts = c.take<int>(nb); w.start = c.take<int>((size_t)nb + 1); w.cursor = c.take<int>(nb); w.order = c.take<unsigned>(n);
    return isb_align(c.off);
}

__device__ __forceinline__ int reflect_px(int i, int n)
{
    if (n == 1) return 0;
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

// op 0: erosion (all pixels under the disc set), op 1: dilation (any pixel under the disc set); borders reflected
__global__ void __launch_bounds__(256) k_morph_disk(const unsigned char* __restrict__ src, int H, int W, int radius, int op, unsigned char* __restrict__ dst)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int r2 = radius * radius;
    bool acc = op == 0;
    for (int dy = -radius; dy <= radius; ++dy) {
        const int yy = reflect_px(y + dy, H);
        for (int dx = -radius; dx <= radius; ++dx) {
            if (dy * dy + dx * dx > r2) continue;
            const bool v = src[(size_t)yy * W + reflect_px(x + dx, W)] != 0;
            if (op == 0) acc = acc && v; else acc = acc || v;
        }
    }
    dst[(size_t)y * W + x] = acc ? 1 : 0;
}

} // namespace

extern "C" size_t isb_segment_median_workspace_bytes(long long n_px, int nb)
{
    MedWs w;
    return carve_med(w, nullptr, 0, (size_t)n_px, nb);
}

extern "C" int isb_segment_median(const void* img, int dtype, const int32_t* seg, long long n_px, int channels, int nb, double* out, void* ws,
                                  size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(img && seg && out && ws, "null pointer");
    ISB_REQUIRE(n_px > 0 && n_px < (1LL << 32) && channels > 0 && channels <= 65535 && nb > 0, "bad sizes");
    ISB_REQUIRE(dtype >= ISB_U8 && dtype <= ISB_F64, "bad dtype");
    MedWs w;
    const size_t need = carve_med(w, ws, ws_bytes, (size_t)n_px, nb);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_STATS, st);
    const size_t n = (size_t)n_px;
    ISB_CUDA_CHECK(cudaMemsetAsync(w.counts, 0, sizeof(int) * (size_t)nb, st));
    k_med_count<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seg, n, nb, w.counts);
    ISB_LAUNCH_CHECK();
    k_med_scan<<<1, 1024, 0, st>>>(w.counts, nb, w.start, w.cursor);
    ISB_LAUNCH_CHECK();
    k_med_scatter<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seg, n, nb, w.cursor, w.order);
    ISB_LAUNCH_CHECK();
    k_med_select<<<dim3(nb, channels), MT, 0, st>>>(img, dtype, channels, w.start, w.order, out);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" int isb_binary_opening_disk(const uint8_t* mask, int H, int W, int radius, uint8_t* tmp, uint8_t* out, isb_stream_t stream)
{
    ISB_REQUIRE(mask && tmp && out, "null pointer");
    ISB_REQUIRE(H > 0 && W > 0 && radius >= 0 && radius <= 512, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((W + 31) / 32, (H + 7) / 8);
    k_morph_disk<<<grid, 256, 0, st>>>(mask, H, W, radius, 0, tmp);
    ISB_LAUNCH_CHECK();
    k_morph_disk<<<grid, 256, 0, st>>>(tmp, H, W, radius, 1, out);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
