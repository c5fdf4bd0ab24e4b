// slic_kmeans.cu -- the k-means sweeps of SLIC, pixel-centric and bit-exact with oracle/slic_oracle.c.
//
// Replaces skimage.segmentation._slic._slic_cython as called through imsegm/superpixels.py:61-63.
//
// The original is cluster-centric: for k ascending, scan the +-2*step window and take the pixel when
// `distance > d` (strict) -- i.e. every pixel ends with argmin over {clusters whose window holds it} of
// (d, k) in lexicographic order.  That per-pixel minimum is order-independent, so it is evaluated here
// pixel-centric: one CTA per 32x32 pixel tile gathers the clusters whose integer window intersects the tile
// (from a uniform bin grid over the current centroids), every pixel loops over that list in shared memory.
// The distance is computed with the oracle's exact operation order in IEEE double (no FMA).
//
// The centroid update of the original is a raster-order SEQUENTIAL double sum per cluster; a tree/atomic
// reduction changes the last ulp and flips exact ties (flat image regions tie all the time).  It is
// reproduced exactly: one warp per cluster walks the cluster's window in raster order, ballots the member
// pixels of each 32-wide chunk, compacts their colours into shared memory in lane (= raster) order, and three
// lanes add them one at a time.  Coordinate sums are integers (exact in any order).
//
// Per sweep HBM traffic (algorithmic): read Lab 24 B/px + write label 4 B/px (assign); the update re-reads
// labels and member colours through L2.
#include "common.cuh"
#include <float.h>

namespace {

constexpr int TILE = 32;   // pixel tile edge of the assignment kernel
constexpr int ACAP = 96;   // candidate clusters staged per round
constexpr int AROWS = 4;   // pixels per thread (same column, rows ty, ty+8, ty+16, ty+24)

struct __align__(16) Cand {
    double cy, cx, c0, c1, c2;
    int y0, y1, x0, x1;
    int k, pad;
};
static_assert(sizeof(Cand) == 64, "Cand must be 64 bytes");

struct KmState {
    // cluster state (SoA)
    double* cy; double* cx; double* c0; double* c1; double* c2;
    int4* win;       // [n] (y0, y1, x0, x1); empty (0,0,0,0) when dead
    int4* obb;       // [n] bbox of orphan pixels (ymin, ymax, xmin, xmax), empty = (INT_MAX, -1, INT_MAX, -1)
    double* sums;    // [n][3] colour sums of the current update
    long long* isum; // [n][3] count, sum y, sum x
    int* bin_start;  // [nbins + 1]
    int* bin_fill;   // [nbins]
    int* bin_items;  // [n]
    int* bin_of;     // [n]
    int n, H, W, step_y, step_x, B, nby, nbx;
    double sw;       // spatial weight 1/step^2
};

__device__ __forceinline__ int4 make_window(double cy, double cx, int step_y, int step_x, int H, int W)
{
    // <Py_ssize_t>max(c - 2*step, 0) / <Py_ssize_t>min(c + 2*step + 1, size): truncation of a non-negative double
    double lo, hi;
    int4 w;
    lo = __dsub_rn(cy, (double)(2 * step_y)); if (0.0 > lo) lo = 0.0;
    hi = __dadd_rn(__dadd_rn(cy, (double)(2 * step_y)), 1.0); if ((double)H < hi) hi = (double)H;
    w.x = (int)lo; w.y = (int)hi;
    lo = __dsub_rn(cx, (double)(2 * step_x)); if (0.0 > lo) lo = 0.0;
    hi = __dadd_rn(__dadd_rn(cx, (double)(2 * step_x)), 1.0); if ((double)W < hi) hi = (double)W;
    w.z = (int)lo; w.w = (int)hi;
    return w;
}

// single CTA: (re)compute centroids + windows, then bin the live clusters by centroid position.
// first == 1: take centres from the seed grid (colour part 0);  else divide the update sums by the counts.
__global__ void __launch_bounds__(1024) k_finalize_bin(KmState s, const double* seeds_yx, int first)
{
    const int nbins = s.nby * s.nbx;
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) s.bin_fill[b] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < s.n; k += blockDim.x) {
        int4 w = make_int4(0, 0, 0, 0);
        int bin = -1;
        bool dead = false;
        double cy, cx;
        if (first) {
            cy = seeds_yx[2 * k]; cx = seeds_yx[2 * k + 1];
            s.cy[k] = cy; s.cx[k] = cx; s.c0[k] = 0.0; s.c1[k] = 0.0; s.c2[k] = 0.0;
        } else {
            int4 pw = s.win[k];
            long long cnt = s.isum[3 * k];
            if (pw.y <= pw.x && pw.w <= pw.z && cnt == 0) dead = true; // was already dead
            else if (cnt == 0) dead = true;                              // lost every pixel: dead for good
            else {
                double dn = (double)cnt;
                cy = __ddiv_rn((double)s.isum[3 * k + 1], dn);
                cx = __ddiv_rn((double)s.isum[3 * k + 2], dn);
                s.cy[k] = cy; s.cx[k] = cx;
                s.c0[k] = __ddiv_rn(s.sums[3 * k], dn);
                s.c1[k] = __ddiv_rn(s.sums[3 * k + 1], dn);
                s.c2[k] = __ddiv_rn(s.sums[3 * k + 2], dn);
            }
        }
        if (!dead) {
            w = make_window(cy, cx, s.step_y, s.step_x, s.H, s.W);
            int by = min(max((int)cy / s.B, 0), s.nby - 1), bx = min(max((int)cx / s.B, 0), s.nbx - 1);
            bin = by * s.nbx + bx;
            atomicAdd(&s.bin_fill[bin], 1);
        }
        s.win[k] = w;
        s.bin_of[k] = bin;
        s.obb[k] = make_int4(INT_MAX, -1, INT_MAX, -1);
    }
    __syncthreads();
    // exclusive scan of bin counts (block-wide, chunked)
    __shared__ int s_part[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nbins; base += blockDim.x) {
        int b = base + threadIdx.x;
        int v = b < nbins ? s.bin_fill[b] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < (int)blockDim.x; o <<= 1) {
            int t = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = s_part[threadIdx.x];
        int carry = s_carry;
        if (b < nbins) { s.bin_start[b] = carry + incl - v; s.bin_fill[b] = 0; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) s.bin_start[nbins] = s_carry;
    __syncthreads();
    for (int k = threadIdx.x; k < s.n; k += blockDim.x) {
        int bin = s.bin_of[k];
        if (bin >= 0) s.bin_items[s.bin_start[bin] + atomicAdd(&s.bin_fill[bin], 1)] = k;
    }
}

// assignment: one CTA per 32x32 tile, 256 threads, 4 pixels per thread
__global__ void __launch_bounds__(256) k_assign(KmState s, const double* __restrict__ lab, int* __restrict__ labels)
{
    __shared__ Cand cand[ACAP];
    __shared__ int s_ncand, s_done, s_row, s_off;
    const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE;
    const int tx1 = min(tx0 + TILE, s.W), ty1 = min(ty0 + TILE, s.H);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t HW = (size_t)s.H * s.W;
    const int x = tx0 + lane;
    const bool xin = x < s.W;

    double p0[AROWS], p1[AROWS], p2[AROWS], best[AROWS];
    int bestk[AROWS];
#pragma unroll
    for (int j = 0; j < AROWS; ++j) {
        int y = ty0 + warp + 8 * j;
        best[j] = DBL_MAX; bestk[j] = -1;
        if (xin && y < s.H) {
            size_t p = (size_t)y * s.W + x;
            p0[j] = lab[p]; p1[j] = lab[HW + p]; p2[j] = lab[2 * HW + p];
        } else { p0[j] = p1[j] = p2[j] = 0.0; }
    }
    // bins that can hold a centroid whose window reaches this tile (superset; the exact window test follows)
    const int by0 = max(ty0 - 2 * s.step_y - 2, 0) / s.B, by1 = min(ty1 + 2 * s.step_y + 1, s.H - 1) / s.B;
    const int bx0 = max(tx0 - 2 * s.step_x - 2, 0) / s.B, bx1 = min(tx1 + 2 * s.step_x + 1, s.W - 1) / s.B;
    if (threadIdx.x == 0) { s_row = by0; s_off = 0; s_done = 0; }
    __syncthreads();

    while (true) {
        if (warp == 0) {
            // deterministic, resumable scan of the bin rows: fill up to ACAP candidates
            int n = 0;
            int row = s_row, off = s_off;
            while (row <= min(by1, s.nby - 1) && n < ACAP) {
                int beg = s.bin_start[row * s.nbx + bx0] + off;
                int end = s.bin_start[row * s.nbx + min(bx1, s.nbx - 1) + 1];
                while (beg < end && n < ACAP) {
                    int room = ACAP - n;
                    int i = beg + lane;
                    bool ok = false;
                    int k = -1;
                    int4 w;
                    if (i < end && lane < room) {
                        k = s.bin_items[i];
                        w = s.win[k];
                        ok = (w.x < ty1) && (w.y > ty0) && (w.z < tx1) && (w.w > tx0);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        int pos = n + __popc(m & ((1u << lane) - 1u));
                        Cand c;
                        c.cy = s.cy[k]; c.cx = s.cx[k]; c.c0 = s.c0[k]; c.c1 = s.c1[k]; c.c2 = s.c2[k];
                        c.y0 = w.x; c.y1 = w.y; c.x0 = w.z; c.x1 = w.w; c.k = k; c.pad = 0;
                        cand[pos] = c;
                    }
                    n += __popc(m);
                    int adv = min(min(32, room), end - beg);
                    beg += adv; off += adv;
                }
                if (beg >= end) { ++row; off = 0; }
            }
            if (lane == 0) { s_ncand = n; s_row = row; s_off = off; s_done = row > min(by1, s.nby - 1); }
        }
        __syncthreads();
        const int nc = s_ncand;
        const int done = s_done;
        if (xin) {
            const double xd = (double)x;
            for (int c = 0; c < nc; ++c) {
                const int cx0 = cand[c].x0, cx1 = cand[c].x1;
                if (x < cx0 || x >= cx1) continue;
                const int cy0 = cand[c].y0, cy1 = cand[c].y1, ck = cand[c].k;
                const double ccy = cand[c].cy, cc0 = cand[c].c0, cc1 = cand[c].c1, cc2 = cand[c].c2;
                const double tx = __dsub_rn(cand[c].cx, xd);
                const double dx2 = __dmul_rn(tx, tx);
#pragma unroll
                for (int j = 0; j < AROWS; ++j) {
                    const int y = ty0 + warp + 8 * j;
                    if (y < cy0 || y >= cy1) continue;
                    const double ty = __dsub_rn(ccy, (double)y);
                    double dc = __dmul_rn(__dadd_rn(__dmul_rn(ty, ty), dx2), s.sw);
                    const double d0 = __dsub_rn(p0[j], cc0), d1 = __dsub_rn(p1[j], cc1), d2 = __dsub_rn(p2[j], cc2);
                    double dcol = __dmul_rn(d0, d0);
                    dcol = __dadd_rn(dcol, __dmul_rn(d1, d1));
                    dcol = __dadd_rn(dcol, __dmul_rn(d2, d2));
                    dc = __dadd_rn(dc, dcol);
                    if (dc < best[j] || (dc == best[j] && bestk[j] >= 0 && ck < bestk[j])) { best[j] = dc; bestk[j] = ck; }
                }
            }
        }
        if (done) break;
        __syncthreads();
    }
    if (xin) {
#pragma unroll
        for (int j = 0; j < AROWS; ++j) {
            int y = ty0 + warp + 8 * j;
            if (y >= s.H) continue;
            size_t p = (size_t)y * s.W + x;
            if (bestk[j] >= 0) labels[p] = bestk[j];
            else {
                // no window holds this pixel: it keeps its label (the original leaves nearest_segments untouched);
                // tell that cluster's update where to look
                int k = labels[p];
                atomicMin(&s.obb[k].x, y); atomicMax(&s.obb[k].y, y);
                atomicMin(&s.obb[k].z, x); atomicMax(&s.obb[k].w, x);
            }
        }
    }
}

// centroid sums: one warp per cluster, raster-order sequential double adds (see header)
__global__ void __launch_bounds__(256) k_update(KmState s, const double* __restrict__ lab, const int* __restrict__ labels)
{
    __shared__ double buf[8][3][32];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int k = blockIdx.x * 8 + wl;
    if (k >= s.n) return;
    int4 w = s.win[k];
    int4 o = s.obb[k];
    int y0 = w.x, y1 = w.y, x0 = w.z, x1 = w.w;
    if (o.y >= o.x) { // orphans: grow the scan box
        if (y1 <= y0 || x1 <= x0) { y0 = o.x; y1 = o.y + 1; x0 = o.z; x1 = o.w + 1; }
        else { y0 = min(y0, o.x); y1 = max(y1, o.y + 1); x0 = min(x0, o.z); x1 = max(x1, o.w + 1); }
    }
    const size_t HW = (size_t)s.H * s.W;
    double acc = 0.0;
    long long cnt = 0, sy = 0, sx = 0;
    for (int y = y0; y < y1; ++y) {
        const size_t rowp = (size_t)y * s.W;
        for (int xb = x0; xb < x1; xb += 32) {
            const int x = xb + lane;
            const bool m = (x < x1) && (labels[rowp + x] == k);
            const unsigned mask = __ballot_sync(0xffffffffu, m);
            if (!mask) continue;
            const int nm = __popc(mask);
            if (m) {
                int pos = __popc(mask & ((1u << lane) - 1u));
                buf[wl][0][pos] = lab[rowp + x];
                buf[wl][1][pos] = lab[HW + rowp + x];
                buf[wl][2][pos] = lab[2 * HW + rowp + x];
                sx += x;
            }
            cnt += nm;
            sy += (long long)y * nm;
            __syncwarp();
            if (lane < 3)
                for (int i = 0; i < nm; ++i) acc = __dadd_rn(acc, buf[wl][lane][i]);
            __syncwarp();
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, d);
    if (lane < 3) s.sums[3 * k + lane] = acc;
    if (lane == 0) { s.isum[3 * k] = cnt; s.isum[3 * k + 1] = sy; s.isum[3 * k + 2] = sx; }
}

__global__ void k_export_centroids(KmState s, double* out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s.n) return;
    out[5 * k] = s.cy[k]; out[5 * k + 1] = s.cx[k];
    out[5 * k + 2] = s.c0[k]; out[5 * k + 3] = s.c1[k]; out[5 * k + 4] = s.c2[k];
}

static size_t carve(KmState& s, void* ws, size_t bytes, int H, int W, int n, int step_y, int step_x)
{
    WsCarver c(ws, bytes);
    s.n = n; s.H = H; s.W = W; s.step_y = step_y; s.step_x = step_x;
    s.B = step_y > step_x ? step_y : step_x;
    if (s.B < 8) s.B = 8;
    s.nby = (H + s.B - 1) / s.B; s.nbx = (W + s.B - 1) / s.B;
    s.cy = c.take<double>(n); s.cx = c.take<double>(n);
    s.c0 = c.take<double>(n); s.c1 = c.take<double>(n); s.c2 = c.take<double>(n);
    s.win = c.take<int4>(n); s.obb = c.take<int4>(n);
    s.sums = c.take<double>(3 * (size_t)n); s.isum = c.take<long long>(3 * (size_t)n);
    s.bin_start = c.take<int>((size_t)s.nby * s.nbx + 1);
    s.bin_fill = c.take<int>((size_t)s.nby * s.nbx);
    s.bin_items = c.take<int>(n); s.bin_of = c.take<int>(n);
    return isb_align(c.off);
}

} // namespace

extern "C" size_t isb_slic_kmeans_workspace_bytes(int H, int W, int n_seeds, int step_y, int step_x)
{
    KmState s;
    return carve(s, nullptr, 0, H, W, n_seeds, step_y, step_x);
}

extern "C" int isb_slic_kmeans(const double* lab_planar, int H, int W, const double* seeds_yx, int n_seeds, int step_y,
                               int step_x, double step, int max_iter, int slic_zero, int32_t* labels, double* centroids,
                               void* ws, size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(lab_planar && seeds_yx && labels && ws, "null pointer");
    ISB_REQUIRE(H > 0 && W > 0 && n_seeds > 0 && step_y > 0 && step_x > 0 && step > 0, "bad sizes");
    if (slic_zero) { isb_set_error("slic_zero (SLICO) is not implemented on the device path"); return ISB_ERR_UNSUPPORTED; }
    KmState s;
    size_t need = carve(s, ws, ws_bytes, H, W, n_seeds, step_y, step_x);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    s.sw = 1.0 / (step * step);
    cudaStream_t st = (cudaStream_t)stream;
    ISB_CUDA_CHECK(cudaMemsetAsync(labels, 0, sizeof(int32_t) * (size_t)H * W, st));
    k_finalize_bin<<<1, 1024, 0, st>>>(s, seeds_yx, 1);
    ISB_LAUNCH_CHECK();
    dim3 agrid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE);
    for (int it = 0; it < max_iter; ++it) {
        { ProfScope p(ISB_PROF_ASSIGN, st); k_assign<<<agrid, 256, 0, st>>>(s, lab_planar, labels); }
        ISB_LAUNCH_CHECK();
        { ProfScope p(ISB_PROF_UPDATE, st); k_update<<<(n_seeds + 7) / 8, 256, 0, st>>>(s, lab_planar, labels); }
        ISB_LAUNCH_CHECK();
        { ProfScope p(ISB_PROF_FINALIZE, st); k_finalize_bin<<<1, 1024, 0, st>>>(s, seeds_yx, 0); }
        ISB_LAUNCH_CHECK();
    }
    if (centroids) {
        k_export_centroids<<<(n_seeds + 255) / 256, 256, 0, st>>>(s, centroids);
        ISB_LAUNCH_CHECK();
    }
    return ISB_OK;
}
