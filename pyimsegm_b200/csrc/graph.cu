// graph.cu -- superpixel adjacency graph, GraphCut energies and the final LUT gathers.
//
// Replaces (reference Python, no native code):
//   imsegm/superpixels.py:115-177  make_graph_segm_connect_grid2d_conn4 / get_segment_diffs_2d_conn4 /
//                                  make_graph_segment_connect_edges   (per-pixel dict loop + np.unique)
//   imsegm/graph_cuts.py:303-336   compute_spatial_dist
//   imsegm/graph_cuts.py:383-439   compute_edge_model
//   imsegm/graph_cuts.py:523-540   compute_unary_cost
//   imsegm/graph_cuts.py:574-657   compute_edge_weights (clamp to [1e-3, 1e3])
//   pyGCO cut_general_graph        float -> int conversion (see oracle/gc_oracle.cpp header)
//   imsegm/pipelines.py:104,109    proba[slic], graph_labels[slic]
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

// ------------------------------------------------------------------ adjacency ------------------------------------------------------

struct AdjWs {
    unsigned long long* table; // [slots] keys (b << 32 | a), empty = ~0
    int* deg;                  // [nb]    number of edges whose larger endpoint is b
    int* off;                  // [nb+1]
    int* fill;                 // [nb]
    int* tmp_a;                // [cap]
    int* ctr;                  // [4] 0: unique edges, 1: overflow flag
    int slots;
};

__device__ __forceinline__ unsigned hash64(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

__device__ void edge_insert(const AdjWs& w, int l0, int l1)
{
    int a = min(l0, l1), b = max(l0, l1);
    unsigned long long key = ((unsigned long long)(unsigned)b << 32) | (unsigned)a;
    unsigned mask = (unsigned)w.slots - 1u;
    unsigned h = hash64(key) & mask;
    for (int probe = 0; probe < w.slots; ++probe) {
        unsigned long long cur = w.table[h];
        if (cur == key) return;
        if (cur == ~0ull) {
            unsigned long long old = atomicCAS(&w.table[h], ~0ull, key);
            if (old == ~0ull) { atomicAdd(&w.deg[b], 1); atomicAdd(&w.ctr[0], 1); return; }
            if (old == key) return;
        }
        h = (h + 1) & mask;
    }
    atomicExch(&w.ctr[1], 1); // table full
}

__global__ void __launch_bounds__(256) k_edge_scan(const int* __restrict__ seg, int H, int W, AdjWs w)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)H * W) return;
    int y = (int)(p / W), x = (int)(p % W);
    int l = seg[p];
    if (x + 1 < W) {
        int r = seg[p + 1];
        // skip when the pixel above saw the same pair
        if (r != l && !(y > 0 && seg[p - W] == l && seg[p - W + 1] == r)) edge_insert(w, l, r);
    }
    if (y + 1 < H) {
        int d = seg[p + W];
        if (d != l && !(x > 0 && seg[p - 1] == l && seg[p + W - 1] == d)) edge_insert(w, l, d);
    }
}

// the same for a volume: 6-connectivity = the pairs with the x+1, y+1 and z+1 neighbour (reference superpixels.py:145-154)
__global__ void __launch_bounds__(256) k_edge_scan3d(const int* __restrict__ seg, int D, int H, int W, AdjWs w)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)D * H * W) return;
    const int x = (int)(p % W), y = (int)((p / W) % H), z = (int)(p / ((size_t)H * W));
    const int l = seg[p];
    if (x + 1 < W) { const int r = seg[p + 1]; if (r != l) edge_insert(w, l, r); }
    if (y + 1 < H) { const int d = seg[p + W]; if (d != l) edge_insert(w, l, d); }
    if (z + 1 < D) { const int b = seg[p + (size_t)H * W]; if (b != l) edge_insert(w, l, b); }
}

// centroids (z, y, x) of the labels of a volume, (-1, -1, -1) for absent labels (superpixels.py:205-242 for 3-D input)
__global__ void k_centroid3d_acc(const int* __restrict__ seg, int D, int H, int W, unsigned long long* acc)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)D * H * W) return;
    const int x = (int)(p % W), y = (int)((p / W) % H), z = (int)(p / ((size_t)H * W));
    unsigned long long* a = acc + 4 * (size_t)seg[p];
    atomicAdd(a, 1ull); atomicAdd(a + 1, (unsigned long long)z); atomicAdd(a + 2, (unsigned long long)y); atomicAdd(a + 3, (unsigned long long)x);
}

__global__ void k_centroid3d_fin(int nb, const unsigned long long* acc, double* centres)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nb) return;
    const double c = (double)acc[4 * (size_t)k];
    for (int d = 0; d < 3; ++d) centres[3 * (size_t)k + d] = c > 0 ? (double)acc[4 * (size_t)k + 1 + d] / c : -1.0;
}

// exclusive scan of deg -> off (single CTA)
__global__ void __launch_bounds__(1024) k_edge_offsets(int nb, AdjWs w, int cap, int* n_edges_out)
{
    __shared__ int s_part[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? w.deg[i] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = s_part[threadIdx.x], carry = s_carry;
        if (i < nb) { w.off[i] = carry + incl - v; w.fill[i] = 0; }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        w.off[nb] = s_carry;
        *n_edges_out = (w.ctr[1] || s_carry > cap) ? cap + 1 : s_carry;
    }
}

__global__ void k_edge_fill(AdjWs w, int cap)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.slots) return;
    unsigned long long key = w.table[i];
    if (key == ~0ull) return;
    int b = (int)(key >> 32), a = (int)(key & 0xffffffffu);
    int pos = w.off[b] + atomicAdd(&w.fill[b], 1);
    if (pos < cap) w.tmp_a[pos] = a;
}

// per larger endpoint b: sort the smaller endpoints and emit (a, b) rows -> edges sorted by (b, a)
__global__ void k_edge_emit(int nb, AdjWs w, int cap, int* __restrict__ edges)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int beg = w.off[b], end = w.off[b + 1];
    if (end > cap) return;
    for (int i = beg + 1; i < end; ++i) { // insertion sort (degree is small)
        int v = w.tmp_a[i], j = i - 1;
        while (j >= beg && w.tmp_a[j] > v) { w.tmp_a[j + 1] = w.tmp_a[j]; --j; }
        w.tmp_a[j + 1] = v;
    }
    for (int i = beg; i < end; ++i) { edges[2 * (size_t)i] = w.tmp_a[i]; edges[2 * (size_t)i + 1] = b; }
}

static int pow2_at_least(long long v) { int p = 1024; while (p < v) p <<= 1; return p; }

static size_t carve_adj(AdjWs& w, void* ws, size_t bytes, int nb, int cap)
{
    WsCarver c(ws, bytes);
    w.slots = pow2_at_least(2LL * cap);
    w.table = c.take<unsigned long long>((size_t)w.slots);
    w.deg = c.take<int>(nb);
    w.off = c.take<int>((size_t)nb + 1);
    w.fill = c.take<int>(nb);
    w.tmp_a = c.take<int>(cap);
    w.ctr = c.take<int>(4);
    return isb_align(c.off);
}

// ------------------------------------------------------------------ energies -------------------------------------------------------

__device__ double block_sum(double v, double* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
    return t;
}

__device__ double block_max(double v, double* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = s_red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, s_red[i]);
    return t;
}

constexpr int ECL = 8;   // CTAs of the energy kernel's thread-block cluster

// the four global quantities of the energy construction (largest unary, mean / deviation of the edge distances, largest weight) are
// reduced over the cluster through distributed shared memory: partials added / compared in rank order, the same value in every CTA
__device__ double cluster_reduce(double v, bool is_max, double* s_red, double* s_x)
{
    const double t = is_max ? block_max(v, s_red) : block_sum(v, s_red);
    cg::cluster_group cl = cg::this_cluster();
    if (threadIdx.x == 0) *s_x = t;
    cl.sync();
    double tot = is_max ? -1.0 : 0.0;
    for (int r = 0; r < ECL; ++r) { const double pr = *cl.map_shared_rank(s_x, r); tot = is_max ? fmax(tot, pr) : tot + pr; }
    cl.sync();
    return tot;
}

// one cluster of ECL CTAs per graph.  vfeat [N, D]: the per-vertex vectors the edge metric compares (proba for 'model').
__global__ void __cluster_dims__(ECL, 1, 1) __launch_bounds__(1024) k_gc_energies(const double* __restrict__ proba, int N_in, const int* n_nodes_dev, int K, const int* __restrict__ edges, int E_in,
                                                      const int* n_edges_dev, const double* __restrict__ centres,
                                                      const double* __restrict__ vfeat, int D, int metric, int spatial,
                                                      double edge_cost, const double* __restrict__ pairwise, double* unary,
                                                      double* edge_w, int* unary_i, int* edge_wi, int* smooth_i, double* sp)
{
    __shared__ double s_red[32];
    __shared__ double s_x;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;   // the cluster is the whole grid
    // an overflowed edge table (count > capacity) holds unspecified rows: no edge is read, the host redoes the image
    const int E = n_edges_dev ? (*n_edges_dev > E_in ? 0 : *n_edges_dev) : E_in;
    const int N = n_nodes_dev ? min(*n_nodes_dev, N_in) : N_in;
    // unary = |-log(clip(p, 0.01, 0.99))|
    double umax = 0.0;
    for (int i = tid; i < N * K; i += nth) {
        double p = proba[i];
        if (p < 0.01) p = 0.01;
        if (p > 1.0 - 0.01) p = 1.0 - 0.01;
        double u = fabs(-log(p));
        unary[i] = u;
        umax = fmax(umax, fabs(u));
    }
    umax = cluster_reduce(umax, true, s_red, &s_x);
    // edge distances
    double dsum = 0.0, ssum = 0.0;
    for (int e = tid; e < E; e += nth) {
        int a = edges[2 * e], b = edges[2 * e + 1];
        double dist = 0.0;
        if (metric != 0) {
            const double* va = vfeat + (size_t)a * D;
            const double* vb = vfeat + (size_t)b * D;
            for (int k = 0; k < D; ++k) {
                double df = va[k] - vb[k];
                if (metric == 1) dist = fmax(dist, df * df);        // lT: max_k (dp)^2
                else if (metric == 2) dist += fabs(df);               // l1
                else dist += df * df;                                 // l2 (sqrt below)
            }
            if (metric == 3) dist = sqrt(dist);
        }
        edge_w[e] = dist;
        dsum += dist;
        if (spatial) {
            double cy = centres[2 * a] - centres[2 * b], cx = centres[2 * a + 1] - centres[2 * b + 1];
            double s = sqrt(cy * cy + cx * cx);
            sp[e] = s;
            ssum += s;
        }
    }
    dsum = cluster_reduce(dsum, false, s_red, &s_x);
    ssum = cluster_reduce(ssum, false, s_red, &s_x);
    const double dmean = E > 0 ? dsum / E : 0.0, smean = E > 0 ? ssum / E : 1.0;
    double vsum = 0.0;
    if (metric != 0)
        for (int e = tid; e < E; e += nth) { double t = edge_w[e] - dmean; vsum += t * t; }
    vsum = cluster_reduce(vsum, false, s_red, &s_x);
    const double sd = sqrt(E > 0 ? vsum / E : 0.0);
    const double denom = 2.0 * (sd * sd);
    double wmax = 0.0;
    for (int e = tid; e < E; e += nth) {
        double wv = metric != 0 ? exp(-edge_w[e] / denom) : 1.0;
        if (spatial) wv = wv / (sp[e] / smean);
        if (wv < 1e-3) wv = 1e-3;
        if (wv > 1e3) wv = 1e3;
        wv *= edge_cost;
        edge_w[e] = wv;
        wmax = fmax(wmax, fabs(wv));
    }
    wmax = cluster_reduce(wmax, true, s_red, &s_x);
    double pmax = pairwise[0];
    for (int i = 1; i < K * K; ++i) pmax = fmax(pmax, pairwise[i]);
    // pyGCO: down_weight_factor = max(|unary|.max(), |w|.max() * pairwise.max()) + 1e-10
    const double dwf = fmax(umax, wmax * pmax) + 1e-10;
    for (int i = tid; i < N * K; i += nth) unary_i[i] = (int)((unary[i] / dwf) * 100000.0);
    for (int e = tid; e < E; e += nth) edge_wi[e] = (int)((edge_w[e] / dwf) * 1000.0);
    for (int i = tid; i < K * K; i += nth) smooth_i[i] = (int)(pairwise[i] * 100.0);
}

// ------------------------------------------------------------------ gathers --------------------------------------------------------

__global__ void __launch_bounds__(256) k_gather(const int* __restrict__ seg, long long n, const int* __restrict__ lut_i,
                                                const double* __restrict__ lut_p, int K, int* __restrict__ out_i, double* __restrict__ out_p)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int l = seg[p];
    if (out_i) out_i[p] = lut_i[l];
    if (out_p) {
        const double* src = lut_p + (size_t)l * K;
        double* dst = out_p + (size_t)p * K;
        for (int k = 0; k < K; ++k) dst[k] = src[k];
    }
}

} // namespace

extern "C" size_t isb_adjacency_workspace_bytes(int nb, int cap)
{
    AdjWs w;
    return carve_adj(w, nullptr, 0, nb, cap);
}

extern "C" int isb_adjacency_edges(const int32_t* seg, int H, int W, int nb, int32_t* edges, int cap, int32_t* n_edges_out, void* ws,
                                   size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(seg && edges && n_edges_out && ws, "null pointer");
    ISB_REQUIRE(H > 0 && W > 0 && nb > 0 && cap > 0, "bad sizes");
    AdjWs w;
    size_t need = carve_adj(w, ws, ws_bytes, nb, cap);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_ADJ, st);
    ISB_CUDA_CHECK(cudaMemsetAsync(w.table, 0xFF, sizeof(unsigned long long) * (size_t)w.slots, st));
    ISB_CUDA_CHECK(cudaMemsetAsync(w.deg, 0, sizeof(int) * (size_t)nb, st));
    ISB_CUDA_CHECK(cudaMemsetAsync(w.ctr, 0, sizeof(int) * 4, st));
    size_t n = (size_t)H * W;
    k_edge_scan<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seg, H, W, w);
    ISB_LAUNCH_CHECK();
    k_edge_offsets<<<1, 1024, 0, st>>>(nb, w, cap, n_edges_out);
    ISB_LAUNCH_CHECK();
    k_edge_fill<<<(w.slots + 255) / 256, 256, 0, st>>>(w, cap);
    ISB_LAUNCH_CHECK();
    k_edge_emit<<<(nb + 127) / 128, 128, 0, st>>>(nb, w, cap, edges);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" int isb_adjacency_edges_3d(const int32_t* seg, int D, int H, int W, int nb, int32_t* edges, int cap, int32_t* n_edges_out,
                                      void* ws, size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(seg && edges && n_edges_out && ws, "null pointer");
    ISB_REQUIRE(D > 0 && H > 0 && W > 0 && nb > 0 && cap > 0, "bad sizes");
    AdjWs w;
    size_t need = carve_adj(w, ws, ws_bytes, nb, cap);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_ADJ, st);
    ISB_CUDA_CHECK(cudaMemsetAsync(w.table, 0xFF, sizeof(unsigned long long) * (size_t)w.slots, st));
    ISB_CUDA_CHECK(cudaMemsetAsync(w.deg, 0, sizeof(int) * (size_t)nb, st));
    ISB_CUDA_CHECK(cudaMemsetAsync(w.ctr, 0, sizeof(int) * 4, st));
    size_t n = (size_t)D * H * W;
    k_edge_scan3d<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seg, D, H, W, w);
    ISB_LAUNCH_CHECK();
    k_edge_offsets<<<1, 1024, 0, st>>>(nb, w, cap, n_edges_out);
    ISB_LAUNCH_CHECK();
    k_edge_fill<<<(w.slots + 255) / 256, 256, 0, st>>>(w, cap);
    ISB_LAUNCH_CHECK();
    k_edge_emit<<<(nb + 127) / 128, 128, 0, st>>>(nb, w, cap, edges);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" int isb_centroids_3d(const int32_t* seg, int D, int H, int W, int nb, double* centres, void* ws, size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(seg && centres && ws, "null pointer");
    ISB_REQUIRE(D > 0 && H > 0 && W > 0 && nb > 0, "bad sizes");
    ISB_REQUIRE(ws_bytes >= sizeof(unsigned long long) * 4 * (size_t)nb, "workspace too small (4 * nb uint64)");
    cudaStream_t st = (cudaStream_t)stream;
    ISB_CUDA_CHECK(cudaMemsetAsync(ws, 0, sizeof(unsigned long long) * 4 * (size_t)nb, st));
    size_t n = (size_t)D * H * W;
    k_centroid3d_acc<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seg, D, H, W, (unsigned long long*)ws);
    ISB_LAUNCH_CHECK();
    k_centroid3d_fin<<<(nb + 255) / 256, 256, 0, st>>>(nb, (const unsigned long long*)ws, centres);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" size_t isb_gc_energies_workspace_bytes(int N, int K, int E) { return isb_align(sizeof(double) * (size_t)(E > 0 ? E : 1)); }

extern "C" int isb_gc_energies(const double* proba, int N, const int32_t* n_nodes_dev, int K, const int32_t* edges, int E, const int32_t* n_edges_dev,
                               const double* centres, int metric, int spatial, double edge_cost, const double* pairwise, double* unary,
                               double* edge_w, int32_t* unary_i, int32_t* edge_wi, int32_t* smooth_i, void* ws, size_t ws_bytes,
                               isb_stream_t stream)
{
    ISB_REQUIRE(proba && edges && pairwise && unary && edge_w && unary_i && edge_wi && smooth_i && ws, "null pointer");
    ISB_REQUIRE(N > 0 && K > 0 && E >= 0, "bad sizes");
    ISB_REQUIRE(metric >= 0 && metric <= 3, "metric must be 0..3");
    ISB_REQUIRE(!spatial || centres, "centres are required for spatially normalised edge weights");
    ISB_REQUIRE(ws_bytes >= isb_gc_energies_workspace_bytes(N, K, E), "workspace too small");
    ProfScope prof(ISB_PROF_ENERGY, (cudaStream_t)stream);
    k_gc_energies<<<ECL, 1024, 0, (cudaStream_t)stream>>>(proba, N, n_nodes_dev, K, edges, E, n_edges_dev, centres, proba, K, metric, spatial, edge_cost,
                                                         pairwise, unary, edge_w, unary_i, edge_wi, smooth_i, (double*)ws);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

__global__ void k_fill_i32(int* p, long long n, int v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int isb_fill_i32(int32_t* dst, long long n, int32_t value, isb_stream_t stream)
{
    ISB_REQUIRE(dst && n > 0, "bad arguments");
    k_fill_i32<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, n, value);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

// dst = dst (op) src on 8-byte words: the in-process stand-in for the collectives of the row-band mode (several bands on one GPU)
__global__ void k_combine(void* dst, const void* src, long long n, int op)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op == 0) ((long long*)dst)[i] += ((const long long*)src)[i];
    else if (op == 1) { long long a = ((long long*)dst)[i], b = ((const long long*)src)[i]; ((long long*)dst)[i] = a > b ? a : b; }
    else if (op == 2) ((double*)dst)[i] = fmin(((double*)dst)[i], ((const double*)src)[i]);
    else if (op == 3) ((double*)dst)[i] = fmax(((double*)dst)[i], ((const double*)src)[i]);
    else ((double*)dst)[i] = ((double*)dst)[i] + ((const double*)src)[i];
}

extern "C" int isb_combine(void* dst, const void* src, long long n, int op, isb_stream_t stream)
{
    ISB_REQUIRE(dst && src && n > 0 && op >= 0 && op <= 4, "bad arguments");
    k_combine<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, src, n, op);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" int isb_gather(const int32_t* seg, long long npx, const int32_t* lut_i, const double* lut_p, int K, int32_t* out_i,
                          double* out_p, isb_stream_t stream)
{
    ISB_REQUIRE(seg && npx > 0, "bad arguments");
    ISB_REQUIRE((!out_i || lut_i) && (!out_p || (lut_p && K > 0)), "LUT missing for a requested output");
    ProfScope prof(ISB_PROF_GATHER, (cudaStream_t)stream);
    k_gather<<<(unsigned)((npx + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seg, npx, lut_i, lut_p, K, out_i, out_p);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
