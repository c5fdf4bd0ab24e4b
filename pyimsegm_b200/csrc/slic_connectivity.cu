// slic_connectivity.cu -- SLIC connectivity enforcement, bit-exact with oracle_enforce_connectivity().
//
// Replaces skimage.segmentation._slic._enforce_label_connectivity_cython (called from slic(),
// imsegm/superpixels.py:61-63).  The original is one sequential raster scan: every unlabelled pixel starts a
// BFS (neighbour order +x,-x,+y,-y) over same-label unlabelled pixels, truncated at max_size; a component
// smaller than min_size takes the label of the LAST already-labelled foreign neighbour the BFS looked at
// ("adjacent", 0 when none), otherwise it gets the next new label.
//
// Parallel decomposition with the same result:
//   1. 4-connected component labelling by union-find; component id = its first pixel in raster order
//      (row runs are linked without atomics, only run-to-run vertical links use atomicMin unions).
//   2. components >= max_size ("oversize", rare) are cut exactly as the truncated BFS would cut them: the
//      pieces depend only on the component's own shape, so one thread replays the BFS per such component.
//   3. a piece is "labelled before C" iff its first pixel precedes C's first pixel, so new labels of kept
//      pieces are a raster-order prefix count of kept roots, and each small piece C replays its own BFS to
//      find the last foreign neighbour pixel whose piece id is < C; chains small->small are followed to a
//      kept piece (or to label 0 when the chain ends without one).
// HBM traffic: a handful of 4 B/px passes over the label map and the scratch arrays.
#include "common.cuh"

namespace {

struct ConnWs {
    int* comp;     // [HW] component id (root pixel index); ~id while an oversize split is being written
    int* size;     // [HW] size, valid at roots
    int* aux;      // [HW] at kept roots: new label; at small roots: adjacent piece id (-1 none)
    int* queue;    // [HW] BFS queues
    int* list;     // [HW] oversize list, later small-root list
    int* row_cnt;  // [H + 1]
    int* ctr;      // [8] counters: 0 n_oversize, 1 n_small, 2 queue cursor
    int4* bbox;    // [n_oversize_max]
};

constexpr int VISBIT = 1 << 30; // 'visited by the small-piece BFS' flag kept inside comp[] (pixel indices stay below 2^30)
__device__ __forceinline__ int dec(int c) { return (c < 0 ? ~c : c) & ~VISBIT; }

__device__ __forceinline__ int find_root(const int* parent, int i)
{
    while (true) {
        int p = parent[i];
        if (p == i) return i;
        i = p;
    }
}

__device__ __forceinline__ void unite(int* parent, int a, int b)
{
    while (true) {
        a = find_root(parent, a);
        b = find_root(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

// one CTA per row: parent[p] = index of the start of p's run of equal labels
__global__ void __launch_bounds__(256) k_row_runs(const int* __restrict__ lab, int H, int W, int* __restrict__ parent)
{
    const int y = blockIdx.x;
    const int* row = lab + (size_t)y * W;
    int* prow = parent + (size_t)y * W;
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < W; base += 256) {
        int x = base + threadIdx.x;
        int v = -1; // start index if this pixel starts a run
        if (x < W) v = (x == 0 || row[x - 1] != row[x]) ? x : -1;
        // inclusive max-scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v = max(v, t);
        }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        int pre = s_carry;
        for (int w = 0; w < warp; ++w) pre = max(pre, s_warp[w]);
        v = max(v, pre);
        if (x < W) prow[x] = y * W + v;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = v;
        __syncthreads();
    }
}

__global__ void k_merge_vertical(const int* __restrict__ lab, int H, int W, int* parent)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)H * W || p < (size_t)W) return;
    int x = (int)(p % W);
    int l = lab[p];
    if (lab[p - W] != l) return;
    // skip when the left neighbour already made the same link
    if (x > 0 && lab[p - 1] == l && lab[p - W - 1] == l) return;
    unite(parent, (int)p, (int)(p - W));
}

__global__ void k_flatten_sizes(const int* __restrict__ lab, int H, int W, const int* __restrict__ parent, int* __restrict__ comp,
                                int* size)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)H * W) return;
    int x = (int)(p % W);
    int rs = parent[p];                       // run start (or, for the run start itself, its tree parent)
    int root = find_root(parent, rs);
    comp[p] = root;
    if (x == W - 1 || lab[p + 1] != lab[p]) {  // run end: add the run length once
        // parent[p] is still the run start unless p starts the run itself (then it may point up the tree)
        int xs = (x > 0 && lab[p - 1] == lab[p]) ? rs % W : x;
        atomicAdd(&size[root], x - xs + 1);
    }
}

__global__ void k_collect_oversize(int n, const int* __restrict__ comp, const int* __restrict__ size, int max_size, int* list,
                                   int* ctr, int* aux)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (comp[p] == p && size[p] >= max_size) {
        int slot = atomicAdd(&ctr[0], 1);
        list[slot] = p;
        aux[p] = slot;
    }
}

__global__ void k_oversize_bbox(int H, int W, const int* __restrict__ comp, const int* __restrict__ size, int max_size,
                                const int* __restrict__ aux, const int* __restrict__ ctr, int4* bbox)
{
    if (ctr[0] == 0) return;
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)H * W) return;
    int r = comp[p];
    if (size[r] < max_size) return;
    int slot = aux[r];
    int y = (int)(p / W), x = (int)(p % W);
    atomicMin(&bbox[slot].x, y); atomicMax(&bbox[slot].y, y);
    atomicMin(&bbox[slot].z, x); atomicMax(&bbox[slot].w, x);
}

// one thread per oversize component: replay the truncated BFS; assigned pixels get comp = ~piece_root
__global__ void k_oversize_split(int H, int W, int* comp, int* size, int max_size, const int* __restrict__ list,
                                 const int* __restrict__ ctr, const int4* __restrict__ bbox, int* queue)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ctr[0]) return;
    const int C = list[i];
    const int4 bb = bbox[i];
    int* q = queue + (size_t)i * max_size;
    const int total = size[C];
    int assigned = 0;
    for (int y = bb.x; y <= bb.y && assigned < total; ++y)
        for (int x = bb.z; x <= bb.w && assigned < total; ++x) {
            int p = y * W + x;
            if (comp[p] != C) continue; // foreign, or already assigned (negative)
            int start = p;
            comp[p] = ~start;
            q[0] = p;
            int n = 1, v = 0;
            while (v < n && n < max_size) {
                int cp = q[v];
                int cy = cp / W, cx = cp - cy * W;
                const int nx[4] = { cx + 1, cx - 1, cx, cx };
                const int ny[4] = { cy, cy, cy + 1, cy - 1 };
                for (int d = 0; d < 4; ++d) {
                    if (nx[d] < 0 || nx[d] >= W || ny[d] < 0 || ny[d] >= H) continue;
                    int np_ = ny[d] * W + nx[d];
                    if (comp[np_] == C) {
                        comp[np_] = ~start;
                        q[n++] = np_;
                        if (n >= max_size) break;
                    }
                }
                ++v;
            }
            size[start] = n;
            assigned += n;
        }
}

// rank of kept roots in raster order: per-row counts -> scan -> per-row assignment
__global__ void __launch_bounds__(256) k_row_count_kept(int H, int W, const int* __restrict__ comp, const int* __restrict__ size,
                                                        int min_size, int* row_cnt)
{
    const int y = blockIdx.x;
    int c = 0;
    for (int x = threadIdx.x; x < W; x += 256) {
        int p = y * W + x;
        c += (dec(comp[p]) == p && size[p] >= min_size) ? 1 : 0;
    }
    __shared__ int s[256];
    s[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) row_cnt[y] = s[0];
}

__global__ void __launch_bounds__(1024) k_scan_rows(int H, int* row_cnt, int* n_labels_out)
{
    __shared__ int s_part[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < H; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < H ? row_cnt[i] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = s_part[threadIdx.x], carry = s_carry;
        if (i < H) row_cnt[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + incl;
        __syncthreads();
    }
    // when no component reaches min_size everything is merged into label 0: the map still holds one label
    if (threadIdx.x == 0) { row_cnt[H] = s_carry; *n_labels_out = s_carry > 0 ? s_carry : 1; }
}

__global__ void __launch_bounds__(256) k_row_assign_labels(int H, int W, const int* __restrict__ comp, const int* __restrict__ size,
                                                           int min_size, const int* __restrict__ row_cnt, int* aux, int* list, int* ctr)
{
    const int y = blockIdx.x;
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = row_cnt[y];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < W; base += 256) {
        int x = base + threadIdx.x;
        int p = y * W + x;
        bool root = x < W && dec(comp[p]) == p;
        bool kept = root && size[p] >= min_size;
        unsigned m = __ballot_sync(0xffffffffu, kept);
        int pre = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = s_carry;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (kept) aux[p] = off + pre;
        else if (root) { aux[p] = -1; list[atomicAdd(&ctr[1], 1)] = p; }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_carry += t; }
        __syncthreads();
    }
}

// one thread per small piece: replay its BFS, remember the last foreign earlier-labelled neighbour piece.
// The kernel is a chain of dependent loads per piece (queue entry -> four neighbours), as long as the largest small piece: the queue of
// a piece (fewer than min_size entries) lives in SHARED memory when it fits (SQ: 32 threads per CTA, entry j of lane l at q[32 j + l]),
// which takes one of the two global round trips out of every BFS step.
template <bool SQ>
__global__ void __launch_bounds__(SQ ? 32 : 256) k_small_adjacent(int H, int W, int* comp, const int* __restrict__ size, int max_size,
                                                                   const int* __restrict__ list, int* ctr, int* aux, int* queue)
{
    extern __shared__ int s_q[];
    const int n_small = ctr[1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_small; i += gridDim.x * blockDim.x) {
        const int C = list[i];
        int* q = SQ ? s_q + threadIdx.x : queue + atomicAdd(&ctr[2], size[C]);
        constexpr int QS = SQ ? 32 : 1;
        int adjacent = -1;
        q[0] = C;
        // the visited flag lives in comp[] itself (only this thread writes the pixels of its own piece; every reader masks the
        // flag with dec()), so one round of four independent loads per BFS step is all the memory latency there is
        comp[C] = (comp[C] < 0 ? ~(dec(comp[C]) | VISBIT) : (comp[C] | VISBIT));
        int n = 1, v = 0;
        while (v < n && n < max_size) {
            const int cp = q[v * QS];
            const int cy = cp / W, cx = cp - cy * W;
            // the four neighbours in the original's order (+x, -x, +y, -y)
            const int np4[4] = { cx + 1 < W ? cp + 1 : -1, cx > 0 ? cp - 1 : -1, cy + 1 < H ? cp + W : -1, cy > 0 ? cp - W : -1 };
            int raw4[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) raw4[d] = np4[d] >= 0 ? comp[np4[d]] : 0;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (np4[d] < 0) continue;
                const int raw = raw4[d];
                const int r = dec(raw);
                if (r == C) {
                    const bool seen = ((raw < 0 ? ~raw : raw) & VISBIT) != 0; // the four neighbours of one pixel are distinct
                    if (!seen) {
                        comp[np4[d]] = raw < 0 ? ~((~raw) | VISBIT) : (raw | VISBIT);
                        q[(n++) * QS] = np4[d];
                        if (n >= max_size) break;
                    }
                } else if (r < C) {
                    adjacent = r;
                }
            }
            ++v;
        }
        aux[C] = adjacent >= 0 ? -2 - adjacent : -1; // small roots store -2-adjacent (kept roots store label >= 0)
    }
}

__global__ void k_write_labels(int n, const int* __restrict__ comp, const int* __restrict__ size, int min_size,
                               const int* __restrict__ aux, int* __restrict__ out)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int r = dec(comp[p]);
    int a = aux[r];
    // follow small -> adjacent chains (each hop goes to a piece with a smaller root index)
    while (a < -1) { r = -2 - a; a = aux[r]; }
    out[p] = a < 0 ? 0 : a;
}

static size_t carve(ConnWs& w, void* ws, size_t bytes, int H, int W)
{
    WsCarver c(ws, bytes);
    size_t n = (size_t)H * W;
    w.comp = c.take<int>(n); w.size = c.take<int>(n); w.aux = c.take<int>(n);
    w.queue = c.take<int>(n); w.list = c.take<int>(n);
    w.row_cnt = c.take<int>((size_t)H + 1);
    w.ctr = c.take<int>(8);
    w.bbox = c.take<int4>(n / 16 + 16);
    return isb_align(c.off);
}

__global__ void k_init_bbox(int n, int4* bbox)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bbox[i] = make_int4(INT_MAX, -1, INT_MAX, -1);
}

} // namespace

extern "C" size_t isb_connectivity_workspace_bytes(int H, int W)
{
    ConnWs w;
    return carve(w, nullptr, 0, H, W);
}

extern "C" int isb_enforce_connectivity(const int32_t* labels, int H, int W, int min_size, int max_size, int32_t* out,
                                        int32_t* n_labels_out, void* ws, size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(labels && out && n_labels_out && ws, "null pointer");
    ISB_REQUIRE(H > 0 && W > 0 && (long long)H * W < (1LL << 30), "bad image size (at most 2^30 pixels)");
    if (max_size < 1) max_size = 1;
    ISB_REQUIRE(max_size >= 16, "max_size < 16 is not supported on the device path (oversize table bound)");
    ConnWs w;
    size_t need = carve(w, ws, ws_bytes, H, W);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(ISB_PROF_CONN, st);
    const int n = H * W;
    const int nb = (n + 255) / 256;
    ISB_CUDA_CHECK(cudaMemsetAsync(w.size, 0, sizeof(int) * (size_t)n, st));
    ISB_CUDA_CHECK(cudaMemsetAsync(w.ctr, 0, sizeof(int) * 8, st));
    k_row_runs<<<H, 256, 0, st>>>(labels, H, W, w.comp);
    ISB_LAUNCH_CHECK();
    k_merge_vertical<<<nb, 256, 0, st>>>(labels, H, W, w.comp);
    ISB_LAUNCH_CHECK();
    // comp doubles as the union-find parent array; flatten writes roots into aux first, then swap roles
    k_flatten_sizes<<<nb, 256, 0, st>>>(labels, H, W, w.comp, w.aux, w.size);
    ISB_LAUNCH_CHECK();
    int* comp = w.aux;   // flattened component ids
    int* aux = w.comp;   // parent array is dead now: reuse as aux
    const int n_over_max = n / max_size + 1;
    k_init_bbox<<<(n_over_max + 255) / 256, 256, 0, st>>>(n_over_max, w.bbox);
    ISB_LAUNCH_CHECK();
    k_collect_oversize<<<nb, 256, 0, st>>>(n, comp, w.size, max_size, w.list, w.ctr, aux);
    ISB_LAUNCH_CHECK();
    k_oversize_bbox<<<nb, 256, 0, st>>>(H, W, comp, w.size, max_size, aux, w.ctr, w.bbox);
    ISB_LAUNCH_CHECK();
    k_oversize_split<<<(n_over_max + 63) / 64, 64, 0, st>>>(H, W, comp, w.size, max_size, w.list, w.ctr, w.bbox, w.queue);
    ISB_LAUNCH_CHECK();
    k_row_count_kept<<<H, 256, 0, st>>>(H, W, comp, w.size, min_size, w.row_cnt);
    ISB_LAUNCH_CHECK();
    k_scan_rows<<<1, 1024, 0, st>>>(H, w.row_cnt, n_labels_out);
    ISB_LAUNCH_CHECK();
    k_row_assign_labels<<<H, 256, 0, st>>>(H, W, comp, w.size, min_size, w.row_cnt, aux, w.list, w.ctr);
    ISB_LAUNCH_CHECK();
    // the kernel reads the real count of small roots and strides over them
    {
        int dev = 0, sms = 0;
        ISB_CUDA_CHECK(cudaGetDevice(&dev));
        ISB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (min_size >= 1 && min_size <= 512) {
            // a small piece has fewer than min_size pixels: its queue fits min_size entries of shared memory per thread
            const size_t smem = sizeof(int) * 32 * (size_t)min_size;
            ISB_CUDA_CHECK(cudaFuncSetAttribute(k_small_adjacent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_small_adjacent<true><<<sms * 16, 32, smem, st>>>(H, W, comp, w.size, max_size, w.list, w.ctr, aux, w.queue);
        } else {
            k_small_adjacent<false><<<sms * 8, 256, 0, st>>>(H, W, comp, w.size, max_size, w.list, w.ctr, aux, w.queue);
        }
        ISB_LAUNCH_CHECK();
    }
    k_write_labels<<<nb, 256, 0, st>>>(n, comp, w.size, min_size, aux, out);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
