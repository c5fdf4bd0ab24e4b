// alpha_expansion.cu -- multi-label GraphCut by alpha-expansion on the superpixel adjacency graph.
//
// Replaces gco.cut_general_graph(edges, edge_weights, unary, pairwise, algorithm='expansion', n_iter)
// as called from imsegm/graph_cuts.py:735-744 (and region_growing.py:148,1698,1715), on the integer energies
// pyGCO builds (see oracle/gc_oracle.cpp for the restated contract).
//
// One CTA (1024 threads) owns one graph for the whole optimisation -- GCO's cycle bookkeeping, every
// expansion move and every max-flow run inside a single launch, no host round trips.
//
// A move on alpha is the exact minimum of a binary submodular energy (Kolmogorov-Zabih construction):
//   x_i = 0 take alpha / x_i = 1 keep;   source->i capacity = cost(x_i=1), i->sink = cost(x_i=0),
//   i->j capacity P_ij = E01 + E10 - E00 - E11 >= 0.
// Max-flow is phase-1 push-relabel (preflow, min cut only):
//   * global relabel = level-synchronous BACKWARD BFS from the sink over residual arcs (exact distance labels;
//     nodes that cannot reach the sink drop out),
//   * node-parallel push/relabel sweeps in between (lock-free pushes with atomics, Hong & He style),
//   * terminates when no node with excess can reach the sink.
// The site keeps its label iff it can reach the sink in the final residual graph -- BK's SINK segment, the
// unique minimiser with the most sites switched -- so labels equal the oracle's whatever the flow algorithm.
//
// State: flow f_e in [0, P_e] per undirected edge (residual a->b = P_e - f_e, b->a = f_e), excess and sink
// capacity per node, heights.  These mutable arrays live in SHARED memory when they fit (N = 5k, E = 15k needs
// 220 KB of the 227 KB), otherwise in the global workspace (L2 resident); the read-only CSR stays in global/L1.
// This stage is latency/SMEM bound, not HBM bound: report time, not a roofline fraction.
#include "common.cuh"

namespace {

constexpr int NT = 1024;
constexpr int HINF = 0x3fffffff;
constexpr int KMAX_S = 16; // smooth-cost table cached in smem up to K = 16

struct GcArgs {
    int N, K, E_cap;
    const int* n_edges_dev; const int* n_nodes_dev;
    const int* edges; const int* w; const int* D; const int* V;
    int n_iter;
    int* labels;
    long long* energy_out;
    int* stats;
    // workspace (global)
    int* off; int* fill; int* adj_v; int* adj_e;
    long long* u0; long long* u1;
    int* newlab;
    // mutable flow state in global memory (used when it does not fit in smem)
    int* g_flow; int* g_cap; int* g_height; long long* g_excess; long long* g_tcap;
    int dyn_bytes;
};

__device__ long long block_sum_ll(long long v, long long* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
    for (int i = 0; i < NT / 32; ++i) t += s_red[i];
    return t;
}

struct Ctx {
    int N, K, E;
    const int* edges; const int* w; const int* D; const int* V;
    const int* off; const int* adj_v; const int* adj_e;
    volatile int* flow; volatile int* cap; volatile int* height;
    volatile long long* excess; volatile long long* tcap;
    int* s_V; long long* s_red;
};

__device__ __forceinline__ int smooth(const Ctx& c, int la, int lb) { return c.K <= KMAX_S ? c.s_V[la * c.K + lb] : c.V[la * c.K + lb]; }

__device__ long long energy_of(const Ctx& c, const int* lab)
{
    long long e = 0;
    for (int i = threadIdx.x; i < c.N; i += NT) e += c.D[(size_t)i * c.K + lab[i]];
    for (int k = threadIdx.x; k < c.E; k += NT) e += (long long)c.w[k] * smooth(c, lab[c.edges[2 * k]], lab[c.edges[2 * k + 1]]);
    return block_sum_ll(e, c.s_red);
}

// backward BFS from the sink: height = 1 + distance to a node with sink capacity; unreachable = HINF.
// returns (block-uniform) whether any node with excess can reach the sink
__device__ bool global_relabel(const Ctx& c, int* stats_relabels)
{
    for (int v = threadIdx.x; v < c.N; v += NT) c.height[v] = (c.tcap[v] > 0) ? 1 : HINF;
    __syncthreads();
    for (int level = 1;; ++level) {
        int changed = 0;
        for (int u = threadIdx.x; u < c.N; u += NT) {
            if (c.height[u] != HINF) continue;
            for (int i = c.off[u]; i < c.off[u + 1]; ++i) {
                int v = c.adj_v[i];
                if (c.height[v] != level) continue;
                int ed = c.adj_e[i];
                int e = ed >> 1;
                // residual of arc u -> v: u is 'a' (dir 0): cap - flow ; u is 'b' (dir 1): flow
                int res = (ed & 1) ? c.flow[e] : c.cap[e] - c.flow[e];
                if (res > 0) { c.height[u] = level + 1; changed = 1; break; }
            }
        }
        // a node relabelled in this pass has height level+1, never == level, so the pass is race-free
        if (!__syncthreads_or(changed)) break;
    }
    int active = 0;
    for (int v = threadIdx.x; v < c.N; v += NT) active |= (c.excess[v] > 0 && c.height[v] != HINF);
    if (threadIdx.x == 0 && stats_relabels) ++*stats_relabels;
    return __syncthreads_or(active) != 0;
}

// one synchronous push-relabel sweep: (A) pushes against frozen heights, barrier, (B) relabels.
// With heights frozen during (A) every push goes from height h to h-1, so the labelling stays valid
// (h[u] <= h[v] + 1 on every residual arc); in (B) a neighbour height read while it is being raised is only
// ever too LOW, which keeps validity.  Returns whether any node is still active.
__device__ bool sweep(const Ctx& c)
{
    for (int u = threadIdx.x; u < c.N; u += NT) {
        long long ex = c.excess[u];
        const int hu = c.height[u];
        if (ex <= 0 || hu == HINF) continue;
        long long tc = c.tcap[u];
        if (tc > 0) { // the sink arc first (only u touches it)
            long long d = ex < tc ? ex : tc;
            c.tcap[u] = tc - d;
            ex -= d;
            atomicAdd((unsigned long long*)&c.excess[u], (unsigned long long)(-d));
        }
        for (int i = c.off[u]; i < c.off[u + 1] && ex > 0; ++i) {
            const int v = c.adj_v[i];
            if (c.height[v] >= hu) continue;
            const int ed = c.adj_e[i];
            const int e = ed >> 1;
            const int res = (ed & 1) ? c.flow[e] : c.cap[e] - c.flow[e];
            if (res <= 0) continue;
            const int d = ex < (long long)res ? (int)ex : res;
            atomicAdd((int*)&c.flow[e], (ed & 1) ? -d : d);
            atomicAdd((unsigned long long*)&c.excess[v], (unsigned long long)(long long)d);
            atomicAdd((unsigned long long*)&c.excess[u], (unsigned long long)(-(long long)d));
            ex -= d;
        }
    }
    __syncthreads();
    int active = 0;
    for (int u = threadIdx.x; u < c.N; u += NT) {
        const int hu = c.height[u];
        if (c.excess[u] <= 0 || hu == HINF) continue;
        active = 1;
        if (c.tcap[u] > 0) continue; // still sink-adjacent: height 1 is exact
        int hmin = HINF;
        for (int i = c.off[u]; i < c.off[u + 1]; ++i) {
            const int ed = c.adj_e[i];
            const int e = ed >> 1;
            const int res = (ed & 1) ? c.flow[e] : c.cap[e] - c.flow[e];
            if (res <= 0) continue;
            const int hv = c.height[c.adj_v[i]];
            if (hv < hmin) hmin = hv;
        }
        if (hmin == HINF) c.height[u] = HINF;          // no residual way out: source side for good
        else if (hmin >= hu) c.height[u] = hmin + 1;    // relabel
    }
    return __syncthreads_or(active) != 0;
}

__global__ void __launch_bounds__(NT, 1) k_alpha_expansion(GcArgs a)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ int s_V[KMAX_S * KMAX_S];
    __shared__ long long s_red[NT / 32];
    __shared__ int s_scan[NT];
    __shared__ int s_carry;
    __shared__ int s_table[64], s_queue[64], s_qn;
    __shared__ int s_stats[4];

    Ctx c;
    c.N = a.n_nodes_dev ? min(*a.n_nodes_dev, a.N) : a.N; c.K = a.K;
    c.E = a.n_edges_dev ? min(*a.n_edges_dev, a.E_cap) : a.E_cap;
    c.edges = a.edges; c.w = a.w; c.D = a.D; c.V = a.V;
    c.off = a.off; c.adj_v = a.adj_v; c.adj_e = a.adj_e;
    c.s_V = s_V; c.s_red = s_red;
    const size_t smem_need = sizeof(long long) * 2 * (size_t)c.N + sizeof(int) * (2 * (size_t)c.E + (size_t)c.N);
    if (smem_need <= (size_t)a.dyn_bytes) {
        // layout: excess[N] ll | tcap[N] ll | flow[E] | cap[E] | height[N]
        long long* p = (long long*)dyn;
        c.excess = p; c.tcap = p + c.N;
        int* q = (int*)(p + 2 * (size_t)c.N);
        c.flow = q; c.cap = q + c.E; c.height = q + 2 * (size_t)c.E;
    } else {
        c.excess = a.g_excess; c.tcap = a.g_tcap; c.flow = a.g_flow; c.cap = a.g_cap; c.height = a.g_height;
    }
    const int N = c.N, K = c.K, E = c.E;
    if (threadIdx.x < 4) s_stats[threadIdx.x] = 0;
    if (K <= KMAX_S) for (int i = threadIdx.x; i < K * K; i += NT) s_V[i] = a.V[i];

    // ---- CSR of the undirected graph (read-only afterwards) ----
    for (int v = threadIdx.x; v < N; v += NT) a.fill[v] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += NT) { atomicAdd(&a.fill[a.edges[2 * e]], 1); atomicAdd(&a.fill[a.edges[2 * e + 1]], 1); }
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < N; base += NT) {
        int i = base + threadIdx.x;
        int v = i < N ? __ldcg(&a.fill[i]) : 0;
        s_scan[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < NT; o <<= 1) {
            int t = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = s_scan[threadIdx.x], carry = s_carry;
        if (i < N) { a.off[i] = carry + incl - v; a.fill[i] = 0; }
        __syncthreads();
        if (threadIdx.x == NT - 1) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.off[N] = s_carry;
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += NT) {
        int va = a.edges[2 * e], vb = a.edges[2 * e + 1];
        int pa = a.off[va] + atomicAdd(&a.fill[va], 1);
        a.adj_v[pa] = vb; a.adj_e[pa] = 2 * e;
        int pb = a.off[vb] + atomicAdd(&a.fill[vb], 1);
        a.adj_v[pb] = va; a.adj_e[pb] = 2 * e + 1;
    }
    __syncthreads();

    int* lab = a.labels;
    long long cur_energy;

    auto expand = [&](int alpha) -> bool {
        // any active site?
        int any = 0;
        for (int v = threadIdx.x; v < N; v += NT) any |= (lab[v] != alpha);
        if (!__syncthreads_or(any)) return false;
        // unary part of the move energy
        for (int v = threadIdx.x; v < N; v += NT) {
            int l = lab[v];
            if (l != alpha) { a.u0[v] = c.D[(size_t)v * K + alpha]; a.u1[v] = c.D[(size_t)v * K + l]; }
            else { a.u0[v] = 0; a.u1[v] = 0; }
        }
        __syncthreads();
        const int Vaa = smooth(c, alpha, alpha);
        int bad = 0;
        for (int e = threadIdx.x; e < E; e += NT) {
            int va = c.edges[2 * e], vb = c.edges[2 * e + 1];
            int la = lab[va], lb = lab[vb];
            long long wk = c.w[e];
            int P = 0;
            if (la != alpha && lb != alpha) {
                long long A = wk * Vaa, B = wk * smooth(c, alpha, lb), C = wk * smooth(c, la, alpha), Dd = wk * smooth(c, la, lb);
                atomicAdd((unsigned long long*)&a.u0[va], (unsigned long long)A);
                atomicAdd((unsigned long long*)&a.u1[va], (unsigned long long)C);
                atomicAdd((unsigned long long*)&a.u1[vb], (unsigned long long)(Dd - C));
                long long Pl = B + C - A - Dd;
                if (Pl < 0) bad = 1;
                P = (int)Pl;
            } else if (la != alpha) {
                atomicAdd((unsigned long long*)&a.u0[va], (unsigned long long)(wk * Vaa));
                atomicAdd((unsigned long long*)&a.u1[va], (unsigned long long)(wk * smooth(c, la, alpha)));
            } else if (lb != alpha) {
                atomicAdd((unsigned long long*)&a.u0[vb], (unsigned long long)(wk * Vaa));
                atomicAdd((unsigned long long*)&a.u1[vb], (unsigned long long)(wk * smooth(c, alpha, lb)));
            }
            c.cap[e] = P;
            c.flow[e] = 0;
        }
        if (__syncthreads_or(bad)) return false; // non-submodular move (GCO refuses it)
        for (int v = threadIdx.x; v < N; v += NT) {
            long long x0 = __ldcg(&a.u0[v]), x1 = __ldcg(&a.u1[v]);
            long long m = x0 < x1 ? x0 : x1;
            bool act = lab[v] != alpha;
            c.excess[v] = act ? x1 - m : 0;
            c.tcap[v] = act ? x0 - m : 0;
        }
        __syncthreads();
        // ---- max-flow (phase 1) ----
        if (threadIdx.x == 0) ++s_stats[1];
        while (global_relabel(c, &s_stats[3])) {
            for (int s = 0; s < 64; ++s) {
                if (threadIdx.x == 0) ++s_stats[2];
                if (!sweep(c)) break;
            }
        }
        // ---- candidate labeling: keep iff the site can still reach the sink ----
        for (int v = threadIdx.x; v < N; v += NT) a.newlab[v] = (lab[v] != alpha && c.height[v] == HINF) ? alpha : lab[v];
        __syncthreads();
        long long e_new = energy_of(c, a.newlab);
        if (e_new < cur_energy) {
            for (int v = threadIdx.x; v < N; v += NT) lab[v] = a.newlab[v];
            cur_energy = e_new;
            if (threadIdx.x == 0) ++s_stats[0];
            __syncthreads();
            return true;
        }
        return false;
    };

    if (E == 0) {
        // no smoothness: GCO's special case, independent argmin per site
        for (int v = threadIdx.x; v < N; v += NT) {
            int best = 0;
            for (int l = 1; l < K; ++l) if (c.D[(size_t)v * K + l] < c.D[(size_t)v * K + best]) best = l;
            lab[v] = best;
        }
        __syncthreads();
    } else {
        cur_energy = energy_of(c, lab);
        const int KT = K < 64 ? K : 64; // label table lives in smem; K > 64 is rejected on the host
        if (threadIdx.x == 0) { for (int l = 0; l < KT; ++l) s_table[l] = l; s_qn = 1; s_queue[0] = KT; }
        __syncthreads();
        if (a.n_iter == -1) {
            // GCO adaptive cycles (see oracle/gc_oracle.cpp)
            while (true) {
                __syncthreads();
                if (s_qn == 0) break;
                const int qsz = s_queue[s_qn - 1];
                int start = KT - qsz;
                for (int next = start; next < KT; ++next) {
                    __syncthreads();
                    const int alpha = s_table[next];
                    bool ok = expand(alpha);
                    __syncthreads();
                    if (!ok) {
                        if (threadIdx.x == 0) { int t = s_table[next]; s_table[next] = s_table[start]; s_table[start] = t; }
                        ++start;
                    }
                }
                __syncthreads();
                const int nsz = KT - start;
                if (threadIdx.x == 0) {
                    if (nsz == qsz) { /* all succeeded: run the same queue again */ }
                    else if (nsz > 0) { if (s_qn < 64) s_queue[s_qn++] = nsz; }
                    else --s_qn;
                }
            }
        } else {
            for (int cycle = 0; cycle < a.n_iter; ++cycle) {
                long long before = cur_energy;
                for (int l = 0; l < KT; ++l) { __syncthreads(); expand(l); }
                if (cur_energy == before) break;
            }
        }
        __syncthreads();
    }
    long long e_fin = energy_of(c, lab);
    if (threadIdx.x == 0) {
        if (a.energy_out) *a.energy_out = e_fin;
        if (a.stats) for (int i = 0; i < 4; ++i) a.stats[i] = s_stats[i];
    }
}

struct GcWs {
    int* off; int* fill; int* adj_v; int* adj_e; long long* u0; long long* u1; int* newlab;
    int* g_flow; int* g_cap; int* g_height; long long* g_excess; long long* g_tcap;
};

static size_t carve_gc(GcWs& w, void* ws, size_t bytes, int N, int E)
{
    WsCarver c(ws, bytes);
    size_t e = E > 0 ? E : 1;
    w.off = c.take<int>((size_t)N + 1); w.fill = c.take<int>(N);
    w.adj_v = c.take<int>(2 * e); w.adj_e = c.take<int>(2 * e);
    w.u0 = c.take<long long>(N); w.u1 = c.take<long long>(N);
    w.newlab = c.take<int>(N);
    w.g_flow = c.take<int>(e); w.g_cap = c.take<int>(e); w.g_height = c.take<int>(N);
    w.g_excess = c.take<long long>(N); w.g_tcap = c.take<long long>(N);
    return isb_align(c.off);
}

} // namespace

extern "C" size_t isb_alpha_expansion_workspace_bytes(int N, int K, int E)
{
    GcWs w;
    return carve_gc(w, nullptr, 0, N, E);
}

extern "C" int isb_alpha_expansion(int N, const int32_t* n_nodes_dev, int K, int E, const int32_t* n_edges_dev, const int32_t* edges, const int32_t* edge_wi,
                                   const int32_t* unary_i, const int32_t* smooth_i, int n_iter, int32_t* labels, int64_t* energy_out,
                                   int32_t* stats_out, void* ws, size_t ws_bytes, isb_stream_t stream)
{
    ISB_REQUIRE(edges && edge_wi && unary_i && smooth_i && labels && ws, "null pointer");
    ISB_REQUIRE(N > 0 && K > 0 && K <= 64 && E >= 0, "bad sizes (K must be <= 64)");
    ISB_REQUIRE(n_iter == -1 || n_iter > 0, "n_iter must be -1 (adaptive cycles) or positive");
    GcWs w;
    size_t need = carve_gc(w, ws, ws_bytes, N, E);
    ISB_REQUIRE(need <= ws_bytes, "workspace too small");
    GcArgs a;
    a.N = N; a.K = K; a.E_cap = E; a.n_edges_dev = n_edges_dev; a.n_nodes_dev = n_nodes_dev;
    a.edges = edges; a.w = edge_wi; a.D = unary_i; a.V = smooth_i; a.n_iter = n_iter;
    a.labels = labels; a.energy_out = (long long*)energy_out; a.stats = stats_out;
    a.off = w.off; a.fill = w.fill; a.adj_v = w.adj_v; a.adj_e = w.adj_e; a.u0 = w.u0; a.u1 = w.u1; a.newlab = w.newlab;
    a.g_flow = w.g_flow; a.g_cap = w.g_cap; a.g_height = w.g_height; a.g_excess = w.g_excess; a.g_tcap = w.g_tcap;
    // always launch with the full dynamic smem: the kernel decides from the REAL edge count (device scalar)
    // whether the mutable flow state fits there or stays in the global workspace
    const size_t smem_max = 227 * 1024 - 8 * 1024; // leave room for the static arrays
    const size_t smem = smem_max;
    a.dyn_bytes = (int)smem_max;
    static bool attr_set = false;
    if (!attr_set) {
        ISB_CUDA_CHECK(cudaFuncSetAttribute(k_alpha_expansion, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        attr_set = true;
    }
    ProfScope prof(ISB_PROF_GC, (cudaStream_t)stream);
    k_alpha_expansion<<<1, NT, smem, (cudaStream_t)stream>>>(a);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
