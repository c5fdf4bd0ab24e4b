// umma_selftest.cu -- known-answer test of the tcgen05 plumbing in umma.cuh (descriptor encodings, tensor-memory allocation,
// commit -> mbarrier, tcgen05.ld lane mapping): D[128, N] = A[128, K] * B[N, K]^T in one CTA with TF32 inputs and an FP32
// accumulator in tensor memory.  Test infrastructure for tests/test_gpu_umma.py; the Leung-Malik contraction (lm_texture.cu)
// uses exactly these operand layouts.
#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

// A [128][K] and B [N][K] row-major f32 in global memory (values already representable in tf32), K % 8 == 0, N % 16 == 0
__global__ void __launch_bounds__(128) k_umma_selftest(const float* __restrict__ A, const float* __restrict__ B, int N, int K, int variant,
                                                       float* __restrict__ D)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float* sA = (float*)sm;                 // [K/4][16][8][4]
    float* sB = sA + 128 * K;               // [K/4][N/8][8][4]
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i / K, k = i - m * K;
        sA[(k >> 2) * (16 * 32) + (m >> 3) * 32 + (m & 7) * 4 + (k & 3)] = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i - n * K;
        sB[(k >> 2) * ((N >> 3) * 32) + (n >> 3) * 32 + (n & 7) * 4 + (k & 3)] = B[i];
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    fence_proxy_async();
    const uint32_t ncols = variant == 2 ? 512u : 128u;
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base), ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base;
    constexpr uint32_t A_COL = 256;     // variant 2: A in tensor memory, lane = row, column A_COL + k
    if (variant == 2) {
        for (int kk = 0; kk < K / 8; ++kk) {
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(A[(size_t)tid * K + kk * 8 + j]);
            tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + A_COL + kk * 8, r);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (tid == 0 && variant == 2) {
        const uint32_t idesc = instr_desc(FMT_TF32, 128, N);
        const uint32_t b_lbo = (uint32_t)(N >> 3) * 128;
        for (int kk = 0; kk < K / 8; ++kk)
            mma_tf32_ts(tbase, tbase + A_COL + kk * 8, smem_desc(smem_u32(sB) + kk * 2 * b_lbo, b_lbo, 128), idesc, kk > 0);
        tc_commit(smem_u32(&bar));
    } else if (tid == 0) {
        const uint32_t idesc = instr_desc(FMT_TF32, 128, N);
        const uint32_t a_lbo = 16 * 128, b_lbo = (uint32_t)(N >> 3) * 128, sbo = 128;
        for (int kk = 0; kk < K / 8; ++kk) {
            const uint32_t a_addr = smem_u32(sA) + kk * 2 * a_lbo, b_addr = smem_u32(sB) + kk * 2 * b_lbo;
            const uint64_t ad = variant ? smem_desc(a_addr, sbo, a_lbo) : smem_desc(a_addr, a_lbo, sbo);
            const uint64_t bd = variant ? smem_desc(b_addr, sbo, b_lbo) : smem_desc(b_addr, b_lbo, sbo);
            mma_tf32(tbase, ad, bd, idesc, kk > 0);
        }
        tc_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const int m = tid; // lane of tensor memory = row of D
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)m * N + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, ncols);
}

// issue rate of tcgen05.mma kind::tf32 (M = 128, K = 8) for a given N: `reps` instructions back to back by one thread, then one commit.
// mode bit 0: rotate over the accumulators that fit (instead of one), bit 1: A from tensor memory (instead of shared memory)
__global__ void __launch_bounds__(128) k_umma_rate(int N, int reps, int mode, long long* cycles)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 8; i += 128) ((float*)sm)[i] = 0.f;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base;
    {   // zero the A columns (496..503) so that no NaN pattern is multiplied
        uint32_t z[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + 496, z);
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (tid == 0) {
        const uint32_t idesc = instr_desc(FMT_TF32, 128, N);
        const int nacc = (mode & 1) ? max(1, 496 / N) : 1;
        const uint64_t ad = smem_desc(smem_u32(sm), 16 * 128, 128);
        const uint64_t bd = smem_desc(smem_u32(sm) + 128 * 8 * 4, (uint32_t)(N >> 3) * 128, 128);
        const long long t0 = clock64();
        int acc = 0;
        for (int i = 0; i < reps; ++i) {
            const uint32_t d = tbase + acc * N;
            if (mode & 2) mma_tf32_ts(d, tbase + 496, bd, idesc, 1u); else mma_tf32(d, ad, bd, idesc, 1u);
            if (++acc == nacc) acc = 0;
        }
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// dependent-issue latency of the FP64 pipe: one warp, chains of n dependent operations; out[0..2] = clocks per DADD, DMUL, DFMA, out[3] = sink
__global__ void k_fp64_latency(int n, double seed, double* out)
{
    double a = seed, b = seed * 0.5, c = seed * 0.25;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = __dadd_rn(a, seed);
    long long t1 = clock64();
    for (int i = 0; i < n; ++i) b = __dmul_rn(b, seed);
    long long t2 = clock64();
    for (int i = 0; i < n; ++i) c = __fma_rn(c, seed, seed);
    long long t3 = clock64();
    if (threadIdx.x == 0) {
        out[0] = (double)(t1 - t0) / n; out[1] = (double)(t2 - t1) / n; out[2] = (double)(t3 - t2) / n; out[3] = a + b + c;
    }
}

} // namespace

extern "C" int isb_fp64_latency(int n, double* out, isb_stream_t stream)
{
    ISB_REQUIRE(out && n > 0, "bad arguments");
    k_fp64_latency<<<1, 32, 0, (cudaStream_t)stream>>>(n, 1.0000001, out);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

// dev / profiling aid: cycles[b] = clocks CTA b needed for `reps` instructions (see k_umma_rate); `ctas` CTAs run the same loop side by side
extern "C" int isb_umma_rate(int N, int reps, int mode, int ctas, long long* cycles, isb_stream_t stream)
{
    ISB_REQUIRE(cycles && N >= 16 && N <= 256 && N % 16 == 0 && reps > 0 && ctas > 0, "bad arguments");
    const size_t smem = (128 + 256) * 8 * 4;
    k_umma_rate<<<ctas, 128, smem, (cudaStream_t)stream>>>(N, reps, mode, cycles);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}

extern "C" int isb_umma_selftest(const float* A, const float* B, int N, int K, int variant, float* D, isb_stream_t stream)
{
    ISB_REQUIRE(A && B && D, "null pointer");
    ISB_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 8 && K <= 64 && K % 8 == 0, "N must be a multiple of 16 <= 256, K a multiple of 8 <= 64");
    ISB_REQUIRE(variant >= 0 && variant <= 2, "variant: 0 / 1 = A from shared memory (descriptor field order), 2 = A from tensor memory");
    const size_t smem = sizeof(float) * (size_t)(128 + N) * K;
    ISB_CUDA_CHECK(cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, N, K, variant, D);
    ISB_LAUNCH_CHECK();
    return ISB_OK;
}
