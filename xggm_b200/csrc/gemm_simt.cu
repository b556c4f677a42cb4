// Exact-fp32 SIMT GEMM for the dense node projections (the "fp32" engine).
//
//   C[M,N] (=|+=) A(M,K) * B(K,N) (+ bias[N]) (+ resid[M,N])
//
// Either operand may be stored k-contiguous or mn-contiguous, which covers the three
// products of a Linear layer without materialising a transpose:
//   forward  out = a w^T      : A k-contig (a[M,K]),  B k-contig (w[N,K])
//   dgrad    ga  = g w        : A k-contig (g[M,N']), B n-contig (w[N',K'])
//   wgrad    gw  = g^T a      : A m-contig (g[rows,N]), B n-contig (a[rows,K]) + split-K
// 128x128x8 CTA tile, 256 threads, 8x8 register tile per thread, register-prefetch
// double buffering.  This engine is the bit-for-bit-reproducible fp32 reference
// engine of the library; the tcgen05 engine (gemm_tc.cu) is the fast one.
#include "common.cuh"

namespace xggm {

constexpr int BM = 128, BN = 128, BK = 8, PAD = 4;

template <bool KC>
__device__ __forceinline__ void load_tile(const float* __restrict__ p, int ld, int mn0, int k0,
                                          int MN, int K, bool vec, float (&r)[4], int t) {
    // KC: element (mn,k) at p[mn*ld + k]; thread -> (mn = t/2, k4 = (t&1)*4), 4 along k
    // !KC: element (mn,k) at p[k*ld + mn]; thread -> (k = t/32, mn4 = (t&31)*4), 4 along mn
    if (KC) {
        const int mn = mn0 + (t >> 1), k = k0 + ((t & 1) << 2);
        if (mn < MN && vec && k + 3 < K) {
            const float4 v = *reinterpret_cast<const float4*>(p + (size_t)mn * ld + k);
            r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                r[i] = (mn < MN && k + i < K) ? p[(size_t)mn * ld + k + i] : 0.f;
        }
    } else {
        const int k = k0 + (t >> 5), mn = mn0 + ((t & 31) << 2);
        if (k < K && vec && mn + 3 < MN) {
            const float4 v = *reinterpret_cast<const float4*>(p + (size_t)k * ld + mn);
            r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                r[i] = (k < K && mn + i < MN) ? p[(size_t)k * ld + mn + i] : 0.f;
        }
    }
}

template <bool KC>
__device__ __forceinline__ void store_tile(float (*s)[BM + PAD], const float (&r)[4], int t) {
    if (KC) {
        const int mn = t >> 1, k = (t & 1) << 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) s[k + i][mn] = r[i];
    } else {
        const int k = t >> 5, mn = (t & 31) << 2;
        *reinterpret_cast<float4*>(&s[k][mn]) = make_float4(r[0], r[1], r[2], r[3]);
    }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                 const float* __restrict__ bias, const float* __restrict__ resid,
                 float* __restrict__ C, int M, int N, int K, int lda, int ldb, int ldc,
                 int vecA, int vecB, int accumulate, int k_per_split) {
    pdl_prologue();
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int t = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const bool split = gridDim.z > 1;
    const int ty = t >> 4, tx = t & 15;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[4], rb[4];
    const int ntiles = (kend - kbeg + BK - 1) / BK;
    if (ntiles > 0) {
        load_tile<A_KC>(A, lda, m0, kbeg, M, kend, vecA, ra, t);
        load_tile<B_KC>(Bm, ldb, n0, kbeg, N, kend, vecB, rb, t);
        store_tile<A_KC>(As[0], ra, t);
        store_tile<B_KC>(Bs[0], rb, t);
    }
    __syncthreads();
    for (int it = 0; it < ntiles; ++it) {
        const int cur = it & 1;
        if (it + 1 < ntiles) {
            load_tile<A_KC>(A, lda, m0, kbeg + (it + 1) * BK, M, kend, vecA, ra, t);
            load_tile<B_KC>(Bm, ldb, n0, kbeg + (it + 1) * BK, N, kend, vecB, rb, t);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < ntiles) {
            store_tile<A_KC>(As[cur ^ 1], ra, t);
            store_tile<B_KC>(Bs[cur ^ 1], rb, t);
        }
        __syncthreads();
    }

    const bool first_split = blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= N) continue;
            float v = acc[i][j];
            const size_t o = (size_t)m * ldc + n;
            if (first_split) {
                if (bias) v += bias[n];
                if (resid) v += resid[o];
            }
            if (split) atomicAdd(&C[o], v);
            else C[o] = accumulate ? C[o] + v : v;
        }
    }
}

// ---- optional per-launch timing of the projection GEMMs (bench.py roofline leg) ----
// When enabled, every GEMM launch is bracketed by a CUDA-event pair on its own stream.
void* gemm_prof_begin(double flops, cudaStream_t st, int members);
struct ProfRec { cudaEvent_t e0, e1; double flops; int members; };
static bool g_prof_on = false;
static ProfRec g_prof[4096];
static int g_prof_n = 0;

int gemm_prof_enable(int on) {
    for (int i = 0; i < g_prof_n; ++i) { cudaEventDestroy(g_prof[i].e0); cudaEventDestroy(g_prof[i].e1); }
    g_prof_n = 0;
    g_prof_on = on != 0;
    return XGGM_OK;
}
// Sums the records of the DOMINANT product shape only (FLOPs per group member >= half of the largest): the
// roofline leg is about the [B*N,768] x [768,768] projections, not the few tiny head GEMMs / Gram tiles.
int gemm_prof_read(double* total_ms, long long* launches, double* flops) {
    *total_ms = 0; *launches = 0; *flops = 0;
    double fmax = 0;
    for (int i = 0; i < g_prof_n; ++i) fmax = fmax > g_prof[i].flops / g_prof[i].members ? fmax : g_prof[i].flops / g_prof[i].members;
    for (int i = 0; i < g_prof_n; ++i) {
        XGGM_CUDA_TRY(cudaEventSynchronize(g_prof[i].e1));
        if (g_prof[i].flops / g_prof[i].members < 0.5 * fmax) continue;
        float ms = 0.f;
        XGGM_CUDA_TRY(cudaEventElapsedTime(&ms, g_prof[i].e0, g_prof[i].e1));
        *total_ms += ms; *flops += g_prof[i].flops; *launches += 1;
    }
    return XGGM_OK;
}
void* gemm_prof_begin(double flops, cudaStream_t st, int members) {
    if (!g_prof_on || g_prof_n >= 4096) return nullptr;
    ProfRec* r = &g_prof[g_prof_n++];
    r->flops = flops;
    r->members = members > 0 ? members : 1;
    cudaEventCreate(&r->e0); cudaEventCreate(&r->e1);
    cudaEventRecord(r->e0, st);
    return r;
}
void gemm_prof_end(void* rec, cudaStream_t st) {
    if (rec) cudaEventRecord(static_cast<ProfRec*>(rec)->e1, st);
}
struct ProfScope {
    void* r; cudaStream_t st;
    ProfScope(double flops, cudaStream_t s) : r(gemm_prof_begin(flops, s, 1)), st(s) {}
    ~ProfScope() { gemm_prof_end(r, st); }
};

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// op: 0 forward (a[M,K], w[N,K]); 1 dgrad (g[M,K=N'], w[K=N',N]); 2 wgrad (g[K=rows,M], a[K=rows,N])
int gemm_simt(int op, const float* A, const float* Bm, const float* bias, const float* resid,
              float* C, int M, int N, int K, int accumulate, cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return XGGM_OK;
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), 1);
    int kps = K;
    if (op == 2) {
        // split-K so a [768,768] weight gradient fills the 148 SMs
        const int tiles = grid.x * grid.y;
        int splits = max(1, min(ceil_div(K, 256), ceil_div(2 * 148, tiles)));
        kps = ceil_div(ceil_div(K, splits), BK) * BK;
        splits = ceil_div(K, kps);
        grid.z = splits;
        if (splits > 1 && !accumulate)
            XGGM_CUDA_TRY(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    }
    ProfScope prof(2.0 * M * N * K, st);
    if (op == 0) {
        const int lda = K, ldb = K;
        XGGM_LAUNCH((gemm_simt_kernel<true, true>), grid, 256, 0, st, 
            A, Bm, bias, resid, C, M, N, K, lda, ldb, N, (lda % 4 == 0) && aligned16(A),
            (ldb % 4 == 0) && aligned16(Bm), accumulate, kps);
    } else if (op == 1) {
        const int lda = K, ldb = N;
        XGGM_LAUNCH((gemm_simt_kernel<true, false>), grid, 256, 0, st, 
            A, Bm, bias, resid, C, M, N, K, lda, ldb, N, (lda % 4 == 0) && aligned16(A),
            (ldb % 4 == 0) && aligned16(Bm), accumulate, kps);
    } else {
        const int lda = M, ldb = N;
        XGGM_LAUNCH((gemm_simt_kernel<false, false>), grid, 256, 0, st, 
            A, Bm, bias, resid, C, M, N, K, lda, ldb, N, (lda % 4 == 0) && aligned16(A),
            (ldb % 4 == 0) && aligned16(Bm), accumulate, kps);
    }
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// column sums: out[c] = sum_r g[r,c]  (bias gradients).  A thread owns one column of a row slab and keeps
// eight independent loads in flight; slabs are sized so that ~4 CTAs per SM exist even for a few hundred rows.
__global__ void __launch_bounds__(128)
colsum_kernel(const float* __restrict__ g, float* __restrict__ out, int R, int C, int rows_per_block) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
    float s = 0.f;
    int r = r0;
    for (; r + 8 <= r1; r += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = g[(size_t)(r + i) * C + c];
        s += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    for (; r < r1; ++r) s += g[(size_t)r * C + c];
    atomicAdd(&out[c], s);
}

int colsum(const float* g, float* out, int R, int C, int accumulate, cudaStream_t st) {
    if (!accumulate) XGGM_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
    if (R <= 0) return XGGM_OK;
    const int gx = ceil_div(C, 128);
    int rpb = ceil_div(R, max(1, 592 / gx));   // ~592 CTAs
    rpb = max(8, (rpb + 7) & ~7);
    dim3 grid(gx, ceil_div(R, rpb));
    XGGM_LAUNCH((colsum_kernel), grid, 128, 0, st, g, out, R, C, rpb);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

}  // namespace xggm
