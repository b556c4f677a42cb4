// Per-graph operators: each N x N adjacency is staged whole in shared memory.
//   adj_apply : out = self_w*x + alpha * (adj | adj^T) @ x        (message passing)
//   bmm_nt    : out[b] = p[b] q[b]^T  ([N,H] x [H,N])             (gadj, pair scores)
//   adj_regen : sigmoid(S / colmax) with zero diagonal, fwd + bwd  (ggm.py:225-228)
//   gat_attn  : dense masked attention of GATConv                  (gat.py:25-49)
#include <cuda_bf16.h>

#include "common.cuh"

namespace xggm {

constexpr int MAX_NODES = 104;  // smem sizing of the N x N stages (cfg-4 sweep goes to N=100)

// ----------------------------------------------------------------- adj_apply
typedef __nv_bfloat16 bf16;
__device__ __forceinline__ void split_bf16(float v, bf16& h, bf16& l) {
    h = __float2bfloat16_rn(v);
    const float hf = __bfloat162float(h);
    l = __float2bfloat16_rn((hf - hf == 0.f) ? v - hf : 0.f);
}

// Generic kernel (any N <= MAX_NODES): grid (ceil(H/256), B); thread = one feature column;
// adjacency in smem (broadcast reads); node rows are walked in register tiles of RT.
template <int RT, bool TRANS>
__global__ void __launch_bounds__(256)
adj_apply_kernel(const float* __restrict__ adj, const float* __restrict__ x,
                 float* __restrict__ out, bf16* __restrict__ hi, bf16* __restrict__ lo, int N, int H,
                 float alpha0, const float* __restrict__ alpha_dev, float self_w, int accumulate) {
    pdl_prologue();
    extern __shared__ float s_adj[];  // [N][N], s_adj[i*N+j] = coefficient of x_j in out_i
    const int b = blockIdx.y;
    const float* ab = adj + (size_t)b * N * N;
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int i = e / N, j = e - i * N;
        s_adj[e] = TRANS ? ab[j * N + i] : ab[e];
    }
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= H) return;
    const float alpha = alpha0 + (alpha_dev ? alpha_dev[0] : 0.f);
    const float* xb = x + (size_t)b * N * H + c;
    const size_t ob = (size_t)b * N * H + c;
    for (int i0 = 0; i0 < N; i0 += RT) {
        float acc[RT];
#pragma unroll
        for (int i = 0; i < RT; ++i) acc[i] = 0.f;
        for (int j = 0; j < N; ++j) {
            const float xv = xb[(size_t)j * H];
#pragma unroll
            for (int i = 0; i < RT; ++i)
                if (i0 + i < N) acc[i] = fmaf(s_adj[(i0 + i) * N + j], xv, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            if (i0 + i >= N) break;
            float v = alpha * acc[i];
            if (self_w != 0.f) v = fmaf(self_w, xb[(size_t)(i0 + i) * H], v);
            const size_t o = ob + (size_t)(i0 + i) * H;
            if (out) {
                if (accumulate) v += out[o];
                out[o] = v;
            }
            if (hi) {
                bf16 h, l;
                split_bf16(v, h, l);
                hi[o] = h;
                if (lo) lo[o] = l;
            }
        }
    }
}

// obj36 fast path: N = 36 exactly.  grid (ceil(H/256), B), 128 threads, each thread owns two
// adjacent feature columns and keeps their 36 node values in registers (one coalesced 8-byte
// load per node row).  The coefficient matrix  C = alpha * (adj | adj^T) + self_w * I  is built
// once in shared memory, so the body is a plain C @ x: two output rows per iteration (four
// independent FMA chains), coefficients fetched as broadcast LDS.128 (8 FMAs per LDS).  The row
// loop is NOT unrolled: the body stays inside the instruction cache and ~100 registers allow five
// CTAs per SM.
template <bool TRANS>
__global__ void __launch_bounds__(128, 5)
adj_apply36_kernel(const float* __restrict__ adj, const float* __restrict__ x, float* __restrict__ out,
                   bf16* __restrict__ hi, bf16* __restrict__ lo, int H, float alpha0,
                   const float* __restrict__ alpha_dev, float self_w, int accumulate) {
    pdl_prologue();
    constexpr int N = 36;
    __shared__ __align__(16) float s_c[N * N];
    const int b = blockIdx.y;
    const float* ab = adj + (size_t)b * N * N;
    const float alpha = alpha0 + (alpha_dev ? alpha_dev[0] : 0.f);
    for (int e = threadIdx.x; e < N * N; e += 128) {
        const int i = e / N, j = e - i * N;
        float c = alpha * (TRANS ? ab[j * N + i] : ab[e]);
        if (i == j) c += self_w;
        s_c[e] = c;
    }
    __syncthreads();
    const int c = blockIdx.x * 256 + 2 * threadIdx.x;
    if (c >= H) return;
    const size_t base = (size_t)b * N * H + c;
    float2 xv[N];
#pragma unroll
    for (int j = 0; j < N; ++j) xv[j] = *reinterpret_cast<const float2*>(x + base + (size_t)j * H);
#pragma unroll 1
    for (int i = 0; i < N; i += 2) {
        float v[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        const float* r0 = s_c + i * N;
#pragma unroll
        for (int j4 = 0; j4 < N / 4; ++j4) {
            const float4 a0 = *reinterpret_cast<const float4*>(r0 + 4 * j4);
            const float4 a1 = *reinterpret_cast<const float4*>(r0 + N + 4 * j4);
            const float c0[4] = {a0.x, a0.y, a0.z, a0.w}, c1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                v[0][0] = fmaf(c0[t], xv[4 * j4 + t].x, v[0][0]);
                v[0][1] = fmaf(c0[t], xv[4 * j4 + t].y, v[0][1]);
                v[1][0] = fmaf(c1[t], xv[4 * j4 + t].x, v[1][0]);
                v[1][1] = fmaf(c1[t], xv[4 * j4 + t].y, v[1][1]);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const size_t o = base + (size_t)(i + r) * H;
            float vx = v[r][0], vy = v[r][1];
            if (out) {
                if (accumulate) {
                    const float2 p = *reinterpret_cast<const float2*>(out + o);
                    vx += p.x; vy += p.y;
                }
                *reinterpret_cast<float2*>(out + o) = make_float2(vx, vy);
            }
            if (hi) {
                bf16 hx, lx, hy, ly;
                split_bf16(vx, hx, lx);
                split_bf16(vy, hy, ly);
                *reinterpret_cast<__nv_bfloat162*>(hi + o) = __halves2bfloat162(hx, hy);
                if (lo) *reinterpret_cast<__nv_bfloat162*>(lo + o) = __halves2bfloat162(lx, ly);
            }
        }
    }
}

// out (fp32, optional) and / or hi+lo (bf16 planes, optional) receive
//   self_w * x + alpha * (adj | adj^T) @ x      (+ previous out when accumulate)
int adj_apply(const float* adj, const float* x, float* out, bf16* hi, bf16* lo, int B, int N, int H,
              float alpha0, const float* alpha_dev, float self_w, bool trans, int accumulate, cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(N >= 1 && N <= MAX_NODES && H >= 1 && (out || hi));
    auto al8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; };
    if (N == 36 && H % 2 == 0 && al8(x) && al8(out) && al8(hi) && al8(lo)) {
        dim3 grid(ceil_div(H, 256), B);
        if (trans)
            XGGM_LAUNCH((adj_apply36_kernel<true>), grid, 128, 0, st, adj, x, out, hi, lo, H, alpha0, alpha_dev, self_w, accumulate);
        else
            XGGM_LAUNCH((adj_apply36_kernel<false>), grid, 128, 0, st, adj, x, out, hi, lo, H, alpha0, alpha_dev, self_w, accumulate);
    } else {
        dim3 grid(ceil_div(H, 256), B);
        const size_t smem = sizeof(float) * N * N;
        if (trans)
            XGGM_LAUNCH((adj_apply_kernel<12, true>), grid, 256, smem, st, adj, x, out, hi, lo, N, H, alpha0, alpha_dev, self_w, accumulate);
        else
            XGGM_LAUNCH((adj_apply_kernel<12, false>), grid, 256, smem, st, adj, x, out, hi, lo, N, H, alpha0, alpha_dev, self_w, accumulate);
    }
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ---------------------------------------------------------------- adj_ln_fwd
// GCNConv tail with the message passing folded in (src/module/gcn.py:22-29, with W.(adj @ h) re-associated
// as adj @ (W.h), P = h W^T coming from the projection GEMM):
//     u = resid + adj_b @ P_b ;  h = LayerNorm(u)          (also xhat, rstd and the bf16 planes of h)
// Each graph is staged whole in shared memory: persistent CTAs (one per SM) take one graph at a time; thread 0
// streams the graph's [N,H] P tile in with bulk async copies (six row groups, one mbarrier each) into one of
// TWO tile buffers, so the tile of graph g+1 arrives under the math of graph g; the residual rows are pulled
// into L2 by a bulk prefetch at the start of the graph and read with 16-byte loads after the accumulation.  Work split: a warp owns ALN_RW complete node rows, a lane
// owns the columns 128 q + 4 lane + {0..3} (the row kernels' layout), so the accumulation
//     acc[r][c] += adj[i_r][j] * P[j][c]      (j = 0..N-1: one real loop, 3 broadcast + NV vector LDS, 12 NV FMA)
// needs no register indexing by row, the LayerNorm statistics are plain warp reductions, and every global
// store is a 16-byte (fp32) or 8-byte (bf16 plane) vector.
constexpr int ALN_THREADS = 384;
constexpr int ALN_WARPS = ALN_THREADS / 32;
constexpr int ALN_RW = 3;                          // node rows per warp
constexpr int ALN_MAXN = ALN_WARPS * ALN_RW;       // 36 = obj36
constexpr int ALN_JU = 6;                          // P rows per copy group / mbarrier
constexpr int ALN_GROUPS = ALN_MAXN / ALN_JU;
template <int NV>
__global__ void __launch_bounds__(ALN_THREADS, 1)
adj_ln_fwd_kernel(const float* __restrict__ adj, const float* __restrict__ P, const float* __restrict__ resid,
                  const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ h,
                  float* __restrict__ xhat, float* __restrict__ rstd_out, bf16* __restrict__ hi,
                  bf16* __restrict__ lo, int B, int N, float eps) {
    pdl_prologue();
    constexpr int H = NV * 128;
    extern __shared__ __align__(16) float aln_sm[];
    const int tile = N * H;                                    // floats per staged tile
    float* Ps = aln_sm;                                        // [2][tile]: double-buffered P tiles
    float* adjS = Ps + 2 * (size_t)tile;                       // [2][ALN_MAXN * ALN_MAXN]: adjacency, row stride N
    uint64_t* bar = reinterpret_cast<uint64_t*>(adjS + 2 * ALN_MAXN * ALN_MAXN);   // [2][GROUPS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int groups = (N + ALN_JU - 1) / ALN_JU;
    // the N x N adjacency rides with the first P row group when a bulk copy can address it
    const bool adj_bulk = ((N * N * 4) % 16 == 0) && ((reinterpret_cast<uintptr_t>(adj) & 15) == 0);
    if (tid == 0) {
        for (int g = 0; g < 2 * ALN_GROUPS; ++g) ptx::mbar_init(&bar[g], 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    auto issue_p = [&](int buf, int b) {      // thread 0 only
        ptx::fence_proxy_async();             // the buffer was last read through the generic proxy
        for (int g = 0; g < groups; ++g) {
            const int r0 = g * ALN_JU, rows = min(ALN_JU, N - r0);
            const uint32_t bytes = (uint32_t)(rows * H * 4);
            const uint32_t abytes = (g == 0 && adj_bulk) ? (uint32_t)(N * N * 4) : 0u;
            ptx::mbar_expect_tx(&bar[buf * ALN_GROUPS + g], bytes + abytes);
            if (abytes)
                ptx::bulk_g2s(adjS + buf * ALN_MAXN * ALN_MAXN, adj + (size_t)b * N * N, abytes, &bar[buf * ALN_GROUPS]);
            ptx::bulk_g2s(Ps + (size_t)buf * tile + (size_t)r0 * H, P + ((size_t)b * N + r0) * H, bytes,
                          &bar[buf * ALN_GROUPS + g]);
        }
    };
    if (tid == 0 && (int)blockIdx.x < B) issue_p(0, blockIdx.x);
    const int i0 = warp * ALN_RW;                              // this warp's rows: i0 .. i0 + ALN_RW - 1
    const float inv_h = 1.0f / (float)H;
    int it = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t phase = (it >> 1) & 1;
        const int next_b = b + gridDim.x;
        if (tid == 0) {
            // the other buffer was drained one graph ago: the next graph's P streams in under this graph's math;
            // this graph's residual tile is pulled into L2 so that the reads after the accumulation hit there
            if (next_b < B) issue_p(buf ^ 1, next_b);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(resid + (size_t)b * tile), "r"(tile * 4) : "memory");
        }
        float* adjB = adjS + buf * ALN_MAXN * ALN_MAXN;
        if (!adj_bulk) {
            const float* ab = adj + (size_t)b * N * N;
            for (int e = tid; e < N * N; e += ALN_THREADS) adjB[e] = ab[e];
            __syncthreads();
        }
        const float* Pb = Ps + (size_t)buf * tile;
        float acc[ALN_RW][NV * 4];
#pragma unroll
        for (int r = 0; r < ALN_RW; ++r)
#pragma unroll
            for (int c = 0; c < NV * 4; ++c) acc[r][c] = 0.f;
        for (int g = 0; g < groups; ++g) {
            ptx::mbar_wait(&bar[buf * ALN_GROUPS + g], phase);
            const int j1 = min(N, (g + 1) * ALN_JU);
#pragma unroll 2
            for (int j = g * ALN_JU; j < j1; ++j) {
                float a[ALN_RW];
#pragma unroll
                for (int r = 0; r < ALN_RW; ++r) a[r] = adjB[min(i0 + r, N - 1) * N + j];   // (rows >= N: idle warps)
                const float* pj = Pb + (size_t)j * H + 4 * lane;
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const float4 p = *reinterpret_cast<const float4*>(pj + 128 * q);
#pragma unroll
                    for (int r = 0; r < ALN_RW; ++r) {
                        acc[r][4 * q] = fmaf(a[r], p.x, acc[r][4 * q]);
                        acc[r][4 * q + 1] = fmaf(a[r], p.y, acc[r][4 * q + 1]);
                        acc[r][4 * q + 2] = fmaf(a[r], p.z, acc[r][4 * q + 2]);
                        acc[r][4 * q + 3] = fmaf(a[r], p.w, acc[r][4 * q + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ALN_RW; ++r) {
            const int i = i0 + r;
            if (i < N) {                                   // warp-uniform
                const size_t ro = ((size_t)b * N + i) * H + 4 * lane;
                float4 t[NV];
#pragma unroll
                for (int q = 0; q < NV; ++q) t[q] = *reinterpret_cast<const float4*>(resid + ro + 128 * q);
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    acc[r][4 * q] += t[q].x; acc[r][4 * q + 1] += t[q].y; acc[r][4 * q + 2] += t[q].z; acc[r][4 * q + 3] += t[q].w;
                    s += (acc[r][4 * q] + acc[r][4 * q + 1]) + (acc[r][4 * q + 2] + acc[r][4 * q + 3]);
                }
                const float mean = warp_sum(s) * inv_h;
                float qs = 0.f;
#pragma unroll
                for (int c = 0; c < NV * 4; ++c) {
                    acc[r][c] -= mean;
                    qs = fmaf(acc[r][c], acc[r][c], qs);
                }
                const float rstd = 1.0f / sqrtf(warp_sum(qs) * inv_h + eps);
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const float4 gq = *reinterpret_cast<const float4*>(gamma + 128 * q + 4 * lane);
                    const float4 bq = *reinterpret_cast<const float4*>(beta + 128 * q + 4 * lane);
                    const float x0 = acc[r][4 * q] * rstd, x1 = acc[r][4 * q + 1] * rstd;
                    const float x2 = acc[r][4 * q + 2] * rstd, x3 = acc[r][4 * q + 3] * rstd;
                    const float4 hv = make_float4(fmaf(x0, gq.x, bq.x), fmaf(x1, gq.y, bq.y), fmaf(x2, gq.z, bq.z),
                                                  fmaf(x3, gq.w, bq.w));
                    *reinterpret_cast<float4*>(h + ro + 128 * q) = hv;
                    if (xhat) *reinterpret_cast<float4*>(xhat + ro + 128 * q) = make_float4(x0, x1, x2, x3);
                    if (hi) split_store4(hi, lo, ro + 128 * q, hv.x, hv.y, hv.z, hv.w);
                }
                if (lane == 0 && rstd_out) rstd_out[(size_t)b * N + i] = rstd;
            }
        }
        __syncthreads();   // every warp is done with this P / adjacency buffer before it is re-armed
    }
}
static size_t aln_smem(int N, int H) {
    return sizeof(float) * (2 * (size_t)N * H + 2 * ALN_MAXN * ALN_MAXN) + sizeof(uint64_t) * (2 * ALN_GROUPS);
}
// obj36-sized graphs whose two fp32 tiles fit in shared memory (N = 36, H = 768: 221 KB); other shapes take the
// unfused kernels
bool adj_ln_supported(int N, int H) {
    return N >= 1 && N <= ALN_MAXN && H % 128 == 0 && H >= 128 && H <= 768 && aln_smem(N, H) <= 227 * 1024;
}

// h = LN(resid + adj @ P) per graph; hi/lo (optional): bf16 planes of h
int adj_ln_fwd(const float* adj, const float* P, const float* resid, const float* gamma, const float* beta, float* h,
               float* xhat, float* rstd, bf16* hi, bf16* lo, int B, int N, int H, float eps, cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    XGGM_REQUIRE(adj && P && resid && gamma && beta && h && adj_ln_supported(N, H));
    XGGM_REQUIRE(al16(P) && al16(resid) && al16(gamma) && al16(beta) && al16(h) && al16(xhat) && al16(hi) && al16(lo));
    const size_t smem = aln_smem(N, H);
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = min(B, sms);
#define XGGM_ALN_LAUNCH(NVV)                                                                                         \
    case NVV: {                                                                                                      \
        static bool attr = false;                                                                                    \
        if (!attr) {                                                                                                 \
            XGGM_CUDA_TRY(cudaFuncSetAttribute(adj_ln_fwd_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                               227 * 1024));                                                         \
            attr = true;                                                                                             \
        }                                                                                                            \
        XGGM_LAUNCH((adj_ln_fwd_kernel<NVV>), grid, ALN_THREADS, smem, st, adj, P, resid, gamma, beta, h, xhat, rstd, hi, \
                    lo, B, N, eps);                                                                                  \
    } break;
    switch (H / 128) {
        XGGM_ALN_LAUNCH(1)
        XGGM_ALN_LAUNCH(2)
        XGGM_ALN_LAUNCH(3)
        XGGM_ALN_LAUNCH(4)
        XGGM_ALN_LAUNCH(5)
        XGGM_ALN_LAUNCH(6)
        default: return XGGM_ERR_ARG;
    }
#undef XGGM_ALN_LAUNCH
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// -------------------------------------------------------------------- bmm_nt
// One CTA per graph.  Feature columns are streamed through smem in chunks of CK;
// each thread owns up to MAXB 3x3 blocks of the N x N result.  The k-order of every
// dot product is identical for (i,j) and (j,i), so p == q gives a bitwise-symmetric S.
constexpr int CK = 64;
constexpr int MAXB = 5;

struct PairTile {
    float acc[MAXB][9];
};

__device__ __forceinline__ void pair_scores(const float* __restrict__ pb, const float* __restrict__ qb,
                                            int N, int H, float* ps, float* qs, PairTile& tile) {
    const int nb = (N + 2) / 3, nblk = nb * nb;
    const bool same = (pb == qb);
#pragma unroll
    for (int u = 0; u < MAXB; ++u)
#pragma unroll
        for (int e = 0; e < 9; ++e) tile.acc[u][e] = 0.f;
    for (int c0 = 0; c0 < H; c0 += CK) {
        const int cw = min(CK, H - c0);
        for (int e = threadIdx.x; e < N * CK; e += blockDim.x) {
            const int i = e / CK, c = e - i * CK;
            const float pv = c < cw ? pb[(size_t)i * H + c0 + c] : 0.f;
            ps[i * (CK + 1) + c] = pv;
            if (!same) qs[i * (CK + 1) + c] = c < cw ? qb[(size_t)i * H + c0 + c] : 0.f;
        }
        __syncthreads();
        const float* qsrc = same ? ps : qs;
#pragma unroll
        for (int u = 0; u < MAXB; ++u) {
            const int blk = threadIdx.x + u * blockDim.x;
            if (blk < nblk) {
                const int ib = blk / nb, jb = blk - ib * nb;
                const int i0 = min(ib * 3, N - 1), i1 = min(ib * 3 + 1, N - 1), i2 = min(ib * 3 + 2, N - 1);
                const int j0 = min(jb * 3, N - 1), j1 = min(jb * 3 + 1, N - 1), j2 = min(jb * 3 + 2, N - 1);
                const float* p0 = ps + i0 * (CK + 1); const float* p1 = ps + i1 * (CK + 1); const float* p2 = ps + i2 * (CK + 1);
                const float* q0 = qsrc + j0 * (CK + 1); const float* q1 = qsrc + j1 * (CK + 1); const float* q2 = qsrc + j2 * (CK + 1);
                float* a = tile.acc[u];
#pragma unroll 8
                for (int c = 0; c < CK; ++c) {
                    const float x0 = p0[c], x1 = p1[c], x2 = p2[c];
                    const float y0 = q0[c], y1 = q1[c], y2 = q2[c];
                    a[0] = fmaf(x0, y0, a[0]); a[1] = fmaf(x0, y1, a[1]); a[2] = fmaf(x0, y2, a[2]);
                    a[3] = fmaf(x1, y0, a[3]); a[4] = fmaf(x1, y1, a[4]); a[5] = fmaf(x1, y2, a[5]);
                    a[6] = fmaf(x2, y0, a[6]); a[7] = fmaf(x2, y1, a[7]); a[8] = fmaf(x2, y2, a[8]);
                }
            }
        }
        __syncthreads();
    }
}

// scatter the register tiles into an [N][N] array (smem or global)
__device__ __forceinline__ void pair_store(const PairTile& tile, float* dst, int N) {
    const int nb = (N + 2) / 3, nblk = nb * nb;
#pragma unroll
    for (int u = 0; u < MAXB; ++u) {
        const int blk = threadIdx.x + u * blockDim.x;
        if (blk < nblk) {
            const int ib = blk / nb, jb = blk - ib * nb;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int i = ib * 3 + r, j = jb * 3 + c;
                    if (i < N && j < N) dst[i * N + j] = tile.acc[u][r * 3 + c];
                }
        }
    }
}

// out[b] (=|+=) alpha * p[b] q[b]^T ;  dot_out? += sum_b <p[b] q[b]^T, dot_ref[b]>
__global__ void __launch_bounds__(256)
bmm_nt_kernel(const float* __restrict__ p, const float* __restrict__ q, float* __restrict__ out,
              int N, int H, float alpha0, const float* __restrict__ alpha_dev, int accumulate,
              const float* __restrict__ dot_ref, float* __restrict__ dot_out) {
    pdl_prologue();
    extern __shared__ float sm[];
    __shared__ float part[8];
    float* ps = sm;
    float* qs = sm + N * (CK + 1);
    const int b = blockIdx.x;
    PairTile tile;
    pair_scores(p + (size_t)b * N * H, q + (size_t)b * N * H, N, H, ps, qs, tile);
    const float alpha = alpha0 + (alpha_dev ? alpha_dev[0] : 0.f);
    float* ob = out + (size_t)b * N * N;
    const float* rb = dot_ref ? dot_ref + (size_t)b * N * N : nullptr;
    const int nb = (N + 2) / 3, nblk = nb * nb;
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < MAXB; ++u) {
        const int blk = threadIdx.x + u * blockDim.x;
        if (blk < nblk) {
            const int ib = blk / nb, jb = blk - ib * nb;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int i = ib * 3 + r, j = jb * 3 + c;
                    if (i < N && j < N) {
                        const float v = tile.acc[u][r * 3 + c];
                        if (rb) dot = fmaf(v, rb[i * N + j], dot);
                        const float o = alpha * v;
                        ob[i * N + j] = accumulate ? ob[i * N + j] + o : o;
                    }
                }
        }
    }
    if (dot_out) {
        dot = warp_sum(dot);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = dot;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
            v = warp_sum(v);
            if (threadIdx.x == 0) atomicAdd(dot_out, v);
        }
    }
}

int bmm_nt(const float* p, const float* q, float* out, int B, int N, int H, float alpha0,
           const float* alpha_dev, int accumulate, const float* dot_ref, float* dot_out,
           cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(N >= 1 && N <= MAX_NODES);
    const size_t smem = sizeof(float) * 2 * N * (CK + 1);
    if (smem > 48 * 1024)
        XGGM_CUDA_TRY(cudaFuncSetAttribute(bmm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XGGM_LAUNCH((bmm_nt_kernel), B, 256, smem, st, p, q, out, N, H, alpha0, alpha_dev, accumulate, dot_ref, dot_out);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ----------------------------------------------------------------- adj_regen
__global__ void __launch_bounds__(256)
adj_regen_fwd_kernel(const float* __restrict__ x, float* __restrict__ adj_out,
                     float* __restrict__ S_out, int32_t* __restrict__ amax_out, int N, int H,
                     int squash) {
    pdl_prologue();
    extern __shared__ float sm[];
    float* ps = sm;                       // [N][CK+1]
    float* S = sm + N * (CK + 1);         // [N][N]
    float* m = S + N * N;                 // [N]
    const int b = blockIdx.x;
    PairTile tile;
    pair_scores(x + (size_t)b * N * H, x + (size_t)b * N * H, N, H, ps, ps, tile);
    pair_store(tile, S, N);
    __syncthreads();
    // column max, first index on ties (torch.max(dim=1) semantics), ggm.py:226
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float best = S[i];
        int arg = 0;
        for (int k = 1; k < N; ++k) {
            const float v = S[k * N + i];
            if (v > best) { best = v; arg = k; }
        }
        m[i] = best;
        if (amax_out) amax_out[(size_t)b * N + i] = arg;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int i = e / N, j = e - i * N;
        const float s = S[e];
        if (S_out) S_out[(size_t)b * N * N + e] = s;
        float v = s / m[i];
        if (squash) v = sigmoidf_(v);
        adj_out[(size_t)b * N * N + e] = (i == j) ? 0.f : v;
    }
}

int adj_regen_fwd(const float* x, float* adj_out, float* S, int32_t* amax, int B, int N, int H,
                  int squash, cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(N >= 1 && N <= MAX_NODES);
    const size_t smem = sizeof(float) * ((size_t)N * (CK + 1) + (size_t)N * N + N);
    if (smem > 48 * 1024)
        XGGM_CUDA_TRY(cudaFuncSetAttribute(adj_regen_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XGGM_LAUNCH((adj_regen_fwd_kernel), B, 256, smem, st, x, adj_out, S, amax, N, H, squash);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// Second half of adj_regen_fwd when S = x x^T was produced by the tensor-core Gram kernel:
// column max / arg-max (first index on ties), divide, sigmoid, zero diagonal.  One CTA per graph.
__global__ void __launch_bounds__(128)
regen_from_s_kernel(const float* __restrict__ S_in, float* __restrict__ adj_out,
                    int32_t* __restrict__ amax_out, int N, int squash) {
    pdl_prologue();
    extern __shared__ float sm[];
    float* S = sm;          // [N][N]
    float* m = sm + N * N;  // [N]
    const int b = blockIdx.x;
    const float* Sb = S_in + (size_t)b * N * N;
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) S[e] = Sb[e];
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float best = S[i];
        int arg = 0;
        for (int k = 1; k < N; ++k) {
            const float v = S[k * N + i];
            if (v > best) { best = v; arg = k; }
        }
        m[i] = best;
        if (amax_out) amax_out[(size_t)b * N + i] = arg;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int i = e / N, j = e - i * N;
        float v = S[e] / m[i];
        if (squash) v = sigmoidf_(v);
        adj_out[(size_t)b * N * N + e] = (i == j) ? 0.f : v;
    }
}

int adj_regen_from_s(const float* S, float* adj_out, int32_t* amax, int B, int N, int squash, cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(N >= 1 && N <= 128);
    const size_t smem = sizeof(float) * ((size_t)N * N + N);
    if (smem > 48 * 1024)
        XGGM_CUDA_TRY(cudaFuncSetAttribute(regen_from_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XGGM_LAUNCH((regen_from_s_kernel), B, 128, smem, st, S, adj_out, amax, N, squash);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// out (=|+=) alpha * S ;  dot_out? += <S, dot_ref>   (tail of bmm_nt when S came from the Gram kernel)
__global__ void __launch_bounds__(256)
scale_accum_kernel(const float* __restrict__ S, float* __restrict__ out, long long n, float alpha0,
                   const float* __restrict__ alpha_dev, int accumulate, const float* __restrict__ dot_ref,
                   float* __restrict__ dot_out) {
    pdl_prologue();
    __shared__ float part[8];
    const float alpha = alpha0 + (alpha_dev ? alpha_dev[0] : 0.f);
    float dot = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = S[i];
        if (dot_ref) dot = fmaf(v, dot_ref[i], dot);
        const float o = alpha * v;
        out[i] = accumulate ? out[i] + o : o;
    }
    if (dot_out) {
        dot = warp_sum(dot);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = dot;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
            v = warp_sum(v);
            if (threadIdx.x == 0) atomicAdd(dot_out, v);
        }
    }
}

int scale_accum(const float* S, float* out, long long n, float alpha0, const float* alpha_dev, int accumulate,
                const float* dot_ref, float* dot_out, cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    const int grid = (int)min((long long)296, (n + 255) / 256);
    XGGM_LAUNCH((scale_accum_kernel), grid, 256, 0, st, S, out, n, alpha0, alpha_dev, accumulate, dot_ref, dot_out);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// d(adj_out)/dS folded into a symmetric coefficient matrix D = dS + dS^T so that
// gx += D x.  One CTA per graph.
__global__ void __launch_bounds__(256)
adj_regen_bwd_kernel(const float* __restrict__ gadj, const float* __restrict__ S_in,
                     const int32_t* __restrict__ amax, float* __restrict__ D_out, int N,
                     int squash) {
    pdl_prologue();
    extern __shared__ float sm[];
    float* S = sm;             // [N][N]
    float* dS = S + N * N;     // [N][N]
    float* m = dS + N * N;     // [N]
    float* dm = m + N;         // [N]
    const int b = blockIdx.x;
    const float* Sb = S_in + (size_t)b * N * N;
    const float* gb = gadj + (size_t)b * N * N;
    const int32_t* ab = amax + (size_t)b * N;
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) S[e] = Sb[e];
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) m[i] = S[ab[i] * N + i];
    __syncthreads();
    // dt and dS (without the max path)
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int i = e / N, j = e - i * N;
        float dt = 0.f;
        if (i != j) {
            dt = gb[e];
            if (squash) {
                const float y = sigmoidf_(S[e] / m[i]);
                dt *= y * (1.f - y);
            }
        }
        dS[e] = dt / m[i];
    }
    __syncthreads();
    // dm_i = - sum_j dt_ij S_ij / m_i^2 = - sum_j dS_ij * S_ij / m_i   (one warp per row)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = warp; i < N; i += nwarp) {
        float s = 0.f;
        for (int j = lane; j < N; j += 32) s = fmaf(dS[i * N + j], S[i * N + j], s);
        s = warp_sum(s);
        if (lane == 0) dm[i] = -s / m[i];
    }
    __syncthreads();
    // route dm_i to the arg-max entry of column i (distinct columns -> no write conflicts)
    for (int i = threadIdx.x; i < N; i += blockDim.x) dS[ab[i] * N + i] += dm[i];
    __syncthreads();
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int i = e / N, j = e - i * N;
        D_out[(size_t)b * N * N + e] = dS[e] + dS[j * N + i];
    }
}

// D[b] = dS + dS^T (the symmetric coefficient matrix of gx (+)= D x)
int adj_regen_bwd_coeffs(const float* gadj, const float* S, const int32_t* amax, float* D, int B, int N, int squash,
                         cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(N >= 1 && N <= MAX_NODES);
    const size_t smem = sizeof(float) * (2 * (size_t)N * N + 2 * N);
    if (smem > 48 * 1024)
        XGGM_CUDA_TRY(cudaFuncSetAttribute(adj_regen_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XGGM_LAUNCH((adj_regen_bwd_kernel), B, 256, smem, st, gadj, S, amax, D, N, squash);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

int adj_regen_bwd(const float* gadj, const float* x, const float* S, const int32_t* amax,
                  float* gx, float* work, int B, int N, int H, int squash, int accumulate_gx,
                  cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_TRY(adj_regen_bwd_coeffs(gadj, S, amax, work, B, N, squash, st));
    return adj_apply(work, x, gx, nullptr, nullptr, B, N, H, 1.f, nullptr, 0.f, false, accumulate_gx, st);
}

// ------------------------------------------------------------------ GAT attn
// scores + mask + row softmax: one CTA per graph
__global__ void __launch_bounds__(256)
gat_scores_kernel(const float* __restrict__ h, const float* __restrict__ a,
                  const float* __restrict__ adj, float* __restrict__ att, int N, int H,
                  float slope) {
    pdl_prologue();
    extern __shared__ float sm[];
    float* s1 = sm;        // [N]  a[:H] . h_i
    float* s2 = sm + N;    // [N]  a[H:] . h_j
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const float* hb = h + (size_t)b * N * H;
    for (int i = warp; i < N; i += nwarp) {
        float u = 0.f, v = 0.f;
        for (int c = lane; c < H; c += 32) {
            const float hv = hb[(size_t)i * H + c];
            u = fmaf(hv, a[c], u);
            v = fmaf(hv, a[H + c], v);
        }
        u = warp_sum(u); v = warp_sum(v);
        if (lane == 0) { s1[i] = u; s2[i] = v; }
    }
    __syncthreads();
    const float* ab = adj + (size_t)b * N * N;
    float* ob = att + (size_t)b * N * N;
    for (int i = warp; i < N; i += nwarp) {
        float mx = -INFINITY;
        for (int j = lane; j < N; j += 32) {
            float e = s1[i] + s2[j];
            e = e > 0.f ? e : slope * e;
            if (ab[i * N + j] == 0.f) e = -9e15f;
            mx = fmaxf(mx, e);
        }
        mx = warp_max(mx);
        float den = 0.f;
        for (int j = lane; j < N; j += 32) {
            float e = s1[i] + s2[j];
            e = e > 0.f ? e : slope * e;
            if (ab[i * N + j] == 0.f) e = -9e15f;
            den += expf(e - mx);
        }
        den = warp_sum(den);
        for (int j = lane; j < N; j += 32) {
            float e = s1[i] + s2[j];
            e = e > 0.f ? e : slope * e;
            if (ab[i * N + j] == 0.f) e = -9e15f;
            ob[i * N + j] = expf(e - mx) / den;
        }
    }
}

__global__ void elu_fwd_kernel(const float* __restrict__ pre, float* __restrict__ out, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = pre[i]; out[i] = v > 0.f ? v : expm1f(v); }
}
__global__ void elu_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ pre,
                               float* __restrict__ gpre, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = pre[i]; gpre[i] = gout[i] * (v > 0.f ? 1.f : expf(v)); }
}

int gat_attn_fwd(const float* h, const float* a, const float* adj, float* out, float* att,
                 float* pre, int B, int N, int H, float slope, int apply_elu, cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(N >= 1 && N <= MAX_NODES);
    XGGM_LAUNCH((gat_scores_kernel), B, 256, sizeof(float) * 2 * N, st, h, a, adj, att, N, H, slope);
    XGGM_LAUNCH_CHECK();
    if (!apply_elu) return adj_apply(att, h, out, nullptr, nullptr, B, N, H, 1.f, nullptr, 0.f, false, 0, st);
    XGGM_TRY(adj_apply(att, h, pre, nullptr, nullptr, B, N, H, 1.f, nullptr, 0.f, false, 0, st));
    const long long n = (long long)B * N * H;
    XGGM_LAUNCH((elu_fwd_kernel), ceil_div(n, 256), 256, 0, st, pre, out, n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// softmax / LeakyReLU / mask backward + the rank-1 score terms.  One CTA per graph.
// In: gatt[N][N] = gpre @ h^T.  Out: gh_i += gs1_i a1 + gs2_i a2 ; ga += [sum gs1_i h_i | sum gs2_j h_j]
__global__ void __launch_bounds__(256)
gat_scores_bwd_kernel(const float* __restrict__ h, const float* __restrict__ a,
                      const float* __restrict__ adj, const float* __restrict__ att,
                      const float* __restrict__ gatt, float* __restrict__ gh,
                      float* __restrict__ ga, int N, int H, float slope) {
    pdl_prologue();
    extern __shared__ float sm[];
    float* s1 = sm;            // [N]
    float* s2 = s1 + N;        // [N]
    float* g1 = s2 + N;        // [N]  gs1
    float* g2 = g1 + N;        // [N]  gs2
    float* ge = g2 + N;        // [N][N] gradient w.r.t. the pre-LeakyReLU score
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const float* hb = h + (size_t)b * N * H;
    for (int i = warp; i < N; i += nwarp) {
        float u = 0.f, v = 0.f;
        for (int c = lane; c < H; c += 32) {
            const float hv = hb[(size_t)i * H + c];
            u = fmaf(hv, a[c], u);
            v = fmaf(hv, a[H + c], v);
        }
        u = warp_sum(u); v = warp_sum(v);
        if (lane == 0) { s1[i] = u; s2[i] = v; }
    }
    __syncthreads();
    const float* ab = adj + (size_t)b * N * N;
    const float* pb = att + (size_t)b * N * N;
    const float* gb = gatt + (size_t)b * N * N;
    for (int i = warp; i < N; i += nwarp) {
        float dot = 0.f;
        for (int j = lane; j < N; j += 32) dot = fmaf(pb[i * N + j], gb[i * N + j], dot);
        dot = warp_sum(dot);
        float rs = 0.f;
        for (int j = lane; j < N; j += 32) {
            float g = pb[i * N + j] * (gb[i * N + j] - dot);
            if (ab[i * N + j] == 0.f) g = 0.f;           // masked_fill blocks the gradient
            else g *= (s1[i] + s2[j]) > 0.f ? 1.f : slope;
            ge[i * N + j] = g;
            rs += g;
        }
        rs = warp_sum(rs);
        if (lane == 0) g1[i] = rs;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float cs = 0.f;
        for (int i = 0; i < N; ++i) cs += ge[i * N + j];
        g2[j] = cs;
    }
    __syncthreads();
    float* ghb = gh + (size_t)b * N * H;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        const float a1 = a[c], a2 = a[H + c];
        float d1 = 0.f, d2 = 0.f;
        for (int i = 0; i < N; ++i) {
            const float hv = hb[(size_t)i * H + c];
            ghb[(size_t)i * H + c] += g1[i] * a1 + g2[i] * a2;
            d1 = fmaf(g1[i], hv, d1);
            d2 = fmaf(g2[i], hv, d2);
        }
        atomicAdd(&ga[c], d1);
        atomicAdd(&ga[H + c], d2);
    }
}

// work: [B,N,H] (gpre) + [B,N,N] (gatt)
int gat_attn_bwd(const float* gout, const float* h, const float* a, const float* adj,
                 const float* att, const float* pre, float* gh, float* ga, float* work, int B,
                 int N, int H, float slope, int apply_elu, cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(N >= 1 && N <= MAX_NODES);
    const long long n = (long long)B * N * H;
    const float* gpre = gout;
    float* gatt = work + n;
    if (apply_elu) {
        XGGM_LAUNCH((elu_bwd_kernel), ceil_div(n, 256), 256, 0, st, gout, pre, work, n);
        XGGM_LAUNCH_CHECK();
        gpre = work;
    }
    XGGM_TRY(adj_apply(att, gpre, gh, nullptr, nullptr, B, N, H, 1.f, nullptr, 0.f, true, 0, st));   // gh = att^T gpre
    XGGM_TRY(bmm_nt(gpre, h, gatt, B, N, H, 1.f, nullptr, 0, nullptr, nullptr, st));                                   // gatt = gpre h^T
    const size_t smem = sizeof(float) * (4 * (size_t)N + (size_t)N * N);
    XGGM_LAUNCH((gat_scores_bwd_kernel), B, 256, smem, st, h, a, adj, att, gatt, gh, ga, N, H, slope);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

}  // namespace xggm
