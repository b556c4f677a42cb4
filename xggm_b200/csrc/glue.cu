// Trainer-glue kernels of the GGM step: diagonal strip, upper-triangle scatter,
// Gaussian perturbation + score target, score-matching MSE, symmetric row-softmax KL,
// fusion read-out, sigmoid, Philox keep-masks.  All are one-pass, coalesced,
// HBM-bound element/row kernels.
#include "common.cuh"
#include "kernels.cuh"

namespace xggm {

static inline int grid1d(long long n, int block = 256) { return ceil_div(n, block); }

// ----------------------------------------------------------------- strip_diag
__global__ void strip_diag_kernel(const float* __restrict__ a, float* __restrict__ out,
                                  long long total, int N) {
    pdl_prologue();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int r = (int)(e % ((long long)N * N));
    out[e] = (r / N == r % N) ? 0.f : a[e];
}
int strip_diag(const float* a, float* out, int B, int N, cudaStream_t st) {
    const long long total = (long long)B * N * N;
    if (total <= 0) return XGGM_OK;
    XGGM_LAUNCH((strip_diag_kernel), grid1d(total), 256, 0, st, a, out, total, N);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// --------------------------------------------------------------- triu scatter
// k(i<j) = i(2N-i-1)/2 + (j-i-1): row-major order of the boolean-mask assignment.
__device__ __forceinline__ int triu_index(int i, int j, int N) { return i * (2 * N - i - 1) / 2 + (j - i - 1); }

__global__ void triu_scatter_fwd_kernel(const float* __restrict__ v, float* __restrict__ adj,
                                        long long total, int N) {
    pdl_prologue();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int NN = N * N, E = N * (N - 1) / 2;
    const long long b = e / NN;
    const int r = (int)(e - b * NN), i = r / N, j = r - i * N;
    adj[e] = (i == j) ? 0.f : v[b * E + triu_index(min(i, j), max(i, j), N)];
}
__global__ void triu_scatter_bwd_kernel(const float* __restrict__ gadj, float* __restrict__ gv,
                                        long long total, int N) {
    pdl_prologue();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int NN = N * N, E = N * (N - 1) / 2;
    const long long b = e / NN;
    const int r = (int)(e - b * NN), i = r / N, j = r - i * N;
    if (i < j) gv[b * E + triu_index(i, j, N)] = gadj[e] + gadj[b * NN + j * N + i];
}
int triu_scatter_fwd(const float* v, float* adj, int B, int N, cudaStream_t st) {
    const long long total = (long long)B * N * N;
    if (total <= 0) return XGGM_OK;
    XGGM_LAUNCH((triu_scatter_fwd_kernel), grid1d(total), 256, 0, st, v, adj, total, N);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
int triu_scatter_bwd(const float* gadj, float* gv, int B, int N, cudaStream_t st) {
    const long long total = (long long)B * N * N;
    if (total <= 0) return XGGM_OK;
    XGGM_LAUNCH((triu_scatter_bwd_kernel), grid1d(total), 256, 0, st, gadj, gv, total, N);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// --------------------------------------------------------------------- noise
// n = randn.triu(1)*sigma ; n += n^T ; target = -n/sigma^2 ; noisy = adj + n
__global__ void edge_noise_kernel(const float* __restrict__ adj, const float* __restrict__ randn,
                                  float sigma, float sigma2, float* __restrict__ noisy,
                                  float* __restrict__ target, long long total, int N) {
    pdl_prologue();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int NN = N * N;
    const long long b = e / NN;
    const int r = (int)(e - b * NN), i = r / N, j = r - i * N;
    // (triu(1)*sigma) + transpose: exactly one of the two addends is non-zero off the diagonal
    const float up = (i < j) ? randn[e] * sigma : 0.f;
    const float lo = (j < i) ? randn[b * NN + j * N + i] * sigma : 0.f;
    const float n = up + lo;
    target[e] = -n / sigma2;
    noisy[e] = adj[e] + n;
}
// sigma2 = (float)(sigma*sigma) evaluated in double by the caller, as Python's `sigma ** 2`
int edge_noise(const float* adj, const float* randn, float sigma, float sigma2, float* noisy,
               float* target, int B, int N, cudaStream_t st) {
    const long long total = (long long)B * N * N;
    if (total <= 0) return XGGM_OK;
    XGGM_LAUNCH((edge_noise_kernel), grid1d(total), 256, 0, st, adj, randn, sigma, sigma2, noisy, target, total, N);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

__global__ void feat_noise_kernel(const float* __restrict__ f, const float* __restrict__ randn,
                                  float sigma, float sigma2, float* __restrict__ noisy,
                                  float* __restrict__ target, long long total, int N, int H,
                                  int bcast) {
    pdl_prologue();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const float n = randn[e] * sigma;
    float fv;
    if (bcast) {
        const long long b = e / ((long long)N * H);
        fv = f[b * H + (e % H)];
    } else {
        fv = f[e];
    }
    noisy[e] = fv + n;
    target[e] = -n / sigma2;
}
// float4 path: grid (chunks of one graph, B); no 64-bit division per element
__global__ void __launch_bounds__(256)
feat_noise_vec_kernel(const float* __restrict__ f, const float* __restrict__ randn, float sigma, float sigma2,
                      float* __restrict__ noisy, float* __restrict__ target, __nv_bfloat16* __restrict__ hi,
                      __nv_bfloat16* __restrict__ lo, int NH4, int H, int bcast) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // float4 index inside graph b
    if (i >= NH4) return;
    const int b = blockIdx.y;
    const size_t o = (size_t)b * NH4 + i;
    const float4 r = reinterpret_cast<const float4*>(randn)[o];
    const float4 fv = bcast ? *reinterpret_cast<const float4*>(f + (size_t)b * H + (4 * i) % H)
                            : reinterpret_cast<const float4*>(f)[o];
    const float4 n = make_float4(r.x * sigma, r.y * sigma, r.z * sigma, r.w * sigma);
    const float4 out = make_float4(fv.x + n.x, fv.y + n.y, fv.z + n.z, fv.w + n.w);
    reinterpret_cast<float4*>(noisy)[o] = out;
    reinterpret_cast<float4*>(target)[o] = make_float4(-n.x / sigma2, -n.y / sigma2, -n.z / sigma2, -n.w / sigma2);
    if (hi) split_store4(hi, lo, 4 * o, out.x, out.y, out.z, out.w);   // operand planes for the first GNN layer
}
// The Gaussian draw INSIDE the kernel (no randn tensor: one launch and 8 B/element less): Philox4x32-10, key = seed,
// counter = float4 index, subsequence = stream + (*epoch << 32); the four 32-bit words give two Box-Muller pairs.
__device__ __forceinline__ float4 philox_normal4(const Philox& rng, uint64_t counter, uint64_t stream) {
    const uint4 r = rng(counter, stream);
    const float u1 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u2 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
    const float u3 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u4 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
    // fast intrinsics: |error| ~1e-6 on log / sin / cos, irrelevant for a noise draw (u in (0,1): the log argument is never 0)
    const float ra = sqrtf(-2.0f * __logf(u1)), rb = sqrtf(-2.0f * __logf(u3));
    float sa, ca, sb, cb;
    __sincosf(6.283185307179586f * u2, &sa, &ca);
    __sincosf(6.283185307179586f * u4, &sb, &cb);
    return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}
__global__ void __launch_bounds__(256)
feat_noise_philox_kernel(const float* __restrict__ f, uint64_t seed, uint64_t stream, const uint64_t* __restrict__ epoch,
                         float sigma, float sigma2, float* __restrict__ noisy, float* __restrict__ target,
                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int NH4, int H, int bcast) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NH4) return;
    const int b = blockIdx.y;
    const size_t o = (size_t)b * NH4 + i;
    const float4 r = philox_normal4(Philox(seed), (uint64_t)o, stream + (epoch ? (*epoch << 32) : 0ull));
    const float4 fv = bcast ? *reinterpret_cast<const float4*>(f + (size_t)b * H + (4 * i) % H)
                            : reinterpret_cast<const float4*>(f)[o];
    const float4 n = make_float4(r.x * sigma, r.y * sigma, r.z * sigma, r.w * sigma);
    const float4 out = make_float4(fv.x + n.x, fv.y + n.y, fv.z + n.z, fv.w + n.w);
    reinterpret_cast<float4*>(noisy)[o] = out;
    reinterpret_cast<float4*>(target)[o] = make_float4(-n.x / sigma2, -n.y / sigma2, -n.z / sigma2, -n.w / sigma2);
    if (hi) split_store4(hi, lo, 4 * o, out.x, out.y, out.z, out.w);
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
int feat_noise_philox(const float* f, uint64_t seed, uint64_t stream, const uint64_t* epoch, float sigma, float sigma2,
                      float* noisy, float* target, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int N, int H, int bcast,
                      cudaStream_t st) {
    const long long total = (long long)B * N * H;
    if (total <= 0) return XGGM_OK;
    XGGM_REQUIRE(H % 4 == 0 && al16(f) && al16(noisy) && al16(target) && (long long)N * H / 4 < (1LL << 30));
    const int NH4 = N * H / 4;
    XGGM_LAUNCH((feat_noise_philox_kernel), dim3(ceil_div(NH4, 256), B), 256, 0, st, f, seed, stream, epoch, sigma, sigma2, noisy,
                target, hi, lo, NH4, H, bcast);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
int feat_noise(const float* f, const float* randn, float sigma, float sigma2, float* noisy,
               float* target, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int N, int H, int bcast, cudaStream_t st) {
    const long long total = (long long)B * N * H;
    if (total <= 0) return XGGM_OK;
    if (H % 4 == 0 && al16(f) && al16(randn) && al16(noisy) && al16(target) && (long long)N * H / 4 < (1LL << 30)) {
        const int NH4 = N * H / 4;
        XGGM_LAUNCH((feat_noise_vec_kernel), dim3(ceil_div(NH4, 256), B), 256, 0, st, f, randn, sigma, sigma2, noisy, target, hi, lo, NH4, H, bcast);
        XGGM_LAUNCH_CHECK();
        return XGGM_OK;
    }
    XGGM_LAUNCH((feat_noise_kernel), grid1d(total), 256, 0, st, f, randn, sigma, sigma2, noisy, target, total, N, H, bcast);
    XGGM_LAUNCH_CHECK();
    if (hi) {   // (rare shapes: planes through the stand-alone splitter)
        const float* src[1] = {noisy};
        __nv_bfloat16* h[1] = {hi};
        __nv_bfloat16* l[1] = {lo};
        return split_planes(src, h, lo ? l : nullptr, &total, 1, st);
    }
    return XGGM_OK;
}

__global__ void sum_nodes_kernel(const float* __restrict__ g, float* __restrict__ out, int B,
                                 int N, int H) {
    pdl_prologue();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)B * H) return;
    const long long b = e / H;
    const int c = (int)(e - b * H);
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += g[(b * N + n) * H + c];
    out[e] = s;
}
int sum_nodes(const float* g, float* out, int B, int N, int H, cudaStream_t st) {
    const long long total = (long long)B * H;
    if (total <= 0) return XGGM_OK;
    XGGM_LAUNCH((sum_nodes_kernel), grid1d(total), 256, 0, st, g, out, B, N, H);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ------------------------------------------------------------ score-matching
__global__ void __launch_bounds__(256)
score_mse_fwd_kernel(const float* __restrict__ s, const float* __restrict__ t, float coef,
                     float* __restrict__ loss, long long n) {
    pdl_prologue();
    __shared__ float part[8];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float d = s[i] - t[i];
        acc = fmaf(d, d, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss, v * coef);
    }
}
__global__ void score_mse_bwd_kernel(const float* __restrict__ s, const float* __restrict__ t,
                                     const float* __restrict__ gloss, float coef,
                                     float* __restrict__ gs, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gs[i] = gloss[0] * coef * (s[i] - t[i]);
}
int score_mse_fwd(const float* s, const float* t, float sigma, float* loss, long long n,
                  cudaStream_t st) {
    XGGM_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
    if (n <= 0) return XGGM_OK;
    const float coef = 0.5f * sigma * sigma / (float)n;
    const int grid = (int)min((long long)148 * 8, (n + 255) / 256);
    XGGM_LAUNCH((score_mse_fwd_kernel), grid, 256, 0, st, s, t, coef, loss, n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
__global__ void __launch_bounds__(256)
score_mse_bwd_vec_kernel(const float4* __restrict__ s, const float4* __restrict__ t, const float* __restrict__ gloss,
                         float coef, float4* __restrict__ gs, long long n4) {
    pdl_prologue();
    const float c = gloss[0] * coef;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = s[i], b = t[i];
        gs[i] = make_float4(c * (a.x - b.x), c * (a.y - b.y), c * (a.z - b.z), c * (a.w - b.w));
    }
}
int score_mse_bwd(const float* s, const float* t, const float* gloss, float sigma, float* gs,
                  long long n, cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    const float coef = sigma * sigma / (float)n;
    if (n % 4 == 0 && al16(s) && al16(t) && al16(gs)) {
        const long long n4 = n / 4;
        const int grid = (int)min((long long)148 * 16, (n4 + 255) / 256);
        XGGM_LAUNCH((score_mse_bwd_vec_kernel), grid, 256, 0, st, reinterpret_cast<const float4*>(s), reinterpret_cast<const float4*>(t), gloss, coef,
                                                       reinterpret_cast<float4*>(gs), n4);
        XGGM_LAUNCH_CHECK();
        return XGGM_OK;
    }
    XGGM_LAUNCH((score_mse_bwd_kernel), grid1d(n), 256, 0, st, s, t, gloss, coef, gs, n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ---------------------------------------------------------------- symmetric KL
// One warp per row.  L_row = sum_c (py-px)(log py - log px).
//
// Node-branch tail (src/vqa/vqacpv2.py:236-246): the same pass can also carry the score-matching term of
// loss_sm = kl_w * KL(x, y) + sm_w * loss_func(x, target) and, in the backward direction, the read-out
// gradient that is identical for all `rows_per_group` nodes of a graph -- so the generated node features
// are read once per direction and their gradient is written once (TailExtras; all-zero = plain KL).
struct TailExtras {
    const float* sm_target;   // [R,C] score-matching target, or null
    const float* row_add;     // [R / rows_per_group, C] added to every gx row of its group, or null
    float kl_w;               // weight of the KL term (1 for the plain entry points)
    float sm_fwd;             // sm_w * 0.5 sigma^2 / (R*C)
    float sm_bwd;             // sm_w * sigma^2 / (R*C)
    int rows_per_group;
};
static inline TailExtras tail_none() { return TailExtras{nullptr, nullptr, 1.f, 0.f, 0.f, 1}; }
struct RowSoftmax {
    float mx, lse;  // log p_c = v_c - mx - lse
};
__device__ __forceinline__ RowSoftmax row_softmax_stats(const float* __restrict__ v, int C, int lane) {
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, v[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(v[c] - mx);
    s = warp_sum(s);
    return {mx, logf(s)};
}

__global__ void __launch_bounds__(256)
sym_kl_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                  float* __restrict__ loss, int R, int C, float inv_count, const TailExtras ex) {
    pdl_prologue();
    __shared__ float part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f, acc_sm = 0.f;
    for (int r = blockIdx.x * 8 + warp; r < R; r += gridDim.x * 8) {
        const float* xr = x + (size_t)r * C;
        const float* yr = y + (size_t)r * C;
        const RowSoftmax sx = row_softmax_stats(xr, C, lane), sy = row_softmax_stats(yr, C, lane);
        for (int c = lane; c < C; c += 32) {
            const float lpx = xr[c] - sx.mx - sx.lse, lpy = yr[c] - sy.mx - sy.lse;
            const float px = expf(lpx), py = expf(lpy);
            // F.kl_div(log_px, py) + F.kl_div(log_py, px), reduction='none'
            acc += py * (lpy - lpx) + px * (lpx - lpy);
        }
        if (ex.sm_target) {
            const float* tr = ex.sm_target + (size_t)r * C;
            for (int c = lane; c < C; c += 32) {
                const float d = xr[c] - tr[c];
                acc_sm = fmaf(d, d, acc_sm);
            }
        }
    }
    acc = acc * (inv_count * ex.kl_w) + acc_sm * ex.sm_fwd;
    acc = warp_sum(acc);
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss, v);
    }
}

__global__ void __launch_bounds__(256)
sym_kl_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                  const float* __restrict__ gloss, float* __restrict__ gx,
                  float* __restrict__ gy, int R, int C, float inv_count, const TailExtras ex) {
    pdl_prologue();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float g = gloss[0] * inv_count * ex.kl_w, gsm = gloss[0] * ex.sm_bwd;
    for (int r = blockIdx.x * 8 + warp; r < R; r += gridDim.x * 8) {
        const float* xr = x + (size_t)r * C;
        const float* yr = y + (size_t)r * C;
        const RowSoftmax sx = row_softmax_stats(xr, C, lane), sy = row_softmax_stats(yr, C, lane);
        float ax = 0.f, ay = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float lpx = xr[c] - sx.mx - sx.lse, lpy = yr[c] - sy.mx - sy.lse;
            const float e = lpy - lpx;
            ax = fmaf(expf(lpx), e, ax);
            ay = fmaf(expf(lpy), e, ay);
        }
        ax = warp_sum(ax);
        ay = warp_sum(ay);
        for (int c = lane; c < C; c += 32) {
            const float lpx = xr[c] - sx.mx - sx.lse, lpy = yr[c] - sy.mx - sy.lse;
            const float px = expf(lpx), py = expf(lpy);
            const float e = lpy - lpx, d = py - px;
            float ox = g * (px * (ax - e) - d);
            if (ex.sm_target) ox = fmaf(gsm, xr[c] - ex.sm_target[(size_t)r * C + c], ox);
            if (ex.row_add) ox += ex.row_add[(size_t)(r / ex.rows_per_group) * C + c];
            if (gx) gx[(size_t)r * C + c] = ox;
            if (gy) gy[(size_t)r * C + c] = g * (py * (e - ay) + d);
        }
    }
}
// Fast path (C % 128 == 0, C <= 1024): both rows live in registers (lane l holds the float4s at columns
// 128 i + 4 l), every exponential is evaluated once, p = e / sum and log p = (v - max) - log(sum).
template <int NV>
__device__ __forceinline__ void kl_row_load(const float* __restrict__ p, int lane, float (&d)[NV * 4], float (&e)[NV * 4],
                                            float& inv_s, float& lse) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(p + 128 * i + 4 * lane);
        d[4 * i] = t.x; d[4 * i + 1] = t.y; d[4 * i + 2] = t.z; d[4 * i + 3] = t.w;
    }
    float mx = d[0];
#pragma unroll
    for (int i = 1; i < NV * 4; ++i) mx = fmaxf(mx, d[i]);
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) {
        d[i] -= mx;
        e[i] = expf(d[i]);
        s += e[i];
    }
    s = warp_sum(s);
    inv_s = 1.0f / s;
    lse = logf(s);
}

template <int NV, bool BWD>
__global__ void __launch_bounds__(256)
sym_kl_fast_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ loss,
                   const float* __restrict__ gloss, float* __restrict__ gx, float* __restrict__ gy, int R,
                   float inv_count, const TailExtras tx) {
    pdl_prologue();
    constexpr int C = NV * 128;
    __shared__ float part[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float g = BWD ? gloss[0] * inv_count * tx.kl_w : 0.f;
    const float gsm = BWD ? gloss[0] * tx.sm_bwd : 0.f;
    float acc = 0.f, acc_sm = 0.f;
    for (int r = blockIdx.x * 8 + warp; r < R; r += gridDim.x * 8) {
        float dx[NV * 4], ex[NV * 4], dy[NV * 4], ey[NV * 4];
        float isx, lsx, isy, lsy;
        kl_row_load<NV>(x + (size_t)r * C, lane, dx, ex, isx, lsx);
        kl_row_load<NV>(y + (size_t)r * C, lane, dy, ey, isy, lsy);
        if (!BWD) {
#pragma unroll
            for (int i = 0; i < NV * 4; ++i) {
                const float lpx = dx[i] - lsx, lpy = dy[i] - lsy;
                acc += (ey[i] * isy - ex[i] * isx) * (lpy - lpx);   // py (lpy-lpx) + px (lpx-lpy)
            }
            if (tx.sm_target) {
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const size_t o = (size_t)r * C + 128 * i + 4 * lane;
                    const float4 xv = *reinterpret_cast<const float4*>(x + o);
                    const float4 tv = *reinterpret_cast<const float4*>(tx.sm_target + o);
                    const float d0 = xv.x - tv.x, d1 = xv.y - tv.y, d2 = xv.z - tv.z, d3 = xv.w - tv.w;
                    acc_sm += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
                }
            }
        } else {
            float ax = 0.f, ay = 0.f;
#pragma unroll
            for (int i = 0; i < NV * 4; ++i) {
                ex[i] *= isx;   // px
                ey[i] *= isy;   // py
                dx[i] = (dy[i] - lsy) - (dx[i] - lsx);   // e = lpy - lpx
                ax = fmaf(ex[i], dx[i], ax);
                ay = fmaf(ey[i], dx[i], ay);
            }
            ax = warp_sum(ax);
            ay = warp_sum(ay);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                float ox[4], oy[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = 4 * i + j;
                    const float d = ey[k] - ex[k];
                    ox[j] = g * (ex[k] * (ax - dx[k]) - d);
                    oy[j] = g * (ey[k] * (dx[k] - ay) + d);
                }
                const size_t o = (size_t)r * C + 128 * i + 4 * lane;
                if (tx.sm_target) {
                    const float4 xv = *reinterpret_cast<const float4*>(x + o);
                    const float4 tv = *reinterpret_cast<const float4*>(tx.sm_target + o);
                    ox[0] = fmaf(gsm, xv.x - tv.x, ox[0]); ox[1] = fmaf(gsm, xv.y - tv.y, ox[1]);
                    ox[2] = fmaf(gsm, xv.z - tv.z, ox[2]); ox[3] = fmaf(gsm, xv.w - tv.w, ox[3]);
                }
                if (tx.row_add) {
                    const float4 av = *reinterpret_cast<const float4*>(
                        tx.row_add + (size_t)(r / tx.rows_per_group) * C + 128 * i + 4 * lane);
                    ox[0] += av.x; ox[1] += av.y; ox[2] += av.z; ox[3] += av.w;
                }
                if (gx) *reinterpret_cast<float4*>(gx + o) = make_float4(ox[0], ox[1], ox[2], ox[3]);
                if (gy) *reinterpret_cast<float4*>(gy + o) = make_float4(oy[0], oy[1], oy[2], oy[3]);
            }
        }
    }
    if (!BWD) {
        acc = acc * (inv_count * tx.kl_w) + acc_sm * tx.sm_fwd;
        acc = warp_sum(acc);
        if (lane == 0) part[warp] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
            v = warp_sum(v);
            if (threadIdx.x == 0) atomicAdd(loss, v);
        }
    }
}
#define XGGM_KL_DISPATCH(C, ...)            \
    switch ((C) / 128) {                    \
        case 1: { constexpr int NV = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int NV = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int NV = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int NV = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int NV = 5; __VA_ARGS__; } break; \
        case 6: { constexpr int NV = 6; __VA_ARGS__; } break; \
        case 7: { constexpr int NV = 7; __VA_ARGS__; } break; \
        default: { constexpr int NV = 8; __VA_ARGS__; } break; \
    }

static inline bool kl_fast_ok(int C, const void* a, const void* b, const void* c = nullptr, const void* d = nullptr,
                              const void* e = nullptr, const void* f = nullptr) {
    return C % 128 == 0 && C >= 128 && C <= 1024 && al16(a) && al16(b) && al16(c) && al16(d) && al16(e) && al16(f);
}
static int kl_fwd_launch(const float* x, const float* y, float* loss, int R, int C, const TailExtras& ex, cudaStream_t st) {
    XGGM_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
    if (R <= 0 || C <= 0) return XGGM_OK;
    const float inv = 1.0f / ((float)R * (float)C);
    if (kl_fast_ok(C, x, y, ex.sm_target)) {
        const int grid = min(148 * 4, ceil_div(R, 8));
        XGGM_KL_DISPATCH(C, (XGGM_LAUNCH((sym_kl_fast_kernel<NV, false>), grid, 256, 0, st, x, y, loss, nullptr, nullptr, nullptr, R, inv, ex)));
        XGGM_LAUNCH_CHECK();
        return XGGM_OK;
    }
    const int grid = min(148 * 8, ceil_div(R, 8));
    XGGM_LAUNCH((sym_kl_fwd_kernel), grid, 256, 0, st, x, y, loss, R, C, inv, ex);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
static int kl_bwd_launch(const float* x, const float* y, const float* gloss, float* gx, float* gy, int R, int C,
                         const TailExtras& ex, cudaStream_t st) {
    if (R <= 0 || C <= 0) return XGGM_OK;
    const float inv = 1.0f / ((float)R * (float)C);
    if (kl_fast_ok(C, x, y, gx, gy, ex.sm_target, ex.row_add)) {
        const int grid = min(148 * 4, ceil_div(R, 8));
        XGGM_KL_DISPATCH(C, (XGGM_LAUNCH((sym_kl_fast_kernel<NV, true>), grid, 256, 0, st, x, y, nullptr, gloss, gx, gy, R, inv, ex)));
        XGGM_LAUNCH_CHECK();
        return XGGM_OK;
    }
    const int grid = min(148 * 8, ceil_div(R, 8));
    XGGM_LAUNCH((sym_kl_bwd_kernel), grid, 256, 0, st, x, y, gloss, gx, gy, R, C, inv, ex);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
int sym_kl_fwd(const float* x, const float* y, float* loss, int R, int C, cudaStream_t st) {
    return kl_fwd_launch(x, y, loss, R, C, tail_none(), st);
}
int sym_kl_bwd(const float* x, const float* y, const float* gloss, float* gx, float* gy, int R,
               int C, cudaStream_t st) {
    return kl_bwd_launch(x, y, gloss, gx, gy, R, C, tail_none(), st);
}

// --------------------------------------------------------------- fusion readout
__global__ void fuse_readout_fwd_kernel(const float* __restrict__ xp, const float* __restrict__ nodes,
                                        float* __restrict__ out, int B, int N, int H) {
    pdl_prologue();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)B * H) return;
    const long long b = e / H;
    const int c = (int)(e - b * H);
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += nodes[(b * N + n) * H + c];
    out[b * 2 * H + c] = xp[e];
    out[b * 2 * H + H + c] = tanhf(s / (float)N);
}
__global__ void fuse_readout_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out,
                                        float* __restrict__ gxp, float* __restrict__ gnodes, float* __restrict__ grow,
                                        int B, int N, int H, int accumulate) {
    pdl_prologue();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)B * H) return;
    const long long b = e / H;
    const int c = (int)(e - b * H);
    if (gxp) gxp[e] = gout[b * 2 * H + c];
    const float t = out[b * 2 * H + H + c];
    const float g = gout[b * 2 * H + H + c] * (1.f - t * t) / (float)N;
    if (grow) grow[e] = g;   // the per-graph row every node shares (consumed by the fused node-tail backward)
    if (!gnodes) return;
    for (int n = 0; n < N; ++n) {
        float* p = gnodes + (b * N + n) * H + c;
        *p = accumulate ? *p + g : g;
    }
}
int fuse_readout_fwd(const float* xp, const float* nodes, float* out, int B, int N, int H, cudaStream_t st) {
    const long long total = (long long)B * H;
    if (total <= 0) return XGGM_OK;
    XGGM_LAUNCH((fuse_readout_fwd_kernel), grid1d(total), 256, 0, st, xp, nodes, out, B, N, H);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
int fuse_readout_bwd(const float* gout, const float* out, float* gxp, float* gnodes, int B, int N,
                     int H, int accumulate, cudaStream_t st) {
    const long long total = (long long)B * H;
    if (total <= 0) return XGGM_OK;
    XGGM_LAUNCH((fuse_readout_bwd_kernel), grid1d(total), 256, 0, st, gout, out, gxp, gnodes, nullptr, B, N, H, accumulate);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ------------------------------------------------------------------- node-branch tail
// loss = kl_w * sym_kl(nodes, feat) + sm_w * loss_func(nodes, target, sigma);  cat = [xp | tanh(mean_n nodes)]
int node_tail_fwd(const float* nodes, const float* feat, const float* target, const float* xp, float sigma, float kl_w,
                  float sm_w, float* loss, float* cat, int B, int N, int H, cudaStream_t st) {
    const int R = B * N;
    const float n = (float)R * (float)H;
    const TailExtras ex{target, nullptr, kl_w, R > 0 ? sm_w * 0.5f * sigma * sigma / n : 0.f, 0.f, N};
    XGGM_TRY(kl_fwd_launch(nodes, feat, loss, R, H, ex, st));
    return fuse_readout_fwd(xp, nodes, cat, B, N, H, st);
}
// gnodes = d loss / d nodes * gloss + (read-out gradient of gcat, the same row for every node of a graph);
// gfeat (nullable) = d loss / d feat * gloss; gxp = gcat[:, :H]; grow [B,H] is scratch.
int node_tail_bwd(const float* nodes, const float* feat, const float* target, const float* cat, const float* gloss,
                  const float* gcat, float sigma, float kl_w, float sm_w, float* gnodes, float* gfeat, float* gxp,
                  float* grow, int B, int N, int H, cudaStream_t st) {
    const int R = B * N;
    const long long total = (long long)B * H;
    if (total <= 0) return XGGM_OK;
    XGGM_LAUNCH((fuse_readout_bwd_kernel), grid1d(total), 256, 0, st, gcat, cat, gxp, nullptr, grow, B, N, H, 0);
    XGGM_LAUNCH_CHECK();
    const float n = (float)R * (float)H;
    const TailExtras ex{target, grow, kl_w, 0.f, sm_w * sigma * sigma / n, N};
    return kl_bwd_launch(nodes, feat, gloss, gnodes, gfeat, R, H, ex, st);
}

// ------------------------------------------------------------------- sigmoid
__global__ void sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = sigmoidf_(x[i]);
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                   float* __restrict__ gx, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = y[i]; gx[i] = gy[i] * v * (1.f - v); }
}
int sigmoid_fwd(const float* x, float* y, long long n, cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    XGGM_LAUNCH((sigmoid_fwd_kernel), grid1d(n), 256, 0, st, x, y, n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
int sigmoid_bwd(const float* gy, const float* y, float* gx, long long n, cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    XGGM_LAUNCH((sigmoid_bwd_kernel), grid1d(n), 256, 0, st, gy, y, gx, n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ------------------------------------------------------- GeLU / masked scaling
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = gelu_erf(x[i]);
}
__global__ void gelu_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                float* __restrict__ gx, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gx[i] = gy[i] * gelu_erf_grad(x[i]);
}
int gelu_fwd(const float* x, float* y, long long n, cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    XGGM_LAUNCH((gelu_fwd_kernel), grid1d(n), 256, 0, st, x, y, n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
int gelu_bwd(const float* gy, const float* x, float* gx, long long n, cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    XGGM_LAUNCH((gelu_bwd_kernel), grid1d(n), 256, 0, st, gy, x, gx, n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
// inverted dropout with an explicit mask: y = keep ? x*scale : 0 (its own backward); keep == NULL: y = x*scale
__global__ void mask_scale_kernel(const float* __restrict__ x, const uint8_t* __restrict__ keep,
                                  float scale, float* __restrict__ y, long long n) {
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (!keep || keep[i]) ? x[i] * scale : 0.f;
}
int mask_scale(const float* x, const uint8_t* keep, float scale, float* y, long long n, cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    XGGM_LAUNCH((mask_scale_kernel), grid1d(n), 256, 0, st, x, keep, scale, y, n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// out = dropout((x + y) / 2): the tail of VisualFeatEncoder (src/lxrt/modeling.py:553-555).
// keep == NULL -> no dropout.  4 elements per thread when everything is 16-byte aligned.
__global__ void __launch_bounds__(256)
avg2_drop_kernel(const float* __restrict__ x, const float* __restrict__ y, const uint8_t* __restrict__ keep,
                 float scale, float* __restrict__ out, long long n, int vec) {
    pdl_prologue();
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec) {
        const long long n4 = n >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 a = reinterpret_cast<const float4*>(x)[i], b = reinterpret_cast<const float4*>(y)[i];
            float4 o = make_float4((a.x + b.x) / 2, (a.y + b.y) / 2, (a.z + b.z) / 2, (a.w + b.w) / 2);
            if (keep) {
                const uchar4 k = reinterpret_cast<const uchar4*>(keep)[i];
                o.x = k.x ? o.x * scale : 0.f; o.y = k.y ? o.y * scale : 0.f;
                o.z = k.z ? o.z * scale : 0.f; o.w = k.w ? o.w * scale : 0.f;
            }
            reinterpret_cast<float4*>(out)[i] = o;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const float o = (x[i] + y[i]) / 2;
            out[i] = keep ? (keep[i] ? o * scale : 0.f) : o;
        }
    }
}
int avg2_drop(const float* x, const float* y, const uint8_t* keep, float scale, float* out, long long n,
              cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    const int vec = n % 4 == 0 && al16(x) && al16(y) && al16(out) && (reinterpret_cast<uintptr_t>(keep) & 3) == 0;
    const int grid = (int)min((long long)148 * 16, ((vec ? n / 4 : n) + 255) / 256);
    XGGM_LAUNCH((avg2_drop_kernel), grid, 256, 0, st, x, y, keep, scale, out, n, vec);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ----------------------------------------------------------------- keep mask
// 4 elements per thread, one Philox call, one 4-byte store.  Same draw as DropSpec mode 2.
__global__ void keep_mask_kernel(uint8_t* __restrict__ keep, long long n, uint32_t thresh,
                                 uint64_t seed, uint64_t stream_id, const uint64_t* __restrict__ epoch,
                                 int aligned4) {
    pdl_prologue();
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q * 4 >= n) return;
    const uint64_t stream = stream_id + (epoch ? (*epoch << 32) : 0ull);
    if (thresh == XGGM_COIN_THRESH) {   // one bit per element (common.cuh): the mapping the row kernels use for p = 1/2
        const uint32_t nib = coin_nibble(Philox(seed), stream, (size_t)q * 4);
        if (aligned4 && q * 4 + 3 < n) {
            reinterpret_cast<uchar4*>(keep)[q] = make_uchar4(nib & 1u, (nib >> 1) & 1u, (nib >> 2) & 1u, (nib >> 3) & 1u);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (q * 4 + i < n) keep[q * 4 + i] = (nib >> i) & 1u;
        }
        return;
    }
    const uint4 r = Philox(seed)((uint64_t)q, stream);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    if (aligned4 && q * 4 + 3 < n) {
        uchar4 o;
        o.x = w[0] >= thresh; o.y = w[1] >= thresh; o.z = w[2] >= thresh; o.w = w[3] >= thresh;
        reinterpret_cast<uchar4*>(keep)[q] = o;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (q * 4 + i < n) keep[q * 4 + i] = w[i] >= thresh ? 1 : 0;
    }
}
int keep_mask(uint8_t* keep, long long n, float p, uint64_t seed, uint64_t stream_id, const uint64_t* epoch,
              cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    XGGM_REQUIRE(p >= 0.f && p < 1.f);
    const int aligned4 = (reinterpret_cast<uintptr_t>(keep) & 3) == 0;
    XGGM_LAUNCH((keep_mask_kernel), grid1d((n + 3) / 4), 256, 0, st, keep, n, drop_threshold(p), seed, stream_id, epoch, aligned4);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

}  // namespace xggm
