// Row-wise fused epilogues of the graph block: LayerNorm and the jump-knowledge head
// tail  dropout(LN(GeLU(z))).  One warp owns one row (H = 768 -> 24 values per lane);
// reductions are warp shuffles, statistics are fp32 two-pass (mean, then centred
// sum of squares) like torch's LayerNorm.  Column-wise parameter gradients
// (gamma/beta) are accumulated per warp in shared memory and flushed once per CTA.
#include "common.cuh"

namespace xggm {

constexpr int ROW_WARPS = 4;        // warps per CTA
constexpr int ROWS_PER_WARP = 8;    // rows each warp walks through

__device__ __forceinline__ void row_range(int M, int& r0, int& r1) {
    const int warp = threadIdx.x >> 5;
    const int base = (blockIdx.x * ROW_WARPS + warp) * ROWS_PER_WARP;
    r0 = base;
    r1 = min(M, base + ROWS_PER_WARP);
}

static inline int row_grid(int M) { return ceil_div(M, ROW_WARPS * ROWS_PER_WARP); }

// ---------------------------------------------------------------- LayerNorm fwd
__global__ void __launch_bounds__(ROW_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ u, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ h,
                     float* __restrict__ xhat, float* __restrict__ rstd_out, int M, int H,
                     float eps) {
    const int lane = threadIdx.x & 31;
    int r0, r1;
    row_range(M, r0, r1);
    for (int r = r0; r < r1; ++r) {
        const float* ur = u + (size_t)r * H;
        float s = 0.f;
        for (int c = lane; c < H; c += 32) s += ur[c];
        const float mean = warp_sum(s) / (float)H;
        float q = 0.f;
        for (int c = lane; c < H; c += 32) { const float d = ur[c] - mean; q = fmaf(d, d, q); }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + eps);
        for (int c = lane; c < H; c += 32) {
            const float xh = (ur[c] - mean) * rstd;
            if (xhat) xhat[(size_t)r * H + c] = xh;
            h[(size_t)r * H + c] = fmaf(xh, gamma[c], beta[c]);
        }
        if (lane == 0 && rstd_out) rstd_out[r] = rstd;
    }
}

// ---------------------------------------------------------------- LayerNorm bwd
// smem: [ROW_WARPS][2][H] private column accumulators
__global__ void __launch_bounds__(ROW_WARPS * 32)
layernorm_bwd_kernel(const float* __restrict__ gh, const float* __restrict__ xhat,
                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                     float* __restrict__ gu, float* __restrict__ ggamma,
                     float* __restrict__ gbeta, int M, int H) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sg = sm + (size_t)warp * 2 * H;
    float* sb = sg + H;
    for (int c = lane; c < H; c += 32) { sg[c] = 0.f; sb[c] = 0.f; }
    int r0, r1;
    row_range(M, r0, r1);
    for (int r = r0; r < r1; ++r) {
        const float* gr = gh + (size_t)r * H;
        const float* xr = xhat + (size_t)r * H;
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < H; c += 32) {
            const float d = gr[c] * gamma[c];
            s1 += d;
            s2 = fmaf(d, xr[c], s2);
        }
        const float c1 = warp_sum(s1) / (float)H, c2 = warp_sum(s2) / (float)H;
        const float rs = rstd[r];
        for (int c = lane; c < H; c += 32) {
            const float g = gr[c], xh = xr[c];
            gu[(size_t)r * H + c] = rs * (g * gamma[c] - c1 - xh * c2);
            sg[c] = fmaf(g, xh, sg[c]);
            sb[c] += g;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < ROW_WARPS; ++w) { a += sm[(size_t)w * 2 * H + c]; b += sm[(size_t)w * 2 * H + H + c]; }
        atomicAdd(&ggamma[c], a);
        atomicAdd(&gbeta[c], b);
    }
}

// ------------------------------------------------- dropout(LN(GeLU(z))) forward
__global__ void __launch_bounds__(ROW_WARPS * 32)
gelu_ln_drop_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const uint8_t* __restrict__ keep,
                        float scale, float* __restrict__ out, float* __restrict__ mean_out,
                        float* __restrict__ rstd_out, int M, int H, float eps, int accumulate) {
    const int lane = threadIdx.x & 31;
    int r0, r1;
    row_range(M, r0, r1);
    for (int r = r0; r < r1; ++r) {
        const float* zr = z + (size_t)r * H;
        float s = 0.f;
        for (int c = lane; c < H; c += 32) s += gelu_erf(zr[c]);
        const float mean = warp_sum(s) / (float)H;
        float q = 0.f;
        for (int c = lane; c < H; c += 32) { const float d = gelu_erf(zr[c]) - mean; q = fmaf(d, d, q); }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + eps);
        for (int c = lane; c < H; c += 32) {
            const size_t o = (size_t)r * H + c;
            float y = fmaf((gelu_erf(zr[c]) - mean) * rstd, gamma[c], beta[c]);
            if (keep) y = keep[o] ? y * scale : 0.f;
            out[o] = accumulate ? out[o] + y : y;
        }
        if (lane == 0) {
            if (mean_out) mean_out[r] = mean;
            if (rstd_out) rstd_out[r] = rstd;
        }
    }
}

// ------------------------------------------------ dropout(LN(GeLU(z))) backward
__global__ void __launch_bounds__(ROW_WARPS * 32)
gelu_ln_drop_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ z,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ gamma, const uint8_t* __restrict__ keep,
                        float scale, float* __restrict__ gz, float* __restrict__ ggamma,
                        float* __restrict__ gbeta, int M, int H) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sg = sm + (size_t)warp * 2 * H;
    float* sb = sg + H;
    for (int c = lane; c < H; c += 32) { sg[c] = 0.f; sb[c] = 0.f; }
    int r0, r1;
    row_range(M, r0, r1);
    for (int r = r0; r < r1; ++r) {
        const float* gr = gout + (size_t)r * H;
        const float* zr = z + (size_t)r * H;
        const uint8_t* kr = keep ? keep + (size_t)r * H : nullptr;
        const float mu = mean[r], rs = rstd[r];
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < H; c += 32) {
            float gy = gr[c];
            if (kr) gy = kr[c] ? gy * scale : 0.f;
            const float yh = (gelu_erf(zr[c]) - mu) * rs;
            const float d = gy * gamma[c];
            s1 += d;
            s2 = fmaf(d, yh, s2);
        }
        const float c1 = warp_sum(s1) / (float)H, c2 = warp_sum(s2) / (float)H;
        for (int c = lane; c < H; c += 32) {
            float gy = gr[c];
            if (kr) gy = kr[c] ? gy * scale : 0.f;
            const float zv = zr[c];
            const float yh = (gelu_erf(zv) - mu) * rs;
            const float gg = rs * (gy * gamma[c] - c1 - yh * c2);
            gz[(size_t)r * H + c] = gg * gelu_erf_grad(zv);
            sg[c] = fmaf(gy, yh, sg[c]);
            sb[c] += gy;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < ROW_WARPS; ++w) { a += sm[(size_t)w * 2 * H + c]; b += sm[(size_t)w * 2 * H + H + c]; }
        atomicAdd(&ggamma[c], a);
        atomicAdd(&gbeta[c], b);
    }
}

// ------------------------------------------------------------------ launchers
int layernorm_fwd(const float* u, const float* gamma, const float* beta, float* h, float* xhat,
                  float* rstd, int M, int H, float eps, cudaStream_t st) {
    if (M <= 0) return XGGM_OK;
    layernorm_fwd_kernel<<<row_grid(M), ROW_WARPS * 32, 0, st>>>(u, gamma, beta, h, xhat, rstd, M, H, eps);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

static int col_smem(int H, size_t& bytes) {
    bytes = sizeof(float) * (size_t)ROW_WARPS * 2 * H;
    return bytes <= 48 * 1024 ? XGGM_OK : XGGM_ERR_ARG;
}

int layernorm_bwd(const float* gh, const float* xhat, const float* rstd, const float* gamma,
                  float* gu, float* ggamma, float* gbeta, int M, int H, cudaStream_t st) {
    if (M <= 0) return XGGM_OK;
    size_t smem;
    XGGM_TRY(col_smem(H, smem));
    layernorm_bwd_kernel<<<row_grid(M), ROW_WARPS * 32, smem, st>>>(gh, xhat, rstd, gamma, gu, ggamma, gbeta, M, H);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

int gelu_ln_drop_fwd(const float* z, const float* gamma, const float* beta, const uint8_t* keep,
                     float scale, float* out, float* mean, float* rstd, int M, int H, float eps,
                     int accumulate, cudaStream_t st) {
    if (M <= 0) return XGGM_OK;
    gelu_ln_drop_fwd_kernel<<<row_grid(M), ROW_WARPS * 32, 0, st>>>(z, gamma, beta, keep, scale, out, mean, rstd, M, H, eps, accumulate);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

int gelu_ln_drop_bwd(const float* gout, const float* z, const float* mean, const float* rstd,
                     const float* gamma, const uint8_t* keep, float scale, float* gz,
                     float* ggamma, float* gbeta, int M, int H, cudaStream_t st) {
    if (M <= 0) return XGGM_OK;
    size_t smem;
    XGGM_TRY(col_smem(H, smem));
    gelu_ln_drop_bwd_kernel<<<row_grid(M), ROW_WARPS * 32, smem, st>>>(gout, z, mean, rstd, gamma, keep, scale, gz, ggamma, gbeta, M, H);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

}  // namespace xggm
