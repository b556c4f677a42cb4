// Training-step tail of the trainers, as fused HBM-bound passes over the flat
// parameter / gradient buffers of xggm_b200.ddp.FlatGrads:
//   * BCEWithLogitsLoss (mean) and its gradient            src/vqa/vqacpv2.py:110,173,220,249
//   * the squared gradient norm of clip_grad_norm_(., 5.)  src/vqa/vqacpv2.py:175,223,252
//   * BertAdam.step with the clip coefficient applied on the fly (no pass that rescales the gradients)
//                                                           src/lxrt/optimization.py:116-203
// The reference walks the parameter list in Python (three elementwise kernels per tensor for the optimizer,
// two for the clip); here a step is one reduction kernel and one update kernel whatever the number of tensors.
#include "common.cuh"
#include "kernels.cuh"

namespace xggm {

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int stream_grid(long long items, int per_block) {
    const long long b = (items + per_block - 1) / per_block;
    return (int)(b < 1 ? 1 : (b > 148LL * 16 ? 148LL * 16 : b));
}
__device__ __forceinline__ float block_sum_256(float v, float* part) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
        r = warp_sum(r);
    }
    return r;   // valid in thread 0
}

// ---------------------------------------------------------------- BCE with logits
// l(x, t) = max(x, 0) - x t + log(1 + exp(-|x|))   (torch's stable form); loss = scale / n * sum l
__global__ void __launch_bounds__(256)
bce_logits_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t, float* __restrict__ loss,
                      long long n, float coef) {
    pdl_prologue();
    __shared__ float part[8];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float xv = x[i];
        acc += fmaxf(xv, 0.f) - xv * t[i] + log1pf(expf(-fabsf(xv)));
    }
    const float s = block_sum_256(acc, part);
    if (threadIdx.x == 0) atomicAdd(loss, s * coef);
}
// gx = gloss * scale / n * (sigmoid(x) - t)
__global__ void __launch_bounds__(256)
bce_logits_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ gloss,
                      float* __restrict__ gx, long long n, float coef) {
    pdl_prologue();
    const float c = gloss[0] * coef;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        gx[i] = c * (sigmoidf_(x[i]) - t[i]);
}
int bce_logits_fwd(const float* x, const float* t, float scale, float* loss, long long n, cudaStream_t st) {
    XGGM_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
    if (n <= 0) return XGGM_OK;
    XGGM_LAUNCH((bce_logits_fwd_kernel), stream_grid(n, 1024), 256, 0, st, x, t, loss, n, scale / (float)n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}
int bce_logits_bwd(const float* x, const float* t, const float* gloss, float scale, float* gx, long long n,
                   cudaStream_t st) {
    if (n <= 0) return XGGM_OK;
    XGGM_LAUNCH((bce_logits_bwd_kernel), stream_grid(n, 1024), 256, 0, st, x, t, gloss, gx, n, scale / (float)n);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ---------------------------------------------------------------- gradient norm
// out[0] (+)= sum g^2   (4 B/element read: HBM bound)
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out, int vec) {
    pdl_prologue();
    __shared__ float part[8];
    float acc = 0.f;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    if (vec) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        const long long n4 = n >> 2;
        for (long long i = tid; i < n4; i += nth) {
            const float4 v = g4[i];
            acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        for (long long i = (n4 << 2) + tid; i < n; i += nth) acc = fmaf(g[i], g[i], acc);
    } else {
        for (long long i = tid; i < n; i += nth) acc = fmaf(g[i], g[i], acc);
    }
    const float s = block_sum_256(acc, part);
    if (threadIdx.x == 0) atomicAdd(out, s);
}
int grad_sumsq(const float* g, long long n, float* out, int accumulate, cudaStream_t st) {
    if (!accumulate) XGGM_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float), st));
    if (n <= 0) return XGGM_OK;
    XGGM_LAUNCH((sumsq_kernel), stream_grid(n, 4096), 256, 0, st, g, n, out, aligned16(g) ? 1 : 0);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ---------------------------------------------------------------- BertAdam
// One pass: 16 B read + 12 B written per parameter.
//   g' = clip * g,  clip = min(1, max_norm / (sqrt(sumsq) + 1e-6))            (torch.nn.utils.clip_grad_norm_)
//   m <- b1 m + (1-b1) g' ; v <- b2 v + (1-b2) g'^2 ; p <- p - lr (m / (sqrt(v) + e) + wd p)   (no bias correction)
struct AdamArgs {
    float lr, b1, one_b1, b2, one_b2, eps, wd, max_norm;
    const float* sumsq;   // device scalar: squared norm of ALL gradients being clipped together, or null
    // device-side schedule (CUDA-graph-safe steps): lr is multiplied by schedule(step / t_total, warmup) with `step`
    // read from the device counter; with `advance` the LAST CTA of the grid increments the counter (every CTA has read
    // it before taking its ticket, so the increment cannot race with a read of the same launch)
    long long* step_dev;
    unsigned int* ticket;
    double warmup;
    long long t_total;
    int schedule, advance;
};
// src/lxrt/optimization.py:28-49 (x = step / t_total)
__device__ __forceinline__ double schedule_factor(int schedule, double x, double warmup) {
    if (x < warmup) return x / warmup;
    if (schedule == XGGM_SCHED_COSINE) return 0.5 * (1.0 + cos(3.14159265358979323846 * x));
    if (schedule == XGGM_SCHED_CONSTANT) return 1.0;
    const double l = (x - 1.0) / (warmup - 1.0);
    return l > 0.0 ? l : 0.0;
}
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamArgs& a, float lr, float clip) {
    g *= clip;
    m = m * a.b1 + a.one_b1 * g;
    v = v * a.b2 + a.one_b2 * (g * g);
    float upd = m / (sqrtf(v) + a.eps);
    if (a.wd > 0.f) upd += a.wd * p;
    p -= lr * upd;
}
__global__ void __launch_bounds__(256)
bertadam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                long long n, const AdamArgs a, int vec) {
    pdl_prologue();
    float clip = 1.f;
    if (a.sumsq && a.max_norm > 0.f) {
        const float c = a.max_norm / (sqrtf(a.sumsq[0]) + 1e-6f);
        clip = c < 1.f ? c : 1.f;
    }
    float lr = a.lr;
    if (a.step_dev && a.t_total > 0) {
        const long long step = *reinterpret_cast<volatile long long*>(a.step_dev);
        lr = (float)((double)a.lr * schedule_factor(a.schedule, (double)step / (double)a.t_total, a.warmup));
    }
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    long long done = 0;
    if (vec) {
        const long long n4 = n >> 2;
        float4* p4 = reinterpret_cast<float4*>(p);
        const float4* g4 = reinterpret_cast<const float4*>(g);
        float4* m4 = reinterpret_cast<float4*>(m);
        float4* v4 = reinterpret_cast<float4*>(v);
        for (long long i = tid; i < n4; i += nth) {
            float4 pv = p4[i], mv = m4[i], vv = v4[i];
            const float4 gv = g4[i];
            adam1(pv.x, gv.x, mv.x, vv.x, a, lr, clip);
            adam1(pv.y, gv.y, mv.y, vv.y, a, lr, clip);
            adam1(pv.z, gv.z, mv.z, vv.z, a, lr, clip);
            adam1(pv.w, gv.w, mv.w, vv.w, a, lr, clip);
            p4[i] = pv; m4[i] = mv; v4[i] = vv;
        }
        done = n4 << 2;
    }
    for (long long i = done + tid; i < n; i += nth) adam1(p[i], g[i], m[i], v[i], a, lr, clip);
    if (a.step_dev && a.advance) {
        __syncthreads();                       // every thread of this CTA has read the counter
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) {
                *a.step_dev += 1;
                *a.ticket = 0u;
                __threadfence();
            }
        }
    }
}
int bertadam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double b1, double b2,
                  double eps, double wd, const float* sumsq, double max_norm, const xggm_lr_schedule_t* sched,
                  cudaStream_t st) {
    if (n <= 0 && !(sched && sched->advance)) return XGGM_OK;
    AdamArgs a{(float)lr, (float)b1, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), (float)eps, (float)wd,
               (float)max_norm, sumsq, nullptr, nullptr, 0.0, 0, 0, 0};
    if (sched) {
        XGGM_REQUIRE(sched->step_dev && sched->ticket_dev && sched->schedule >= 0 && sched->schedule <= 2);
        a.step_dev = sched->step_dev;
        a.ticket = sched->ticket_dev;
        a.warmup = sched->warmup;
        a.t_total = sched->t_total;
        a.schedule = sched->schedule;
        a.advance = sched->advance;
    }
    const int vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v);
    XGGM_LAUNCH((bertadam_kernel), stream_grid(n, 2048), 256, 0, st, p, g, m, v, n, a, vec);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

}  // namespace xggm
