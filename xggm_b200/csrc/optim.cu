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



// ================================================================================================
// Data-parallel optimiser step over NVLink / NVSwitch peer memory (one process per GPU):
//     gradient all-reduce (AVG)  +  clip_grad_norm_  +  BertAdam  +  parameter all-gather
// as ONE fused sequence in which every rank owns 1/W of the flat buffers:
//   dp_reduce  publish "my gradient bucket is complete" (epoch flag) to every peer and wait for theirs; reduce-SCATTER:
//              average MY slice over the W peer buckets (16-byte loads straight from peer memory), accumulate its
//              squared norm; the last CTA publishes that partial norm to every peer
//   dp_adam    wait; total norm = sum of the W partials in rank order (bit-identical on every rank); BertAdam on MY
//                 slice only (1/W of the update work) and PUSH the new parameters into every peer's parameter buffer
//                 (all-gather by remote stores); the last CTA publishes "my slice has landed"
//   dp_end     wait until every peer's slice has landed here; advance the epoch
// Compared with ncclAllReduce + sumsq + update this moves the same gradient bytes once over the fabric (reduce-scatter)
// plus the parameters once (all-gather) -- the volume of one ring all-reduce -- but drops the separate norm pass,
// does 1/W of the Adam work per GPU, and has no collective-launch latency: flags are plain system-scope words in peer
// memory (torch symmetric memory allocates and exchanges the buffers; this file only sees raw pointers).
// Every wait is bounded: a peer that never arrives becomes a CUDA error after 2 s, never a hung GPU.
// ================================================================================================
constexpr int DP_MAX_RANKS = 16;
constexpr int DP_CTL_EPOCH = 0, DP_CTL_TICKET = 1, DP_CTL_SLICE_SUMSQ = 2, DP_CTL_TICKET2 = 3, DP_CTL_FLAG_A = 16, DP_CTL_FLAG_B = 32, DP_CTL_FLAG_C = 48,
              DP_CTL_PARTIAL = 64;   // uint32 word offsets inside a rank's control block (XGGM_DP_CTL_BYTES)
struct DpPeers {
    int rank, world;
    float* grad[DP_MAX_RANKS];
    float* param[DP_MAX_RANKS];
    unsigned int* ctl[DP_MAX_RANKS];
    float* grad_mc;     // NVSwitch MULTICAST address of the gradient buckets (all ranks' copies at once), or null
    float* param_mc;    // ... of the parameter buffers
};
struct DpRanges {
    int count;
    long long lo[XGGM_DP_MAX_RANGES], hi[XGGM_DP_MAX_RANGES];
    float lr[XGGM_DP_MAX_RANGES];     // base learning rate of each range (parameter groups sharing one bucket)
};
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {   // peer memory: bypass L1, no read-only path
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
// NVLS: one load that returns the SUM of the addressed 16 bytes over every GPU of the multicast group (the reduction
// happens inside the switch), and one store that lands in every GPU's copy.
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_peer_f4(float* p, const float4& v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// block-wide: wait until every peer's flag word (base + k) has reached `epoch`
__device__ __forceinline__ void dp_wait_flags(const unsigned int* ctl, int base, int world, unsigned int epoch) {
    if ((int)threadIdx.x < world) {
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(ld_acquire_sys(ctl + base + threadIdx.x) - epoch) < 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 2000000000ull) __trap();
        }
    }
    __syncthreads();
}
// K2: publish "my gradient bucket is complete", wait for every peer's, reduce-scatter my slice, and (last CTA) publish
// the slice's squared norm.  The epoch word only changes in dp_end_kernel, so every CTA of K2 / K4 reads the same value.
__global__ void __launch_bounds__(256) dp_reduce_kernel(const DpPeers pe, long long slice_lo, long long slice_hi) {
    unsigned int* ctl = pe.ctl[pe.rank];
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(ctl + DP_CTL_EPOCH) + 1u;
    if (blockIdx.x == 0) {
        __threadfence_system();     // the backward kernels' gradient writes (earlier in this stream) before the flag
        if ((int)threadIdx.x < pe.world) st_release_sys(pe.ctl[threadIdx.x] + DP_CTL_FLAG_A + pe.rank, epoch);
    }
    dp_wait_flags(ctl, DP_CTL_FLAG_A, pe.world, epoch);
    __shared__ float part[8];
    const float inv = 1.f / (float)pe.world;
    float acc = 0.f;
    float* mine = pe.grad[pe.rank];
    const long long n4 = (slice_hi - slice_lo) >> 2;     // slices are multiples of 4 floats
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long e = slice_lo + 4 * i;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pe.grad_mc) {
            s = multimem_ld_reduce_f4(pe.grad_mc + e);   // summed inside the NVSwitch: 1/W of the fabric reads
        } else {
            for (int k = 0; k < pe.world; ++k) {         // fixed rank order: the same sum on whichever rank owns the slice
                const float4 v = ld_peer_f4(pe.grad[k] + e);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
        *reinterpret_cast<float4*>(mine + e) = s;
        acc += (s.x * s.x + s.y * s.y) + (s.z * s.z + s.w * s.w);
    }
    const float r = block_sum_256(acc, part);
    __shared__ unsigned int last_s;
    if (threadIdx.x == 0) {
        atomicAdd(reinterpret_cast<float*>(ctl) + DP_CTL_SLICE_SUMSQ, r);
        __threadfence();
        last_s = (atomicAdd(ctl + DP_CTL_TICKET, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (last_s) {       // every CTA's partial is in: publish the slice norm to all peers (one lane per peer)
        const float v = *reinterpret_cast<volatile float*>(reinterpret_cast<float*>(ctl) + DP_CTL_SLICE_SUMSQ);
        __syncthreads();
        if ((int)threadIdx.x < pe.world) {
            unsigned int* peer = pe.ctl[threadIdx.x];
            asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(reinterpret_cast<float*>(peer) + DP_CTL_PARTIAL + pe.rank), "f"(v) : "memory");
            __threadfence_system();
            st_release_sys(peer + DP_CTL_FLAG_B + pe.rank, epoch);
        }
        if (threadIdx.x == 0) {   // reset for the next step
            ctl[DP_CTL_TICKET] = 0u;
            reinterpret_cast<float*>(ctl)[DP_CTL_SLICE_SUMSQ] = 0.f;
        }
    }
}
__global__ void __launch_bounds__(256)
dp_adam_kernel(const DpPeers pe, long long slice_lo, long long slice_hi, const DpRanges rg, float* __restrict__ m,
               float* __restrict__ v, const AdamArgs a, float* __restrict__ sumsq_out) {
    unsigned int* ctl = pe.ctl[pe.rank];
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(ctl + DP_CTL_EPOCH) + 1u;
    dp_wait_flags(ctl, DP_CTL_FLAG_B, pe.world, epoch);
    float total = 0.f;
    for (int k = 0; k < pe.world; ++k) total += *reinterpret_cast<volatile float*>(reinterpret_cast<float*>(ctl) + DP_CTL_PARTIAL + k);
    if (blockIdx.x == 0 && threadIdx.x == 0 && sumsq_out) sumsq_out[0] = total;
    float clip = 1.f;
    if (a.max_norm > 0.f) {
        const float c = a.max_norm / (sqrtf(total) + 1e-6f);
        clip = c < 1.f ? c : 1.f;
    }
    double sched = 1.0;
    if (a.step_dev && a.t_total > 0) {
        const long long step = *reinterpret_cast<volatile long long*>(a.step_dev);
        sched = schedule_factor(a.schedule, (double)step / (double)a.t_total, a.warmup);
    }
    float* p_mine = pe.param[pe.rank];
    const float* g_mine = pe.grad[pe.rank];
    for (int r = 0; r < rg.count; ++r) {                  // the active ranges of the bucket, clipped to my slice
        const long long lo = rg.lo[r] > slice_lo ? rg.lo[r] : slice_lo, hi = rg.hi[r] < slice_hi ? rg.hi[r] : slice_hi;
        if (hi <= lo) continue;
        const float lr = (float)((double)rg.lr[r] * sched);
        const long long n4 = (hi - lo) >> 2;              // range bounds are multiples of 32 floats (FlatGrads alignment)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            const long long e = lo + 4 * i;
            float4 pv = *reinterpret_cast<const float4*>(p_mine + e);
            const float4 gv = *reinterpret_cast<const float4*>(g_mine + e);
            float4 mv = *reinterpret_cast<float4*>(m + e), vv = *reinterpret_cast<float4*>(v + e);
            adam1(pv.x, gv.x, mv.x, vv.x, a, lr, clip);
            adam1(pv.y, gv.y, mv.y, vv.y, a, lr, clip);
            adam1(pv.z, gv.z, mv.z, vv.z, a, lr, clip);
            adam1(pv.w, gv.w, mv.w, vv.w, a, lr, clip);
            *reinterpret_cast<float4*>(m + e) = mv;
            *reinterpret_cast<float4*>(v + e) = vv;
            if (pe.param_mc) {
                multimem_st_f4(pe.param_mc + e, pv);      // one store, every rank's copy (own included)
            } else {
                for (int k = 0; k < pe.world; ++k) {      // all-gather by remote stores (own copy included)
                    const int kk = (pe.rank + k) % pe.world;  // start at home: spreads the fabric traffic over the peers
                    st_peer_f4(pe.param[kk] + e, pv);
                }
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(ctl + DP_CTL_TICKET2, 1u) == gridDim.x - 1) {   // last CTA of this rank: every slice element is on its way
            __threadfence_system();
            ctl[DP_CTL_TICKET2] = 0u;
            if (a.step_dev && a.advance) *a.step_dev += 1;
            for (int k = 0; k < pe.world; ++k) st_release_sys(pe.ctl[k] + DP_CTL_FLAG_C + pe.rank, epoch);
        }
    }
}
__global__ void __launch_bounds__(32) dp_end_kernel(const DpPeers pe) {
    unsigned int* ctl = pe.ctl[pe.rank];
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(ctl + DP_CTL_EPOCH) + 1u;
    dp_wait_flags(ctl, DP_CTL_FLAG_C, pe.world, epoch);
    __threadfence_system();
    if (threadIdx.x == 0) ctl[DP_CTL_EPOCH] = epoch;    // the step is over on this rank: next call, next epoch
}

int dp_bertadam_step(const xggm_dp_peers_t* peers, float* m, float* v, long long n, const long long* range_lo,
                     const long long* range_hi, const double* range_lr, int n_ranges, double lr, double b1, double b2, double eps, double wd,
                     double max_norm, const xggm_lr_schedule_t* sched, float* sumsq_out, cudaStream_t st) {
    XGGM_REQUIRE(peers && m && v && n > 0 && peers->world >= 1 && peers->world <= DP_MAX_RANKS && peers->rank >= 0 &&
                 peers->rank < peers->world && n_ranges >= 0 && n_ranges <= XGGM_DP_MAX_RANGES && (n_ranges == 0 || (range_lo && range_hi)));
    DpPeers pe;
    pe.rank = peers->rank; pe.world = peers->world;
    pe.grad_mc = static_cast<float*>(peers->grad_multicast);
    pe.param_mc = static_cast<float*>(peers->param_multicast);
    XGGM_REQUIRE(aligned16(pe.grad_mc) && aligned16(pe.param_mc));
    for (int k = 0; k < DP_MAX_RANKS; ++k) {
        const bool live = k < pe.world;
        pe.grad[k] = live ? static_cast<float*>(peers->grad[k]) : nullptr;
        pe.param[k] = live ? static_cast<float*>(peers->param[k]) : nullptr;
        pe.ctl[k] = live ? static_cast<unsigned int*>(peers->ctl[k]) : nullptr;
        if (live) XGGM_REQUIRE(pe.grad[k] && pe.param[k] && pe.ctl[k] && aligned16(pe.grad[k]) && aligned16(pe.param[k]));
    }
    XGGM_REQUIRE(aligned16(m) && aligned16(v));
    DpRanges rg;
    rg.count = n_ranges;
    for (int r = 0; r < n_ranges; ++r) {
        XGGM_REQUIRE(range_lo[r] % 4 == 0 && range_hi[r] % 4 == 0 && range_lo[r] >= 0 && range_hi[r] <= n);
        rg.lo[r] = range_lo[r]; rg.hi[r] = range_hi[r];
        rg.lr[r] = (float)(range_lr ? range_lr[r] : lr);
    }
    // equal slices, 4-float granularity (the bucket length is a multiple of 32)
    const long long per = ((n + pe.world - 1) / pe.world + 3) & ~3LL;
    const long long lo = per * pe.rank < n ? per * pe.rank : n, hi = lo + per < n ? lo + per : n;
    XGGM_REQUIRE(n % 4 == 0);
    AdamArgs a{(float)lr, (float)b1, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), (float)eps, (float)wd,
               (float)max_norm, nullptr, nullptr, nullptr, 0.0, 0, 0, 0};
    if (sched) {
        XGGM_REQUIRE(sched->step_dev && sched->schedule >= 0 && sched->schedule <= 2);
        a.step_dev = sched->step_dev; a.warmup = sched->warmup; a.t_total = sched->t_total;
        a.schedule = sched->schedule; a.advance = sched->advance;
    }
    const int grid = stream_grid(hi - lo, 2048);
    XGGM_LAUNCH((dp_reduce_kernel), grid, 256, 0, st, pe, lo, hi);
    XGGM_LAUNCH_CHECK();
    XGGM_LAUNCH((dp_adam_kernel), grid, 256, 0, st, pe, lo, hi, rg, m, v, a, sumsq_out);
    XGGM_LAUNCH_CHECK();
    XGGM_LAUNCH((dp_end_kernel), 1, 32, 0, st, pe);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

}  // namespace xggm
