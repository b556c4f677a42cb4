// Row-wise fused epilogues of the graph block: LayerNorm and the jump-knowledge head
// tail  dropout(LN(GeLU(z))), forward and backward.
//
// Fast path (H % 128 == 0, H <= 1024; H = 768 in X-GGM): one warp owns one row and keeps it in
// registers -- lane l holds the float4 at columns 128*i + 4*l -- so every tensor is read exactly
// once with 512-byte coalesced warp accesses, GeLU/erf is evaluated once per element, statistics
// are fp32 two-pass (mean, then centred sum of squares, as torch's LayerNorm) via warp shuffles,
// and the column-wise parameter gradients (gamma / beta / bias) accumulate in registers across
// the rows a warp walks, are combined across the CTA's warps in shared memory and flushed with
// one atomicAdd per column per CTA.  Each kernel can also emit the bf16 hi/lo planes of its
// output (the tensor-core operand format of gemm_tc.cu), which removes a separate split pass.
// Generic path (any H): the same math with strided loops.
#include <cuda_bf16.h>

#include "common.cuh"

namespace xggm {

typedef __nv_bfloat16 bf16;

constexpr int ROW_WARPS = 8;  // warps per CTA (fast path)

template <int NV>
__device__ __forceinline__ void load_row(const float* __restrict__ p, int lane, float (&v)[NV * 4]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(p + 128 * i + 4 * lane);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
template <int NV>
__device__ __forceinline__ void store_row(float* __restrict__ p, int lane, const float (&v)[NV * 4]) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
        *reinterpret_cast<float4*>(p + 128 * i + 4 * lane) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// bf16 STORAGE rows (bf16 engine: saved z / xhat are kept as bf16 only): lane l holds elements 128 i + 4 l .. + 3
template <int NV>
__device__ __forceinline__ void load_row_bf16(const bf16* __restrict__ p, int lane, float (&v)[NV * 4]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const uint2 t = *reinterpret_cast<const uint2*>(p + 128 * i + 4 * lane);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x), b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
        v[4 * i] = __low2float(a); v[4 * i + 1] = __high2float(a); v[4 * i + 2] = __low2float(b); v[4 * i + 3] = __high2float(b);
    }
}
template <int NV>
__device__ __forceinline__ void store_row_bf16(bf16* __restrict__ p, int lane, const float (&v)[NV * 4]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[4 * i], v[4 * i + 1]), b = __floats2bfloat162_rn(v[4 * i + 2], v[4 * i + 3]);
        *reinterpret_cast<uint2*>(p + 128 * i + 4 * lane) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
}
template <int NV>
__device__ __forceinline__ void store_planes(bf16* hi, bf16* lo, size_t row_off, int lane, const float (&v)[NV * 4]) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
        split_store4(hi, lo, row_off + 128 * i + 4 * lane, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <int NV>
__device__ __forceinline__ void row_stats(const float (&v)[NV * 4], float inv_h, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) s += v[i];
    mean = warp_sum(s) * inv_h;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    rstd = 1.0f / sqrtf(warp_sum(q) * inv_h + eps);
}
// combine per-warp register column accumulators across the CTA and add them to dst[H]
template <int NV>
__device__ __forceinline__ void flush_columns(const float (&acc)[NV * 4], float* __restrict__ dst, float* sm) {
    constexpr int H = NV * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i)
        *reinterpret_cast<float4*>(sm + warp * H + 128 * i + 4 * lane) =
            make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < ROW_WARPS; ++w) a += sm[w * H + c];
        atomicAdd(&dst[c], a);
    }
}

// ================================================================== fast kernels
// Per-warp two-stage row prefetch: lane 0 launches 1-D bulk async copies (cp.async.bulk + mbarrier
// complete_tx) of the NEXT row of up to NT input tensors into this warp's shared-memory stage while
// the warp computes on the current one, so a warp never sits on a DRAM round trip between rows.
// E1: bytes per element of tensor 1 (4, or 2 when that tensor is a bf16 storage row; its stage slot keeps H floats).
template <int NV, int NT, int E1 = 4>
struct RowPrefetch {
    static constexpr int H = NV * 128;
    static constexpr int STAGE = NT * H;                          // floats per stage
    static constexpr int WARP_FLOATS = 2 * STAGE;
    static constexpr size_t SMEM = sizeof(float) * ROW_WARPS * WARP_FLOATS + sizeof(uint64_t) * ROW_WARPS * 2;
    float* buf;
    uint64_t* bar;
    __device__ __forceinline__ void init(float* sm, int warp, int lane) {
        buf = sm + (size_t)warp * WARP_FLOATS;
        bar = reinterpret_cast<uint64_t*>(sm + (size_t)ROW_WARPS * WARP_FLOATS) + warp * 2;
        if (lane == 0) {
            ptx::mbar_init(&bar[0], 1);
            ptx::mbar_init(&bar[1], 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
    }
    // src[t] = row pointer of tensor t (null -> skipped)
    __device__ __forceinline__ void issue(int stage, const float* const (&src)[NT], int lane) {
        if (lane == 0) {
            uint32_t bytes = 0;
#pragma unroll
            for (int t = 0; t < NT; ++t) bytes += src[t] ? H * (t == 1 ? E1 : 4) : 0;
            ptx::fence_proxy_async();   // the stage was last read through the generic proxy
            ptx::mbar_expect_tx(&bar[stage], bytes);
#pragma unroll
            for (int t = 0; t < NT; ++t)
                if (src[t]) ptx::bulk_g2s(buf + stage * STAGE + t * H, src[t], H * (t == 1 ? E1 : 4), &bar[stage]);
        }
    }
    __device__ __forceinline__ void wait(int stage, uint32_t phase) { ptx::mbar_wait(&bar[stage], phase); }
    __device__ __forceinline__ void read(int stage, int t, int lane, float (&v)[NV * 4]) {
        if (t == 1 && E1 == 2) load_row_bf16<NV>(reinterpret_cast<const bf16*>(buf + stage * STAGE + t * H), lane, v);
        else load_row<NV>(buf + stage * STAGE + t * H, lane, v);
    }
};

// XB: xhat is stored as bf16 (bf16 engine).  h may be null when only its operand planes are wanted.
template <int NV, bool XB>
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_fwd_fast(const float* __restrict__ u, const float* __restrict__ gamma, const float* __restrict__ beta,
            float* __restrict__ h, float* __restrict__ xhat, float* __restrict__ rstd_out,
            bf16* __restrict__ hi, bf16* __restrict__ lo, int M, float eps) {
    pdl_prologue();
    constexpr int H = NV * 128;
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wid = blockIdx.x * ROW_WARPS + warp, nw = gridDim.x * ROW_WARPS;
    RowPrefetch<NV, 1> pf;
    pf.init(sm, warp, lane);
    if (wid < M) {
        const float* src[1] = {u + (size_t)wid * H};
        pf.issue(0, src, lane);
    }
    // gamma / beta stay in registers for all rows of this warp (re-loading them per row cost a third of the
    // kernel's stall samples: an L1 round trip in front of every row's stores)
    float gam[NV * 4], bet[NV * 4];
    load_row<NV>(gamma, lane, gam);
    load_row<NV>(beta, lane, bet);
    int it = 0;
    for (int r = wid; r < M; r += nw, ++it) {
        const int stage = it & 1;
        if (r + nw < M) {
            const float* src[1] = {u + (size_t)(r + nw) * H};
            pf.issue(stage ^ 1, src, lane);
        }
        pf.wait(stage, (it >> 1) & 1);
        float v[NV * 4];
        pf.read(stage, 0, lane, v);
        __syncwarp();
        float mean, rstd;
        row_stats<NV>(v, 1.0f / (float)H, eps, mean, rstd);
#pragma unroll
        for (int i = 0; i < NV * 4; ++i) v[i] = (v[i] - mean) * rstd;
        if (xhat) {
            if (XB) store_row_bf16<NV>(reinterpret_cast<bf16*>(xhat) + (size_t)r * H, lane, v);
            else store_row<NV>(xhat + (size_t)r * H, lane, v);
        }
#pragma unroll
        for (int i = 0; i < NV * 4; ++i) v[i] = fmaf(v[i], gam[i], bet[i]);
        if (h) store_row<NV>(h + (size_t)r * H, lane, v);
        if (hi) store_planes<NV>(hi, lo, (size_t)r * H, lane, v);
        if (lane == 0 && rstd_out) rstd_out[r] = rstd;
    }
}

template <int NV, bool XB>
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_bwd_fast(const float* __restrict__ gh, const float* __restrict__ xhat, const float* __restrict__ rstd,
            const float* __restrict__ gamma, float* __restrict__ gu, float* __restrict__ ggamma,
            float* __restrict__ gbeta, bf16* __restrict__ hi, bf16* __restrict__ lo, int M) {
    pdl_prologue();
    constexpr int H = NV * 128;
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wid = blockIdx.x * ROW_WARPS + warp, nw = gridDim.x * ROW_WARPS;
    RowPrefetch<NV, 2, XB ? 2 : 4> pf;
    constexpr int XS = XB ? 2 : 1;   // xhat rows advance by H elements of 2 (bf16) or 4 bytes: index in half-floats
    pf.init(sm, warp, lane);
    if (wid < M) {
        const float* src[2] = {gh + (size_t)wid * H, xhat + (size_t)wid * H / XS};
        pf.issue(0, src, lane);
    }
    float gam[NV * 4], ag[NV * 4], ab[NV * 4];
    load_row<NV>(gamma, lane, gam);
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
    int it = 0;
    for (int r = wid; r < M; r += nw, ++it) {
        const int stage = it & 1;
        if (r + nw < M) {
            const float* src[2] = {gh + (size_t)(r + nw) * H, xhat + (size_t)(r + nw) * H / XS};
            pf.issue(stage ^ 1, src, lane);
        }
        const float rs = rstd[r];
        pf.wait(stage, (it >> 1) & 1);
        float g[NV * 4], xh[NV * 4];
        pf.read(stage, 0, lane, g);
        pf.read(stage, 1, lane, xh);
        __syncwarp();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV * 4; ++i) {
            const float d = g[i] * gam[i];
            s1 += d;
            s2 = fmaf(d, xh[i], s2);
        }
        const float c1 = warp_sum(s1) / (float)H, c2 = warp_sum(s2) / (float)H;
#pragma unroll
        for (int i = 0; i < NV * 4; ++i) {
            ag[i] = fmaf(g[i], xh[i], ag[i]);
            ab[i] += g[i];
            g[i] = rs * (g[i] * gam[i] - c1 - xh[i] * c2);
        }
        store_row<NV>(gu + (size_t)r * H, lane, g);
        if (hi) store_planes<NV>(hi, lo, (size_t)r * H, lane, g);
    }
    flush_columns<NV>(ag, ggamma, sm);
    flush_columns<NV>(ab, gbeta, sm);
}

__device__ __forceinline__ void load_keep4(const uint8_t* __restrict__ k, size_t off, bool (&m)[4]) {
    const uchar4 t = *reinterpret_cast<const uchar4*>(k + off);
    m[0] = t.x != 0; m[1] = t.y != 0; m[2] = t.z != 0; m[3] = t.w != 0;
}

// keep decision for the 4 consecutive elements starting at flat offset `off` (off % 4 == 0):
// mode 1 reads the uint8 mask, mode 2 draws them with Philox exactly as keep_mask_kernel would
// (counter = off / 4, subsequence = stream), so a mask tensor never has to exist.
__device__ __forceinline__ void drop_bits4(const DropSpec& d, uint64_t stream, size_t off, bool (&m)[4]) {
    if (d.mode == 1) {
        load_keep4(d.keep, off, m);
    } else if (d.thresh == XGGM_COIN_THRESH) {
        const uint32_t nib = coin_nibble(Philox(d.seed), stream, off);
        m[0] = nib & 1; m[1] = nib & 2; m[2] = nib & 4; m[3] = nib & 8;
    } else {
        const uint4 r = Philox(d.seed)((uint64_t)(off >> 2), stream);
        m[0] = r.x >= d.thresh; m[1] = r.y >= d.thresh; m[2] = r.z >= d.thresh; m[3] = r.w >= d.thresh;
    }
}
// Fast kernels, coin mapping: a row of H = 128 NV elements needs NV Philox blocks in total.  Lane i < NV
// draws block i of the row; every lane then fetches, for each i, the word that holds its 4 bits with warp
// shuffles (4 SHFL + 3 SEL per block instead of a 10-round Philox per lane).  ro = row offset (ro % 128 == 0).
template <int NV>
__device__ __forceinline__ void coin_row(const DropSpec& d, uint64_t stream, size_t ro, int lane, uint32_t (&nib)[NV]) {
    const uint4 r = Philox(d.seed)((uint64_t)(ro >> 7) + (uint64_t)(lane < NV ? lane : 0), stream);
    const uint32_t w = (uint32_t)lane >> 3, sh = ((uint32_t)lane & 7u) * 4u;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const uint32_t x = __shfl_sync(0xffffffffu, r.x, i), y = __shfl_sync(0xffffffffu, r.y, i);
        const uint32_t z = __shfl_sync(0xffffffffu, r.z, i), t = __shfl_sync(0xffffffffu, r.w, i);
        const uint32_t word = w == 0 ? x : (w == 1 ? y : (w == 2 ? z : t));
        nib[i] = (word >> sh) & 0xFu;
    }
}
__device__ __forceinline__ bool drop_bit1(const DropSpec& d, uint64_t stream, size_t off) {
    if (d.mode == 1) return d.keep[off] != 0;
    if (d.thresh == XGGM_COIN_THRESH) return (coin_nibble(Philox(d.seed), stream, off & ~(size_t)3) >> (off & 3)) & 1u;
    const uint4 r = Philox(d.seed)((uint64_t)(off >> 2), stream);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    return w[off & 3] >= d.thresh;
}
__device__ __forceinline__ uint64_t drop_stream(const DropSpec& d) {
    return d.stream + ((d.mode == 2 && d.epoch) ? (*d.epoch << 32) : 0ull);
}

// ZB: z is a bf16 storage row (bf16 engine)
template <int NV, bool ZB>
__global__ void __launch_bounds__(ROW_WARPS * 32)
gld_fwd_fast(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
             const DropSpec drop, float* __restrict__ out,
             float* __restrict__ mean_out, float* __restrict__ rstd_out, bf16* __restrict__ hi,
             bf16* __restrict__ lo, int M, float eps, int accumulate) {
    pdl_prologue();
    constexpr int H = NV * 128;
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t dstream = drop_stream(drop);
    const float scale = drop.scale;
    const int wid = blockIdx.x * ROW_WARPS + warp, nw = gridDim.x * ROW_WARPS;
    RowPrefetch<NV, 2, ZB ? 2 : 4> pf;   // tensor 0 = previous out (accumulate), tensor 1 = z
    constexpr int ZS = ZB ? 2 : 1;
    pf.init(sm, warp, lane);
    if (wid < M) {
        const float* src[2] = {accumulate ? out + (size_t)wid * H : nullptr, z + (size_t)wid * H / ZS};
        pf.issue(0, src, lane);
    }
    int it = 0;
    for (int r = wid; r < M; r += nw, ++it) {
        const int stage = it & 1;
        const size_t ro = (size_t)r * H;
        if (r + nw < M) {
            const float* src[2] = {accumulate ? out + (size_t)(r + nw) * H : nullptr, z + (size_t)(r + nw) * H / ZS};
            pf.issue(stage ^ 1, src, lane);
        }
        pf.wait(stage, (it >> 1) & 1);
        float v[NV * 4], o[NV * 4];
        pf.read(stage, 1, lane, v);
        if (accumulate) pf.read(stage, 0, lane, o);
        __syncwarp();   // every lane has drained the stage before lane 0 refills it next iteration
#pragma unroll
        for (int i = 0; i < NV * 4; ++i) v[i] = gelu_fast(v[i]);
        float mean, rstd;
        row_stats<NV>(v, 1.0f / (float)H, eps, mean, rstd);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 g = *reinterpret_cast<const float4*>(gamma + 128 * i + 4 * lane);
            const float4 b = *reinterpret_cast<const float4*>(beta + 128 * i + 4 * lane);
            v[4 * i] = fmaf((v[4 * i] - mean) * rstd, g.x, b.x);
            v[4 * i + 1] = fmaf((v[4 * i + 1] - mean) * rstd, g.y, b.y);
            v[4 * i + 2] = fmaf((v[4 * i + 2] - mean) * rstd, g.z, b.z);
            v[4 * i + 3] = fmaf((v[4 * i + 3] - mean) * rstd, g.w, b.w);
        }
        if (drop.mode == 2 && drop.thresh == XGGM_COIN_THRESH) {
            uint32_t nib[NV];
            coin_row<NV>(drop, dstream, ro, lane, nib);
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[4 * i + j] = ((nib[i] >> j) & 1u) ? v[4 * i + j] * scale : 0.f;
        } else if (drop.mode) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                bool m[4];
                drop_bits4(drop, dstream, ro + 128 * i + 4 * lane, m);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[4 * i + j] = m[j] ? v[4 * i + j] * scale : 0.f;
            }
        }
        if (accumulate) {
#pragma unroll
            for (int i = 0; i < NV * 4; ++i) v[i] += o[i];
        }
        store_row<NV>(out + ro, lane, v);
        if (hi) store_planes<NV>(hi, lo, ro, lane, v);
        if (lane == 0) {
            if (mean_out) mean_out[r] = mean;
            if (rstd_out) rstd_out[r] = rstd;
        }
    }
}

template <int NV, bool ZB>
__global__ void __launch_bounds__(ROW_WARPS * 32, 1)
gld_bwd_fast(const float* __restrict__ gout, const float* __restrict__ z, const float* __restrict__ mean,
             const float* __restrict__ rstd, const float* __restrict__ gamma, const DropSpec drop,
             float* __restrict__ gz, float* __restrict__ ggamma, float* __restrict__ gbeta,
             float* __restrict__ gbias, bf16* __restrict__ hi, bf16* __restrict__ lo, int M) {
    pdl_prologue();
    constexpr int H = NV * 128;
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t dstream = drop_stream(drop);
    const float scale = drop.scale;
    const int wid = blockIdx.x * ROW_WARPS + warp, nw = gridDim.x * ROW_WARPS;
    RowPrefetch<NV, 2, ZB ? 2 : 4> pf;
    constexpr int ZS = ZB ? 2 : 1;
    pf.init(sm, warp, lane);
    if (wid < M) {
        const float* src[2] = {gout + (size_t)wid * H, z + (size_t)wid * H / ZS};
        pf.issue(0, src, lane);
    }
    float ag[NV * 4], ab[NV * 4], az[NV * 4];
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) { ag[i] = 0.f; ab[i] = 0.f; az[i] = 0.f; }
    int it = 0;
    for (int r = wid; r < M; r += nw, ++it) {
        const int stage = it & 1;
        const size_t ro = (size_t)r * H;
        if (r + nw < M) {
            const float* src[2] = {gout + (size_t)(r + nw) * H, z + (size_t)(r + nw) * H / ZS};
            pf.issue(stage ^ 1, src, lane);
        }
        const float mu = mean[r], rs = rstd[r];
        pf.wait(stage, (it >> 1) & 1);
        float gy[NV * 4], zv[NV * 4], yh[NV * 4];
        pf.read(stage, 0, lane, gy);
        pf.read(stage, 1, lane, zv);
        __syncwarp();
        if (drop.mode == 2 && drop.thresh == XGGM_COIN_THRESH) {
            uint32_t nib[NV];
            coin_row<NV>(drop, dstream, ro, lane, nib);
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) gy[4 * i + j] = ((nib[i] >> j) & 1u) ? gy[4 * i + j] * scale : 0.f;
        } else if (drop.mode) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                bool m[4];
                drop_bits4(drop, dstream, ro + 128 * i + 4 * lane, m);
#pragma unroll
                for (int j = 0; j < 4; ++j) gy[4 * i + j] = m[j] ? gy[4 * i + j] * scale : 0.f;
            }
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 gm = *reinterpret_cast<const float4*>(gamma + 128 * i + 4 * lane);
            const float gmv[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = 4 * i + j;
                // one erf per element: cdf feeds both gelu(z) and its derivative
                float cdf, pdf;
                normal_cdf_pdf(zv[e], cdf, pdf);
                yh[e] = (zv[e] * cdf - mu) * rs;
                zv[e] = cdf + zv[e] * pdf;  // gelu'(z)
                ag[e] = fmaf(gy[e], yh[e], ag[e]);
                ab[e] += gy[e];
                gy[e] *= gmv[j];            // d = gy * gamma
                s1 += gy[e];
                s2 = fmaf(gy[e], yh[e], s2);
            }
        }
        const float c1 = warp_sum(s1) / (float)H, c2 = warp_sum(s2) / (float)H;
#pragma unroll
        for (int i = 0; i < NV * 4; ++i) {
            gy[i] = rs * (gy[i] - c1 - yh[i] * c2) * zv[i];
            az[i] += gy[i];
        }
        if (gz) store_row<NV>(gz + ro, lane, gy);
        if (hi) store_planes<NV>(hi, lo, ro, lane, gy);
    }
    // the prefetch stages are idle now: reuse their shared memory for the column reduction
    flush_columns<NV>(ag, ggamma, sm);
    flush_columns<NV>(ab, gbeta, sm);
    if (gbias) flush_columns<NV>(az, gbias, sm);
}

// ================================================================== generic kernels (any H)
constexpr int GEN_WARPS = 4;
constexpr int GEN_ROWS = 8;
__device__ __forceinline__ void row_range(int M, int& r0, int& r1) {
    const int base = (blockIdx.x * GEN_WARPS + (threadIdx.x >> 5)) * GEN_ROWS;
    r0 = base;
    r1 = min(M, base + GEN_ROWS);
}
static inline int gen_grid(int M) { return ceil_div(M, GEN_WARPS * GEN_ROWS); }

__device__ __forceinline__ void split_store1(bf16* hi, bf16* lo, size_t o, float v) {
    const bf16 h = __float2bfloat16_rn(v);
    hi[o] = h;
    if (lo) {
        const float hf = __bfloat162float(h);
        lo[o] = __float2bfloat16_rn((hf - hf == 0.f) ? v - hf : 0.f);
    }
}

__global__ void __launch_bounds__(GEN_WARPS * 32)
ln_fwd_gen(const float* __restrict__ u, const float* __restrict__ gamma, const float* __restrict__ beta,
           float* __restrict__ h, float* __restrict__ xhat, float* __restrict__ rstd_out,
           bf16* __restrict__ hi, bf16* __restrict__ lo, int M, int H, float eps) {
    pdl_prologue();
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
            const size_t o = (size_t)r * H + c;
            const float xh = (ur[c] - mean) * rstd;
            if (xhat) xhat[o] = xh;
            const float y = fmaf(xh, gamma[c], beta[c]);
            h[o] = y;
            if (hi) split_store1(hi, lo, o, y);
        }
        if (lane == 0 && rstd_out) rstd_out[r] = rstd;
    }
}

// smem: [GEN_WARPS][3][H] private column accumulators
__global__ void __launch_bounds__(GEN_WARPS * 32)
ln_bwd_gen(const float* __restrict__ gh, const float* __restrict__ xhat, const float* __restrict__ rstd,
           const float* __restrict__ gamma, float* __restrict__ gu, float* __restrict__ ggamma,
           float* __restrict__ gbeta, bf16* __restrict__ hi, bf16* __restrict__ lo, int M, int H) {
    pdl_prologue();
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sg = sm + (size_t)warp * 3 * H;
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
            const size_t o = (size_t)r * H + c;
            const float g = gr[c], xh = xr[c];
            const float v = rs * (g * gamma[c] - c1 - xh * c2);
            gu[o] = v;
            if (hi) split_store1(hi, lo, o, v);
            sg[c] = fmaf(g, xh, sg[c]);
            sb[c] += g;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < GEN_WARPS; ++w) { a += sm[(size_t)w * 3 * H + c]; b += sm[(size_t)w * 3 * H + H + c]; }
        atomicAdd(&ggamma[c], a);
        atomicAdd(&gbeta[c], b);
    }
}

__global__ void __launch_bounds__(GEN_WARPS * 32)
gld_fwd_gen(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
            const DropSpec drop, float* __restrict__ out,
            float* __restrict__ mean_out, float* __restrict__ rstd_out, bf16* __restrict__ hi,
            bf16* __restrict__ lo, int M, int H, float eps, int accumulate) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const uint64_t dstream = drop_stream(drop);
    const float scale = drop.scale;
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
            if (drop.mode) y = drop_bit1(drop, dstream, o) ? y * scale : 0.f;
            if (accumulate) y += out[o];
            out[o] = y;
            if (hi) split_store1(hi, lo, o, y);
        }
        if (lane == 0) {
            if (mean_out) mean_out[r] = mean;
            if (rstd_out) rstd_out[r] = rstd;
        }
    }
}

__global__ void __launch_bounds__(GEN_WARPS * 32)
gld_bwd_gen(const float* __restrict__ gout, const float* __restrict__ z, const float* __restrict__ mean,
            const float* __restrict__ rstd, const float* __restrict__ gamma, const DropSpec drop,
            float* __restrict__ gz, float* __restrict__ ggamma, float* __restrict__ gbeta,
            float* __restrict__ gbias, bf16* __restrict__ hi, bf16* __restrict__ lo, int M, int H) {
    pdl_prologue();
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t dstream = drop_stream(drop);
    const float scale = drop.scale;
    float* sg = sm + (size_t)warp * 3 * H;
    float* sb = sg + H;
    float* sz = sb + H;
    for (int c = lane; c < H; c += 32) { sg[c] = 0.f; sb[c] = 0.f; sz[c] = 0.f; }
    int r0, r1;
    row_range(M, r0, r1);
    for (int r = r0; r < r1; ++r) {
        const float* gr = gout + (size_t)r * H;
        const float* zr = z + (size_t)r * H;
        const float mu = mean[r], rs = rstd[r];
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < H; c += 32) {
            float gy = gr[c];
            if (drop.mode) gy = drop_bit1(drop, dstream, (size_t)r * H + c) ? gy * scale : 0.f;
            const float yh = (gelu_erf(zr[c]) - mu) * rs;
            const float d = gy * gamma[c];
            s1 += d;
            s2 = fmaf(d, yh, s2);
        }
        const float c1 = warp_sum(s1) / (float)H, c2 = warp_sum(s2) / (float)H;
        for (int c = lane; c < H; c += 32) {
            const size_t o = (size_t)r * H + c;
            float gy = gr[c];
            if (drop.mode) gy = drop_bit1(drop, dstream, o) ? gy * scale : 0.f;
            const float zv = zr[c];
            const float yh = (gelu_erf(zv) - mu) * rs;
            const float v = rs * (gy * gamma[c] - c1 - yh * c2) * gelu_erf_grad(zv);
            if (gz) gz[o] = v;
            if (hi) split_store1(hi, lo, o, v);
            sg[c] = fmaf(gy, yh, sg[c]);
            sb[c] += gy;
            sz[c] += v;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int w = 0; w < GEN_WARPS; ++w) {
            a += sm[(size_t)w * 3 * H + c];
            b += sm[(size_t)w * 3 * H + H + c];
            d += sm[(size_t)w * 3 * H + 2 * H + c];
        }
        atomicAdd(&ggamma[c], a);
        atomicAdd(&gbeta[c], b);
        if (gbias) atomicAdd(&gbias[c], d);
    }
}

// ================================================================== launchers
static int row_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}
static int fast_grid(int M, int ctas_per_sm = 2) {
    return max(1, min(ceil_div(M, ROW_WARPS), ctas_per_sm * row_sms()));
}
static inline bool fast_ok(int H, const void* a, const void* b = nullptr, const void* c = nullptr) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return H % 128 == 0 && H >= 128 && H <= 1024 && al(a) && al(b) && al(c);
}
#define XGGM_ROW_DISPATCH(H, ...)           \
    switch ((H) / 128) {                    \
        case 1: { constexpr int NV = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int NV = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int NV = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int NV = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int NV = 5; __VA_ARGS__; } break; \
        case 6: { constexpr int NV = 6; __VA_ARGS__; } break; \
        case 7: { constexpr int NV = 7; __VA_ARGS__; } break; \
        default: { constexpr int NV = 8; __VA_ARGS__; } break; \
    }

template <typename K>
static int gen_smem(K kernel, int H, size_t& bytes) {
    bytes = sizeof(float) * (size_t)GEN_WARPS * 3 * H;
    if (bytes > 227 * 1024) return XGGM_ERR_ARG;
    if (bytes > 48 * 1024)
        XGGM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return XGGM_OK;
}

// hi/lo: optional bf16 planes of the output h (lo may be null with hi set: bf16 engine)
// xhat_bf16: xhat is written as bf16 (fast path only; the bf16 engine's storage mode).  h may be null (fast path).
int layernorm_fwd(const float* u, const float* gamma, const float* beta, float* h, float* xhat, float* rstd,
                  bf16* hi, bf16* lo, int M, int H, float eps, cudaStream_t st, int xhat_bf16) {
    if (M <= 0) return XGGM_OK;
    if (fast_ok(H, u, h, xhat) && fast_ok(H, gamma, beta, hi) && fast_ok(H, lo)) {
        XGGM_ROW_DISPATCH(H, {
            constexpr size_t smem = RowPrefetch<NV, 1>::SMEM;
            if (xhat_bf16) {
                static bool attr = false;
                if (!attr) { XGGM_CUDA_TRY(cudaFuncSetAttribute(ln_fwd_fast<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
                XGGM_LAUNCH((ln_fwd_fast<NV, true>), fast_grid(M, 2), ROW_WARPS * 32, smem, st, u, gamma, beta, h, xhat, rstd, hi, lo, M, eps);
            } else {
                static bool attr = false;
                if (!attr) { XGGM_CUDA_TRY(cudaFuncSetAttribute(ln_fwd_fast<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
                XGGM_LAUNCH((ln_fwd_fast<NV, false>), fast_grid(M, 2), ROW_WARPS * 32, smem, st, u, gamma, beta, h, xhat, rstd, hi, lo, M, eps);
            }
        });
    } else {
        XGGM_REQUIRE(!xhat_bf16 && h);
        XGGM_LAUNCH((ln_fwd_gen), gen_grid(M), GEN_WARPS * 32, 0, st, u, gamma, beta, h, xhat, rstd, hi, lo, M, H, eps);
    }
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ggamma / gbeta are ACCUMULATED into (caller zeroes them)
int layernorm_bwd(const float* gh, const float* xhat, const float* rstd, const float* gamma, float* gu,
                  float* ggamma, float* gbeta, bf16* hi, bf16* lo, int M, int H, cudaStream_t st, int xhat_bf16) {
    if (M <= 0) return XGGM_OK;
    if (fast_ok(H, gh, xhat, gu) && fast_ok(H, gamma, hi, lo)) {
        XGGM_ROW_DISPATCH(H, {
            constexpr size_t smem = RowPrefetch<NV, 2>::SMEM;   // (>= the ROW_WARPS*H floats flush_columns needs)
            if (xhat_bf16) {
                static bool attr = false;
                if (!attr) { XGGM_CUDA_TRY(cudaFuncSetAttribute(ln_bwd_fast<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
                XGGM_LAUNCH((ln_bwd_fast<NV, true>), fast_grid(M, 1), ROW_WARPS * 32, smem, st, gh, xhat, rstd, gamma, gu, ggamma, gbeta, hi, lo, M);
            } else {
                static bool attr = false;
                if (!attr) { XGGM_CUDA_TRY(cudaFuncSetAttribute(ln_bwd_fast<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
                XGGM_LAUNCH((ln_bwd_fast<NV, false>), fast_grid(M, 1), ROW_WARPS * 32, smem, st, gh, xhat, rstd, gamma, gu, ggamma, gbeta, hi, lo, M);
            }
        });
    } else {
        XGGM_REQUIRE(!xhat_bf16);
        size_t smem;
        XGGM_TRY(gen_smem(ln_bwd_gen, H, smem));
        XGGM_LAUNCH((ln_bwd_gen), gen_grid(M), GEN_WARPS * 32, smem, st, gh, xhat, rstd, gamma, gu, ggamma, gbeta, hi, lo, M, H);
    }
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

int gelu_ln_drop_fwd(const float* z, const float* gamma, const float* beta, const DropSpec& drop,
                     float* out, float* mean, float* rstd, bf16* hi, bf16* lo, int M, int H, float eps,
                     int accumulate, cudaStream_t st, int z_bf16) {
    if (M <= 0) return XGGM_OK;
    if (fast_ok(H, z, out, hi) && fast_ok(H, gamma, beta, lo) && (reinterpret_cast<uintptr_t>(drop.keep) & 3) == 0) {
        XGGM_ROW_DISPATCH(H, {
            constexpr size_t smem = RowPrefetch<NV, 2>::SMEM;
            if (z_bf16) {
                static bool attr = false;
                if (!attr) { XGGM_CUDA_TRY(cudaFuncSetAttribute(gld_fwd_fast<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
                XGGM_LAUNCH((gld_fwd_fast<NV, true>), fast_grid(M), ROW_WARPS * 32, smem, st, z, gamma, beta, drop, out, mean, rstd, hi, lo, M, eps, accumulate);
            } else {
                static bool attr = false;
                if (!attr) { XGGM_CUDA_TRY(cudaFuncSetAttribute(gld_fwd_fast<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
                XGGM_LAUNCH((gld_fwd_fast<NV, false>), fast_grid(M), ROW_WARPS * 32, smem, st, z, gamma, beta, drop, out, mean, rstd, hi, lo, M, eps, accumulate);
            }
        });
    } else {
        XGGM_REQUIRE(!z_bf16);
        XGGM_LAUNCH((gld_fwd_gen), gen_grid(M), GEN_WARPS * 32, 0, st, z, gamma, beta, drop, out, mean, rstd, hi, lo, M, H, eps, accumulate);
    }
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// gz may be null (only the planes are wanted); ggamma / gbeta / gbias? are ACCUMULATED into
int gelu_ln_drop_bwd(const float* gout, const float* z, const float* mean, const float* rstd, const float* gamma,
                     const DropSpec& drop, float* gz, float* ggamma, float* gbeta, float* gbias,
                     bf16* hi, bf16* lo, int M, int H, cudaStream_t st, int z_bf16) {
    if (M <= 0) return XGGM_OK;
    if (fast_ok(H, gout, z, gz) && fast_ok(H, gamma, hi, lo) && (reinterpret_cast<uintptr_t>(drop.keep) & 3) == 0) {
        XGGM_ROW_DISPATCH(H, {
            constexpr size_t smem = RowPrefetch<NV, 2>::SMEM;   // (>= the ROW_WARPS*H floats flush_columns needs)
            if (z_bf16) {
                static bool attr = false;
                if (!attr) { XGGM_CUDA_TRY(cudaFuncSetAttribute(gld_bwd_fast<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
                XGGM_LAUNCH((gld_bwd_fast<NV, true>), fast_grid(M, 1), ROW_WARPS * 32, smem, st, gout, z, mean, rstd, gamma, drop, gz, ggamma, gbeta, gbias, hi, lo, M);
            } else {
                static bool attr = false;
                if (!attr) { XGGM_CUDA_TRY(cudaFuncSetAttribute(gld_bwd_fast<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
                XGGM_LAUNCH((gld_bwd_fast<NV, false>), fast_grid(M, 1), ROW_WARPS * 32, smem, st, gout, z, mean, rstd, gamma, drop, gz, ggamma, gbeta, gbias, hi, lo, M);
            }
        });
    } else {
        XGGM_REQUIRE(!z_bf16);
        size_t smem;
        XGGM_TRY(gen_smem(gld_bwd_gen, H, smem));
        XGGM_LAUNCH((gld_bwd_gen), gen_grid(M), GEN_WARPS * 32, smem, st, gout, z, mean, rstd, gamma, drop, gz, ggamma, gbeta, gbias, hi, lo, M, H);
    }
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// ================================================================== VisualFeatEncoder tail
// out = dropout( (LN(z) + LN(box_fc(boxes))) / 2 ), src/lxrt/modeling.py:546-556, with z = visn_fc(feats) from the
// projection GEMM: the 4 -> H box projection (4 FMAs per element), both BertLayerNorms (eps 1e-12), the average and the
// dropout in ONE pass over z (was: SIMT GEMM + 2 LayerNorm + avg/dropout = 4 launches and 5 extra [M,H] round trips).
// Parameters (box weight [H,4], box bias, two gamma / beta pairs) live in shared memory; a warp owns a row.
struct VisnTailSmem {
    // floats: Wb[H*4] | bb[H] | g1[H] | b1[H] | g2[H] | b2[H]
    static __host__ __device__ constexpr int floats(int H) { return 9 * H; }
};
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
visn_tail_fwd_fast(const float* __restrict__ z, const float* __restrict__ boxes, const float* __restrict__ Wb,
                   const float* __restrict__ bb, const float* __restrict__ g1, const float* __restrict__ b1,
                   const float* __restrict__ g2, const float* __restrict__ b2, const DropSpec drop,
                   float* __restrict__ out, float* __restrict__ xhat1, float* __restrict__ rstd1,
                   float* __restrict__ mean2, float* __restrict__ rstd2, int M, float eps) {
    pdl_prologue();
    constexpr int H = NV * 128;
    extern __shared__ __align__(16) float sm[];
    float* sW = sm;
    float* sb = sW + 4 * H;
    float* sg1 = sb + H; float* sb1 = sg1 + H; float* sg2 = sb1 + H; float* sb2 = sg2 + H;
    for (int i = threadIdx.x; i < 4 * H; i += blockDim.x) sW[i] = Wb[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) { sb[i] = bb[i]; sg1[i] = g1[i]; sb1[i] = b1[i]; sg2[i] = g2[i]; sb2[i] = b2[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t dstream = drop_stream(drop);
    const float scale = drop.scale;
    const int wid = blockIdx.x * ROW_WARPS + warp, nw = gridDim.x * ROW_WARPS;
    for (int r = wid; r < M; r += nw) {
        const size_t ro = (size_t)r * H;
        float v[NV * 4], t[NV * 4];
        load_row<NV>(z + ro, lane, v);
        const float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)r * 4);
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 128 * i + 4 * lane + j;
                const float4 w = *reinterpret_cast<const float4*>(sW + 4 * c);
                t[4 * i + j] = fmaf(w.w, bx.w, fmaf(w.z, bx.z, fmaf(w.y, bx.y, fmaf(w.x, bx.x, sb[c]))));
            }
        float m1, r1, m2, r2;
        row_stats<NV>(v, 1.0f / (float)H, eps, m1, r1);
        row_stats<NV>(t, 1.0f / (float)H, eps, m2, r2);
#pragma unroll
        for (int i = 0; i < NV * 4; ++i) v[i] = (v[i] - m1) * r1;
        if (xhat1) store_row<NV>(xhat1 + ro, lane, v);
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 128 * i + 4 * lane + j, e = 4 * i + j;
                const float y1 = fmaf(v[e], sg1[c], sb1[c]);
                const float y2 = fmaf((t[e] - m2) * r2, sg2[c], sb2[c]);
                v[e] = 0.5f * (y1 + y2);
            }
        if (drop.mode) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                bool m[4];
                drop_bits4(drop, dstream, ro + 128 * i + 4 * lane, m);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[4 * i + j] = m[j] ? v[4 * i + j] * scale : 0.f;
            }
        }
        store_row<NV>(out + ro, lane, v);
        if (lane == 0) {
            if (rstd1) rstd1[r] = r1;
            if (mean2) mean2[r] = m2;
            if (rstd2) rstd2[r] = r2;
        }
    }
}

// backward: gz (gradient of z = visn_fc output), gt (gradient of the box projection's output), gamma / beta gradients of
// both LayerNorms (ACCUMULATED into).  The box projection and its normalised value are recomputed from the 16-byte box row.
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32, 1)
visn_tail_bwd_fast(const float* __restrict__ gout, const float* __restrict__ xhat1, const float* __restrict__ rstd1,
                   const float* __restrict__ boxes, const float* __restrict__ Wb, const float* __restrict__ bb,
                   const float* __restrict__ mean2, const float* __restrict__ rstd2, const float* __restrict__ g1,
                   const float* __restrict__ g2, const DropSpec drop, float* __restrict__ gz, float* __restrict__ gt,
                   float* __restrict__ gg1, float* __restrict__ gb1, float* __restrict__ gg2, float* __restrict__ gb2, int M) {
    pdl_prologue();
    constexpr int H = NV * 128;
    extern __shared__ __align__(16) float sm[];
    float* sW = sm;
    float* sb = sW + 4 * H;
    float* sg1 = sb + H; float* sg2 = sg1 + H;
    float* red = sg2 + H;     // ROW_WARPS * H floats for flush_columns
    for (int i = threadIdx.x; i < 4 * H; i += blockDim.x) sW[i] = Wb[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) { sb[i] = bb[i]; sg1[i] = g1[i]; sg2[i] = g2[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t dstream = drop_stream(drop);
    const float scale = 0.5f * drop.scale;
    const int wid = blockIdx.x * ROW_WARPS + warp, nw = gridDim.x * ROW_WARPS;
    float a1[NV * 4], a2[NV * 4], ab[NV * 4];
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) { a1[i] = 0.f; a2[i] = 0.f; ab[i] = 0.f; }
    for (int r = wid; r < M; r += nw) {
        const size_t ro = (size_t)r * H;
        float g[NV * 4], xh[NV * 4];
        load_row<NV>(gout + ro, lane, g);
        load_row<NV>(xhat1 + ro, lane, xh);
        const float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)r * 4);
        const float r1 = rstd1[r], m2 = mean2[r], r2 = rstd2[r];
        if (drop.mode) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                bool m[4];
                drop_bits4(drop, dstream, ro + 128 * i + 4 * lane, m);
#pragma unroll
                for (int j = 0; j < 4; ++j) g[4 * i + j] = m[j] ? g[4 * i + j] * scale : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NV * 4; ++i) g[i] *= scale;
        }
        // LayerNorm 1 (visual features)
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 128 * i + 4 * lane + j, e = 4 * i + j;
                const float d = g[e] * sg1[c];
                s1 += d;
                s2 = fmaf(d, xh[e], s2);
            }
        float c1 = warp_sum(s1) / (float)H, c2 = warp_sum(s2) / (float)H;
        float o[NV * 4];
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 128 * i + 4 * lane + j, e = 4 * i + j;
                a1[e] = fmaf(g[e], xh[e], a1[e]);
                ab[e] += g[e];
                o[e] = r1 * (g[e] * sg1[c] - c1 - xh[e] * c2);
            }
        store_row<NV>(gz + ro, lane, o);
        // LayerNorm 2 (boxes): recompute the projection and its normalised value
        s1 = 0.f; s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 128 * i + 4 * lane + j, e = 4 * i + j;
                const float4 w = *reinterpret_cast<const float4*>(sW + 4 * c);
                const float tv = fmaf(w.w, bx.w, fmaf(w.z, bx.z, fmaf(w.y, bx.y, fmaf(w.x, bx.x, sb[c]))));
                xh[e] = (tv - m2) * r2;
                const float d = g[e] * sg2[c];
                s1 += d;
                s2 = fmaf(d, xh[e], s2);
            }
        c1 = warp_sum(s1) / (float)H; c2 = warp_sum(s2) / (float)H;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 128 * i + 4 * lane + j, e = 4 * i + j;
                a2[e] = fmaf(g[e], xh[e], a2[e]);
                o[e] = r2 * (g[e] * sg2[c] - c1 - xh[e] * c2);
            }
        store_row<NV>(gt + ro, lane, o);
    }
    flush_columns<NV>(a1, gg1, red);
    flush_columns<NV>(a2, gg2, red);
    flush_columns<NV>(ab, gb1, red);
    flush_columns<NV>(ab, gb2, red);
}

bool visn_tail_supported(int H, int pos_dim) { return pos_dim == 4 && H % 128 == 0 && H >= 128 && H <= 1024; }

int visn_tail_fwd(const float* z, const float* boxes, const float* Wb, const float* bb, const float* g1, const float* b1,
                  const float* g2, const float* b2, const DropSpec& drop, float* out, float* xhat1, float* rstd1, float* mean2,
                  float* rstd2, int M, int H, float eps, cudaStream_t st) {
    if (M <= 0) return XGGM_OK;
    XGGM_REQUIRE(visn_tail_supported(H, 4) && fast_ok(H, z, out, xhat1) && (reinterpret_cast<uintptr_t>(boxes) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(Wb) & 15) == 0 && (reinterpret_cast<uintptr_t>(drop.keep) & 3) == 0);
    XGGM_ROW_DISPATCH(H, {
        const size_t smem = sizeof(float) * VisnTailSmem::floats(H);
        XGGM_LAUNCH((visn_tail_fwd_fast<NV>), fast_grid(M, 4), ROW_WARPS * 32, smem, st, z, boxes, Wb, bb, g1, b1, g2, b2, drop, out,
                    xhat1, rstd1, mean2, rstd2, M, eps);
    });
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

int visn_tail_bwd(const float* gout, const float* xhat1, const float* rstd1, const float* boxes, const float* Wb,
                  const float* bb, const float* mean2, const float* rstd2, const float* g1, const float* g2,
                  const DropSpec& drop, float* gz, float* gt, float* gg1, float* gb1, float* gg2, float* gb2, int M, int H,
                  cudaStream_t st) {
    if (M <= 0) return XGGM_OK;
    XGGM_REQUIRE(visn_tail_supported(H, 4) && fast_ok(H, gout, xhat1, gz) && fast_ok(H, gt) &&
                 (reinterpret_cast<uintptr_t>(boxes) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wb) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(drop.keep) & 3) == 0);
    XGGM_ROW_DISPATCH(H, {
        const size_t smem = sizeof(float) * (7 * H + ROW_WARPS * H);
        static bool attr = false;
        if (!attr) { XGGM_CUDA_TRY(cudaFuncSetAttribute(visn_tail_bwd_fast<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
        XGGM_LAUNCH((visn_tail_bwd_fast<NV>), fast_grid(M, 1), ROW_WARPS * 32, smem, st, gout, xhat1, rstd1, boxes, Wb, bb, mean2, rstd2,
                    g1, g2, drop, gz, gt, gg1, gb1, gg2, gb2, M);
    });
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

}  // namespace xggm
