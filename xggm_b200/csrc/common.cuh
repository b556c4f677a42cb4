// Shared device helpers for the X-GGM graph-block kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <atomic>

#include "../../include/xggm_b200.h"

namespace xggm {

constexpr int WARP = 32;

// ---- error plumbing (never throw across the C ABI) -------------------------
void set_cuda_error(cudaError_t e, const char* where);
#define XGGM_CUDA_TRY(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) {                              \
            ::xggm::set_cuda_error(_e, #expr);                \
            return XGGM_ERR_CUDA;                             \
        }                                                     \
    } while (0)
// one XGGM_LAUNCH_CHECK per kernel launch: also feeds xggm_launch_count()
extern std::atomic<unsigned long long> g_kernel_launches;
#define XGGM_LAUNCH_CHECK()                  \
    do {                                     \
        ++::xggm::g_kernel_launches;         \
        XGGM_CUDA_TRY(cudaGetLastError());   \
    } while (0)
#define XGGM_TRY(expr)                 \
    do {                               \
        int _r = (expr);               \
        if (_r != XGGM_OK) return _r;  \
    } while (0)
#define XGGM_REQUIRE(cond)                 \
    do {                                   \
        if (!(cond)) return XGGM_ERR_ARG;  \
    } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and
// starts with pdl_prologue(): `griddepcontrol.launch_dependents` lets the NEXT kernel of the stream (or
// captured graph) be scheduled as soon as all CTAs of this one are resident, and `griddepcontrol.wait`
// blocks until the PREVIOUS grid has completed and its memory is visible -- so launch latency, CTA
// scheduling and (in the GEMM kernel) barrier / TMEM / tensor-map setup overlap the predecessor's tail.
// No kernel touches global memory before its wait, hence ordering is exactly that of a plain stream.
// Measured SLOWER than plain launches inside the captured step on B200 (api.cu: pdl_mode), so the attribute is
// only set when XGGM_PDL asks for it; without it both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
    pdl_launch_dependents();
    pdl_wait();
}
bool pdl_enabled();
int pdl_mode();
template <typename... KArgs, typename... Args>
static inline void launch_kernel(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // errors surface in XGGM_LAUNCH_CHECK
}
#define XGGM_LAUNCH(kern, grid, block, smem, st, ...) ::xggm::launch_kernel(::xggm::pdl_enabled(), kern, grid, block, smem, st, __VA_ARGS__)

static inline cudaStream_t as_stream(xggm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- warp / block reductions -----------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + 1-D bulk async copy (per-warp row prefetch in rowops.cu) -----------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a lost completion becomes a CUDA error (trap) after 2 s, never a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!mbar_try_wait(bar, parity)) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 2000000000ull) __trap();
    }
}
// global -> shared bulk copy (bytes % 16 == 0, both 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace ptx

// ---- fp32 -> bf16 hi/lo operand planes (four consecutive elements, 8-byte stores; lo may be null) ----
__device__ __forceinline__ void split_store4(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, float a, float b, float c, float d) {
    const float v[4] = {a, b, c, d};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = __float2bfloat16_rn(v[i]);
        const float hf = __bfloat162float(h[i]);
        l[i] = __float2bfloat16_rn((hf - hf == 0.f) ? v[i] - hf : 0.f);
    }
    __nv_bfloat162 h01 = __halves2bfloat162(h[0], h[1]), h23 = __halves2bfloat162(h[2], h[3]);
    *reinterpret_cast<uint2*>(hi + off) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
    if (lo) {
        __nv_bfloat162 l01 = __halves2bfloat162(l[0], l[1]), l23 = __halves2bfloat162(l[2], l[3]);
        *reinterpret_cast<uint2*>(lo + off) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
    }
}


// ---- math ------------------------------------------------------------------
// Exact-erf GeLU (src/lxrt/modeling.py:116-124 of the reference): x*0.5*(1+erf(x/sqrt2)).
__device__ __forceinline__ float gelu_erf(float x) {
    return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}
// d/dx gelu_erf = Phi(x) + x*phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}
// Normal CDF and PDF for the GeLU of the hot row kernels: Phi(x) = 0.5 erfc(-x/sqrt2) through the
// Abramowitz-Stegun 7.1.26 rational form (|error| <= 1.5e-7 in erf, i.e. the same 1e-7 class as
// erff's own 2-ulp error), sharing one exp(-x^2/2) with the PDF: ~14 instructions instead of ~33
// for erff + expf.  The tail is formed without cancellation (0.5*poly*e on the small side).
__device__ __forceinline__ void normal_cdf_pdf(float x, float& cdf, float& pdf) {
    const float u = fabsf(x) * 0.70710678118654752440f;
    const float e = __expf(-u * u);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, u, 1.0f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(t, p, 1.421413741f);
    p = fmaf(t, p, -0.284496736f);
    p = fmaf(t, p, 0.254829592f);
    const float half_tail = 0.5f * (p * t) * e;          // 0.5 * erfc(|x|/sqrt2)
    cdf = x >= 0.f ? 1.0f - half_tail : half_tail;
    pdf = 0.39894228040143267794f * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
    float c, p;
    normal_cdf_pdf(x, c, p);
    return x * c;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- Philox4x32-10 counter RNG (dropout / noise without mask tensors) --------
struct Philox {
    uint32_t key0, key1;
    __device__ __forceinline__ Philox(uint64_t seed) : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t counter, uint64_t stream_id) const {
        uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32);
        uint32_t c2 = (uint32_t)stream_id, c3 = (uint32_t)(stream_id >> 32);
        uint32_t k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// Dropout of the read-out heads (F.dropout, src/module/gcn.py:72-76), three ways:
//   mode 0 none; mode 1 explicit uint8 keep-mask (parity runs inject the reference's masks);
//   mode 2 Philox4x32-10 drawn inside the consuming kernel: key = seed, counter = element/4,
//          subsequence = stream + (*epoch << 32).  `epoch` is an optional DEVICE counter so that a
//          captured CUDA graph draws fresh masks on every replay.
struct DropSpec {
    const uint8_t* keep;
    const uint64_t* epoch;
    uint64_t seed, stream;
    uint32_t thresh;   // keep iff random word >= thresh
    float scale;       // 1 / (1 - p)
    int mode;
};
// p = 0.5 (the only rate the reference uses) needs ONE random bit per element: element e keeps iff bit (e % 128)
// of the 128-bit Philox block with counter e / 128 is set -- 32x fewer Philox blocks than a 32-bit word per
// element.  Every consumer (the row kernels, xggm_keep_mask) switches to this mapping when the threshold is
// exactly 1/2; other rates keep the word-per-element mapping.
constexpr uint32_t XGGM_COIN_THRESH = 0x80000000u;
__device__ __forceinline__ uint32_t philox_word(const uint4& r, uint32_t w) {
    return w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
}
// the 4 keep bits of elements off .. off+3 (off % 4 == 0) under the coin mapping, bit j = element off + j
__device__ __forceinline__ uint32_t coin_nibble(const Philox& rng, uint64_t stream, size_t off) {
    const uint4 r = rng((uint64_t)(off >> 7), stream);
    const uint32_t b = (uint32_t)(off & 127);
    return (philox_word(r, b >> 5) >> (b & 31)) & 0xFu;
}
static inline uint32_t drop_threshold(float p) {
    const double t = (double)p * 4294967296.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}
static inline DropSpec drop_none() { return DropSpec{nullptr, nullptr, 0, 0, 0u, 1.f, 0}; }
static inline DropSpec drop_mask(const uint8_t* keep, float scale) {
    return keep ? DropSpec{keep, nullptr, 0, 0, 0u, scale, 1} : drop_none();
}

}  // namespace xggm
