// extern "C" surface of libxggm_b200.so (see include/xggm_b200.h) and the composite
// GCN / GIN layer drivers that sequence the kernels on the caller's stream.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace xggm {

// ---- error state -------------------------------------------------------------
static thread_local char g_cuda_err[256] = "";
std::atomic<unsigned long long> g_kernel_launches{0};
void set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e), where);
}

// XGGM_PDL: 0 (default) plain stream-ordered launches; 1 PDL on every kernel; 2 GEMM kernels only; 3 all but GEMMs.
// Measured at B=256 inside the captured step: 2.01 ms (0) vs 2.12 (1), 2.07 (2), 2.02 (3) -> off by default.
int pdl_mode() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("XGGM_PDL");
        on = e ? atoi(e) : 0;
    }
    return on;
}
bool pdl_enabled() { return pdl_mode() == 1 || pdl_mode() == 3; }
constexpr float LN_EPS = 1e-5f;  // nn.LayerNorm default (src/module/gcn.py:14,47)

// chunk sizes are padded to 8 floats so that every bf16 plane starts 16-byte aligned (TMA)
static inline long long pad8(long long n) { return (n + 7) & ~7LL; }

// ---- projection engine selection -------------------------------------------------
// precision: XGGM_PREC_FP32 (tcgen05, split-bf16 x3), XGGM_PREC_BF16 (tcgen05, single pass),
// XGGM_PREC_FP32_SIMT (exact fp32 FMA kernel).  Shapes TMA cannot address (row pitch not a
// multiple of 16 bytes) always take the SIMT kernel.
int g_precision = XGGM_PREC_FP32;

typedef __nv_bfloat16 bf16;
struct Operand {          // one GEMM operand: the fp32 tensor and (tcgen05 engine) its bf16 planes
    const float* f32;
    const bf16* hi;
    const bf16* lo;
};
static inline bool use_tc(int M, int N, int K) {
    return g_precision != XGGM_PREC_FP32_SIMT && gemm_tc_supported(M, N, K);
}
static inline int npass() { return g_precision == XGGM_PREC_BF16 ? 1 : 3; }
// bf16 STORAGE (BASELINE configs[2]): under the single-pass bf16 engine the saved pre-activations z and normalised
// values xhat of a GCN / GIN layer are kept as bf16 only (the projection writes z as its hi plane and nothing else; the
// row kernels read / write bf16 rows) -- half the bytes of the fp32 copies the fp32-parity engine needs.  Fast row
// kernels only (H % 128 == 0, H <= 1024); XGGM_BF16_STORAGE=0 keeps fp32 storage for A/B runs.
static inline bool bf16_storage(int H) {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("XGGM_BF16_STORAGE");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on && g_precision == XGGM_PREC_BF16 && H % 128 == 0 && H >= 128 && H <= 1024;
}
// planes of an [n]-element fp32 array stored in a region of pad8(n) floats: hi | lo
static inline Operand planes_at(const float* f32, float* region, long long n) {
    bf16* hi = reinterpret_cast<bf16*>(region);
    return Operand{f32, hi, hi + pad8(n) };
}
static inline bf16* mut(const bf16* p) { return const_cast<bf16*>(p); }
static inline bf16* lo_or_null(const Operand& o) { return npass() == 3 ? const_cast<bf16*>(o.lo) : nullptr; }
static int split_one(const Operand& o, long long n, cudaStream_t st) {
    const float* src[1] = {o.f32};
    bf16* hi[1] = {const_cast<bf16*>(o.hi)};
    bf16* lo[1] = {const_cast<bf16*>(o.lo)};
    return split_planes(src, hi, npass() == 3 ? lo : nullptr, &n, 1, st);
}
// out[M,N] = a[M,K] w[N,K]^T + bias + resid
static int proj_fwd(bool tc, const Operand& a, const Operand& w, const float* bias, const float* resid,
                    float* out, int M, int N, int K, cudaStream_t st) {
    if (tc) return gemm_tc(false, false, a.hi, a.lo, w.hi, w.lo, bias, resid, out, nullptr, nullptr, M, N, K, 0, 0, npass(), st);
    return gemm_simt(0, a.f32, w.f32, bias, resid, out, M, N, K, 0, st);
}
// ---- saved-activation layout of one GCN / GIN layer ---------------------------
struct GnnLayout {
    long long MH, Mr, HH;   // padded sizes of an [M,H], an [M] and an [H,H] chunk
    int n_convs, n_heads, kind;
    long long conv_stride, head_stride, head_base, xplanes, total;
    // GCN conv k: agg | xhat | h_next | rstd | (pad) | P(agg) | P(h_next)
    // GIN conv k: pre | z    | h_next | mean | rstd  | P(pre) | P(h_next)
    GnnLayout(int kind_, long long M, int H, int nc) : kind(kind_) {
        MH = pad8(M * H);
        Mr = pad8(M);
        HH = pad8((long long)H * H);
        n_convs = nc;
        n_heads = nc + 1;
        conv_stride = 5 * MH + 2 * Mr;
        head_stride = MH + 2 * Mr;  // z | mean | rstd
        head_base = conv_stride * nc;
        xplanes = head_base + head_stride * n_heads;   // P(x)
        total = xplanes + MH;
    }
    long long conv(int k, int slot) const {  // slots 0..2 fp32 [M,H]; 3,4 [M]; 5,6 plane regions
        const long long base = conv_stride * k;
        if (slot < 3) return base + slot * MH;
        if (slot < 5) return base + 3 * MH + (slot - 3) * Mr;
        return base + 3 * MH + 2 * Mr + (slot - 5) * MH;
    }
    long long head(int j, int slot) const {
        const long long base = head_base + head_stride * j;
        return slot == 0 ? base : base + MH + (slot - 1) * Mr;
    }
    // fwd: u | weight planes | block-diagonal coefficient planes
    long long work_fwd(long long coef) const { return MH + (2 * n_convs + 1) * HH + pad8(coef); }
    // bwd: buf0 | buf1 | gt | gq | R0 | P(gq) | weight planes | S scratch [B,N,N] | coefficient planes | R1 | R2  (R* = gradient regions)
    long long work_bwd(long long bnn, long long coef) const { return 8 * MH + (2 * n_convs + 1) * HH + pad8(bnn) + pad8(coef); }
};

// weight planes for one layer live at the tail of the work buffer: conv k -> slot k, head j -> slot nc + j
// transposed = true (backward pass): the planes hold W^T, the layout proj_dgrad wants.
static int split_weights(int kind, const float* const* cp, const float* const* hp, float* wregion,
                         const GnnLayout& L, int H, Operand* wconv, Operand* whead, bool tc, bool transposed,
                         cudaStream_t st, const void* const* wplanes = nullptr) {
    const int nc = L.n_convs;
    if (tc && wplanes) {
        // prepared weight planes (xggm_weight_planes_build): conv k -> entry k, head j -> entry nc + j; nothing to split
        for (int i = 0; i < 2 * nc + 1; ++i) {
            XGGM_REQUIRE(wplanes[i]);
            const float* W = i < nc ? ((kind == XGGM_KIND_GCN) ? cp[3 * i] : cp[5 * i + 1]) : hp[4 * (i - nc)];
            bf16 *hi, *lo, *thi, *tlo;
            weight_planes_views(const_cast<void*>(wplanes[i]), H, H, &hi, &lo, &thi, &tlo);
            const Operand o = transposed ? Operand{W, thi, tlo} : Operand{W, hi, lo};
            if (i < nc) wconv[i] = o; else whead[i - nc] = o;
        }
        return XGGM_OK;
    }
    const float* src[16];
    bf16* hi[16];
    bf16* lo[16];
    long long n[16];
    int cnt = 0;
    for (int k = 0; k < nc; ++k) {
        const float* W = (kind == XGGM_KIND_GCN) ? cp[3 * k] : cp[5 * k + 1];
        wconv[k] = planes_at(W, wregion + (long long)k * L.HH, (long long)H * H);
    }
    for (int j = 0; j <= nc; ++j)
        whead[j] = planes_at(hp[4 * j], wregion + (long long)(nc + j) * L.HH, (long long)H * H);
    if (!tc) return XGGM_OK;
    if (transposed) {
        for (int i = 0; i < 2 * nc + 1; ++i) {
            const Operand& o = i < nc ? wconv[i] : whead[i - nc];
            src[i] = o.f32; hi[i] = const_cast<bf16*>(o.hi); lo[i] = const_cast<bf16*>(o.lo);
        }
        return split_planes_t(src, hi, npass() == 3 ? lo : nullptr, H, H, 2 * nc + 1, st);
    }
    for (int i = 0; i < 2 * nc + 1; ++i) {
        const Operand& o = i < nc ? wconv[i] : whead[i - nc];
        src[cnt] = o.f32; hi[cnt] = const_cast<bf16*>(o.hi); lo[cnt] = const_cast<bf16*>(o.lo);
        n[cnt] = (long long)H * H;
        if (++cnt == 16) { XGGM_TRY(split_planes(src, hi, npass() == 3 ? lo : nullptr, n, cnt, st)); cnt = 0; }
    }
    if (cnt) XGGM_TRY(split_planes(src, hi, npass() == 3 ? lo : nullptr, n, cnt, st));
    return XGGM_OK;
}

constexpr int MAX_CONVS = 7;

// XGGM_FUSED_ADJ_LN=1: GCN forward message passing inside the LayerNorm kernel (adj_ln_fwd, one CTA per graph) instead
// of the tensor-core kernel + LayerNorm.  Measured slower at B=256 (140 vs 24 + 18 us: one graph per SM leaves too
// little memory parallelism), so it is only the path of the exact-fp32 engine and of shapes the tensor cores skip.
static bool fused_adj_ln() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("XGGM_FUSED_ADJ_LN");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on != 0;
}

// dropout of read-out head j: explicit masks win, then in-kernel Philox, else none
static inline DropSpec head_drop(const uint8_t* const* keeps, const xggm_philox_t* ph, float drop_p, int j) {
    const float scale = 1.f / (1.f - drop_p);
    if (keeps) return drop_mask(keeps[j], scale);
    if (ph && drop_p > 0.f)
        return DropSpec{nullptr, ph->dev_epoch, ph->seed, ph->stream0 + (uint64_t)j, drop_threshold(drop_p), scale, 2};
    return drop_none();
}

// ---- grouped projections -------------------------------------------------------------------------
// Independent products of the same shape share ONE tensor-core launch (gemm_tc_group): the context
// projection of conv k with read-out head k (both only need h_k), and in the backward pass their two
// weight gradients and their two input gradients.  A lone 768x768 projection at B=256 is two tiles per
// CTA pair, so its per-launch fixed cost (prologue, exposed last epilogue, teardown) is ~1/4 of its
// time; a group pays it once.  The exact-fp32 engine runs the members back to back.
struct Lin {
    Operand a, w;              // forward: activations, weight; dgrad: gradient, W^T planes; wgrad: gradient, activations
    const float* bias;
    const float* resid;
    float* out;
    const Operand* out_planes; // also emit the result as bf16 planes (tensor-core engine only)
    int accumulate;
};
static inline GemmProb as_prob(const Lin& l) {
    return GemmProb{l.a.hi, l.a.lo, l.w.hi, l.w.lo, l.bias, l.resid, l.out,
                    l.out_planes ? const_cast<bf16*>(l.out_planes->hi) : nullptr,
                    l.out_planes ? const_cast<bf16*>(l.out_planes->lo) : nullptr, l.accumulate};
}
static int fwd_group(bool tc, const Lin* l, int n, int M, int N, int K, cudaStream_t st) {
    if (!tc) {
        for (int i = 0; i < n; ++i)
            XGGM_TRY(gemm_simt(0, l[i].a.f32, l[i].w.f32, l[i].bias, l[i].resid, l[i].out, M, N, K, 0, st));
        return XGGM_OK;
    }
    GemmProb q[3];
    for (int i = 0; i < n; ++i) q[i] = as_prob(l[i]);
    return gemm_tc_group(false, false, q, n, M, N, K, 0, npass(), st);
}
// ga[M,K] (+)= g[M,N] w[N,K]   (tensor-core engine: `w` carries the planes of W^T)
static int dgrad_group(bool tc, const Lin* l, int n, int M, int N, int K, cudaStream_t st) {
    if (!tc) {
        for (int i = 0; i < n; ++i)
            XGGM_TRY(gemm_simt(1, l[i].a.f32, l[i].w.f32, nullptr, nullptr, l[i].out, M, K, N, l[i].accumulate, st));
        return XGGM_OK;
    }
    GemmProb q[3];
    for (int i = 0; i < n; ++i) q[i] = as_prob(l[i]);
    return gemm_tc_group(false, false, q, n, M, K, N, 0, npass(), st);
}
// gw[N,K] (+)= g[M,N]^T a[M,K]
static int wgrad_group(bool tc, const Lin* l, int n, int M, int N, int K, cudaStream_t st) {
    if (!tc) {
        for (int i = 0; i < n; ++i)
            XGGM_TRY(gemm_simt(2, l[i].a.f32, l[i].w.f32, nullptr, nullptr, l[i].out, N, K, M, l[i].accumulate, st));
        return XGGM_OK;
    }
    GemmProb q[3];
    for (int i = 0; i < n; ++i) q[i] = as_prob(l[i]);
    return gemm_tc_group(true, true, q, n, N, K, M, 1, npass(), st);
}

// ga[M,K] (+)= g0 w0 + g1 w1 as ONE product whose contraction runs over both gradients (same output, one epilogue)
static int dgrad_kcat(bool tc, const Lin* l, int M, int N, int K, cudaStream_t st) {
    if (!tc) return dgrad_group(false, l, 2, M, N, K, st);
    GemmProb q[2] = {as_prob(l[0]), as_prob(l[1])};
    return gemm_tc_group(false, false, q, 2, M, K, N, 0, npass(), st, true);
}

static int gnn_fwd(int kind, const float* x, const float* adj, const float* const* cp,
                   const float* const* hp, const uint8_t* const* keeps, const xggm_philox_t* philox, float drop_p, float* out,
                   float* saved, float* work, const void* x_planes, void* out_planes, int B, int N, int H, int nc,
                   cudaStream_t st, const void* const* wplanes = nullptr) {
    XGGM_REQUIRE(kind == XGGM_KIND_GCN || kind == XGGM_KIND_GIN);
    XGGM_REQUIRE(B >= 0 && N > 0 && H > 0 && nc >= 0 && nc <= MAX_CONVS && drop_p >= 0.f && drop_p < 1.f);
    const int M = B * N;
    if (M == 0) return XGGM_OK;
    XGGM_REQUIRE(x && adj && cp && hp && out && saved && work);
    const GnnLayout L(kind, M, H, nc);
    const bool tc = use_tc(M, H, H);
    const long long MHn = (long long)M * H;
    Operand wconv[MAX_CONVS], whead[MAX_CONVS + 1];
    XGGM_TRY(split_weights(kind, cp, hp, work + L.MH, L, H, wconv, whead, tc, false, st, wplanes));
    // current node features as a GEMM operand: planes handed over by the producer of x, else built here
    Operand hop = x_planes ? planes_at(x, static_cast<float*>(const_cast<void*>(x_planes)), MHn)
                           : planes_at(x, saved + L.xplanes, MHn);
    if (tc && !x_planes) XGGM_TRY(split_one(hop, MHn, st));
    Operand hops[MAX_CONVS + 1];
    hops[0] = hop;
    const float* h = x;
    // message passing on the tensor cores: block-diagonal coefficient tiles (built once per call for GCN,
    // per conv for GIN whose (1+eps) differs) x the node planes
    const bool adjtc = tc && adj_tc_supported(N, H);
    bf16* coef_hi = reinterpret_cast<bf16*>(work + L.MH + (2 * nc + 1) * L.HH);
    bf16* coef_lo = coef_hi + pad8(adjtc ? adj_tc_coef_elems(B, N) : 0);
    if (adjtc && kind == XGGM_KIND_GCN && nc > 0 && !fused_adj_ln())
        XGGM_TRY(build_blockdiag(adj, coef_hi, npass() == 3 ? coef_lo : nullptr, B, N, 1.f, nullptr, 0.f, 0, st));
    const bool bst = tc && bf16_storage(H);
    Operand zplane[MAX_CONVS + 1];   // bf16 storage: z_j exists only as the hi plane the projection writes
    for (int j = 0; j <= nc; ++j) zplane[j] = Operand{nullptr, reinterpret_cast<const bf16*>(saved + L.head(j, 0)), nullptr};
    auto head_lin = [&](int j) -> Lin {   // z_j = h_j W_j^T + b_j
        if (bst) return Lin{hops[j], whead[j], hp[4 * j + 1], nullptr, nullptr, &zplane[j], 0};
        return Lin{hops[j], whead[j], hp[4 * j + 1], nullptr, saved + L.head(j, 0), nullptr, 0};
    };
    // out (+)= dropout(LN(GeLU(z_j))); the last accumulation can also emit `out` as operand planes for its consumers
    const Operand out_op = planes_at(out, static_cast<float*>(out_planes), MHn);
    auto head_post = [&](int j) -> int {
        const bool emit = tc && out_planes && j == nc;
        return gelu_ln_drop_fwd(saved + L.head(j, 0), hp[4 * j + 2], hp[4 * j + 3], head_drop(keeps, philox, drop_p, j), out,
                                saved + L.head(j, 1), saved + L.head(j, 2), emit ? mut(out_op.hi) : nullptr,
                                emit ? lo_or_null(out_op) : nullptr, M, H, LN_EPS, j > 0, st, bst ? 1 : 0);
    };
    for (int k = 0; k < nc; ++k) {
        float* h_next = saved + L.conv(k, 2);
        const Operand next_op = planes_at(h_next, saved + L.conv(k, 6), MHn);
        const bool gcn = kind == XGGM_KIND_GCN;
        Lin g[2];
        if (gcn && adjtc && !fused_adj_ln()) {
            // GCNConv with W.(adj @ h) re-associated as adj @ (W.h): P = h Wc^T shares its launch AND its A operand
            // with read-out head k and leaves the GEMM as operand planes; the message passing then runs on the tensor
            // cores with the residual in its epilogue, u = h + adj @ P, and LayerNorm follows.
            const Operand P_op = planes_at(saved + L.conv(k, 0), saved + L.conv(k, 5), MHn);
            g[0] = Lin{hops[k], wconv[k], nullptr, nullptr, nullptr, &P_op, 0};
            g[1] = head_lin(k);
            XGGM_TRY(fwd_group(tc, g, 2, M, H, H, st));
            // the residual h_k comes from its operand planes (hi + lo = 16 mantissa bits, ~1e-5 like the products), so
            // h_{k+1} never has to exist as an fp32 tensor: LayerNorm writes xhat and the planes only
            const bool fast_ln = H % 128 == 0 && H >= 128 && H <= 1024;
            if (adj_ln_tc_supported(N, H)) {
                // H = 768: message passing + LayerNorm in ONE cluster kernel (four CTAs x 192 columns share a row block and
                // exchange the row statistics through distributed shared memory); u never reaches global memory
                XGGM_TRY(adj_ln_tc(coef_hi, coef_lo, P_op.hi, P_op.lo, hops[k].hi, hops[k].lo, cp[3 * k + 1], cp[3 * k + 2],
                                   saved + L.conv(k, 1), bst ? 1 : 0, saved + L.conv(k, 3), nullptr, mut(next_op.hi),
                                   lo_or_null(next_op), B, N, H, LN_EPS, npass(), st));
            } else {
                XGGM_TRY(adj_apply_tc(coef_hi, coef_lo, P_op.hi, P_op.lo, work /* u */, nullptr, nullptr, B, N, H, 0, npass(), st,
                                      nullptr, hops[k].hi, hops[k].lo));
                XGGM_TRY(layernorm_fwd(work, cp[3 * k + 1], cp[3 * k + 2], fast_ln ? nullptr : h_next, saved + L.conv(k, 1),
                                       saved + L.conv(k, 3), mut(next_op.hi), lo_or_null(next_op), M, H, LN_EPS, st, bst ? 1 : 0));
            }
        } else if (gcn && adj_ln_supported(N, H)) {
            // same algebra with the message passing folded into the LayerNorm kernel (one CTA per graph):
            //     h_next = LN(h + adj @ P)
            float* P = saved + L.conv(k, 0);
            g[0] = Lin{hops[k], wconv[k], nullptr, nullptr, P, nullptr, 0};
            g[1] = head_lin(k);
            XGGM_TRY(fwd_group(tc, g, 2, M, H, H, st));
            XGGM_TRY(adj_ln_fwd(adj, P, h, cp[3 * k + 1], cp[3 * k + 2], h_next, saved + L.conv(k, 1), saved + L.conv(k, 3),
                                tc ? mut(next_op.hi) : nullptr, tc ? lo_or_null(next_op) : nullptr, B, N, H, LN_EPS, st));
        } else if (gcn) {
            // shapes the fused kernel does not take: the same re-associated algebra from separate kernels
            float* P = saved + L.conv(k, 0);
            g[0] = Lin{hops[k], wconv[k], nullptr, nullptr, P, nullptr, 0};
            g[1] = head_lin(k);
            XGGM_TRY(fwd_group(tc, g, 2, M, H, H, st));
            XGGM_CUDA_TRY(cudaMemcpyAsync(work, h, sizeof(float) * (size_t)MHn, cudaMemcpyDeviceToDevice, st));
            XGGM_TRY(adj_apply(adj, P, work, nullptr, nullptr, B, N, H, 1.f, nullptr, 0.f, false, 1, st));   // u = h + adj @ P
            XGGM_TRY(layernorm_fwd(work, cp[3 * k + 1], cp[3 * k + 2], h_next, saved + L.conv(k, 1), saved + L.conv(k, 3),
                                   tc ? mut(next_op.hi) : nullptr, tc ? lo_or_null(next_op) : nullptr, M, H, LN_EPS, st));
        } else {
            float* pre = saved + L.conv(k, 0);   // GIN: pre = h + (1+eps) adj @ h
            Operand pre_op = planes_at(pre, saved + L.conv(k, 5), MHn);
            const float* eps = cp[5 * k];
            // tensor-core engine: the aggregate is only ever a GEMM operand -> bf16 planes, no fp32 copy
            if (adjtc) {
                XGGM_TRY(build_blockdiag(adj, coef_hi, npass() == 3 ? coef_lo : nullptr, B, N, 1.f, eps, 1.f, 0, st));
                XGGM_TRY(adj_apply_tc(coef_hi, coef_lo, hops[k].hi, hops[k].lo, nullptr, mut(pre_op.hi), lo_or_null(pre_op),
                                      B, N, H, 0, npass(), st));
            } else {
                XGGM_TRY(adj_apply(adj, h, tc ? nullptr : pre, tc ? mut(pre_op.hi) : nullptr, tc ? lo_or_null(pre_op) : nullptr,
                                   B, N, H, 1.f, eps, 1.f, false, 0, st));
            }
            // conv projection + read-out head k in one launch
            const Operand zc = Operand{nullptr, reinterpret_cast<const bf16*>(saved + L.conv(k, 1)), nullptr};
            g[0] = bst ? Lin{pre_op, wconv[k], cp[5 * k + 2], nullptr, nullptr, &zc, 0}
                       : Lin{pre_op, wconv[k], cp[5 * k + 2], nullptr, saved + L.conv(k, 1) /* z */, nullptr, 0};
            g[1] = head_lin(k);
            XGGM_TRY(fwd_group(tc, g, 2, M, H, H, st));
            XGGM_TRY(gelu_ln_drop_fwd(saved + L.conv(k, 1), cp[5 * k + 3], cp[5 * k + 4], drop_none(), h_next,
                                      saved + L.conv(k, 3), saved + L.conv(k, 4), tc ? mut(next_op.hi) : nullptr,
                                      tc ? lo_or_null(next_op) : nullptr, M, H, LN_EPS, 0, st, bst ? 1 : 0));
        }
        XGGM_TRY(head_post(k));
        hops[k + 1] = next_op;
        h = h_next;
    }
    const Lin last = head_lin(nc);
    XGGM_TRY(fwd_group(tc, &last, 1, M, H, H, st));
    return head_post(nc);
}

static int gnn_bwd(int kind, const float* gout, const float* x, const float* adj,
                   const float* const* cp, const float* const* hp, const uint8_t* const* keeps,
                   const xggm_philox_t* philox, float drop_p, const float* saved_c, float* work, float* gx, float* gadj,
                   float* const* cg, float* const* hg, int acc, const void* x_planes, int B, int N, int H, int nc,
                   cudaStream_t st, const void* const* wplanes = nullptr) {
    XGGM_REQUIRE(kind == XGGM_KIND_GCN || kind == XGGM_KIND_GIN);
    XGGM_REQUIRE(B >= 0 && N > 0 && H > 0 && nc >= 0 && nc <= MAX_CONVS && cp && hp && cg && hg);
    const int M = B * N;
    if (M == 0 && acc) return XGGM_OK;
    if (M == 0) {  // empty batch: parameter gradients are zero
        const int per = (kind == XGGM_KIND_GCN) ? 3 : 5;
        for (int k = 0; k < nc; ++k)
            for (int q = 0; q < per; ++q) {
                const size_t n = (kind == XGGM_KIND_GCN) ? (q == 0 ? (size_t)H * H : H)
                                                         : (q == 0 ? 1 : (q == 1 ? (size_t)H * H : H));
                XGGM_CUDA_TRY(cudaMemsetAsync(cg[per * k + q], 0, sizeof(float) * n, st));
            }
        for (int j = 0; j <= nc; ++j)
            for (int q = 0; q < 4; ++q)
                XGGM_CUDA_TRY(cudaMemsetAsync(hg[4 * j + q], 0, sizeof(float) * (q == 0 ? (size_t)H * H : H), st));
        return XGGM_OK;
    }
    XGGM_REQUIRE(gout && x && adj && saved_c && work && gx);   // gadj == NULL: the adjacency needs no gradient
    XGGM_REQUIRE(gadj || kind == XGGM_KIND_GCN);               // (GIN still needs gq h^T for d eps)
    float* saved = const_cast<float*>(saved_c);  // plane regions are read-only here; the cast only feeds planes_at
    const GnnLayout L(kind, M, H, nc);
    const bool tc = use_tc(M, H, H);
    const bool gcn = kind == XGGM_KIND_GCN;
    const bool bst = tc && bf16_storage(H);      // must mirror gnn_fwd: z / xhat were saved as bf16
    const long long MH = L.MH, MHn = (long long)M * H;
    float* buf[2] = {work, work + MH};
    float* gt = work + 2 * MH;  // GIN: gz of the conv (exact engine only; planes otherwise)
    float* gq = work + 3 * MH;
    Operand wconv[MAX_CONVS], whead[MAX_CONVS + 1];
    XGGM_TRY(split_weights(kind, cp, hp, work + 6 * MH, L, H, wconv, whead, tc, true, st, wplanes));
    if (gadj) XGGM_CUDA_TRY(cudaMemsetAsync(gadj, 0, sizeof(float) * (size_t)B * N * N, st));
    // per-graph products gq h^T go through the tensor-core Gram kernel when it can address them
    const bool gram = tc && gram_tc_supported(N, H);
    const Operand gq_op = planes_at(gq, work + 5 * MH, MHn);
    float* s_scratch = work + 6 * MH + (2 * nc + 1) * L.HH;
    const long long BNN = (long long)B * N * N;
    const bool adjtc = gram && adj_tc_supported(N, H);   // needs the planes of gq the Gram path already emits
    bf16* coef_hi = reinterpret_cast<bf16*>(s_scratch + pad8(BNN));
    bf16* coef_lo = coef_hi + pad8(adjtc ? adj_tc_coef_elems(B, N) : 0);
    if (adjtc && gcn && nc > 0)   // adj^T coefficients, shared by every conv of the layer
        XGGM_TRY(build_blockdiag(adj, coef_hi, npass() == 3 ? coef_lo : nullptr, B, N, 1.f, nullptr, 0.f, 1, st));

    auto act = [&](int j) -> Operand {   // h_j as a GEMM operand (planes saved by the forward pass)
        if (j == 0)
            return x_planes ? planes_at(x, static_cast<float*>(const_cast<void*>(x_planes)), MHn)
                            : planes_at(x, saved + L.xplanes, MHn);
        return planes_at(saved + L.conv(j - 1, 2), saved + L.conv(j - 1, 6), MHn);
    };
    // Three gradient regions of [M,H] floats each: bf16 planes under the tensor-core engine (the fp32
    // tensor never exists), the fp32 tensor itself under the exact engine.
    //   R0 conv-level gradient (gu / gz of conv k), R1 gz of head k, R2 gz of the last head
    const long long coef_floats = N <= 128 ? pad8(adj_tc_coef_elems(B, N)) : 0;   // as in xggm_gnn_work_floats
    float* R[3] = {work + 4 * MH, s_scratch + pad8(BNN) + coef_floats, s_scratch + pad8(BNN) + coef_floats + MH};
    // head j: gz_j (as planes or fp32 in `region`) and the bias / gamma / beta gradients
    auto head_gz = [&](int j, float* region) -> int {
        if (!acc) XGGM_CUDA_TRY(cudaMemsetAsync(hg[4 * j + 2], 0, sizeof(float) * H, st));
        if (!acc) XGGM_CUDA_TRY(cudaMemsetAsync(hg[4 * j + 3], 0, sizeof(float) * H, st));
        if (!acc) XGGM_CUDA_TRY(cudaMemsetAsync(hg[4 * j + 1], 0, sizeof(float) * H, st));
        const Operand g = planes_at(region, region, MHn);
        return gelu_ln_drop_bwd(gout, saved + L.head(j, 0), saved + L.head(j, 1), saved + L.head(j, 2),
                                hp[4 * j + 2], head_drop(keeps, philox, drop_p, j), tc ? nullptr : region,
                                hg[4 * j + 2], hg[4 * j + 3], hg[4 * j + 1], tc ? mut(g.hi) : nullptr,
                                tc ? lo_or_null(g) : nullptr, M, H, st, bst ? 1 : 0);
    };

    int cur = 0;
    float* gh = (nc == 0) ? gx : buf[cur];
    // last head: its input gradient starts the chain; its weight gradient rides with the first conv level's
    XGGM_TRY(head_gz(nc, R[2]));
    const Operand gzN = planes_at(R[2], R[2], MHn);
    Lin pending_w = Lin{gzN, act(nc), nullptr, nullptr, hg[4 * nc], nullptr, acc};
    {
        const Lin d = Lin{gzN, whead[nc], nullptr, nullptr, gh, nullptr, 0};
        XGGM_TRY(dgrad_group(tc, &d, 1, M, H, H, st));
    }
    bool have_pending = true;
    for (int k = nc - 1; k >= 0; --k) {
        const float* hk = (k == 0) ? x : saved + L.conv(k - 1, 2);
        float* gnext = (k == 0) ? gx : buf[cur ^ 1];
        // conv-level gradient g0: GCN gu = LN backward (also the residual path: written into gnext);
        // GIN gz of the conv's GeLU+LN
        const float* g0_f32 = gcn ? gnext : gt;
        const Operand g0 = planes_at(g0_f32, R[0], MHn);
        const float* eps = gcn ? nullptr : cp[5 * k];
        if (gcn) {
            if (!acc) XGGM_CUDA_TRY(cudaMemsetAsync(cg[3 * k + 1], 0, sizeof(float) * H, st));
            if (!acc) XGGM_CUDA_TRY(cudaMemsetAsync(cg[3 * k + 2], 0, sizeof(float) * H, st));
            const bool xb = bst && adj_tc_supported(N, H) && !fused_adj_ln();   // (the path of gnn_fwd that wrote bf16 xhat)
            XGGM_TRY(layernorm_bwd(gh, saved + L.conv(k, 1), saved + L.conv(k, 3), cp[3 * k + 1], gnext,
                                   cg[3 * k + 1], cg[3 * k + 2], tc ? mut(g0.hi) : nullptr,
                                   tc ? lo_or_null(g0) : nullptr, M, H, st, xb ? 1 : 0));
        } else {
            if (!acc) XGGM_CUDA_TRY(cudaMemsetAsync(cg[5 * k], 0, sizeof(float), st));
            if (!acc) XGGM_CUDA_TRY(cudaMemsetAsync(cg[5 * k + 3], 0, sizeof(float) * H, st));
            if (!acc) XGGM_CUDA_TRY(cudaMemsetAsync(cg[5 * k + 4], 0, sizeof(float) * H, st));
            if (!acc) XGGM_CUDA_TRY(cudaMemsetAsync(cg[5 * k + 2], 0, sizeof(float) * H, st));
            XGGM_TRY(gelu_ln_drop_bwd(gh, saved + L.conv(k, 1), saved + L.conv(k, 3), saved + L.conv(k, 4),
                                      cp[5 * k + 3], drop_none(), tc ? nullptr : gt, cg[5 * k + 3], cg[5 * k + 4], cg[5 * k + 2],
                                      tc ? mut(g0.hi) : nullptr, tc ? lo_or_null(g0) : nullptr, M, H, st, bst ? 1 : 0));
        }
        XGGM_TRY(head_gz(k, R[1]));
        const Operand g1 = planes_at(R[1], R[1], MHn);
        if (gcn) {
            // re-associated conv (see gnn_fwd): u = h_k + adj @ P, P = h_k Wc^T, so with gu = LN backward
            //   gP = adj^T gu ;  gWc = gP^T h_k ;  gadj += gu P^T ;  grad h_k = gu + gP Wc + gz_k W_k
            // gP lives in the `gq` region: bf16 planes (tensor-core engine) or fp32 (exact engine).
            const Operand gP = planes_at(gq, gq, MHn);
            if (adjtc)
                XGGM_TRY(adj_apply_tc(coef_hi, coef_lo, g0.hi, g0.lo, nullptr, mut(gP.hi), lo_or_null(gP), B, N, H, 0,
                                      npass(), st));
            else
                XGGM_TRY(adj_apply(adj, gnext, tc ? nullptr : gq, tc ? mut(gP.hi) : nullptr, tc ? lo_or_null(gP) : nullptr,
                                   B, N, H, 1.f, nullptr, 0.f, true, 0, st));
            if (gadj) {   // (a constant adjacency, e.g. the ground-truth graph of the node branch, needs none)
                const float* Pf = saved + L.conv(k, 0);
                if (gram) {
                    // planes of P: saved by the forward tensor-core path, else built here from the fp32 copy
                    const bool saved_planes = adjtc && !fused_adj_ln();
                    const Operand Pop = saved_planes ? planes_at(Pf, saved + L.conv(k, 5), MHn) : planes_at(Pf, work + 5 * MH, MHn);
                    if (!saved_planes) XGGM_TRY(split_one(Pop, MHn, st));
                    XGGM_TRY(gram_tc(g0.hi, g0.lo, Pop.hi, Pop.lo, s_scratch, B, N, H, npass(), st));
                    XGGM_TRY(scale_accum(s_scratch, gadj, BNN, 1.f, nullptr, 1, nullptr, nullptr, st));
                } else {
                    XGGM_TRY(bmm_nt(gnext, Pf, gadj, B, N, H, 1.f, nullptr, 1, nullptr, nullptr, st));
                }
            }
            {   // both weight gradients read h_k (+ the pending last head)
                Lin w[3];
                int n = 0;
                w[n++] = Lin{gP, act(k), nullptr, nullptr, cg[3 * k], nullptr, acc};
                w[n++] = Lin{g1, act(k), nullptr, nullptr, hg[4 * k], nullptr, acc};
                if (have_pending) { w[n++] = pending_w; have_pending = false; }
                XGGM_TRY(wgrad_group(tc, w, n, M, H, H, st));
            }
            {   // gnext (= gu) += [gP | gz_k] [Wc ; W_k]: ONE product with the two contractions concatenated
                Lin d[2];
                d[0] = Lin{gP, wconv[k], nullptr, nullptr, gnext, nullptr, 1};
                d[1] = Lin{g1, whead[k], nullptr, nullptr, gnext, nullptr, 1};
                XGGM_TRY(dgrad_kcat(tc, d, M, H, H, st));
            }
            gh = gnext;
            cur ^= 1;
            continue;
        }
        const Operand pre_op = planes_at(saved + L.conv(k, 0), saved + L.conv(k, 5), MHn);
        // weight gradients of the conv projection and of head k (+ the pending last head) in one launch
        {
            Lin w[3];
            int n = 0;
            w[n++] = Lin{g0, pre_op, nullptr, nullptr, cg[5 * k + 1], nullptr, acc};
            w[n++] = Lin{g1, act(k), nullptr, nullptr, hg[4 * k], nullptr, acc};
            if (have_pending) { w[n++] = pending_w; have_pending = false; }
            XGGM_TRY(wgrad_group(tc, w, n, M, H, H, st));
        }
        // input gradients: gq = gz Wc (with planes for the Gram / message-passing kernels) and gnext = gz_k W_k
        {
            Lin d[2];
            d[0] = Lin{g0, wconv[k], nullptr, nullptr, gq, gram ? &gq_op : nullptr, 0};
            d[1] = Lin{g1, whead[k], nullptr, nullptr, gnext, nullptr, 0};
            XGGM_TRY(dgrad_group(tc, d, 2, M, H, H, st));
        }
        // gadj += (1+eps) gpre h^T ; geps += <gpre h^T, adj>
        if (gram) {
            const Operand hk_op = act(k);
            XGGM_TRY(gram_tc(gq_op.hi, gq_op.lo, hk_op.hi, hk_op.lo, s_scratch, B, N, H, npass(), st));
            XGGM_TRY(scale_accum(s_scratch, gadj, BNN, 1.f, eps, 1, adj, cg[5 * k], st));
        } else {
            XGGM_TRY(bmm_nt(gq, hk, gadj, B, N, H, 1.f, eps, 1, adj, cg[5 * k], st));
        }
        // grad h_k += gpre + (1+eps) adj^T gpre
        if (adjtc) {
            XGGM_TRY(build_blockdiag(adj, coef_hi, npass() == 3 ? coef_lo : nullptr, B, N, 1.f, eps, 1.f, 1, st));
            XGGM_TRY(adj_apply_tc(coef_hi, coef_lo, gq_op.hi, gq_op.lo, gnext, nullptr, nullptr, B, N, H, 1, npass(), st));
        } else {
            XGGM_TRY(adj_apply(adj, gq, gnext, nullptr, nullptr, B, N, H, 1.f, eps, 1.f, true, 1, st));
        }
        gh = gnext;
        cur ^= 1;
    }
    if (have_pending) XGGM_TRY(wgrad_group(tc, &pending_w, 1, M, H, H, st));
    return XGGM_OK;
}

}  // namespace xggm

using namespace xggm;

extern "C" {

int xggm_abi_version(void) { return XGGM_ABI_VERSION; }

const char* xggm_strerror(int code) {
    switch (code) {
        case XGGM_OK: return "ok";
        case XGGM_ERR_ARG: return "invalid argument (shape, NULL pointer or unsupported size)";
        case XGGM_ERR_CUDA: return "CUDA runtime error (see xggm_last_cuda_error)";
        case XGGM_ERR_ARCH: return "device is not compute capability 10.x (B200 / sm_100a required)";
        case XGGM_ERR_UNSUPPORTED: return "not implemented in this build";
        default: return "unknown error";
    }
}

const char* xggm_last_cuda_error(void) { return g_cuda_err; }
int xggm_prof_enable(int on) { return gemm_prof_enable(on); }
int xggm_debug_timeline(unsigned long long* dev_buf) {
    gemm_tc_set_debug(dev_buf);
    return XGGM_OK;
}
int xggm_prof_read(double* total_ms, long long* launches, double* flops) {
    XGGM_REQUIRE(total_ms && launches && flops);
    return gemm_prof_read(total_ms, launches, flops);
}
unsigned long long xggm_launch_count(void) { return g_kernel_launches.load(); }

int xggm_set_device(int device) {
    XGGM_CUDA_TRY(cudaSetDevice(device));
    return XGGM_OK;
}

int xggm_device_check(int device) {
    cudaDeviceProp prop;
    XGGM_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    return prop.major == 10 ? XGGM_OK : XGGM_ERR_ARCH;
}

int xggm_set_precision(int mode) {
    XGGM_REQUIRE(mode == XGGM_PREC_FP32 || mode == XGGM_PREC_BF16 || mode == XGGM_PREC_FP32_SIMT);
    g_precision = mode;
    return XGGM_OK;
}
int xggm_get_precision(void) { return g_precision; }

// work layout of the Linear entry points: P(a)[M,K] | P(w)[N,K] or P(w^T)[K,Np] | P(g)[M,Np]
// Np = N rounded up to 8: planes whose rows have length N are stored with that pitch and zero columns, so a
// ragged output width (encoder_adj: 630, answer head: 2274) stays on the tensor cores.
static inline int pad8i(int n) { return (n + 7) & ~7; }
long long xggm_linear_work_bytes(int M, int N, int K) {
    if (M < 0 || N <= 0 || K <= 0) return -1;
    const long long Np = pad8i(N);
    return 4 * (pad8((long long)M * K) + pad8(Np * K) + pad8((long long)M * Np));
}
static inline bool linear_tc(const void* work, int M, int N, int K) {
    return work != nullptr && M > 0 && g_precision != XGGM_PREC_FP32_SIMT && gemm_tc_ragged_ok(M, N, K);
}

int xggm_linear_fwd_ex(const float* a, const float* w, const float* bias, const float* resid,
                       float* out, int M, int N, int K, void* work, const void* w_planes, xggm_stream_t s) {
    if (M == 0) return XGGM_OK;
    XGGM_REQUIRE(a && w && out && M >= 0 && N > 0 && K > 0);
    const bool tc = linear_tc(work, M, N, K);
    float* wk = static_cast<float*>(work);
    Operand ao{a, nullptr, nullptr}, wo{w, nullptr, nullptr};
    if (tc) {
        ao = planes_at(a, wk, (long long)M * K);
        if (w_planes) {   // prepared weight planes: only the activations are split here
            bf16 *hi, *lo, *thi, *tlo;
            weight_planes_views(const_cast<void*>(w_planes), N, K, &hi, &lo, &thi, &tlo);
            wo = Operand{w, hi, lo};
            XGGM_TRY(split_one(ao, (long long)M * K, as_stream(s)));
        } else {
            wo = planes_at(w, wk + pad8((long long)M * K), (long long)pad8i(N) * K);
            const float* src[2] = {a, w};
            bf16* hi[2] = {const_cast<bf16*>(ao.hi), const_cast<bf16*>(wo.hi)};
            bf16* lo[2] = {const_cast<bf16*>(ao.lo), const_cast<bf16*>(wo.lo)};
            const long long n[2] = {(long long)M * K, (long long)N * K};
            XGGM_TRY(split_planes(src, hi, npass() == 3 ? lo : nullptr, n, 2, as_stream(s)));
        }
    }
    return proj_fwd(tc, ao, wo, bias, resid, out, M, N, K, as_stream(s));
}
int xggm_linear_fwd(const float* a, const float* w, const float* bias, const float* resid,
                    float* out, int M, int N, int K, void* work, xggm_stream_t s) {
    return xggm_linear_fwd_ex(a, w, bias, resid, out, M, N, K, work, nullptr, s);
}
int xggm_linear_bwd_input(const float* g, const float* w, float* ga, int M, int N, int K,
                          int accumulate, void* work, xggm_stream_t s) {
    return xggm_linear_bwd_input_ex(g, w, ga, M, N, K, accumulate, work, nullptr, s);
}
int xggm_linear_bwd_input_ex(const float* g, const float* w, float* ga, int M, int N, int K,
                             int accumulate, void* work, const void* w_planes, xggm_stream_t s) {
    if (M == 0) return XGGM_OK;
    XGGM_REQUIRE(g && w && ga && M >= 0 && N > 0 && K > 0);
    const bool tc = linear_tc(work, M, N, K);
    if (!tc) return gemm_simt(1, g, w, nullptr, nullptr, ga, M, K, N, accumulate, as_stream(s));
    float* wk = static_cast<float*>(work);
    const int Np = pad8i(N);
    Operand wo = planes_at(w, wk + pad8((long long)M * K), (long long)Np * K);                                        // w^T [K,Np]
    const Operand go = planes_at(g, wk + pad8((long long)M * K) + pad8((long long)Np * K), (long long)M * Np);        // g [M,Np]
    if (Np == N) XGGM_TRY(split_one(go, (long long)M * N, as_stream(s)));   // dense rows: the vectorised splitter
    else XGGM_TRY(split_planes_pitched(g, mut(go.hi), lo_or_null(go), M, N, Np, as_stream(s)));
    if (w_planes) {
        bf16 *hi, *lo, *thi, *tlo;
        weight_planes_views(const_cast<void*>(w_planes), N, K, &hi, &lo, &thi, &tlo);
        wo = Operand{w, thi, tlo};
    } else {
        const float* src[1] = {w};
        bf16* hi[1] = {mut(wo.hi)};
        bf16* lo[1] = {mut(wo.lo)};
        XGGM_TRY(split_planes_t(src, hi, npass() == 3 ? lo : nullptr, N, K, 1, as_stream(s), Np));
    }
    // ga[M,K] (+)= g[M,Np] (w^T[K,Np])^T : the zero columns N..Np-1 add nothing
    return gemm_tc(false, false, go.hi, go.lo, wo.hi, wo.lo, nullptr, nullptr, ga, nullptr, nullptr, M, K, Np, accumulate, 0,
                   npass(), as_stream(s));
}
int xggm_linear_bwd_weight(const float* g, const float* a, float* gw, float* gbias, int M, int N,
                           int K, int accumulate, void* work, xggm_stream_t s) {
    XGGM_REQUIRE(gw && M >= 0 && N > 0 && K > 0);
    if (M == 0 && accumulate) return XGGM_OK;
    if (M == 0) {
        XGGM_CUDA_TRY(cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)N * K, as_stream(s)));
        if (gbias) XGGM_CUDA_TRY(cudaMemsetAsync(gbias, 0, sizeof(float) * N, as_stream(s)));
        return XGGM_OK;
    }
    XGGM_REQUIRE(g && a);
    const bool tc = linear_tc(work, M, N, K);
    float* wk = static_cast<float*>(work);
    if (tc) {
        const int Np = pad8i(N);
        const Operand ao = planes_at(a, wk, (long long)M * K);
        const Operand go = planes_at(g, wk + pad8((long long)M * K) + pad8((long long)Np * K), (long long)M * Np);
        if (Np == N) {
            const float* src[2] = {g, a};
            bf16* hi[2] = {mut(go.hi), mut(ao.hi)};
            bf16* lo[2] = {mut(go.lo), mut(ao.lo)};
            const long long n[2] = {(long long)M * N, (long long)M * K};
            XGGM_TRY(split_planes(src, hi, npass() == 3 ? lo : nullptr, n, 2, as_stream(s)));
        } else {
            XGGM_TRY(split_planes_pitched(g, mut(go.hi), lo_or_null(go), M, N, Np, as_stream(s)));
            XGGM_TRY(split_one(ao, (long long)M * K, as_stream(s)));
        }
        // gw[N,K] (+)= g[M,N]^T a[M,K]: A = g planes read MN-major with row pitch Np
        XGGM_TRY(gemm_tc(true, true, go.hi, go.lo, ao.hi, ao.lo, nullptr, nullptr, gw, nullptr, nullptr, N, K, M, accumulate,
                         1, npass(), as_stream(s), Np, 0));
    } else {
        XGGM_TRY(gemm_simt(2, g, a, nullptr, nullptr, gw, N, K, M, accumulate, as_stream(s)));
    }
    if (gbias) XGGM_TRY(colsum(g, gbias, M, N, accumulate, as_stream(s)));
    return XGGM_OK;
}

// work layout of the message-passing entry points: P(x or gout)[B,N,H] | coefficient planes
long long xggm_adj_apply_work_bytes(int B, int N, int H) {
    if (B < 0 || N <= 0 || H <= 0) return -1;
    return 4 * (pad8((long long)B * N * H) + (N <= 128 ? pad8(adj_tc_coef_elems(B, N)) : 0));
}
static inline bool adj_tc_ok(const void* work, int N, int H) {
    return work != nullptr && g_precision != XGGM_PREC_FP32_SIMT && adj_tc_supported(N, H);
}
// out (=|+=) self_w * v + alpha * (adj | adj^T) @ v through the tensor cores
static int adj_apply_tc_f32(const float* adj, const float* v, float* out, int B, int N, int H, float alpha0,
                            const float* alpha_dev, float self_w, int trans, int accumulate, float* work,
                            cudaStream_t st, const void* v_planes = nullptr) {
    const long long n = (long long)B * N * H;
    const Operand vo = planes_at(v, v_planes ? static_cast<float*>(const_cast<void*>(v_planes)) : work, n);
    bf16* chi = reinterpret_cast<bf16*>(work + pad8(n));
    bf16* clo = chi + pad8(adj_tc_coef_elems(B, N));
    if (!v_planes) XGGM_TRY(split_one(vo, n, st));
    XGGM_TRY(build_blockdiag(adj, chi, npass() == 3 ? clo : nullptr, B, N, alpha0, alpha_dev, self_w, trans, st));
    return adj_apply_tc(chi, clo, vo.hi, vo.lo, out, nullptr, nullptr, B, N, H, accumulate, npass(), st);
}

int xggm_adj_apply_fwd(const float* adj, const float* x, float* out, int B, int N, int H,
                       float alpha0, const float* alpha_dev, float self_w, void* work, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(adj && x && out && B >= 0);
    if (adj_tc_ok(work, N, H) && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        return adj_apply_tc_f32(adj, x, out, B, N, H, alpha0, alpha_dev, self_w, 0, 0, static_cast<float*>(work), as_stream(s));
    return adj_apply(adj, x, out, nullptr, nullptr, B, N, H, alpha0, alpha_dev, self_w, false, 0, as_stream(s));
}
int xggm_adj_apply_bwd(const float* adj, const float* x, const float* gout, float* gx,
                       float* gadj_raw, int B, int N, int H, float alpha0, const float* alpha_dev,
                       float self_w, int accumulate_gx, void* work, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(adj && x && gout && gx && B >= 0);
    if (adj_tc_ok(work, N, H) && (reinterpret_cast<uintptr_t>(gx) & 15) == 0)
        XGGM_TRY(adj_apply_tc_f32(adj, gout, gx, B, N, H, alpha0, alpha_dev, self_w, 1, accumulate_gx, static_cast<float*>(work), as_stream(s)));
    else
        XGGM_TRY(adj_apply(adj, gout, gx, nullptr, nullptr, B, N, H, alpha0, alpha_dev, self_w, true, accumulate_gx, as_stream(s)));
    if (gadj_raw) XGGM_TRY(bmm_nt(gout, x, gadj_raw, B, N, H, 1.f, nullptr, 0, nullptr, nullptr, as_stream(s)));
    return XGGM_OK;
}

int xggm_layernorm_fwd(const float* u, const float* gamma, const float* beta, float* h,
                       float* xhat, float* rstd, int M, int H, float eps, xggm_stream_t s) {
    if (M == 0) return XGGM_OK;
    XGGM_REQUIRE(u && gamma && beta && h && M >= 0 && H > 0);
    return layernorm_fwd(u, gamma, beta, h, xhat, rstd, nullptr, nullptr, M, H, eps, as_stream(s));
}
int xggm_layernorm_bwd(const float* gh, const float* xhat, const float* rstd, const float* gamma,
                       float* gu, float* ggamma, float* gbeta, int M, int H, xggm_stream_t s) {
    if (M == 0) return XGGM_OK;
    XGGM_REQUIRE(gh && xhat && rstd && gamma && gu && ggamma && gbeta && M >= 0 && H > 0);
    return layernorm_bwd(gh, xhat, rstd, gamma, gu, ggamma, gbeta, nullptr, nullptr, M, H, as_stream(s));
}
int xggm_gelu_ln_drop_fwd(const float* z, const float* gamma, const float* beta,
                          const uint8_t* keep, float scale, float* out, float* mean, float* rstd,
                          int M, int H, float eps, int accumulate, xggm_stream_t s) {
    if (M == 0) return XGGM_OK;
    XGGM_REQUIRE(z && gamma && beta && out && M >= 0 && H > 0);
    return gelu_ln_drop_fwd(z, gamma, beta, drop_mask(keep, scale), out, mean, rstd, nullptr, nullptr, M, H, eps, accumulate, as_stream(s));
}
int xggm_gelu_ln_drop_bwd(const float* gout, const float* z, const float* mean, const float* rstd,
                          const float* gamma, const uint8_t* keep, float scale, float* gz,
                          float* ggamma, float* gbeta, int M, int H, xggm_stream_t s) {
    if (M == 0) return XGGM_OK;
    XGGM_REQUIRE(gout && z && mean && rstd && gamma && gz && ggamma && gbeta && M >= 0 && H > 0);
    return gelu_ln_drop_bwd(gout, z, mean, rstd, gamma, drop_mask(keep, scale), gz, ggamma, gbeta, nullptr, nullptr, nullptr, M, H, as_stream(s));
}

long long xggm_adj_regen_work_bytes(int B, int N, int H) {
    if (B < 0 || N <= 0 || H <= 0) return -1;
    return 4 * pad8((long long)B * N * H);
}
long long xggm_planes_bytes(long long n_elems) { return n_elems < 0 ? -1 : 4 * pad8(n_elems); }
long long xggm_weight_planes_bytes(int N, int K) { return (N <= 0 || K <= 0) ? -1 : 2 * weight_planes_elems(N, K); }
int xggm_weight_planes_build(const float* const* weights, void* const* planes, const int* N, const int* K, int count,
                             xggm_stream_t s) {
    if (count == 0) return XGGM_OK;
    XGGM_REQUIRE(weights && planes && N && K && count > 0);
    if (g_precision == XGGM_PREC_FP32_SIMT) return XGGM_OK;   // the exact engine reads the fp32 weights
    return weight_planes_build(weights, planes, N, K, count, npass() == 3, as_stream(s));
}
int xggm_adj_regen_fwd_ex(const float* x, float* adj_out, float* S, int32_t* amax, int B, int N,
                          int H, int squash, void* work, const void* x_planes, xggm_stream_t s) {
    XGGM_REQUIRE(B >= 0 && H > 0);
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(x && adj_out);
    if ((work || x_planes) && S && g_precision != XGGM_PREC_FP32_SIMT && gram_tc_supported(N, H)) {
        // pair scores on the tensor cores: planes of x -> Gram kernel -> S, then the per-graph tail
        const long long n = (long long)B * N * H;
        const Operand xo = planes_at(x, static_cast<float*>(x_planes ? const_cast<void*>(x_planes) : work), n);
        if (!x_planes) XGGM_TRY(split_one(xo, n, as_stream(s)));
        XGGM_TRY(gram_tc(xo.hi, xo.lo, xo.hi, xo.lo, S, B, N, H, npass(), as_stream(s)));
        return adj_regen_from_s(S, adj_out, amax, B, N, squash, as_stream(s));
    }
    return adj_regen_fwd(x, adj_out, S, amax, B, N, H, squash, as_stream(s));
}
int xggm_adj_regen_fwd(const float* x, float* adj_out, float* S, int32_t* amax, int B, int N,
                       int H, int squash, void* work, xggm_stream_t s) {
    return xggm_adj_regen_fwd_ex(x, adj_out, S, amax, B, N, H, squash, work, nullptr, s);
}
int xggm_adj_regen_bwd_ex(const float* gadj, const float* x, const float* S, const int32_t* amax,
                          float* gx, float* work, int B, int N, int H, int squash,
                          int accumulate_gx, void* tc_work, const void* x_planes, xggm_stream_t s) {
    XGGM_REQUIRE(B >= 0 && H > 0);
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(gadj && x && S && amax && gx && work);
    if (adj_tc_ok(tc_work, N, H) && (reinterpret_cast<uintptr_t>(gx) & 15) == 0) {
        // D = dS + dS^T per graph (small kernel), then gx (+)= D x on the tensor cores
        XGGM_TRY(adj_regen_bwd_coeffs(gadj, S, amax, work, B, N, squash, as_stream(s)));
        return adj_apply_tc_f32(work, x, gx, B, N, H, 1.f, nullptr, 0.f, 0, accumulate_gx, static_cast<float*>(tc_work),
                                as_stream(s), x_planes);
    }
    return adj_regen_bwd(gadj, x, S, amax, gx, work, B, N, H, squash, accumulate_gx, as_stream(s));
}
int xggm_adj_regen_bwd(const float* gadj, const float* x, const float* S, const int32_t* amax,
                       float* gx, float* work, int B, int N, int H, int squash,
                       int accumulate_gx, void* tc_work, xggm_stream_t s) {
    return xggm_adj_regen_bwd_ex(gadj, x, S, amax, gx, work, B, N, H, squash, accumulate_gx, tc_work, nullptr, s);
}

long long xggm_gnn_saved_floats(int kind, int B, int N, int H, int n_convs) {
    if ((kind != XGGM_KIND_GCN && kind != XGGM_KIND_GIN) || B < 0 || N <= 0 || H <= 0 || n_convs < 0) return -1;
    return GnnLayout(kind, (long long)B * N, H, n_convs).total;
}
long long xggm_gnn_work_floats(int kind, int B, int N, int H, int n_convs) {
    if ((kind != XGGM_KIND_GCN && kind != XGGM_KIND_GIN) || B < 0 || N <= 0 || H <= 0 || n_convs < 0) return -1;
    const GnnLayout L(kind, (long long)B * N, H, n_convs);
    const long long coef = N <= 128 ? adj_tc_coef_elems(B, N) : 0;
    const long long wb = L.work_bwd((long long)B * N * N, coef), wf = L.work_fwd(coef);
    return wb > wf ? wb : wf;
}
int xggm_gnn_fwd_ex(int kind, const float* x, const float* adj, const float* const* conv_params,
                    const float* const* head_params, const uint8_t* const* keeps, const xggm_philox_t* philox,
                    float drop_p, float* out, float* saved, float* work, const void* x_planes, void* out_planes,
                    int B, int N, int H, int n_convs, const void* const* weight_planes, xggm_stream_t s) {
    return gnn_fwd(kind, x, adj, conv_params, head_params, keeps, philox, drop_p, out, saved, work, x_planes, out_planes,
                   B, N, H, n_convs, as_stream(s), weight_planes);
}
int xggm_gnn_fwd(int kind, const float* x, const float* adj, const float* const* conv_params,
                 const float* const* head_params, const uint8_t* const* keeps, const xggm_philox_t* philox,
                 float drop_p, float* out, float* saved, float* work, int B, int N, int H, int n_convs,
                 xggm_stream_t s) {
    return gnn_fwd(kind, x, adj, conv_params, head_params, keeps, philox, drop_p, out, saved, work, nullptr, nullptr,
                   B, N, H, n_convs, as_stream(s));
}
int xggm_gnn_bwd_ex(int kind, const float* gout, const float* x, const float* adj,
                    const float* const* conv_params, const float* const* head_params,
                    const uint8_t* const* keeps, const xggm_philox_t* philox, float drop_p, const float* saved,
                    float* work, float* gx, float* gadj, float* const* conv_grads, float* const* head_grads,
                    int accumulate_param_grads, const void* x_planes, int B, int N, int H, int n_convs,
                    const void* const* weight_planes, xggm_stream_t s) {
    return gnn_bwd(kind, gout, x, adj, conv_params, head_params, keeps, philox, drop_p, saved, work, gx, gadj,
                   conv_grads, head_grads, accumulate_param_grads, x_planes, B, N, H, n_convs, as_stream(s), weight_planes);
}
int xggm_gnn_bwd(int kind, const float* gout, const float* x, const float* adj,
                 const float* const* conv_params, const float* const* head_params,
                 const uint8_t* const* keeps, const xggm_philox_t* philox, float drop_p, const float* saved,
                 float* work, float* gx, float* gadj, float* const* conv_grads, float* const* head_grads,
                 int accumulate_param_grads, int B, int N, int H, int n_convs, xggm_stream_t s) {
    return gnn_bwd(kind, gout, x, adj, conv_params, head_params, keeps, philox, drop_p, saved, work, gx, gadj,
                   conv_grads, head_grads, accumulate_param_grads, nullptr, B, N, H, n_convs, as_stream(s));
}

int xggm_gat_attn_fwd(const float* h, const float* a, const float* adj, float* out, float* att,
                      float* pre, int B, int N, int H, float alpha, int apply_elu, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(h && a && adj && out && att && pre && B >= 0 && H > 0);
    return gat_attn_fwd(h, a, adj, out, att, pre, B, N, H, alpha, apply_elu, as_stream(s));
}
int xggm_gat_attn_bwd(const float* gout, const float* h, const float* a, const float* adj,
                      const float* att, const float* pre, float* gh, float* ga, float* work, int B,
                      int N, int H, float alpha, int apply_elu, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(gout && h && a && adj && att && pre && gh && ga && work && B >= 0 && H > 0);
    return gat_attn_bwd(gout, h, a, adj, att, pre, gh, ga, work, B, N, H, alpha, apply_elu, as_stream(s));
}
int xggm_gelu_fwd(const float* x, float* y, long long n, xggm_stream_t s) {
    if (n == 0) return XGGM_OK;
    XGGM_REQUIRE(x && y && n >= 0);
    return gelu_fwd(x, y, n, as_stream(s));
}
int xggm_gelu_bwd(const float* gy, const float* x, float* gx, long long n, xggm_stream_t s) {
    if (n == 0) return XGGM_OK;
    XGGM_REQUIRE(gy && x && gx && n >= 0);
    return gelu_bwd(gy, x, gx, n, as_stream(s));
}
int xggm_mask_scale(const float* x, const uint8_t* keep, float scale, float* y, long long n,
                    xggm_stream_t s) {
    if (n == 0) return XGGM_OK;
    XGGM_REQUIRE(x && y && n >= 0);
    return mask_scale(x, keep, scale, y, n, as_stream(s));
}
int xggm_avg2_drop(const float* x, const float* y, const uint8_t* keep, float scale, float* out, long long n,
                   xggm_stream_t s) {
    if (n == 0) return XGGM_OK;
    XGGM_REQUIRE(x && y && out && n >= 0);
    return avg2_drop(x, y, keep, scale, out, n, as_stream(s));
}

int xggm_strip_diag(const float* a, float* out, int B, int N, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(a && out && B >= 0 && N > 0);
    return strip_diag(a, out, B, N, as_stream(s));
}
int xggm_triu_scatter_fwd(const float* v, float* adj, int B, int N, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(v && adj && B >= 0 && N > 0);
    return triu_scatter_fwd(v, adj, B, N, as_stream(s));
}
int xggm_triu_scatter_bwd(const float* gadj, float* gv, int B, int N, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(gadj && gv && B >= 0 && N > 0);
    return triu_scatter_bwd(gadj, gv, B, N, as_stream(s));
}
int xggm_edge_noise(const float* adj, const float* randn, double sigma, float* noisy,
                    float* target, int B, int N, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(adj && randn && noisy && target && B >= 0 && N > 0 && sigma != 0.0);
    return edge_noise(adj, randn, (float)sigma, (float)(sigma * sigma), noisy, target, B, N, as_stream(s));
}
int xggm_feat_noise_ex(const float* f, const float* randn, double sigma, float* noisy, float* target,
                       void* noisy_planes, int B, int N, int H, int f_is_broadcast, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(f && randn && noisy && target && B >= 0 && N > 0 && H > 0 && sigma != 0.0);
    const bool emit = noisy_planes && g_precision != XGGM_PREC_FP32_SIMT;
    const Operand o = planes_at(noisy, static_cast<float*>(noisy_planes), (long long)B * N * H);
    return feat_noise(f, randn, (float)sigma, (float)(sigma * sigma), noisy, target, emit ? mut(o.hi) : nullptr,
                      emit ? lo_or_null(o) : nullptr, B, N, H, f_is_broadcast, as_stream(s));
}
int xggm_feat_noise_philox(const float* f, const xggm_philox_t* rng, double sigma, float* noisy, float* target,
                           void* noisy_planes, int B, int N, int H, int f_is_broadcast, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(f && rng && noisy && target && B >= 0 && N > 0 && H > 0 && H % 4 == 0 && sigma != 0.0);
    const bool emit = noisy_planes && g_precision != XGGM_PREC_FP32_SIMT;
    const Operand o = planes_at(noisy, static_cast<float*>(noisy_planes), (long long)B * N * H);
    return feat_noise_philox(f, rng->seed, rng->stream0, rng->dev_epoch, (float)sigma, (float)(sigma * sigma), noisy, target,
                             emit ? mut(o.hi) : nullptr, emit ? lo_or_null(o) : nullptr, B, N, H, f_is_broadcast, as_stream(s));
}
int xggm_feat_noise(const float* f, const float* randn, double sigma, float* noisy, float* target,
                    int B, int N, int H, int f_is_broadcast, xggm_stream_t s) {
    return xggm_feat_noise_ex(f, randn, sigma, noisy, target, nullptr, B, N, H, f_is_broadcast, s);
}
int xggm_sum_nodes(const float* g, float* out, int B, int N, int H, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(g && out && B >= 0 && N > 0 && H > 0);
    return sum_nodes(g, out, B, N, H, as_stream(s));
}
int xggm_score_mse_fwd(const float* score, const float* target, double sigma, float* loss,
                       long long n_elem, xggm_stream_t s) {
    XGGM_REQUIRE(score && target && loss && n_elem >= 0);
    return score_mse_fwd(score, target, (float)sigma, loss, n_elem, as_stream(s));
}
int xggm_score_mse_bwd(const float* score, const float* target, const float* gloss, double sigma,
                       float* gscore, long long n_elem, xggm_stream_t s) {
    if (n_elem == 0) return XGGM_OK;
    XGGM_REQUIRE(score && target && gloss && gscore && n_elem >= 0);
    return score_mse_bwd(score, target, gloss, (float)sigma, gscore, n_elem, as_stream(s));
}
int xggm_sym_kl_fwd(const float* x, const float* y, float* loss, int R, int C, xggm_stream_t s) {
    XGGM_REQUIRE(x && y && loss && R >= 0 && C > 0);
    return sym_kl_fwd(x, y, loss, R, C, as_stream(s));
}
int xggm_sym_kl_bwd(const float* x, const float* y, const float* gloss, float* gx, float* gy,
                    int R, int C, xggm_stream_t s) {
    if (R == 0) return XGGM_OK;
    XGGM_REQUIRE(x && y && gloss && R >= 0 && C > 0);
    return sym_kl_bwd(x, y, gloss, gx, gy, R, C, as_stream(s));
}
int xggm_fuse_readout_fwd(const float* xp, const float* nodes, float* out, int B, int N, int H,
                          xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(xp && nodes && out && B >= 0 && N > 0 && H > 0);
    return fuse_readout_fwd(xp, nodes, out, B, N, H, as_stream(s));
}
int xggm_fuse_readout_bwd(const float* gout, const float* out, float* gxp, float* gnodes, int B,
                          int N, int H, int accumulate_gnodes, xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(gout && out && gnodes && B >= 0 && N > 0 && H > 0);
    return fuse_readout_bwd(gout, out, gxp, gnodes, B, N, H, accumulate_gnodes, as_stream(s));
}
int xggm_node_tail_fwd(const float* nodes, const float* feat, const float* target, const float* xp,
                       double sigma, double kl_w, double sm_w, float* loss, float* cat, int B, int N,
                       int H, xggm_stream_t s) {
    XGGM_REQUIRE(loss && B >= 0 && N > 0 && H > 0);
    if (B == 0) {
        XGGM_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), as_stream(s)));
        return XGGM_OK;
    }
    XGGM_REQUIRE(nodes && feat && target && xp && cat);
    return node_tail_fwd(nodes, feat, target, xp, (float)sigma, (float)kl_w, (float)sm_w, loss, cat, B, N, H, as_stream(s));
}
int xggm_node_tail_bwd(const float* nodes, const float* feat, const float* target, const float* cat,
                       const float* gloss, const float* gcat, double sigma, double kl_w, double sm_w,
                       float* gnodes, float* gfeat, float* gxp, float* grow, int B, int N, int H,
                       xggm_stream_t s) {
    if (B == 0) return XGGM_OK;
    XGGM_REQUIRE(nodes && feat && target && cat && gloss && gcat && gnodes && gxp && grow && B >= 0 && N > 0 && H > 0);
    return node_tail_bwd(nodes, feat, target, cat, gloss, gcat, (float)sigma, (float)kl_w, (float)sm_w, gnodes, gfeat,
                         gxp, grow, B, N, H, as_stream(s));
}
int xggm_bce_logits_fwd(const float* logit, const float* target, double scale, float* loss, long long n,
                        xggm_stream_t s) {
    XGGM_REQUIRE(loss && n >= 0 && (n == 0 || (logit && target)));
    return bce_logits_fwd(logit, target, (float)scale, loss, n, as_stream(s));
}
int xggm_bce_logits_bwd(const float* logit, const float* target, const float* gloss, double scale,
                        float* glogit, long long n, xggm_stream_t s) {
    if (n == 0) return XGGM_OK;
    XGGM_REQUIRE(logit && target && gloss && glogit && n >= 0);
    return bce_logits_bwd(logit, target, gloss, (float)scale, glogit, n, as_stream(s));
}
int xggm_grad_sumsq(const float* g, long long n, float* sumsq, int accumulate, xggm_stream_t s) {
    XGGM_REQUIRE(sumsq && n >= 0 && (n == 0 || g));
    return grad_sumsq(g, n, sumsq, accumulate, as_stream(s));
}
int xggm_bertadam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double b1,
                       double b2, double eps, double weight_decay, const float* sumsq, double max_norm,
                       xggm_stream_t s) {
    if (n == 0) return XGGM_OK;
    XGGM_REQUIRE(p && g && m && v && n >= 0 && b1 >= 0.0 && b1 < 1.0 && b2 >= 0.0 && b2 < 1.0 && eps >= 0.0);
    return bertadam_step(p, g, m, v, n, lr, b1, b2, eps, weight_decay, sumsq, max_norm, nullptr, as_stream(s));
}
int xggm_bertadam_step_ex(float* p, const float* g, float* m, float* v, long long n, double lr, double b1,
                          double b2, double eps, double weight_decay, const float* sumsq, double max_norm,
                          const xggm_lr_schedule_t* sched, xggm_stream_t s) {
    if (n == 0 && !(sched && sched->advance)) return XGGM_OK;
    XGGM_REQUIRE(n >= 0 && (n == 0 || (p && g && m && v)) && b1 >= 0.0 && b1 < 1.0 && b2 >= 0.0 && b2 < 1.0 && eps >= 0.0);
    return bertadam_step(p, g, m, v, n, lr, b1, b2, eps, weight_decay, sumsq, max_norm, sched, as_stream(s));
}
int xggm_visn_tail_supported(int H, int pos_dim) { return visn_tail_supported(H, pos_dim) ? 1 : 0; }
int xggm_visn_tail_fwd(const float* z, const float* boxes, const float* box_w, const float* box_b, const float* gamma1,
                       const float* beta1, const float* gamma2, const float* beta2, const uint8_t* keep, float scale,
                       float* out, float* xhat1, float* rstd1, float* mean2, float* rstd2, int M, int H, float eps,
                       xggm_stream_t s) {
    if (M == 0) return XGGM_OK;
    XGGM_REQUIRE(z && boxes && box_w && box_b && gamma1 && beta1 && gamma2 && beta2 && out && xhat1 && rstd1 && mean2 && rstd2 && M >= 0);
    return visn_tail_fwd(z, boxes, box_w, box_b, gamma1, beta1, gamma2, beta2, drop_mask(keep, scale), out, xhat1, rstd1, mean2, rstd2,
                         M, H, eps, as_stream(s));
}
int xggm_visn_tail_bwd(const float* gout, const float* xhat1, const float* rstd1, const float* boxes, const float* box_w,
                       const float* box_b, const float* mean2, const float* rstd2, const float* gamma1, const float* gamma2,
                       const uint8_t* keep, float scale, float* gz, float* gt, float* ggamma1, float* gbeta1, float* ggamma2,
                       float* gbeta2, int M, int H, xggm_stream_t s) {
    if (M == 0) return XGGM_OK;
    XGGM_REQUIRE(gout && xhat1 && rstd1 && boxes && box_w && box_b && mean2 && rstd2 && gamma1 && gamma2 && gz && gt && ggamma1 &&
                 gbeta1 && ggamma2 && gbeta2 && M >= 0);
    DropSpec d = drop_mask(keep, scale);
    if (!keep) d.scale = scale;      // eval mode / no mask: plain scale (1.0)
    return visn_tail_bwd(gout, xhat1, rstd1, boxes, box_w, box_b, mean2, rstd2, gamma1, gamma2, d, gz, gt, ggamma1, gbeta1, ggamma2,
                         gbeta2, M, H, as_stream(s));
}
int xggm_dp_bertadam_step(const xggm_dp_peers_t* peers, float* m, float* v, long long n, const long long* range_lo,
                          const long long* range_hi, const double* range_lr, int n_ranges, double lr, double b1, double b2, double eps,
                          double weight_decay, double max_norm, const xggm_lr_schedule_t* sched, float* sumsq_out,
                          xggm_stream_t s) {
    XGGM_REQUIRE(b1 >= 0.0 && b1 < 1.0 && b2 >= 0.0 && b2 < 1.0 && eps >= 0.0);
    return dp_bertadam_step(peers, m, v, n, range_lo, range_hi, range_lr, n_ranges, lr, b1, b2, eps, weight_decay, max_norm, sched,
                            sumsq_out, as_stream(s));
}
int xggm_sigmoid_fwd(const float* x, float* y, long long n, xggm_stream_t s) {
    if (n == 0) return XGGM_OK;
    XGGM_REQUIRE(x && y && n >= 0);
    return sigmoid_fwd(x, y, n, as_stream(s));
}
int xggm_sigmoid_bwd(const float* gy, const float* y, float* gx, long long n, xggm_stream_t s) {
    if (n == 0) return XGGM_OK;
    XGGM_REQUIRE(gy && y && gx && n >= 0);
    return sigmoid_bwd(gy, y, gx, n, as_stream(s));
}
int xggm_keep_mask(uint8_t* keep, long long n, float p, uint64_t seed, uint64_t stream_id,
                   const uint64_t* dev_epoch, xggm_stream_t s) {
    if (n == 0) return XGGM_OK;
    XGGM_REQUIRE(keep && n >= 0);
    return keep_mask(keep, n, p, seed, stream_id, dev_epoch, as_stream(s));
}

}  // extern "C"
