// tcgen05 / TMEM / TMA GEMM engine for the dense node projections (sm_100a only).
//
//   C[M,N] (=|+=) sum_k A(m,k) * B(n,k)  (+ bias[n]) (+ resid[m,n])
//
// Operands are bf16 "planes".  fp32 parity mode (NPASS = 3) represents every fp32 value v as
// hi = bf16(v), lo = bf16(v - hi) and issues three tensor-core products per k-step
//   A_lo*B_hi + A_hi*B_lo + A_hi*B_hi      (fp32 accumulation in TMEM)
// which carries 16 mantissa bits per operand (relative error ~1e-5, measured in the tests);
// bf16 mode (NPASS = 1) uses the hi planes only.
//
// Either operand may be K-major (stored [MN, K], k contiguous) or MN-major (stored [K, MN]),
// which covers the three products of a Linear layer without a transpose:
//   forward  out = a w^T : A = a  [M,K]  K-major,   B = w [N,K]   K-major
//   dgrad    ga  = g w   : A = g  [M,N'] K-major,   B = w [N',K'] MN-major
//   wgrad    gw  = g^T a : A = g  [rows,N] MN-major, B = a [rows,K] MN-major  (+ split-K)
//
// Kernel shape: persistent CTAs (one per SM), 320 threads:
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16 per instruction)
//   warps 2..9  epilogue (two per TMEM lane quarter): tcgen05.ld 32x32b -> smem transpose -> coalesced fp32 stores
// Two accumulator buffers in TMEM (2 x BN columns) let the epilogue of tile i overlap the
// main loop of tile i+1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace xggm {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;   // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UK = 16;   // K of one tcgen05.mma for 16-bit operands
constexpr int EPI_WARPS = 8;   // two per TMEM lane quarter: a lone warp's dependent-issue rate bounds the epilogue
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_PITCH = 33;
constexpr int SMEM_LIMIT = 227 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A lost arrival must surface as a CUDA error, never as a hung GPU: trap after 2 s of waiting.
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = global_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (global_ns() - t0 > 2000000000ull) __trap();
    }
}

// ---- cluster helpers (CTA pairs, cta_group::2) --------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a local shared address) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c_inner, int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}
// CTA-pair variant: the data lands in this CTA's smem, the bytes are counted on the barrier at
// `bar_cluster_addr` (the pair leader's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                 int c_inner, int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- tcgen05 -------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 256 x N x 16 across a CTA pair: issued by the leader, A rows / B columns / D rows split over both CTAs
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// pair variant of the commit: arrives on the barrier at the same smem offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}
// arrives on `bar` once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (SWIZZLE_128B, Blackwell version 1).
//   K-major  tile [rows][64 bf16]: 8-row atoms of 1024 B  -> SBO = 1024 (LBO unused)
//   MN-major tile [BK rows][64 bf16] per 64-wide MN atom  -> SBO = 1024 (8 k-rows), LBO = BK*128 (next MN atom)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
// Instruction descriptor: bf16 x bf16 -> fp32, M = 128, N = BN, per-operand major-ness.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void split1(float x, __nv_bfloat16& hi, __nv_bfloat16& lo);
__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const float4& o) {
    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
    split1(o.x, h0, l0); split1(o.y, h1, l1); split1(o.z, h2, l2); split1(o.w, h3, l3);
    __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
    *reinterpret_cast<uint2*>(hi + off) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    if (lo) {
        __nv_bfloat162 c = __halves2bfloat162(l0, l1), d = __halves2bfloat162(l2, l3);
        *reinterpret_cast<uint2*>(lo + off) = make_uint2(*reinterpret_cast<uint32_t*>(&c), *reinterpret_cast<uint32_t*>(&d));
    }
}

// 16-byte vector reduction into global memory (no return value): split-K partial tiles
__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// One launch can carry up to MAX_GROUP independent products of the SAME shape and operand layout
// (e.g. the context projection and the read-out head of one conv level, or their two weight
// gradients): the tile list is the concatenation of the problems' tiles, so the per-launch fixed
// cost (prologue, exposed last epilogue, teardown) is paid once per group instead of once per product.
constexpr int MAX_GROUP = 3;
struct ProbOut {
    const float* bias;   // [N] or null
    const float* resid;  // [M,N] or null
    float* C;
    __nv_bfloat16* c_hi; // optional: also emit C as bf16 planes (vec4 mode only; operand of a later GEMM)
    __nv_bfloat16* c_lo;
    int accumulate;      // C += (non-atomic)
    int pad_;
    const __nv_bfloat16* r_hi;   // optional: the residual given as bf16 operand planes (hi + lo = 16 mantissa bits; lo may be
    const __nv_bfloat16* r_lo;   // null) instead of an fp32 tensor -- lets a producer keep ONLY the planes of a tensor
};
struct MapSet {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
};
struct GroupMaps {
    MapSet m[MAX_GROUP];
};
struct Params {
    int M, N;            // output extent
    int num_kb;          // ceil(K / BK)
    int tiles_m, tiles_n, splits, kb_per_split;
    int group;           // number of problems in this launch (1..MAX_GROUP)
    int tiles_per_prob;  // tiles_m * tiles_n * splits
    int kcat;            // > 0: ONE product whose contraction is the concatenation of two operand pairs (map sets 0 and
                         // 1, kcat k-blocks each): ga = gP Wc + gz Wk as a single K = 2 x 768 GEMM
    ProbOut pr[MAX_GROUP];
    int ldc;
    int atomic;          // split-K partial sums: atomicAdd into a pre-zeroed / pre-initialised C
    int vec4;            // C / resid / bias are 16-byte aligned and N % 4 == 0: float4 epilogue
    // Batched per-graph Gram mode (gram_n > 0): A and B are the SAME-row tiles of two [M,K] K-major
    // tensors, tile t covers the gram_g complete graphs starting at row t*gram_g*gram_n, and only the
    // gram_n x gram_n diagonal blocks  S[b] = P[b] Q[b]^T  are written to C[B][gram_n][gram_n].
    int gram_n, gram_g, gram_b;
    // Block-diagonal mode (bd_stride > 0; message passing  out[b] = C[b] @ x[b]  on the tensor cores):
    // A is a stack of dense 128 x 128 K-major tiles (tile t = the coefficient blocks of the bd_stride/N
    // graphs starting at row t*bd_stride, zero elsewhere), B is x itself read MN-major with its k rows
    // starting at t*bd_stride, and only the first bd_stride rows of each output tile exist.
    int bd_stride;
    // bd_kn > 0: UNALIGNED block-diagonal tiles.  Tile t covers output rows [128 t, 128 t + 128) whatever graphs they
    // fall in; its k rows start at the first row of the first graph it touches, bd_kn * floor(128 t / bd_kn), and
    // the coefficient tile is 128 x (64 num_kb) wide.  No dead rows (128 / 128 instead of 108 / 128 live at N = 36).
    int bd_kn;
    unsigned long long* dbg;   // optional device buffer: CTA 0 records a globaltimer timeline (tools/gemm_timeline.py)
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cta_group::2) per 256 x BN tile; each CTA
// stages its own 128 A rows and HALF of the B tile, so a stage is smaller and the ring deeper.
template <int BN, int NPASS, int CG = 1>
struct Cfg {
    static constexpr int A_TILE = BM * BK * 2;
    static constexpr int B_TILE = (BN / CG) * BK * 2;
    static constexpr int NPLANE = NPASS == 3 ? 2 : 1;
    static constexpr int STAGE_BYTES = NPLANE * (A_TILE + B_TILE);
    static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;
    static constexpr int FIXED = EPI_BYTES + 256 + 1024;  // + barriers + alignment slack
    static constexpr int STAGES_RAW = (SMEM_LIMIT - FIXED) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int SMEM = STAGES * STAGE_BYTES + FIXED;
    static constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
    static_assert(STAGES >= 2, "need at least a double-buffered operand ring");
};

template <int BN, int NPASS, bool A_MN, bool B_MN, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GroupMaps maps, const Params p) {
    using C = Cfg<BN, NPASS, CG>;
    static_assert(CG == 1 || (!A_MN && !B_MN), "the CTA-pair kernel takes K-major operands");
    // K-major A x MN-major B is the message-passing product only: the one kernel whose residual may arrive as operand planes
    constexpr bool PLANE_RESID = !A_MN && B_MN;
    pdl_launch_dependents();   // the next kernel may be scheduled once every CTA of this one is resident
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B, computed as an offset so the pointer keeps its shared-space provenance
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_base = smem;
    float* epi = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);
    uint64_t* full = bars;                    // [STAGES] TMA -> MMA
    uint64_t* empty = bars + C::STAGES;       // [STAGES] MMA -> TMA
    uint64_t* acc_full = bars + 2 * C::STAGES;      // [2] MMA -> epilogue
    uint64_t* acc_empty = bars + 2 * C::STAGES + 2; // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_per_prob * p.group;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;       // 0 = pair leader (issues the MMAs)
    const int worker = blockIdx.x / CG, num_workers = gridDim.x / CG;  // a worker = a CTA or a CTA pair
    const bool trace = p.dbg != nullptr && blockIdx.x == 0;
#define XGGM_TRACE(slot) do { if (trace) p.dbg[slot] = global_ns(); } while (0)
    if (threadIdx.x == 0) XGGM_TRACE(0);

    if (threadIdx.x == 0) {
        for (int g = 0; g < p.group; ++g) {
            prefetch_tmap(&maps.m[g].a_hi);
            prefetch_tmap(&maps.m[g].b_hi);
            if (NPASS == 3) {
                prefetch_tmap(&maps.m[g].a_lo);
                prefetch_tmap(&maps.m[g].b_lo);
            }
        }
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], EPI_WARPS * CG);   // one elected arrival per epilogue warp (of both CTAs)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
        else tmem_alloc(tmem_slot, C::TMEM_COLS);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();   // the peer's barriers must exist before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) XGGM_TRACE(1);
    // everything above (barriers, TMEM, tensor-map prefetch) overlapped the previous kernel's tail; from here
    // on operands / addends written by that kernel are read
    pdl_wait();

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = worker; tile < num_tiles; tile += num_workers) {
                const int prob = tile / p.tiles_per_prob, t = tile - prob * p.tiles_per_prob;
                const int sp = t % p.splits;
                const int mn = t / p.splits;
                int m0 = (mn / p.tiles_n) * (BM * CG) + (int)rank * BM, n0 = (mn % p.tiles_n) * BN;
                if (p.gram_n > 0) m0 = n0 = mn * p.gram_g * p.gram_n;
                int b_krow0 = p.bd_stride > 0 ? (mn / p.tiles_n) * p.bd_stride : 0;   // block-diagonal mode
                if (p.bd_kn > 0) b_krow0 = ((mn / p.tiles_n) * BM / p.bd_kn) * p.bd_kn;
                const int kb0 = sp * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int seg = (p.kcat > 0 && kb >= p.kcat) ? 1 : 0;   // K-concatenated product: second operand pair
                    const MapSet& ms = maps.m[prob + seg];
                    const int kseg = kb - seg * p.kcat;
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (CG == 2) {
                        // pair: both CTAs' bytes are counted on the leader's barrier
                        const uint32_t lead_full = map_to_cta(&full[stage], 0);
                        if (rank == 0) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
                        uint8_t* sa = stage_base + stage * C::STAGE_BYTES;
                        uint8_t* sb = sa + C::NPLANE * C::A_TILE;
                        const int k0 = kseg * BK;
#pragma unroll
                        for (int pl = 0; pl < C::NPLANE; ++pl) {
                            tma_load_2d_pair(sa + pl * C::A_TILE, pl == 0 ? &ms.a_hi : &ms.a_lo, lead_full, k0, m0);
                            tma_load_2d_pair(sb + pl * C::B_TILE, pl == 0 ? &ms.b_hi : &ms.b_lo, lead_full, k0,
                                             n0 + (int)rank * (BN / 2));
                        }
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(&full[stage], C::STAGE_BYTES);
                    uint8_t* sa = stage_base + stage * C::STAGE_BYTES;
                    uint8_t* sb = sa + C::NPLANE * C::A_TILE;
                    const int k0 = kseg * BK;
#pragma unroll
                    for (int pl = 0; pl < C::NPLANE; ++pl) {
                        const CUtensorMap* ma = pl == 0 ? &ms.a_hi : &ms.a_lo;
                        const CUtensorMap* mb = pl == 0 ? &ms.b_hi : &ms.b_lo;
                        if (A_MN) {
#pragma unroll
                            for (int i = 0; i < BM / 64; ++i)
                                tma_load_2d(sa + pl * C::A_TILE + i * (BK * 128), ma, &full[stage], m0 + 64 * i, k0);
                        } else {
                            tma_load_2d(sa + pl * C::A_TILE, ma, &full[stage], k0, m0);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int i = 0; i < BN / 64; ++i)
                                tma_load_2d(sb + pl * C::B_TILE + i * (BK * 128), mb, &full[stage], n0 + 64 * i, b_krow0 + k0);
                        } else {
                            tma_load_2d(sb + pl * C::B_TILE, mb, &full[stage], k0, n0);
                        }
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc(BM * CG, BN, A_MN, B_MN);
            constexpr uint32_t a_lbo = A_MN ? BK * 128 : 16, b_lbo = B_MN ? BK * 128 : 16;
            constexpr uint32_t a_kstep = A_MN ? UK * 128 : UK * 2;  // bytes per 16-wide k-step
            constexpr uint32_t b_kstep = B_MN ? UK * 128 : UK * 2;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = worker; tile < num_tiles; tile += num_workers) {
                const int sp = tile % p.splits;
                const int kb0 = sp * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t first = 1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    if (kb == kb0) XGGM_TRACE(2 + 4 * min(tile / num_workers, 3));      // first operands of the tile landed
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage_base + stage * C::STAGE_BYTES);
                    const uint32_t sb = sa + C::NPLANE * C::A_TILE;
                    // pass order: small cross terms first, hi*hi last
#pragma unroll
                    for (int pass = 0; pass < NPASS; ++pass) {
                        const int apl = (NPASS == 3 && pass == 0) ? 1 : 0;   // A_lo, A_hi, A_hi
                        const int bpl = (NPASS == 3 && pass == 1) ? 1 : 0;   // B_hi, B_lo, B_hi
                        const uint64_t adesc0 = make_sdesc(sa + apl * C::A_TILE, a_lbo, 1024);
                        const uint64_t bdesc0 = make_sdesc(sb + bpl * C::B_TILE, b_lbo, 1024);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            if (CG == 2)
                                umma_bf16_pair(d_tmem, adesc0 + (uint64_t)((k * a_kstep) >> 4),
                                               bdesc0 + (uint64_t)((k * b_kstep) >> 4), idesc, first ? 0u : 1u);
                            else
                                umma_bf16(d_tmem, adesc0 + (uint64_t)((k * a_kstep) >> 4),
                                          bdesc0 + (uint64_t)((k * b_kstep) >> 4), idesc, first ? 0u : 1u);
                            first = 0;
                        }
                    }
                    if (CG == 2) umma_commit_pair(&empty[stage]);   // frees the slot in both CTAs
                    else umma_commit(&empty[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (CG == 2) umma_commit_pair(&acc_full[acc]);
                else umma_commit(&acc_full[acc]);     // accumulator complete -> epilogue
                XGGM_TRACE(3 + 4 * min(tile / num_workers, 3));                          // all MMAs of the tile issued
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ================================ epilogue ====================================
        const int q = warp & 3;  // TMEM lane quarter this warp may read: lanes 32q .. 32q+31
        const int csub = (warp - 2) >> 2;   // which of the quarter's EPI_WARPS/4 warps: takes column chunks c % (EPI_WARPS/4) == csub
        float* st = epi + (warp - 2) * 32 * EPI_PITCH;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = worker; tile < num_tiles; tile += num_workers) {
            const int prob = tile / p.tiles_per_prob, t = tile - prob * p.tiles_per_prob;
            const ProbOut po = p.pr[prob];
            const int mn = t / p.splits;
            const int sp = t % p.splits;
            const int n0 = (mn % p.tiles_n) * BN;
            const int m0 = p.bd_stride > 0 ? (mn / p.tiles_n) * p.bd_stride
                                           : (mn / p.tiles_n) * (BM * CG) + (int)rank * BM;
            const bool lead = (sp == 0);  // bias / residual are added by the first split only
            const int mrow0 = m0 + q * 32;
            const int rows = p.bd_stride > 0 ? min(32, min(p.M - mrow0, p.bd_stride - q * 32)) : min(32, p.M - mrow0);
            const bool vec = p.vec4 && p.gram_n == 0;   // (split-K partial sums use vector reductions)
            // float4 domain: lane -> (row = 4*it + lane/8, 4 columns at 4*(lane%8))
            const int rsub = lane >> 3, c4 = (lane & 7) * 4;
            constexpr int CSTEP = EPI_WARPS / 4;
            // bias + residual + accumulate addends of column chunk `cc` for this lane's 8 float4s.  Issued one
            // chunk AHEAD of their use (the first chunk's before the accumulator is even ready), so their
            // L2/HBM latency hides behind the main loop / the previous chunk instead of stalling the warp.
            // Raw loads only -- no arithmetic on the loaded values here, so the warp does not wait for them.
            // x[] = residual rows, or (when there is no residual) the old C rows of an accumulate.
            auto load_addends = [&](int cc, float4& bv, float4 (&x)[8]) {
                const int nv = n0 + cc * 32 + c4;
                const bool ok = vec && cc < BN / 32 && nv < p.N && rows > 0;   // N % 4 == 0 in vec mode
                bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok && lead && po.bias) bv = __ldg(reinterpret_cast<const float4*>(po.bias + nv));
                const float* src = (lead && po.resid) ? po.resid : ((po.accumulate && !p.atomic) ? po.C : nullptr);
                if constexpr (PLANE_RESID) {
                    if (lead && po.r_hi) {      // residual from bf16 planes: RAW 8-byte loads (hi -> x.xy, lo -> x.zw), decoded at use
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int row = 4 * it + rsub;
                            x[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ok && row < rows) {
                                const size_t off = (size_t)(mrow0 + row) * p.ldc + nv;
                                const uint2 h = *reinterpret_cast<const uint2*>(po.r_hi + off);
                                uint2 l = make_uint2(0u, 0u);
                                if (po.r_lo) l = *reinterpret_cast<const uint2*>(po.r_lo + off);
                                x[it] = make_float4(__uint_as_float(h.x), __uint_as_float(h.y), __uint_as_float(l.x), __uint_as_float(l.y));
                            }
                        }
                        return;
                    }
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = 4 * it + rsub;
                    x[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok && src && row < rows)
                        x[it] = *reinterpret_cast<const float4*>(src + (size_t)(mrow0 + row) * p.ldc + nv);
                }
            };
            constexpr int NCH = (BN / 32) / CSTEP;   // column chunks per epilogue warp
            float4 bvs[NCH + 1];
            float4 xs[NCH + 1][8];
            load_addends(csub, bvs[0], xs[0]);

            mbar_wait(&acc_full[acc], acc_phase);
            if (warp == 2 && lane == 0) XGGM_TRACE(4 + 4 * min(tile / num_workers, 3)); // accumulator ready
            tc_fence_after();
            if (p.gram_n > 0) {
                // thread = accumulator row r of the tile; keep the columns of r's own graph only
                const int gn = p.gram_n, span = p.gram_g * gn;
                const int r = q * 32 + lane;
                const int gl = r / gn;                       // graph slot of this row inside the tile
                const long long b = (long long)mn * p.gram_g + gl;
                const bool row_ok = r < span && b < p.gram_b;
                const int lo_col = (q * 32) / gn * gn;                       // warp-uniform column window
                const int hi_col = min(span, ((q * 32 + 31) / gn + 1) * gn);
                float* orow = po.C + (b * gn + (r - gl * gn)) * gn - gl * gn;  // orow[col] = S[b][i][col - gl*gn]
#pragma unroll 1
                for (int c = csub; c < BN / 32; c += CSTEP) {
                    if (c * 32 + 32 <= lo_col || c * 32 >= hi_col) continue;  // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(tmem_base + acc * BN + c * 32 + ((uint32_t)(q * 32) << 16), v);
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = c * 32 + j;
                            if (col >= gl * gn && col < gl * gn + gn) {
                                if (p.atomic) atomicAdd(&orow[col], __uint_as_float(v[j]));
                                else orow[col] = __uint_as_float(v[j]);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
                continue;
            }
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {   // fully unrolled: the addend registers ping-pong without moves
                const int c = csub + ci * CSTEP;
                const int ncol0 = n0 + c * 32;
                const bool live = ncol0 < p.N && rows > 0;  // warp-uniform
                // (1) next chunk's addends go on the wire now
                load_addends(c + CSTEP, bvs[ci + 1], xs[ci + 1]);
                if (!live) continue;
                // (2) accumulator chunk: TMEM -> registers (thread = row) -> smem transpose
                uint32_t v[32];
                tmem_ld32(tmem_base + acc * BN + c * 32 + ((uint32_t)(q * 32) << 16), v);
                // vec mode: 16-byte accesses into a pitch-32 tile whose float4 index is XOR-swizzled with the row
                // (conflict-free for the row-wise writes here and the 4-row x 8-float4 reads below)
                float4* st4 = reinterpret_cast<float4*>(st);
                if (vec) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st4[lane * 8 + (j ^ (lane & 7))] =
                            make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                        __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) st[lane * EPI_PITCH + j] = __uint_as_float(v[j]);
                }
                __syncwarp();
                // (3) coalesced stores
                if (vec) {
                    const int nv = ncol0 + c4;
                    if (nv < p.N) {
                        const bool both = lead && po.resid && po.accumulate && !p.atomic;   // rare: second addend read late
                        float4 o[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int row = 4 * it + rsub;
                            const float4 sp = st4[row * 8 + ((lane & 7) ^ (row & 7))];
                            const float4 bv = bvs[ci];
                            float4 x = xs[ci][it];
                            if constexpr (PLANE_RESID) {
                                if (lead && po.r_hi) {   // decode the raw plane words: bf16 pairs, value = hi + lo
                                    const uint32_t h0 = __float_as_uint(x.x), h1 = __float_as_uint(x.y), l0 = __float_as_uint(x.z), l1 = __float_as_uint(x.w);
                                    x = make_float4(__uint_as_float(h0 << 16) + __uint_as_float(l0 << 16),
                                                    __uint_as_float(h0 & 0xFFFF0000u) + __uint_as_float(l0 & 0xFFFF0000u),
                                                    __uint_as_float(h1 << 16) + __uint_as_float(l1 << 16),
                                                    __uint_as_float(h1 & 0xFFFF0000u) + __uint_as_float(l1 & 0xFFFF0000u));
                                }
                            }
                            o[it] = make_float4(sp.x + bv.x + x.x, sp.y + bv.y + x.y, sp.z + bv.z + x.z, sp.w + bv.w + x.w);
                        }
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int row = 4 * it + rsub;
                            if (row < rows) {
                                const size_t off = (size_t)(mrow0 + row) * p.ldc + nv;
                                if (both) {
                                    const float4 a = *reinterpret_cast<const float4*>(po.C + off);
                                    o[it].x += a.x; o[it].y += a.y; o[it].z += a.z; o[it].w += a.w;
                                }
                                if (p.atomic) red_add_v4(po.C + off, o[it]);   // split-K partial sum
                                else if (po.C) *reinterpret_cast<float4*>(po.C + off) = o[it];
                                if (po.c_hi) store_planes4(po.c_hi, po.c_lo, off, o[it]);
                            }
                        }
                    }
                } else {
                    const int n = ncol0 + lane;
                    if (n < p.N) {
                        const float bv = (lead && po.bias) ? po.bias[n] : 0.f;
#pragma unroll 4
                        for (int i = 0; i < rows; ++i) {
                            float val = st[i * EPI_PITCH + lane] + bv;
                            const size_t o = (size_t)(mrow0 + i) * p.ldc + n;
                            if (lead && po.resid) val += po.resid[o];
                            if (p.atomic) atomicAdd(&po.C[o], val);
                            else po.C[o] = po.accumulate ? po.C[o] + val : val;
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (warp == 2 && lane == 0) XGGM_TRACE(5 + 4 * min(tile / num_workers, 3)); // epilogue of the tile done
            if (lane == 0) {   // hand the accumulator buffer back to the (pair leader's) MMA thread
                if (CG == 2 && rank != 0) mbar_arrive_cluster(map_to_cta(&acc_empty[acc], 0));
                else mbar_arrive(&acc_empty[acc]);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all();   // neither CTA may free TMEM / exit while the pair's MMAs or signals are in flight
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
        else tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
    if (threadIdx.x == 0) XGGM_TRACE(18);
#undef XGGM_TRACE
}

// ---- fp32 -> bf16 hi/lo planes -----------------------------------------------------------------
__device__ __forceinline__ void split1(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    const float h = __bfloat162float(hi);
    const float r = x - h;
    lo = __float2bfloat16_rn((h - h == 0.f) ? r : 0.f);  // inf/nan: keep them in hi only
}

struct SplitJob {
    const float* src;
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;  // may be null (bf16 mode)
    long long n;
};
constexpr int MAX_SPLIT_JOBS = 8;
struct SplitJobs {
    SplitJob j[MAX_SPLIT_JOBS];
    int count;
};

__global__ void __launch_bounds__(256) split_planes_kernel(const SplitJobs jobs) {
    pdl_prologue();
    const SplitJob job = jobs.j[blockIdx.y];
    const long long n4 = job.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4*>(job.src)[i];
        __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
        split1(v.x, h0, l0); split1(v.y, h1, l1); split1(v.z, h2, l2); split1(v.w, h3, l3);
        __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
        uint2 hv = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        reinterpret_cast<uint2*>(job.hi)[i] = hv;
        if (job.lo) {
            __nv_bfloat162 c = __halves2bfloat162(l0, l1), d = __halves2bfloat162(l2, l3);
            reinterpret_cast<uint2*>(job.lo)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&c), *reinterpret_cast<uint32_t*>(&d));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // tail (n % 4)
        for (long long i = n4 << 2; i < job.n; ++i) {
            __nv_bfloat16 h, l;
            split1(job.src[i], h, l);
            job.hi[i] = h;
            if (job.lo) job.lo[i] = l;
        }
    }
}

// pitched split: src [R,C] fp32 -> hi/lo [R,P] bf16 with zero columns C..P-1 (P % 8 == 0)
__global__ void __launch_bounds__(256)
split_planes_pitched_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                            long long R, int C, int P) {
    pdl_prologue();
    const long long n = R * P, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long r = i / P;
        const int c = (int)(i - r * P);
        __nv_bfloat16 h, l;
        split1(c < C ? src[r * C + c] : 0.f, h, l);
        hi[i] = h;
        if (lo) lo[i] = l;
    }
}

// transposed split: src [R,C] fp32 -> hi/lo [C,R] bf16 (weights only: dgrad reads W^T K-major)
struct SplitTJob {
    const float* src;
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
};
struct SplitTJobs {
    SplitTJob j[MAX_SPLIT_JOBS];
};
__global__ void __launch_bounds__(256) split_planes_t_kernel(const SplitTJobs jobs, int R, int C, int P) {
    pdl_prologue();
    __shared__ float tile[32][33];
    const SplitTJob job = jobs.j[blockIdx.z];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < R && c < C) ? job.src[(size_t)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;   // output row = source column; output pitch P >= R, zero padded
        if (c < C && r < P) {
            __nv_bfloat16 h, l;
            split1(r < R ? tile[tx][i] : 0.f, h, l);
            job.hi[(size_t)c * P + r] = h;
            if (job.lo) job.lo[(size_t)c * P + r] = l;
        }
    }
}

// Block-diagonal coefficient tiles for message passing on the tensor cores.  Tile t (one CTA) is a
// dense [128][128] K-major bf16 matrix: for each of the G = 128/N graphs b = t*G + g it holds
//   C_b = alpha * (adj_b | adj_b^T) + self_w * I     at rows/cols [g*N, g*N + N), zero elsewhere.
// grid (tiles, 4): each CTA writes 32 rows of a tile; a thread produces 8 consecutive columns (one 16-byte store per plane).
__global__ void __launch_bounds__(128)
build_blockdiag_kernel(const float* __restrict__ adj, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                       int B, int N, int G, float alpha0, const float* __restrict__ alpha_dev, float self_w, int trans) {
    pdl_prologue();
    const int t = blockIdx.x;
    const float alpha = alpha0 + (alpha_dev ? alpha_dev[0] : 0.f);
    __nv_bfloat16* th = hi + (size_t)t * BM * BM;
    __nv_bfloat16* tl = lo ? lo + (size_t)t * BM * BM : nullptr;
    for (int v = threadIdx.x; v < 32 * (BM / 8); v += blockDim.x) {
        const int r = blockIdx.y * 32 + v / (BM / 8), c0 = (v % (BM / 8)) * 8;
        const int g = r / N, i = r - g * N;
        const long long b = (long long)t * G + g;
        const bool row_live = g < G && b < B;
        const float* ab = adj + (size_t)(row_live ? b : 0) * N * N;
        __align__(16) __nv_bfloat16 hv[8], lv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = c0 + u - g * N;
            float val = 0.f;
            if (row_live && j >= 0 && j < N) {
                val = alpha * ab[trans ? j * N + i : i * N + j];
                if (i == j) val += self_w;
            }
            split1(val, hv[u], lv[u]);
        }
        *reinterpret_cast<uint4*>(th + (size_t)r * BM + c0) = *reinterpret_cast<const uint4*>(hv);
        if (tl) *reinterpret_cast<uint4*>(tl + (size_t)r * BM + c0) = *reinterpret_cast<const uint4*>(lv);
    }
}

// Unaligned variant (Params::bd_kn): tile t = output rows [128 t, 128 t + 128), columns = the KW k rows starting at
// row N * floor(128 t / N) (the first row of the first graph the tile touches).  Dense [128][KW] K-major tile.
__global__ void __launch_bounds__(256)
build_blockdiag_u_kernel(const float* __restrict__ adj, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                         long long M, int N, int KW, float alpha0, const float* __restrict__ alpha_dev, float self_w, int trans) {
    pdl_prologue();
    const int t = blockIdx.x;
    const float alpha = alpha0 + (alpha_dev ? alpha_dev[0] : 0.f);
    const long long kstart = ((long long)t * BM / N) * N;
    __nv_bfloat16* th = hi + (size_t)t * BM * KW;
    __nv_bfloat16* tl = lo ? lo + (size_t)t * BM * KW : nullptr;
    const int vec_per_row = KW / 8;
    for (int v = threadIdx.x; v < 16 * vec_per_row; v += blockDim.x) {     // grid.y = 8: 16 rows of the tile per CTA
        const int r = blockIdx.y * 16 + v / vec_per_row, c0 = (v % vec_per_row) * 8;
        const long long m = (long long)t * BM + r;
        const long long b = m / N;
        const int i = (int)(m - b * N);
        const bool row_live = m < M;
        const float* ab = adj + (size_t)(row_live ? b : 0) * N * N;
        __align__(16) __nv_bfloat16 hv[8], lv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long j = kstart + c0 + u - b * N;
            float val = 0.f;
            if (row_live && j >= 0 && j < N) {
                val = alpha * ab[trans ? j * N + i : (long long)i * N + j];
                if (i == j) val += self_w;
            }
            split1(val, hv[u], lv[u]);
        }
        *reinterpret_cast<uint4*>(th + (size_t)r * KW + c0) = *reinterpret_cast<const uint4*>(hv);
        if (tl) *reinterpret_cast<uint4*>(tl + (size_t)r * KW + c0) = *reinterpret_cast<const uint4*>(lv);
    }
}

// Prepared weight planes: ONE launch turns up to MAX_WP_JOBS fp32 matrices W [N,K] into BOTH operand layouts the
// projections read -- the K-major planes of W (forward) and the planes of W^T [K,P] (P = N rounded up to 8, zero
// padded; the input-gradient product) -- so that a training step splits its weights once (after the optimiser
// update) instead of once per layer call and direction.  32 x 32 tiles through shared memory: both outputs are
// written with coalesced rows.
constexpr int MAX_WP_JOBS = 16;
struct WPlaneJob {
    const float* src;
    __nv_bfloat16* hi;    // [N,K]
    __nv_bfloat16* lo;    // may be null (bf16 engine)
    __nv_bfloat16* thi;   // [K,P]
    __nv_bfloat16* tlo;
    int N, K, P;
};
struct WPlaneJobs {
    WPlaneJob j[MAX_WP_JOBS];
};
__global__ void __launch_bounds__(256) weight_planes_kernel(const WPlaneJobs jobs) {
    pdl_prologue();
    __shared__ float tile[64][65];
    const WPlaneJob job = jobs.j[blockIdx.z];
    const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;    // W rows n, columns k; K and P are multiples of 8
    if (k0 >= job.K || n0 >= job.P) return;
    // phase 1: a thread takes two consecutive k of one row -> one 4-byte store per plane (128 B per warp and row)
    for (int v = threadIdx.x; v < 64 * 32; v += blockDim.x) {
        const int i = v >> 5, n = n0 + i, k = k0 + 2 * (v & 31);
        float2 w = make_float2(0.f, 0.f);
        if (n < job.N && k < job.K) w = *reinterpret_cast<const float2*>(job.src + (size_t)n * job.K + k);
        tile[i][2 * (v & 31)] = w.x;
        tile[i][2 * (v & 31) + 1] = w.y;
        if (n < job.N && k < job.K) {
            __nv_bfloat16 h0, l0, h1, l1;
            split1(w.x, h0, l0);
            split1(w.y, h1, l1);
            *reinterpret_cast<__nv_bfloat162*>(job.hi + (size_t)n * job.K + k) = __halves2bfloat162(h0, h1);
            if (job.lo) *reinterpret_cast<__nv_bfloat162*>(job.lo + (size_t)n * job.K + k) = __halves2bfloat162(l0, l1);
        }
    }
    __syncthreads();
    // phase 2: W^T row k, two consecutive n (pitch P, zero padded beyond N)
    for (int v = threadIdx.x; v < 64 * 32; v += blockDim.x) {
        const int i = v >> 5, k = k0 + i, nl = 2 * (v & 31), n = n0 + nl;
        if (k < job.K && n < job.P) {
            __nv_bfloat16 h0, l0, h1, l1;
            split1(n < job.N ? tile[nl][i] : 0.f, h0, l0);
            split1(n + 1 < job.N ? tile[nl + 1][i] : 0.f, h1, l1);
            *reinterpret_cast<__nv_bfloat162*>(job.thi + (size_t)k * job.P + n) = __halves2bfloat162(h0, h1);
            if (job.tlo) *reinterpret_cast<__nv_bfloat162*>(job.tlo + (size_t)k * job.P + n) = __halves2bfloat162(l0, l1);
        }
    }
}

// =================================================================================================
// Message passing + LayerNorm in ONE kernel (GCNConv, src/module/gcn.py:28-29, re-associated):
//     h_next = LN( h + adj_b @ P_b ) * gamma + beta            xhat, rstd saved; h_next leaves as operand planes
// A row-complete epilogue needs all H = 768 columns of a row, i.e. 768 fp32 TMEM columns where an SM has 512: the row
// block is therefore split over a CLUSTER OF FOUR CTAs, one 192-column tile each.  Every CTA runs the (tiny, K = 192)
// block-diagonal product of its tile into TMEM, adds the residual (from h's operand planes, staged through shared
// memory into the thread-per-row domain), keeps per-row sum / sum of squares of its 192 columns, writes u back INTO
// TMEM (tcgen05.st), and the four CTAs exchange the partial statistics through distributed shared memory
// (st.shared::cluster + an mbarrier every epilogue warp of the cluster arrives on).  A second pass over TMEM
// normalises and stores xhat, the planes of h_next (and optionally fp32 h_next) with coalesced 16-byte accesses.
// u never touches global memory and the separate LayerNorm kernel disappears.
// One tile per CTA (grid = 4 x row tiles, cluster dimension 4); warp roles as in gemm_tc_kernel.
// =================================================================================================
struct AdjLnParams {
    int M, H, num_kb, bd_stride, bd_kn;
    const __nv_bfloat16* r_hi;   // residual h as operand planes (lo may be null: bf16 engine)
    const __nv_bfloat16* r_lo;
    const float* gamma;
    const float* beta;
    float* xhat;                 // [M,H] fp32, or bf16 when xhat_bf16
    float* rstd;                 // [M]
    float* h_out;                // optional fp32 h_next
    __nv_bfloat16* o_hi;         // planes of h_next
    __nv_bfloat16* o_lo;
    float eps;
    int xhat_bf16;
};
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const uint64_t t0 = global_ns();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (global_ns() - t0 > 2000000000ull) __trap();
    }
}
__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}

constexpr int ALN_BN = 192;
constexpr int ALN_CL = 4;       // CTAs per cluster = column tiles per row block (H = 768)
template <int NPASS>
struct AlnCfg {
    static constexpr int A_TILE = BM * BK * 2;
    static constexpr int B_TILE = ALN_BN * BK * 2;
    static constexpr int NPLANE = NPASS == 3 ? 2 : 1;
    static constexpr int STAGE_BYTES = NPLANE * (A_TILE + B_TILE);
    static constexpr int STAGES = 2;
    static constexpr int EPI_BYTES = EPI_WARPS * 32 * 32 * 4;                  // one swizzled [32][8 float4] tile per warp
    static constexpr int STAT_BYTES = 2 * ALN_CL * BM * 8;                     // [8 sources][128 rows] float2
    static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + STAT_BYTES + 256 + 1024;
    static_assert(SMEM <= SMEM_LIMIT, "adj_ln_tc shared memory");
};

template <int NPASS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
adj_ln_tc_kernel(const __grid_constant__ GroupMaps maps, const AdjLnParams p) {
    using C = AlnCfg<NPASS>;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_base = smem;
    float4* epi = reinterpret_cast<float4*>(smem + C::STAGES * C::STAGE_BYTES);
    float2* stats = reinterpret_cast<float2*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);   // [2*ALN_CL][BM]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES + C::STAT_BYTES);
    uint64_t* full = bars;                  // [STAGES]
    uint64_t* empty = bars + C::STAGES;     // [STAGES]
    uint64_t* acc_full = bars + 2 * C::STAGES;
    uint64_t* stat_bar = bars + 2 * C::STAGES + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();          // column tile of this CTA
    const int t = blockIdx.x / ALN_CL;                 // row tile of this cluster
    const MapSet& ms = maps.m[0];
    if (threadIdx.x == 0) {
        prefetch_tmap(&ms.a_hi);
        prefetch_tmap(&ms.b_hi);
        if (NPASS == 3) { prefetch_tmap(&ms.a_lo); prefetch_tmap(&ms.b_lo); }
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_init(stat_bar, ALN_CL * EPI_WARPS);       // one elected arrival per epilogue warp of the whole cluster
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    cluster_sync_all();                                // the peers' barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    const int n0 = (int)crank * ALN_BN;
    const int m0 = t * BM;                             // rows of the coefficient tile
    const int out0 = p.bd_kn > 0 ? t * BM : t * p.bd_stride;                       // first output row of the tile
    const int live = min(p.bd_kn > 0 ? BM : p.bd_stride, p.M - out0);             // output rows that exist
    const int b_krow0 = p.bd_kn > 0 ? (t * BM / p.bd_kn) * p.bd_kn : t * p.bd_stride;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int stage = kb % C::STAGES;
                if (kb >= C::STAGES) mbar_wait(&empty[stage], ((kb / C::STAGES) - 1) & 1);
                mbar_expect_tx(&full[stage], C::STAGE_BYTES);
                uint8_t* sa = stage_base + stage * C::STAGE_BYTES;
                uint8_t* sb = sa + C::NPLANE * C::A_TILE;
#pragma unroll
                for (int pl = 0; pl < C::NPLANE; ++pl) {
                    tma_load_2d(sa + pl * C::A_TILE, pl == 0 ? &ms.a_hi : &ms.a_lo, &full[stage], kb * BK, m0);
#pragma unroll
                    for (int i = 0; i < ALN_BN / 64; ++i)
                        tma_load_2d(sb + pl * C::B_TILE + i * (BK * 128), pl == 0 ? &ms.b_hi : &ms.b_lo, &full[stage],
                                    n0 + 64 * i, b_krow0 + kb * BK);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, ALN_BN, false, true);
            uint32_t first = 1;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int stage = kb % C::STAGES;
                mbar_wait(&full[stage], (kb / C::STAGES) & 1);
                tc_fence_after();
                const uint32_t sa = smem_u32(stage_base + stage * C::STAGE_BYTES);
                const uint32_t sb = sa + C::NPLANE * C::A_TILE;
#pragma unroll
                for (int pass = 0; pass < NPASS; ++pass) {
                    const int apl = (NPASS == 3 && pass == 0) ? 1 : 0;
                    const int bpl = (NPASS == 3 && pass == 1) ? 1 : 0;
                    const uint64_t adesc0 = make_sdesc(sa + apl * C::A_TILE, 16, 1024);
                    const uint64_t bdesc0 = make_sdesc(sb + bpl * C::B_TILE, BK * 128, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        umma_bf16(tmem_base, adesc0 + (uint64_t)((k * UK * 2) >> 4), bdesc0 + (uint64_t)((k * UK * 128) >> 4), idesc,
                                  first ? 0u : 1u);
                        first = 0;
                    }
                }
                umma_commit(&empty[stage]);
            }
            umma_commit(acc_full);
        }
    } else {
        // ================================ epilogue: two passes over TMEM ====================================
        const int q = warp & 3, csub = (warp - 2) >> 2;
        float4* st4 = epi + (warp - 2) * 32 * 8;
        const int rsub = lane >> 3, c4i = lane & 7;
        const int rows = max(0, min(32, live - q * 32));           // live rows of this warp's lane quarter
        const int mrow0 = out0 + q * 32;
        constexpr int NCH = (ALN_BN / 32) / 2;                      // column chunks per warp (two warps per quarter)
        mbar_wait(acc_full, 0);
        tc_fence_after();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const int c = csub + 2 * ci;
            const int nv = n0 + c * 32 + 4 * c4i;
            // (a) residual chunk: coalesced 8-byte plane loads in the store-domain mapping -> swizzled smem tile
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int row = 4 * it + rsub;
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < rows) {
                    const size_t off = (size_t)(mrow0 + row) * p.H + nv;
                    const uint2 h = *reinterpret_cast<const uint2*>(p.r_hi + off);
                    r = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xFFFF0000u),
                                    __uint_as_float(h.y << 16), __uint_as_float(h.y & 0xFFFF0000u));
                    if (p.r_lo) {
                        const uint2 l = *reinterpret_cast<const uint2*>(p.r_lo + off);
                        r.x += __uint_as_float(l.x << 16); r.y += __uint_as_float(l.x & 0xFFFF0000u);
                        r.z += __uint_as_float(l.y << 16); r.w += __uint_as_float(l.y & 0xFFFF0000u);
                    }
                }
                st4[row * 8 + (c4i ^ (row & 7))] = r;
            }
            __syncwarp();
            // (b) accumulator chunk -> registers (thread = row), (c) u = acc + residual, row statistics, (d) u back to TMEM
            uint32_t v[32];
            const uint32_t taddr = tmem_base + c * 32 + ((uint32_t)(q * 32) << 16);
            tmem_ld32(taddr, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 r = st4[lane * 8 + (j ^ (lane & 7))];
                const float u0 = __uint_as_float(v[4 * j]) + r.x, u1 = __uint_as_float(v[4 * j + 1]) + r.y;
                const float u2 = __uint_as_float(v[4 * j + 2]) + r.z, u3 = __uint_as_float(v[4 * j + 3]) + r.w;
                s1 += (u0 + u1) + (u2 + u3);
                s2 = fmaf(u0, u0, fmaf(u1, u1, fmaf(u2, u2, fmaf(u3, u3, s2))));
                v[4 * j] = __float_as_uint(u0); v[4 * j + 1] = __float_as_uint(u1);
                v[4 * j + 2] = __float_as_uint(u2); v[4 * j + 3] = __float_as_uint(u3);
            }
            tmem_st32(taddr, v);
            __syncwarp();
        }
        // exchange the partial statistics: slot (crank, csub) of every CTA of the cluster, row q*32 + lane
        {
            float2* mine = stats + (crank * 2 + csub) * BM + q * 32 + lane;
#pragma unroll
            for (uint32_t k = 0; k < ALN_CL; ++k) st_cluster_f2(map_to_cta(mine, k), s1, s2);
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (uint32_t k = 0; k < ALN_CL; ++k) mbar_arrive_cluster(map_to_cta(stat_bar, k));
            }
        }
        mbar_wait_cluster(stat_bar, 0);
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * ALN_CL; ++k) {
            const float2 w = stats[k * BM + q * 32 + lane];
            t1 += w.x; t2 += w.y;
        }
        const float inv_h = 1.0f / (float)p.H;
        const float mean = t1 * inv_h;
        const float var = fmaxf(t2 * inv_h - mean * mean, 0.f);
        const float rstd = 1.0f / sqrtf(var + p.eps);
        if (crank == 0 && csub == 0 && lane < rows && p.rstd) p.rstd[mrow0 + lane] = rstd;
        // pass 2: normalise, transpose through smem, coalesced stores of xhat / h_next planes (/ fp32 h_next)
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const int c = csub + 2 * ci;
            const int nv = n0 + c * 32 + 4 * c4i;
            uint32_t v[32];
            tmem_ld32(tmem_base + c * 32 + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                st4[lane * 8 + (j ^ (lane & 7))] =
                    make_float4((__uint_as_float(v[4 * j]) - mean) * rstd, (__uint_as_float(v[4 * j + 1]) - mean) * rstd,
                                (__uint_as_float(v[4 * j + 2]) - mean) * rstd, (__uint_as_float(v[4 * j + 3]) - mean) * rstd);
            __syncwarp();
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + nv));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + nv));
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int row = 4 * it + rsub;
                if (row < rows) {
                    const float4 x = st4[row * 8 + (c4i ^ (row & 7))];
                    const size_t off = (size_t)(mrow0 + row) * p.H + nv;
                    if (p.xhat) {
                        if (p.xhat_bf16) {
                            __nv_bfloat162 a = __floats2bfloat162_rn(x.x, x.y), bb = __floats2bfloat162_rn(x.z, x.w);
                            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.xhat) + off) =
                                make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&bb));
                        } else {
                            *reinterpret_cast<float4*>(p.xhat + off) = x;
                        }
                    }
                    const float4 h = make_float4(fmaf(x.x, g.x, b.x), fmaf(x.y, g.y, b.y), fmaf(x.z, g.z, b.z), fmaf(x.w, g.w, b.w));
                    if (p.h_out) *reinterpret_cast<float4*>(p.h_out + off) = h;
                    if (p.o_hi) store_planes4(p.o_hi, p.o_lo, off, h);
                }
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    tc_fence_before();
    cluster_sync_all();      // no CTA may exit (or free TMEM) while a peer can still write its statistics / barrier
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D bf16 tensor [rows][cols] (cols contiguous), box = [box_rows][64], SWIZZLE_128B, zero OOB fill.
// `pitch` (elements, 0 = cols): row pitch of a padded plane (ragged widths are stored with a multiple-of-8 pitch).
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows, long long pitch = 0) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return XGGM_ERR_UNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)(pitch > 0 ? pitch : cols) * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? XGGM_OK : XGGM_ERR_ARG;
}

// per-launch timing lives in gemm_simt.cu
void* gemm_prof_begin(double flops, cudaStream_t st, int members = 1);
void gemm_prof_end(void* rec, cudaStream_t st);

static unsigned long long* g_tc_dbg = nullptr;   // see xggm_debug_timeline()
void gemm_tc_set_debug(unsigned long long* dev_buf) { g_tc_dbg = dev_buf; }
static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int BN, int NPASS, bool A_MN, bool B_MN, int CG = 1>
static int launch_tc(const tc::GroupMaps& maps, const tc::Params& p, int grid, cudaStream_t st) {
    using C = tc::Cfg<BN, NPASS, CG>;
    static bool attr_set = false;
    auto kern = tc::gemm_tc_kernel<BN, NPASS, A_MN, B_MN, CG>;
    if (!attr_set) {
        XGGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    if (CG == 1) {
        launch_kernel(pdl_mode() == 1 || pdl_mode() == 2, kern, grid, tc::NUM_THREADS, C::SMEM, st, maps, p);
    } else {  // CTA pairs: clusters of 2 (same TPC) so cta_group::2 MMAs can span both SMs
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(tc::NUM_THREADS);
        cfg.dynamicSmemBytes = C::SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (pdl_mode() == 1 || pdl_mode() == 2) ? 2 : 1;
        XGGM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, maps, p));
    }
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// CTA-pair engine on/off (XGGM_TC_PAIR=0 keeps every product on the single-CTA kernel; for A/B runs)
static bool pair_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("XGGM_TC_PAIR");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

template <int BN, int NPASS>
static int dispatch_major(bool a_mn, bool b_mn, const tc::GroupMaps& maps, const tc::Params& p, int grid,
                          cudaStream_t st) {
    if (!a_mn && !b_mn) return launch_tc<BN, NPASS, false, false>(maps, p, grid, st);
    if (!a_mn && b_mn) return launch_tc<BN, NPASS, false, true>(maps, p, grid, st);
    if (a_mn && b_mn) return launch_tc<BN, NPASS, true, true>(maps, p, grid, st);
    return XGGM_ERR_UNSUPPORTED;  // (MN-major A with K-major B is not needed by a Linear layer)
}

// Linear layers whose OUTPUT width N is ragged (encoder_adj: 630, the answer head: 2274) still run on the tensor
// cores: the planes that have N as their row length (the gradient g [M,N] and W^T [K,N]) are stored with the pitch
// padded to a multiple of 8 and zero columns, which the contraction simply runs over.
bool gemm_tc_ragged_ok(int M, int N, int K) { return M > 0 && N > 0 && K > 0 && (K % 8 == 0); }
bool gemm_tc_supported(int M, int N, int K) {
    // TMA needs 16-byte row pitches in every plane a Linear's three products touch
    return M > 0 && N > 0 && K > 0 && (N % 8 == 0) && (K % 8 == 0);
}

static void single_problem(tc::Params& p, const float* bias, const float* resid, float* C, __nv_bfloat16* c_hi,
                           __nv_bfloat16* c_lo, int accumulate) {
    p.group = 1;
    p.kcat = 0;
    p.tiles_per_prob = p.tiles_m * p.tiles_n * p.splits;
    for (int g = 0; g < tc::MAX_GROUP; ++g) p.pr[g] = tc::ProbOut{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr};
    p.pr[0] = tc::ProbOut{bias, resid, C, c_hi, c_lo, accumulate, 0, nullptr, nullptr};
}

// `count` (1..3) products of the same shape in ONE launch.
// A planes: a_mn ? [K,M] : [M,K];  B planes: b_mn ? [K,N] : [N,K].  lo planes may be null when npass == 1.
// kcat: the two entries of `pr` are the two K-segments of ONE product, C = A0 B0^T + A1 B1^T (+ addends of pr[0]).
int gemm_tc_group(bool a_mn, bool b_mn, const GemmProb* pr, int count, int M, int N, int K, int allow_split_k,
                  int npass, cudaStream_t st, bool kcat) {
    if (M <= 0 || N <= 0 || K <= 0 || count <= 0) return XGGM_OK;
    XGGM_REQUIRE(pr && count <= tc::MAX_GROUP && (npass == 1 || npass == 3));
    XGGM_REQUIRE(!kcat || (count == 2 && !a_mn && !b_mn && !allow_split_k));
    const int segments = count;          // operand pairs to build tensor maps for
    if (kcat) count = 1;                 // ... but one output problem
    for (int g = 0; g < segments; ++g)
        XGGM_REQUIRE(pr[g].a_hi && pr[g].b_hi && (pr[g].C || pr[g].c_hi || g >= count) && (npass == 1 || (pr[g].a_lo && pr[g].b_lo)));
    const int sms = num_sms();
    const int seg_kb = ceil_div(K, tc::BK);
    const int num_kb = kcat ? 2 * seg_kb : seg_kb;
    // Few output tiles (a head on [B,768] rows: M = 256 -> 4 pair tiles) cannot fill the machine: those products
    // take the single-CTA kernel with narrow 128 x 64 tiles (3x the CTAs, whole K per tile -> still deterministic).
    const long long pair_tiles = (long long)ceil_div(M, 2 * tc::BM) * ceil_div(N, 192) * count;
    const bool starved = !a_mn && !b_mn && !allow_split_k && pair_tiles * 8 <= sms;
    const bool pair = !a_mn && !b_mn && M > tc::BM && pair_enabled() && sms % 2 == 0 && !starved;
    // K-major x K-major (forward, and dgrad against transposed weight planes): CTA-pair kernel,
    // 256 x 192 tiles, cta_group::2 MMAs, each CTA stages half of the B tile.
    constexpr int PBN = 192;
    int bn = PBN, splits = 1, tiles_m = ceil_div(M, 2 * tc::BM);
    if (!pair) {
        tiles_m = ceil_div(M, tc::BM);
        // tile width: fewest "waves x width"; ties go to the wider tile (less A re-read)
        long long best = -1;
        const int cand[2] = {192, 128};
        int splits_for[2] = {1, 1};
        for (int ci = 0; ci < 2; ++ci) {
            const int w = cand[ci];
            const int tn = ceil_div(N, w);
            int sp = 1;
            if (allow_split_k) sp = max(1, min(num_kb / 4, sms / max(1, tiles_m * tn * count)));
            splits_for[ci] = sp;
            const long long kb_per = ceil_div(num_kb, sp);
            const long long waves = ceil_div((long long)tiles_m * tn * sp * count, sms);
            const long long cost = waves * w * kb_per;
            if (best < 0 || cost < best) { best = cost; bn = w; }
        }
        splits = splits_for[bn == 192 ? 0 : 1];
        if (starved) { bn = 64; splits = 1; }
    }
    const int tiles_n = ceil_div(N, bn);
    const int kb_per_split = ceil_div(num_kb, splits);
    splits = ceil_div(num_kb, kb_per_split);

    const long long a_rows = a_mn ? K : M, a_cols = a_mn ? M : K;
    const long long b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
    const int a_box = a_mn ? tc::BK : tc::BM, b_box = b_mn ? tc::BK : (pair ? PBN / 2 : bn);
    tc::GroupMaps maps;
    tc::Params p;
    p.M = M; p.N = N; p.num_kb = num_kb;
    p.tiles_m = tiles_m; p.tiles_n = tiles_n; p.splits = splits; p.kb_per_split = kb_per_split;
    p.group = count; p.tiles_per_prob = tiles_m * tiles_n * splits;
    p.kcat = kcat ? seg_kb : 0;
    p.ldc = N;
    p.atomic = splits > 1 ? 1 : 0;
    p.vec4 = (N % 4 == 0);
    p.gram_n = p.gram_g = p.gram_b = 0;
    p.bd_stride = 0;
    p.bd_kn = 0;
    p.dbg = g_tc_dbg;
    for (int g = 0; g < tc::MAX_GROUP; ++g) {
        const GemmProb& q = pr[g < segments ? g : 0];
        XGGM_TRY(make_map(&maps.m[g].a_hi, q.a_hi, a_rows, a_cols, a_box, q.lda));
        XGGM_TRY(make_map(&maps.m[g].b_hi, q.b_hi, b_rows, b_cols, b_box, q.ldb));
        if (npass == 3) {
            XGGM_TRY(make_map(&maps.m[g].a_lo, q.a_lo, a_rows, a_cols, a_box, q.lda));
            XGGM_TRY(make_map(&maps.m[g].b_lo, q.b_lo, b_rows, b_cols, b_box, q.ldb));
        } else {
            maps.m[g].a_lo = maps.m[g].a_hi;
            maps.m[g].b_lo = maps.m[g].b_hi;
        }
        p.pr[g] = tc::ProbOut{q.bias, q.resid, q.C, q.c_hi, npass == 3 ? q.c_lo : nullptr, q.accumulate, 0, nullptr, nullptr};
        if (g >= count) continue;
        if (((reinterpret_cast<uintptr_t>(q.C) | reinterpret_cast<uintptr_t>(q.resid) |
              reinterpret_cast<uintptr_t>(q.bias)) & 15) != 0)
            p.vec4 = 0;
    }
    for (int g = 0; g < count; ++g) {
        const GemmProb& q = pr[g];
        if (q.c_hi && (!p.vec4 || p.atomic || (reinterpret_cast<uintptr_t>(q.c_hi) & 7) ||
                       (reinterpret_cast<uintptr_t>(q.c_lo) & 7)))
            return XGGM_ERR_ARG;  // plane emission needs the float4 epilogue
        if (splits > 1 && !q.accumulate && q.C)
            XGGM_CUDA_TRY(cudaMemsetAsync(q.C, 0, sizeof(float) * (size_t)M * N, st));
        if (splits > 1 && !q.C) return XGGM_ERR_ARG;   // planes-only output needs whole-K tiles
    }
    const int total = p.tiles_per_prob * count;
    void* prof = gemm_prof_begin(2.0 * M * N * K * segments, st, segments);
    int rc;
    if (pair) {
        const int grid = 2 * min(sms / 2, total);
        rc = npass == 3 ? launch_tc<PBN, 3, false, false, 2>(maps, p, grid, st)
                        : launch_tc<PBN, 1, false, false, 2>(maps, p, grid, st);
    } else {
        const int grid = min(sms, total);
        if (bn == 64) {
            rc = npass == 3 ? launch_tc<64, 3, false, false>(maps, p, grid, st)
                            : launch_tc<64, 1, false, false>(maps, p, grid, st);
        } else if (bn == 192) {
            rc = npass == 3 ? dispatch_major<192, 3>(a_mn, b_mn, maps, p, grid, st)
                            : dispatch_major<192, 1>(a_mn, b_mn, maps, p, grid, st);
        } else {
            rc = npass == 3 ? dispatch_major<128, 3>(a_mn, b_mn, maps, p, grid, st)
                            : dispatch_major<128, 1>(a_mn, b_mn, maps, p, grid, st);
        }
    }
    gemm_prof_end(prof, st);
    return rc;
}

int gemm_tc(bool a_mn, bool b_mn, const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, const __nv_bfloat16* b_hi,
            const __nv_bfloat16* b_lo, const float* bias, const float* resid, float* C, __nv_bfloat16* c_hi,
            __nv_bfloat16* c_lo, int M, int N, int K, int accumulate, int allow_split_k, int npass, cudaStream_t st,
            int lda, int ldb) {
    const GemmProb q{a_hi, a_lo, b_hi, b_lo, bias, resid, C, c_hi, c_lo, accumulate, lda, ldb};
    return gemm_tc_group(a_mn, b_mn, &q, 1, M, N, K, allow_split_k, npass, st, false);
}

bool gram_tc_supported(int N, int H) { return N >= 1 && N <= tc::BM && H > 0 && H % 8 == 0; }

// S[b] = P[b] Q[b]^T for every graph b: P, Q are [B*N, H] bf16 planes (K-major), S is [B,N,N] fp32 (overwritten).
// One 128 x 128 tensor-core tile covers floor(128/N) whole graphs; only its diagonal blocks are stored.
int gram_tc(const __nv_bfloat16* p_hi, const __nv_bfloat16* p_lo, const __nv_bfloat16* q_hi, const __nv_bfloat16* q_lo,
            float* S, int B, int N, int H, int npass, cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(p_hi && q_hi && S && gram_tc_supported(N, H) && (npass == 1 || (npass == 3 && p_lo && q_lo)));
    const long long M = (long long)B * N;
    const int G = tc::BM / N;
    tc::GroupMaps maps;
    CUtensorMap &ah = maps.m[0].a_hi, &al = maps.m[0].a_lo, &bh = maps.m[0].b_hi, &bl = maps.m[0].b_lo;
    XGGM_TRY(make_map(&ah, p_hi, M, H, tc::BM));
    XGGM_TRY(make_map(&bh, q_hi, M, H, 128));
    if (npass == 3) {
        XGGM_TRY(make_map(&al, p_lo, M, H, tc::BM));
        XGGM_TRY(make_map(&bl, q_lo, M, H, 128));
    } else {
        al = ah;
        bl = bh;
    }
    for (int g = 1; g < tc::MAX_GROUP; ++g) maps.m[g] = maps.m[0];
    tc::Params p;
    p.M = (int)M; p.N = (int)M; p.num_kb = ceil_div(H, tc::BK);
    p.tiles_m = ceil_div(B, G); p.tiles_n = 1;
    // fewer tiles than half the SMs (B = 256 graphs of 36 nodes: 86 tiles on 148 SMs): the contraction is cut in
    // two and the halves meet in S with atomic adds (two addends from zero: the sum does not depend on their order)
    p.splits = (2 * p.tiles_m <= num_sms() && p.num_kb >= 4) ? 2 : 1;
    p.kb_per_split = ceil_div(p.num_kb, p.splits);
    p.splits = ceil_div(p.num_kb, p.kb_per_split);
    p.ldc = N;
    p.atomic = p.splits > 1 ? 1 : 0; p.vec4 = 0;
    if (p.atomic) XGGM_CUDA_TRY(cudaMemsetAsync(S, 0, sizeof(float) * (size_t)B * N * N, st));
    single_problem(p, nullptr, nullptr, S, nullptr, nullptr, 0);
    p.gram_n = N; p.gram_g = G; p.gram_b = B;
    p.bd_stride = 0;
    p.bd_kn = 0;
    p.dbg = nullptr;
    const int grid = min(num_sms(), p.tiles_m * p.splits);
    void* prof = gemm_prof_begin(2.0 * B * N * N * H, st);
    const int rc = npass == 3 ? launch_tc<128, 3, false, false>(maps, p, grid, st)
                              : launch_tc<128, 1, false, false>(maps, p, grid, st);
    gemm_prof_end(prof, st);
    return rc;
}

bool adj_tc_supported(int N, int H) { return N >= 1 && N <= tc::BM && H > 0 && H % 8 == 0; }
// Widest k range a 128-row output tile can touch: the rows of every graph it overlaps.
static inline int bd_span(int N) { return N * ((N - 1 + tc::BM - 1) / N + 1); }
constexpr int BD_KW = 192;   // coefficient-tile width of the unaligned scheme (3 k-blocks)
// Unaligned tiles (all 128 rows of every tile live, ceil(M/128) row tiles) whenever the span fits 192 k rows
// (N <= 48 and N = 64: obj36 graphs touch at most 5 graphs = 180 rows); XGGM_ADJ_ALIGNED=1 keeps the graph-aligned
// tiles (floor(128/N) whole graphs per tile) for A/B runs, and wider graphs always take them.
static bool bd_unaligned(int N) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("XGGM_ADJ_ALIGNED");
        forced = (e && e[0] == '1') ? 1 : 0;
    }
    return !forced && bd_span(N) <= BD_KW;
}
long long adj_tc_coef_elems(int B, int N) {
    const long long aligned = (long long)ceil_div(B, tc::BM / N) * tc::BM * tc::BM;
    const long long unaligned = (long long)ceil_div((long long)B * N, tc::BM) * tc::BM * BD_KW;
    return aligned > unaligned ? aligned : unaligned;
}

// coefficient planes (adj_tc_coef_elems bf16 each) for adj_apply_tc
int build_blockdiag(const float* adj, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int N, float alpha0,
                    const float* alpha_dev, float self_w, int trans, cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(adj && hi && N >= 1 && N <= tc::BM);
    if (bd_unaligned(N)) {
        const long long M = (long long)B * N;
        XGGM_LAUNCH((tc::build_blockdiag_u_kernel), dim3(ceil_div(M, tc::BM), 8), 256, 0, st, adj, hi, lo, M, N, BD_KW, alpha0,
                    alpha_dev, self_w, trans);
        XGGM_LAUNCH_CHECK();
        return XGGM_OK;
    }
    const int G = tc::BM / N;
    XGGM_LAUNCH((tc::build_blockdiag_kernel), dim3(ceil_div(B, G), 4), 128, 0, st, adj, hi, lo, B, N, G, alpha0, alpha_dev, self_w, trans);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// out[b] (=|+=) C[b] @ x[b] for every graph: coefficient planes from build_blockdiag, x as [B*N, H] bf16 planes.
// out (fp32) and / or out planes; accumulate adds to the previous fp32 out; resid (optional, [B*N,H] fp32) is added.
int adj_apply_tc(const __nv_bfloat16* c_hi_in, const __nv_bfloat16* c_lo_in, const __nv_bfloat16* x_hi,
                 const __nv_bfloat16* x_lo, float* out, __nv_bfloat16* o_hi, __nv_bfloat16* o_lo, int B, int N, int H,
                 int accumulate, int npass, cudaStream_t st, const float* resid, const __nv_bfloat16* r_hi,
                 const __nv_bfloat16* r_lo) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(c_hi_in && x_hi && (out || o_hi) && adj_tc_supported(N, H) && (npass == 1 || (c_lo_in && x_lo)));
    XGGM_REQUIRE(!(resid && r_hi) && (reinterpret_cast<uintptr_t>(r_hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(r_lo) & 7) == 0);
    XGGM_REQUIRE(!accumulate || out);
    const long long M = (long long)B * N;
    const bool unal = bd_unaligned(N);
    const int G = tc::BM / N, T = unal ? ceil_div(M, tc::BM) : ceil_div(B, G);
    const int KW = unal ? BD_KW : tc::BM;
    constexpr int ABN = 192;
    tc::GroupMaps maps;
    CUtensorMap &ah = maps.m[0].a_hi, &al = maps.m[0].a_lo, &bh = maps.m[0].b_hi, &bl = maps.m[0].b_lo;
    XGGM_TRY(make_map(&ah, c_hi_in, (long long)T * tc::BM, KW, tc::BM));
    XGGM_TRY(make_map(&bh, x_hi, M, H, tc::BK));
    if (npass == 3) {
        XGGM_TRY(make_map(&al, c_lo_in, (long long)T * tc::BM, KW, tc::BM));
        XGGM_TRY(make_map(&bl, x_lo, M, H, tc::BK));
    } else {
        al = ah;
        bl = bh;
    }
    tc::Params p;
    p.M = (int)M; p.N = H; p.num_kb = KW / tc::BK;
    p.tiles_m = T; p.tiles_n = ceil_div(H, ABN); p.splits = 1; p.kb_per_split = p.num_kb;
    p.ldc = H;
    p.atomic = 0;
    p.vec4 = (H % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    single_problem(p, nullptr, resid, out, o_hi, npass == 3 ? o_lo : nullptr, accumulate);
    p.pr[0].r_hi = r_hi;
    p.pr[0].r_lo = npass == 3 ? r_lo : nullptr;
    if (resid && (reinterpret_cast<uintptr_t>(resid) & 15)) return XGGM_ERR_ARG;
    for (int g = 1; g < tc::MAX_GROUP; ++g) maps.m[g] = maps.m[0];
    p.gram_n = p.gram_g = p.gram_b = 0;
    p.bd_stride = unal ? tc::BM : G * N;
    p.bd_kn = unal ? N : 0;
    p.dbg = g_tc_dbg;
    XGGM_REQUIRE(p.vec4 && (reinterpret_cast<uintptr_t>(o_hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(o_lo) & 7) == 0);
    const int grid = min(num_sms(), p.tiles_m * p.tiles_n);
    // (not a projection: kept out of the GEMM roofline accounting)
    return npass == 3 ? launch_tc<ABN, 3, false, true>(maps, p, grid, st)
                      : launch_tc<ABN, 1, false, true>(maps, p, grid, st);
}

// h_next = LN(h + C[b] @ P[b]) fused (adj_ln_tc_kernel): coefficient planes from build_blockdiag, P and the residual h as
// [B*N, H] operand planes.  H must be 4 x 192 (X-GGM: 768).  xhat [M,H] (fp32, or bf16 when xhat_bf16), rstd [M],
// planes of h_next (o_hi / o_lo) and optionally fp32 h_next.
// OPT-IN (XGGM_ADJ_LN_TC=1).  Parity-green on every fixture and on the B=256 oracle tests, but measured SLOWER than the
// two kernels it replaces at the BASELINE batch: 90 us per launch against 25 + 13 us (step 2.04 vs 1.87 ms, same box;
// profiles/r02n_launches.txt).  One tile per CTA serialises prologue -> TMA -> MMA -> two-pass epilogue (~25-30 us), the
// 72 row blocks of B=256 are 72 clusters of four, and a B200 holds 34 such clusters at a time (GPCs of 18-20 SMs): three
// rounds.  A persistent, double-buffered variant would hide the epilogue behind the next tile at B >= 1024, not at 256.
bool adj_ln_tc_supported(int N, int H) {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("XGGM_ADJ_LN_TC");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on && adj_tc_supported(N, H) && H == tc::ALN_CL * tc::ALN_BN;
}
int adj_ln_tc(const __nv_bfloat16* c_hi_in, const __nv_bfloat16* c_lo_in, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
              const __nv_bfloat16* r_hi, const __nv_bfloat16* r_lo, const float* gamma, const float* beta, float* xhat,
              int xhat_bf16, float* rstd, float* h_out, __nv_bfloat16* o_hi, __nv_bfloat16* o_lo, int B, int N, int H, float eps,
              int npass, cudaStream_t st) {
    if (B <= 0) return XGGM_OK;
    XGGM_REQUIRE(c_hi_in && x_hi && r_hi && gamma && beta && xhat && rstd && adj_ln_tc_supported(N, H) &&
                 (npass == 1 || (c_lo_in && x_lo && r_lo)));
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    XGGM_REQUIRE(al16(gamma) && al16(beta) && al16(xhat) && al16(h_out) && (reinterpret_cast<uintptr_t>(r_hi) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(r_lo) & 7) == 0 && (reinterpret_cast<uintptr_t>(o_hi) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(o_lo) & 7) == 0);
    const long long M = (long long)B * N;
    const bool unal = bd_unaligned(N);
    const int G = tc::BM / N, T = unal ? ceil_div(M, tc::BM) : ceil_div(B, G);
    const int KW = unal ? BD_KW : tc::BM;
    tc::GroupMaps maps;
    CUtensorMap &ah = maps.m[0].a_hi, &al = maps.m[0].a_lo, &bh = maps.m[0].b_hi, &bl = maps.m[0].b_lo;
    XGGM_TRY(make_map(&ah, c_hi_in, (long long)T * tc::BM, KW, tc::BM));
    XGGM_TRY(make_map(&bh, x_hi, M, H, tc::BK));
    if (npass == 3) {
        XGGM_TRY(make_map(&al, c_lo_in, (long long)T * tc::BM, KW, tc::BM));
        XGGM_TRY(make_map(&bl, x_lo, M, H, tc::BK));
    } else {
        al = ah;
        bl = bh;
    }
    for (int g = 1; g < tc::MAX_GROUP; ++g) maps.m[g] = maps.m[0];
    tc::AdjLnParams p;
    p.M = (int)M; p.H = H; p.num_kb = KW / tc::BK;
    p.bd_stride = unal ? tc::BM : G * N;
    p.bd_kn = unal ? N : 0;
    p.r_hi = r_hi; p.r_lo = npass == 3 ? r_lo : nullptr;
    p.gamma = gamma; p.beta = beta;
    p.xhat = xhat; p.xhat_bf16 = xhat_bf16; p.rstd = rstd; p.h_out = h_out;
    p.o_hi = o_hi; p.o_lo = npass == 3 ? o_lo : nullptr;
    p.eps = eps;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(T * tc::ALN_CL);
    cfg.blockDim = dim3(tc::NUM_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = tc::ALN_CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (npass == 3) {
        static bool set3 = false;
        if (!set3) { XGGM_CUDA_TRY(cudaFuncSetAttribute(tc::adj_ln_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::AlnCfg<3>::SMEM)); set3 = true; }
        cfg.dynamicSmemBytes = tc::AlnCfg<3>::SMEM;
        XGGM_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::adj_ln_tc_kernel<3>, maps, p));
    } else {
        static bool set1 = false;
        if (!set1) { XGGM_CUDA_TRY(cudaFuncSetAttribute(tc::adj_ln_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::AlnCfg<1>::SMEM)); set1 = true; }
        cfg.dynamicSmemBytes = tc::AlnCfg<1>::SMEM;
        XGGM_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::adj_ln_tc_kernel<1>, maps, p));
    }
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// Transposed split of `count` [R,C] fp32 matrices into [C,R] bf16 planes (one launch per 8 matrices).
int split_planes_t(const float* const* src, __nv_bfloat16* const* hi, __nv_bfloat16* const* lo, int R, int C,
                   int count, cudaStream_t st, int pitch) {
    const int P = pitch > 0 ? pitch : R;
    int done = 0;
    while (done < count) {
        tc::SplitTJobs jobs;
        const int n = min(tc::MAX_SPLIT_JOBS, count - done);
        for (int i = 0; i < n; ++i) {
            jobs.j[i].src = src[done + i];
            jobs.j[i].hi = hi[done + i];
            jobs.j[i].lo = lo ? lo[done + i] : nullptr;
        }
        if (R > 0 && C > 0) {
            XGGM_LAUNCH((tc::split_planes_t_kernel), dim3(ceil_div(C, 32), ceil_div(P, 32), n), 256, 0, st, jobs, R, C, P);
            XGGM_LAUNCH_CHECK();
        }
        done += n;
    }
    return XGGM_OK;
}

// src [R,C] -> planes [R,P] (P >= C, P % 8 == 0), zero padded
int split_planes_pitched(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long R, int C, int P, cudaStream_t st) {
    if (R <= 0 || C <= 0) return XGGM_OK;
    XGGM_REQUIRE(src && hi && P >= C);
    const int grid = (int)max(1LL, min((long long)num_sms() * 8, (R * P + 255) / 256));
    XGGM_LAUNCH((tc::split_planes_pitched_kernel), grid, 256, 0, st, src, hi, lo, R, C, P);
    XGGM_LAUNCH_CHECK();
    return XGGM_OK;
}

// Split up to MAX_SPLIT_JOBS fp32 arrays into bf16 hi (+ lo) planes with one launch.
int split_planes(const float* const* src, __nv_bfloat16* const* hi, __nv_bfloat16* const* lo, const long long* n,
                 int count, cudaStream_t st) {
    int done = 0;
    while (done < count) {
        tc::SplitJobs jobs;
        jobs.count = min(tc::MAX_SPLIT_JOBS, count - done);
        long long nmax = 0;
        for (int i = 0; i < jobs.count; ++i) {
            jobs.j[i].src = src[done + i];
            jobs.j[i].hi = hi[done + i];
            jobs.j[i].lo = lo ? lo[done + i] : nullptr;
            jobs.j[i].n = n[done + i];
            nmax = n[done + i] > nmax ? n[done + i] : nmax;
        }
        if (nmax > 0) {
            const int gx = (int)max(1LL, min((long long)num_sms() * 8, (nmax / 4 + 255) / 256));
            XGGM_LAUNCH((tc::split_planes_kernel), dim3(gx, jobs.count), 256, 0, st, jobs);
            XGGM_LAUNCH_CHECK();
        }
        done += jobs.count;
    }
    return XGGM_OK;
}

// Layout of one prepared weight buffer (bf16 elements): hi[pad8(N*K)] | lo[pad8(N*K)] | thi[pad8(K*P)] | tlo[pad8(K*P)]
static inline long long wp_pad8(long long n) { return (n + 7) & ~7LL; }
long long weight_planes_elems(int N, int K) {
    const long long P = (N + 7) & ~7;
    return 2 * wp_pad8((long long)N * K) + 2 * wp_pad8((long long)K * P);
}
void weight_planes_views(void* buf, int N, int K, __nv_bfloat16** hi, __nv_bfloat16** lo, __nv_bfloat16** thi,
                         __nv_bfloat16** tlo) {
    const long long P = (N + 7) & ~7;
    __nv_bfloat16* b = static_cast<__nv_bfloat16*>(buf);
    *hi = b;
    *lo = b + wp_pad8((long long)N * K);
    *thi = b + 2 * wp_pad8((long long)N * K);
    *tlo = *thi + wp_pad8((long long)K * P);
}
int weight_planes_build(const float* const* W, void* const* bufs, const int* N, const int* K, int count, bool with_lo,
                        cudaStream_t st) {
    int done = 0;
    while (done < count) {
        tc::WPlaneJobs jobs;
        const int n = min(tc::MAX_WP_JOBS, count - done);
        int kmax = 0, pmax = 0;
        for (int i = 0; i < n; ++i) {
            tc::WPlaneJob& j = jobs.j[i];
            XGGM_REQUIRE(W[done + i] && bufs[done + i] && N[done + i] > 0 && K[done + i] > 0);
            j.src = W[done + i];
            j.N = N[done + i]; j.K = K[done + i]; j.P = (j.N + 7) & ~7;
            weight_planes_views(bufs[done + i], j.N, j.K, &j.hi, &j.lo, &j.thi, &j.tlo);
            if (!with_lo) { j.lo = nullptr; j.tlo = nullptr; }
            kmax = max(kmax, j.K); pmax = max(pmax, j.P);
        }
        XGGM_LAUNCH((tc::weight_planes_kernel), dim3(ceil_div(kmax, 64), ceil_div(pmax, 64), n), 256, 0, st, jobs);
        XGGM_LAUNCH_CHECK();
        done += n;
    }
    return XGGM_OK;
}

}  // namespace xggm

