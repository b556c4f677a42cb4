/* xggm_b200 -- C ABI of the B200-native X-GGM graph block (sm_100a).
 *
 * The reference (jingjing12110/X-GGM) is pure PyTorch and has no FFI; this header
 * is the new native surface its nn.Module boundary binds to (INTEGRATION.md shows
 * the ctypes stub).  Every entry point cites the reference lines it replaces
 * (paths relative to the reference root; "ggm.py" =
 * src/module/graph_generative_modeling.py).
 *
 * Conventions
 *  - All tensors are dense, row-major, contiguous fp32 DEVICE buffers owned by the
 *    caller (PyTorch allocates them); the library never allocates, frees or keeps
 *    a pointer after the call returns.  "?" marks a pointer that may be NULL.
 *  - B graphs, N nodes per graph (36 for obj36), H feature width (768), M = B*N rows.
 *  - Every call enqueues on the caller's stream and returns without synchronising.
 *  - Return value: XGGM_OK or a negative XGGM_ERR_* code.  Nothing throws.
 *  - No per-call state is kept, but three PROCESS-WIDE settings exist and are not synchronised: the projection
 *    engine (xggm_set_precision), the bench instrumentation (xggm_prof_*, xggm_debug_timeline) and the launch
 *    counter (atomic).  Set them before worker threads start issuing calls; entry points themselves may be
 *    called from several host threads (autograd does) as long as each call's buffers are its own.
 *    One CUDA context per process (one process per GPU).
 *  - No CPU fallback: xggm_device_check() fails on anything but compute
 *    capability 10.x and the kernels exist only as sm_100a SASS.
 */
#ifndef XGGM_B200_H
#define XGGM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XGGM_ABI_VERSION 4

#define XGGM_OK 0
#define XGGM_ERR_ARG (-1)         /* bad shape / NULL pointer / unsupported size */
#define XGGM_ERR_CUDA (-2)        /* CUDA runtime error; see xggm_last_cuda_error() */
#define XGGM_ERR_ARCH (-3)        /* device is not compute capability 10.x */
#define XGGM_ERR_UNSUPPORTED (-4) /* valid request this build does not implement */

typedef void* xggm_stream_t; /* a cudaStream_t */

int xggm_abi_version(void);
const char* xggm_strerror(int code);
const char* xggm_last_cuda_error(void);
/* Number of CUDA kernels this library has launched in this process (bench accounting). */
unsigned long long xggm_launch_count(void);
/* Bench instrumentation: while enabled, every projection-GEMM launch is bracketed by a
 * CUDA-event pair on its stream; xggm_prof_read sums the durations (ms), count and algorithmic
 * FLOPs (2*M*N*K) of the launches of the dominant product shape (FLOPs >= half of the largest
 * recorded launch).  Enabling (or disabling) clears the records. */
int xggm_prof_enable(int on);
int xggm_prof_read(double* total_ms, long long* launches, double* flops);
/* Kernel-tuning aid: while dev_buf (>= 32 uint64 of DEVICE memory) is set, CTA 0 of every tcgen05
 * GEMM launch records %globaltimer at its pipeline milestones (tools/gemm_timeline.py decodes them).
 * NULL switches it off. */
int xggm_debug_timeline(unsigned long long* dev_buf);
/* The library links its own (static) CUDA runtime: make `device` current for the calling
 * thread before enqueueing work on one of its streams (cheap; call it per entry). */
int xggm_set_device(int device);
/* XGGM_OK iff `device` is a compute-capability-10.x GPU (B200). */
int xggm_device_check(int device);

/* ------------------------------------------------------------------------- *
 * Projection engine (process-wide setting; default XGGM_PREC_FP32)
 *   XGGM_PREC_FP32      tcgen05 tensor cores, every fp32 operand split into bf16 hi+lo planes and
 *                       multiplied in three passes (lo*hi + hi*lo + hi*hi, fp32 accumulation in
 *                       TMEM): ~1e-5 relative error, inside the 1e-4 fp32 parity budget.
 *   XGGM_PREC_BF16      tcgen05, operands rounded to bf16 once (single pass): the 2e-2 bf16 budget.
 *   XGGM_PREC_FP32_SIMT exact fp32 FMA kernel (no tensor cores); also taken for shapes whose
 *                       row pitch TMA cannot address (N or K not a multiple of 8).
 * ------------------------------------------------------------------------- */
#define XGGM_PREC_FP32 0
#define XGGM_PREC_BF16 1
#define XGGM_PREC_FP32_SIMT 2
int xggm_set_precision(int mode);
int xggm_get_precision(void);

/* ------------------------------------------------------------------------- *
 * Dense node projections  (nn.Linear: src/module/gcn.py:13,44; gin.py:14)
 * `work` is caller-owned scratch of xggm_linear_work_bytes(M,N,K) bytes (16-byte aligned) that
 * receives the bf16 operand planes of the tcgen05 engine; NULL selects the exact SIMT kernel.
 * ------------------------------------------------------------------------- */
long long xggm_linear_work_bytes(int M, int N, int K);
/* out[M,N] = a[M,K] w[N,K]^T + bias[N]? + resid[M,N]? */
int xggm_linear_fwd(const float* a, const float* w, const float* bias, const float* resid,
                    float* out, int M, int N, int K, void* work, xggm_stream_t s);
/* ga[M,K] (+)= g[M,N] w[N,K]   (accumulate != 0 adds into ga) */
int xggm_linear_bwd_input(const float* g, const float* w, float* ga, int M, int N, int K,
                          int accumulate, void* work, xggm_stream_t s);
/* Prepared weight planes.  A weight changes once per optimiser step but is read by several products per step
 * (forward as W, input gradient as W^T, in every layer call): xggm_weight_planes_build turns `count` fp32
 * matrices W_i [N_i,K_i] into both bf16 operand layouts in ONE launch (per 16 matrices) -- buffer i is
 * xggm_weight_planes_bytes(N_i,K_i) bytes, 16-byte aligned, layout hi[N,K] | lo[N,K] | hi(W^T)[K,P] | lo(W^T)[K,P],
 * P = N rounded up to 8 -- and the `_ex` entry points take it instead of splitting W themselves (w_planes /
 * weight_planes NULL = split internally, as the plain entry points do).  The caller rebuilds the planes whenever
 * the fp32 weights change (xggm_b200.optim.BertAdam does, right after its update kernel) and when the engine
 * precision changes; the exact-fp32 engine ignores them. */
long long xggm_weight_planes_bytes(int N, int K);
int xggm_weight_planes_build(const float* const* weights, void* const* planes, const int* N, const int* K, int count,
                             xggm_stream_t s);
int xggm_linear_fwd_ex(const float* a, const float* w, const float* bias, const float* resid,
                       float* out, int M, int N, int K, void* work, const void* w_planes, xggm_stream_t s);
int xggm_linear_bwd_input_ex(const float* g, const float* w, float* ga, int M, int N, int K,
                             int accumulate, void* work, const void* w_planes, xggm_stream_t s);
/* gw[N,K] (+)= g[M,N]^T a[M,K];  gbias[N]? (+)= column sums of g.  accumulate != 0 adds into the
 * buffers (gradient accumulation straight into a parameter's .grad), otherwise they are overwritten. */
int xggm_linear_bwd_weight(const float* g, const float* a, float* gw, float* gbias,
                           int M, int N, int K, int accumulate, void* work, xggm_stream_t s);

/* ------------------------------------------------------------------------- *
 * Adjacency-weighted message passing
 *   GCNConv  torch.bmm(adj, x)                    src/module/gcn.py:28
 *   GINConv  X + (1 + eps) * A @ X                src/module/gin.py:32
 * out[b] = self_w * x[b] + alpha * adj[b] @ x[b],  alpha = alpha0 + (*alpha_dev if non-NULL)
 * `work` (xggm_adj_apply_work_bytes bytes, 16-byte aligned; N <= 128, H % 8 == 0) runs the product on the
 * tensor cores (block-diagonal coefficient tiles x bf16 planes of the node features); NULL selects the
 * SIMT kernel.
 * ------------------------------------------------------------------------- */
long long xggm_adj_apply_work_bytes(int B, int N, int H);
int xggm_adj_apply_fwd(const float* adj, const float* x, float* out, int B, int N, int H,
                       float alpha0, const float* alpha_dev, float self_w, void* work, xggm_stream_t s);
/* gx (+)= self_w*gout + alpha * adj^T @ gout ; gadj_raw? = gout @ x^T (NOT scaled by alpha;
 * the caller scales, and d alpha = <gadj_raw, adj>). */
int xggm_adj_apply_bwd(const float* adj, const float* x, const float* gout, float* gx,
                       float* gadj_raw, int B, int N, int H, float alpha0,
                       const float* alpha_dev, float self_w, int accumulate_gx, void* work, xggm_stream_t s);

/* ------------------------------------------------------------------------- *
 * Row normalisation / activations
 * ------------------------------------------------------------------------- */
/* nn.LayerNorm(H) of src/module/gcn.py:14,29: h = xhat*gamma+beta; saves xhat and rstd. */
int xggm_layernorm_fwd(const float* u, const float* gamma, const float* beta, float* h,
                       float* xhat, float* rstd, int M, int H, float eps, xggm_stream_t s);
/* gu = LN backward; ggamma/gbeta are ACCUMULATED into (caller zeroes them). */
int xggm_layernorm_bwd(const float* gh, const float* xhat, const float* rstd, const float* gamma,
                       float* gu, float* ggamma, float* gbeta, int M, int H, xggm_stream_t s);
/* Jump-knowledge head tail  F.dropout(LN(GeLU(z)), p)  src/module/gcn.py:44-47,72-76
 * (GeLU = exact erf, src/lxrt/modeling.py:116-124).
 * out (=|+=) keep*scale*LN(gelu(z)); keep? is a uint8 mask (1 = keep), scale = 1/(1-p).
 * Saves per-row mean and rstd of gelu(z). */
int xggm_gelu_ln_drop_fwd(const float* z, const float* gamma, const float* beta,
                          const uint8_t* keep, float scale, float* out, float* mean, float* rstd,
                          int M, int H, float eps, int accumulate, xggm_stream_t s);
/* gz = backward of the above w.r.t. z; ggamma/gbeta ACCUMULATED. */
int xggm_gelu_ln_drop_bwd(const float* gout, const float* z, const float* mean, const float* rstd,
                          const float* gamma, const uint8_t* keep, float scale, float* gz,
                          float* ggamma, float* gbeta, int M, int H, xggm_stream_t s);

/* ------------------------------------------------------------------------- *
 * Operand-plane hand-over (tensor-core engines).  The tcgen05 kernels read every fp32 tensor as
 * two bf16 "planes" (hi = bf16(v), lo = bf16(v - hi)).  A kernel that PRODUCES a tensor can emit
 * its planes in the same pass, and the `_ex` entry points below let the caller hand those planes
 * to the CONSUMERS of that tensor (`*_planes` arguments; NULL = build them internally, which is
 * what the plain entry points do), so the [B,N,H] node features are not re-read and re-split once
 * per consumer (next GNN layer, adjacency regeneration forward and backward).  A plane buffer is
 * xggm_planes_bytes(n_elems) bytes, 16-byte aligned, layout hi[pad8(n)] | lo[pad8(n)] bf16; it is
 * only meaningful for the precision it was written under and is ignored by the exact-fp32 engine.
 * ------------------------------------------------------------------------- */
long long xggm_planes_bytes(long long n_elems);

/* ------------------------------------------------------------------------- *
 * Adjacency regeneration   ggm.py:225-228 (GCN), :188-191 (GIN), :261-264 (GAT),
 * :124-126 (EdgeGenerator, squash = 0)
 *   S = x x^T ; m_i = max_k S[k,i] ; adj[i,j] = sigmoid(S[i,j]/m_i) (i != j), 0 on the diagonal
 * Saves S[B,N,N] and the arg-max row index amax[B,N] (first index on ties, as torch.max).
 * ------------------------------------------------------------------------- */
/* `work` (xggm_adj_regen_work_bytes bytes, 16-byte aligned) lets the pair scores run on the
 * tensor cores (bf16 planes of x + batched Gram kernel); NULL selects the per-graph SIMT kernel. */
long long xggm_adj_regen_work_bytes(int B, int N, int H);
int xggm_adj_regen_fwd(const float* x, float* adj_out, float* S, int32_t* amax, int B, int N,
                       int H, int squash, void* work, xggm_stream_t s);
int xggm_adj_regen_fwd_ex(const float* x, float* adj_out, float* S, int32_t* amax, int B, int N,
                          int H, int squash, void* work, const void* x_planes, xggm_stream_t s);
/* gx (+)= (dS + dS^T) x.  `work` is a [B,N,N] scratch buffer; `tc_work`? (xggm_adj_apply_work_bytes bytes)
 * moves the D x product onto the tensor cores. */
int xggm_adj_regen_bwd(const float* gadj, const float* x, const float* S, const int32_t* amax,
                       float* gx, float* work, int B, int N, int H, int squash,
                       int accumulate_gx, void* tc_work, xggm_stream_t s);
int xggm_adj_regen_bwd_ex(const float* gadj, const float* x, const float* S, const int32_t* amax,
                          float* gx, float* work, int B, int N, int H, int squash,
                          int accumulate_gx, void* tc_work, const void* x_planes, xggm_stream_t s);

/* ------------------------------------------------------------------------- *
 * Whole GCN / GIN layers (conv chain + jump-knowledge read-out)
 *   GCN.forward  src/module/gcn.py:64-77   (GCNConv :22-29)
 *   GIN.forward  src/module/gin.py:68-87   (GINConv :21-34)
 * Parameter tables are HOST arrays of DEVICE pointers:
 *   GCN conv k : conv_params[3k+0..2] = ctx_layer.weight[H,H], layer_norm.weight[H], layer_norm.bias[H]
 *   GIN conv k : conv_params[5k+0..4] = eps[1], linear.0.weight[H,H], linear.0.bias[H],
 *                                        linear.2.weight[H], linear.2.bias[H]
 *   head j     : head_params[4j+0..3] = 0.weight[H,H], 0.bias[H], 2.weight[H], 2.bias[H]
 *   keeps[j]?  : uint8 [M,H] keep-mask of head j (NULL table and NULL philox => no dropout)
 * `saved` (xggm_gnn_saved_floats floats) carries activations to the backward call;
 * `work` (xggm_gnn_work_floats floats) is scratch.  kind: 0 = GCN, 1 = GIN.
 * ------------------------------------------------------------------------- */
#define XGGM_KIND_GCN 0
#define XGGM_KIND_GIN 1
/* In-kernel dropout of the read-out heads (used when `keeps` is NULL, `philox` is not, drop_p > 0):
 * Philox4x32-10, key = seed, subsequence = stream0 + head index; counter = element index / 4 and one 32-bit
 * word per element (keep iff word >= p * 2^32), except for drop_p == 0.5 exactly, where counter = element index
 * / 128 and element e keeps iff bit (e % 128) of the 128-bit block is set (one random bit per element).
 * The subsequence gains (*dev_epoch << 32) when dev_epoch, a DEVICE counter, is given -- lets a captured CUDA graph
 * draw fresh masks on every replay).  Forward and backward must receive the same triple.
 * xggm_keep_mask(seed, stream0 + j, dev_epoch) materialises exactly the mask head j uses. */
typedef struct xggm_philox_s {
    uint64_t seed;
    uint64_t stream0;
    const uint64_t* dev_epoch;
} xggm_philox_t;
long long xggm_gnn_saved_floats(int kind, int B, int N, int H, int n_convs);
long long xggm_gnn_work_floats(int kind, int B, int N, int H, int n_convs);
int xggm_gnn_fwd(int kind, const float* x, const float* adj, const float* const* conv_params,
                 const float* const* head_params, const uint8_t* const* keeps, const xggm_philox_t* philox,
                 float drop_p, float* out, float* saved, float* work, int B, int N, int H, int n_convs,
                 xggm_stream_t s);
/* x_planes? : planes of x from its producer; out_planes? : receives the planes of `out` (emitted by the
 * last read-out accumulation).  The backward call must get the same x_planes.
 * weight_planes? : HOST table of 2*n_convs + 1 prepared weight-plane buffers (xggm_weight_planes_build; conv k ->
 * entry k, head j -> entry n_convs + j), or NULL. */
int xggm_gnn_fwd_ex(int kind, const float* x, const float* adj, const float* const* conv_params,
                    const float* const* head_params, const uint8_t* const* keeps, const xggm_philox_t* philox,
                    float drop_p, float* out, float* saved, float* work, const void* x_planes, void* out_planes,
                    int B, int N, int H, int n_convs, const void* const* weight_planes, xggm_stream_t s);
/* Gradient tables mirror the parameter tables (same order); every parameter-gradient buffer is
 * overwritten, or accumulated into when accumulate_param_grads != 0 (the buffers then are the
 * parameters' live .grad tensors).  gx[B,N,H] is always overwritten; gadj[B,N,N] is overwritten, or may be
 * NULL for a GCN layer whose adjacency needs no gradient (the gq h^T products are then skipped; a GIN layer
 * still needs them for d eps and must pass a buffer). */
int xggm_gnn_bwd(int kind, const float* gout, const float* x, const float* adj,
                 const float* const* conv_params, const float* const* head_params,
                 const uint8_t* const* keeps, const xggm_philox_t* philox, float drop_p, const float* saved,
                 float* work, float* gx, float* gadj, float* const* conv_grads, float* const* head_grads,
                 int accumulate_param_grads, int B, int N, int H, int n_convs, xggm_stream_t s);
int xggm_gnn_bwd_ex(int kind, const float* gout, const float* x, const float* adj,
                    const float* const* conv_params, const float* const* head_params,
                    const uint8_t* const* keeps, const xggm_philox_t* philox, float drop_p, const float* saved,
                    float* work, float* gx, float* gadj, float* const* conv_grads, float* const* head_grads,
                    int accumulate_param_grads, const void* x_planes, int B, int N, int H, int n_convs,
                    const void* const* weight_planes, xggm_stream_t s);

/* ------------------------------------------------------------------------- *
 * GAT attention  src/module/gat.py:25-49 (after h = linear_layer(x), which is
 * xggm_linear_fwd).  e_ij = LeakyReLU_alpha(a1.h_i + a2.h_j); masked_fill(adj==0,-9e15);
 * row softmax; out = elu(att @ h) (apply_elu = concat flag, gat.py:46-49).
 * a = attn_layer.weight[2H].  Saves att[B,N,N] and pre-activation pre[B,N,H] = att @ h.
 * ------------------------------------------------------------------------- */
int xggm_gat_attn_fwd(const float* h, const float* a, const float* adj, float* out, float* att,
                      float* pre, int B, int N, int H, float alpha, int apply_elu, xggm_stream_t s);
/* gh overwritten; ga[2H] ACCUMULATED (caller zeroes); work = B*N*H + B*N*N floats scratch. */
int xggm_gat_attn_bwd(const float* gout, const float* h, const float* a, const float* adj,
                      const float* att, const float* pre, float* gh, float* ga, float* work, int B,
                      int N, int H, float alpha, int apply_elu, xggm_stream_t s);

/* ------------------------------------------------------------------------- *
 * Trainer glue  (src/vqa/vqacpv2.py, src/gqa/gqa_ood.py, src/module/graph_utils.py)
 * ------------------------------------------------------------------------- */
/* a.triu(1)+a.tril(-1)                                   src/vqa/vqacpv2.py:188 */
int xggm_strip_diag(const float* a, float* out, int B, int N, xggm_stream_t s);
/* v[B,N(N-1)/2] -> symmetric adj[B,N,N], zero diagonal    src/vqa/vqacpv2.py:195-199 */
int xggm_triu_scatter_fwd(const float* v, float* adj, int B, int N, xggm_stream_t s);
/* gv[b,k] = gadj[b,i,j] + gadj[b,j,i] */
int xggm_triu_scatter_bwd(const float* gadj, float* gv, int B, int N, xggm_stream_t s);
/* add_edge_noise_v2                                       src/module/graph_utils.py:162-168
 * randn = torch.randn_like(adj) drawn by the caller (keeps torch's RNG stream).
 * target = -n / (float)(sigma*sigma), the reference's `-noise / (sigma ** 2)`. */
int xggm_edge_noise(const float* adj, const float* randn, double sigma, float* noisy,
                    float* target, int B, int N, xggm_stream_t s);
/* add_feature_noise_v2                                    src/module/graph_utils.py:144-149
 * f is [B,N,H], or [B,H] broadcast over the N nodes when f_is_broadcast != 0
 * (node_fc on 36 identical rows, src/vqa/vqacpv2.py:228-229). */
int xggm_feat_noise(const float* f, const float* randn, double sigma, float* noisy, float* target,
                    int B, int N, int H, int f_is_broadcast, xggm_stream_t s);
/* same, also emitting the planes of `noisy` (noisy_planes?) for the first GNN layer */
int xggm_feat_noise_ex(const float* f, const float* randn, double sigma, float* noisy, float* target,
                       void* noisy_planes, int B, int N, int H, int f_is_broadcast, xggm_stream_t s);
/* same with the Gaussian draw INSIDE the kernel (no randn tensor): Philox4x32-10, key = rng->seed, counter = float4 index,
 * subsequence = rng->stream0 + (*rng->dev_epoch << 32 if non-NULL), two Box-Muller pairs per counter.  H % 4 == 0.
 * (xggm_philox_t is declared with the GCN / GIN layers below.) */
struct xggm_philox_s;
int xggm_feat_noise_philox(const float* f, const struct xggm_philox_s* rng, double sigma, float* noisy, float* target,
                           void* noisy_planes, int B, int N, int H, int f_is_broadcast, xggm_stream_t s);
/* out[B,H] = sum_n g[B,n,H]  (backward of the broadcast above) */
int xggm_sum_nodes(const float* g, float* out, int B, int N, int H, xggm_stream_t s);
/* loss_func                                               src/vqa/vqacpv2.py:48-51
 * loss[0] = 0.5 sigma^2 / n_elem * sum (score-target)^2 ; n_elem = B*R*C of the [B,R,C] inputs
 * (sum over the last two dims, mean over B, divided by R*C). */
int xggm_score_mse_fwd(const float* score, const float* target, double sigma, float* loss,
                       long long n_elem, xggm_stream_t s);
/* gscore = gloss[0] * sigma^2/n_elem * (score-target) */
int xggm_score_mse_bwd(const float* score, const float* target, const float* gloss, double sigma,
                       float* gscore, long long n_elem, xggm_stream_t s);
/* compute_kl_loss                                         src/vqa/vqacpv2.py:54-61
 * x,y are [R,C], softmax over C; loss[0] = mean over R*C. */
int xggm_sym_kl_fwd(const float* x, const float* y, float* loss, int R, int C, xggm_stream_t s);
/* gx?, gy? overwritten with d loss/dx * gloss[0], d loss/dy * gloss[0]. */
int xggm_sym_kl_bwd(const float* x, const float* y, const float* gloss, float* gx, float* gy,
                    int R, int C, xggm_stream_t s);
/* cat[x, tanh(mean_n nodes)]                              src/vqa/vqacpv2.py:216-218
 * out[B,2H] = [xp[B,H] | tanh(mean_n nodes[B,N,H])] */
int xggm_fuse_readout_fwd(const float* xp, const float* nodes, float* out, int B, int N, int H,
                          xggm_stream_t s);
/* gxp[B,H] = gout[:, :H];  gnodes[B,N,H] (+)= gout[:, H:] * (1-t^2)/N with t = out[:, H:] */
int xggm_fuse_readout_bwd(const float* gout, const float* out, float* gxp, float* gnodes, int B,
                          int N, int H, int accumulate_gnodes, xggm_stream_t s);
/* Node-branch tail in one pass per direction          src/vqa/vqacpv2.py:236-246 (= gqa_ood.py:243-252)
 *   loss[0] = kl_w * compute_kl_loss(nodes, feat) + sm_w * loss_func(nodes, target, sigma)
 *   cat[B,2H] = [xp | tanh(mean_n nodes)]                       (the fusion_fc input)
 * i.e. xggm_sym_kl + xggm_score_mse + xggm_fuse_readout with the generated node features read once
 * and their gradient written once.  kl_w carries the trainer's weights (0.15 * num_answers), sm_w = 6.
 * bwd: gnodes[B,N,H] = gloss[0] * d loss/d nodes + read-out gradient of gcat; gfeat (nullable) =
 * gloss[0] * d loss/d feat; gxp[B,H] = gcat[:, :H]; grow[B,H] is caller scratch. */
int xggm_node_tail_fwd(const float* nodes, const float* feat, const float* target, const float* xp,
                       double sigma, double kl_w, double sm_w, float* loss, float* cat, int B, int N,
                       int H, xggm_stream_t s);
int xggm_node_tail_bwd(const float* nodes, const float* feat, const float* target, const float* cat,
                       const float* gloss, const float* gcat, double sigma, double kl_w, double sm_w,
                       float* gnodes, float* gfeat, float* gxp, float* grow, int B, int N, int H,
                       xggm_stream_t s);
/* ------------------------------------------------------------------------- *
 * Training-step tail: answer loss, gradient clipping, BertAdam -- fused passes over
 * flat fp32 buffers (all parameters / gradients / moments of a parameter group contiguous, same offsets).
 * ------------------------------------------------------------------------- */
/* nn.BCEWithLogitsLoss()(logit, target) * scale           src/vqa/vqacpv2.py:110,173 (scale = target.size(1))
 * loss[0] = scale / n * sum( max(x,0) - x t + log(1 + exp(-|x|)) );  gx = gloss[0] * scale / n * (sigmoid(x) - t) */
int xggm_bce_logits_fwd(const float* logit, const float* target, double scale, float* loss, long long n,
                        xggm_stream_t s);
int xggm_bce_logits_bwd(const float* logit, const float* target, const float* gloss, double scale,
                        float* glogit, long long n, xggm_stream_t s);
/* sumsq[0] (+)= sum g^2 : the squared total norm of nn.utils.clip_grad_norm_ (src/vqa/vqacpv2.py:175,223,252).
 * Call once per flat gradient buffer with accumulate = 1 after the first to clip several buffers (e.g. the
 * block's and the encoder's) by their joint norm. */
int xggm_grad_sumsq(const float* g, long long n, float* sumsq, int accumulate, xggm_stream_t s);
/* BertAdam.step                                           src/lxrt/optimization.py:116-203
 *   g' = g * min(1, max_norm / (sqrt(sumsq[0]) + 1e-6))      (sumsq NULL or max_norm <= 0: no clipping)
 *   m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  p -= lr * (m / (sqrt(v) + eps) + weight_decay * p)
 * `lr` is the SCHEDULED learning rate of this step (warmup_linear etc. are host arithmetic on the step
 * counter, optimization.py:28-55); there is no bias correction, as in the reference.  g is not modified. */
int xggm_bertadam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double b1,
                       double b2, double eps, double weight_decay, const float* sumsq, double max_norm,
                       xggm_stream_t s);
/* CUDA-graph-safe variant: the learning-rate schedule of optimization.py:28-55,176-191 is evaluated ON THE DEVICE
 * from a device step counter, so a captured step keeps following warmup_linear / warmup_cosine / warmup_constant
 * across replays.  lr_scheduled = lr * schedule(step_dev[0] / t_total, warmup) (t_total <= 0: lr unscheduled);
 * with advance != 0 the launch also increments step_dev[0] after every CTA has read it (ticket_dev: a zero-
 * initialised device word the library uses for that hand-shake; it is left at zero).  A parameter group that
 * is updated in several calls (disjoint ranges, see xggm_b200.optim) sets advance on the last call only.
 * n == 0 with advance != 0 only advances the counter.  sched == NULL: identical to xggm_bertadam_step. */
#define XGGM_SCHED_COSINE 0
#define XGGM_SCHED_CONSTANT 1
#define XGGM_SCHED_LINEAR 2
typedef struct {
    long long* step_dev;        /* device int64: optimizer steps taken so far */
    unsigned int* ticket_dev;   /* device uint32, zero between calls */
    double warmup;              /* optimization.py: `warmup` (fraction of t_total) */
    long long t_total;          /* optimization.py: `t_total` (-1 / 0: constant lr) */
    int schedule;               /* XGGM_SCHED_* */
    int advance;                /* != 0: step_dev[0] += 1 at the end of this launch */
} xggm_lr_schedule_t;
int xggm_bertadam_step_ex(float* p, const float* g, float* m, float* v, long long n, double lr, double b1,
                          double b2, double eps, double weight_decay, const float* sumsq, double max_norm,
                          const xggm_lr_schedule_t* sched, xggm_stream_t s);
/* ------------------------------------------------------------------------- *
 * Data-parallel optimiser step over NVLink / NVSwitch PEER MEMORY (one process per GPU): the gradient
 * all-reduce (average), clip_grad_norm_, BertAdam and the parameter all-gather as one fused sequence in which rank r
 * owns the r-th 1/world slice of the flat buffers: it averages its slice of the `world` peer gradient buckets with
 * 16-byte loads from peer memory (reduce-scatter), the squared norms of the slices meet through one word per peer
 * (summed in rank order: bit-identical clip coefficient everywhere), BertAdam runs on the slice only and the new
 * parameters are pushed into every peer's parameter buffer with remote stores (all-gather).  Five short kernels on the
 * caller's stream; the cross-GPU hand-shakes are system-scope flag words in the peers' control blocks (bounded waits:
 * a missing peer is a CUDA error after 2 s).  CUDA-graph capturable (the epoch lives in the control block).
 *   peers->grad[k] / param[k] : rank k's flat gradient / parameter buffer (n floats each, 16-byte aligned), mapped in
 *                               this process (torch symmetric memory / CUDA IPC / VMM -- the library only sees pointers)
 *   peers->grad_multicast / param_multicast : see the struct
 *   peers->ctl[k]             : rank k's control block, XGGM_DP_CTL_BYTES bytes, zeroed once at allocation
 *   m, v                      : this rank's moment buffers (only its slice is touched)
 *   range_lo/hi               : the bucket's ACTIVE element ranges (multiples of 4; parameters without a gradient this
 *                               step are skipped as optimization.py:139-141 does); n_ranges <= XGGM_DP_MAX_RANGES
 *   range_lr?                 : base learning rate per range (parameter groups with different rates sharing one
 *                               bucket, e.g. the trainers' encoder / down-task groups); NULL = `lr` everywhere
 *   sumsq_out?                : receives the squared total norm of the averaged gradient
 * Every rank must make the same call (same n, ranges, hyper-parameters) once per step.  world == 1 is valid. */
#define XGGM_DP_MAX_RANKS 16
#define XGGM_DP_MAX_RANGES 8
#define XGGM_DP_CTL_BYTES 512
typedef struct {
    int rank, world;
    void* grad[XGGM_DP_MAX_RANKS];
    void* param[XGGM_DP_MAX_RANKS];
    void* ctl[XGGM_DP_MAX_RANKS];
    /* optional NVSwitch multicast (NVLS) addresses of the same two symmetric allocations: when non-NULL the reduce-scatter
     * is ONE multimem.ld_reduce per 16 bytes (summed inside the switch) and the all-gather ONE multimem.st, instead of
     * `world` peer loads / stores.  NULL (no multicast support): unicast loops. */
    void* grad_multicast;
    void* param_multicast;
} xggm_dp_peers_t;
int xggm_dp_bertadam_step(const xggm_dp_peers_t* peers, float* m, float* v, long long n, const long long* range_lo,
                          const long long* range_hi, const double* range_lr, int n_ranges, double lr, double b1, double b2, double eps,
                          double weight_decay, double max_norm, const xggm_lr_schedule_t* sched, float* sumsq_out,
                          xggm_stream_t s);
/* elementwise sigmoid (encoder_adj tail, src/vqa/vqacpv2_model.py:91-94) */
int xggm_sigmoid_fwd(const float* x, float* y, long long n, xggm_stream_t s);
int xggm_sigmoid_bwd(const float* gy, const float* y, float* gx, long long n, xggm_stream_t s);
/* GeLU (exact erf), src/lxrt/modeling.py:116-140 */
int xggm_gelu_fwd(const float* x, float* y, long long n, xggm_stream_t s);
int xggm_gelu_bwd(const float* gy, const float* x, float* gx, long long n, xggm_stream_t s);
/* inverted dropout with an explicit mask, y = keep ? x*scale : 0 (GAT input dropout,
 * src/module/gat.py:73); the same call is its backward.  keep == NULL: y = x*scale. */
int xggm_mask_scale(const float* x, const uint8_t* keep, float scale, float* y, long long n,
                    xggm_stream_t s);
/* Tail of VisualFeatEncoder (src/lxrt/modeling.py:553-555): out = dropout((x + y) / 2),
 * keep? a uint8 mask (NULL = eval mode), scale = 1/(1-p).  Its backward w.r.t. x and y is
 * xggm_mask_scale(gout, keep, scale/2). */
int xggm_avg2_drop(const float* x, const float* y, const uint8_t* keep, float scale, float* out, long long n,
                   xggm_stream_t s);
/* The whole tail of VisualFeatEncoder in one pass over z = visn_fc(feats) [M,H] (src/lxrt/modeling.py:546-556):
 *   out = dropout( (LN_eps(z) * gamma1 + beta1  +  LN_eps(boxes box_w^T + box_b) * gamma2 + beta2) / 2 )
 * boxes [M,4], box_w [H,4] (box_fc.weight), keep? uint8 [M,H] (NULL = eval), scale = 1/(1-p).  Saves xhat1 [M,H],
 * rstd1 [M], mean2 [M], rstd2 [M].  Supported for pos_dim == 4, H % 128 == 0, H <= 1024 (xggm_visn_tail_supported);
 * other shapes compose xggm_linear_* / xggm_layernorm_* / xggm_avg2_drop.
 * Backward: gz = d loss / d z, gt = d loss / d (box projection output) (feed xggm_linear_bwd_weight with feats / boxes),
 * and the four LayerNorm parameter gradients, ACCUMULATED into (caller zeroes them). */
int xggm_visn_tail_supported(int H, int pos_dim);
int xggm_visn_tail_fwd(const float* z, const float* boxes, const float* box_w, const float* box_b, const float* gamma1,
                       const float* beta1, const float* gamma2, const float* beta2, const uint8_t* keep, float scale,
                       float* out, float* xhat1, float* rstd1, float* mean2, float* rstd2, int M, int H, float eps,
                       xggm_stream_t s);
int xggm_visn_tail_bwd(const float* gout, const float* xhat1, const float* rstd1, const float* boxes, const float* box_w,
                       const float* box_b, const float* mean2, const float* rstd2, const float* gamma1, const float* gamma2,
                       const uint8_t* keep, float scale, float* gz, float* gt, float* ggamma1, float* gbeta1, float* ggamma2,
                       float* gbeta2, int M, int H, xggm_stream_t s);
/* Philox keep-mask (1 = keep with probability 1-p): counter = element index / 4,
 * subsequence = stream_id + (*dev_epoch << 32 if dev_epoch, a device counter, is non-NULL). */
int xggm_keep_mask(uint8_t* keep, long long n, float p, uint64_t seed, uint64_t stream_id,
                   const uint64_t* dev_epoch, xggm_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* XGGM_B200_H */
