// Internal launcher prototypes shared by the translation units of libxggm_b200.so.
// Every function enqueues on `st`, returns XGGM_OK or a negative XGGM_ERR_* code and never synchronises.
//   gemm_simt.cu  exact-fp32 FMA GEMM, column sums, per-launch GEMM timing
//   gemm_tc.cu    tcgen05 GEMM / Gram / block-diagonal message passing, bf16 plane builders
//   rowops.cu     LayerNorm and GeLU+LayerNorm+dropout row kernels
//   graph_ops.cu  per-graph SIMT kernels (message passing, pair scores, adjacency regeneration, GAT attention)
//   glue.cu       trainer glue (noise, scatter, losses, read-out, masks)
//   optim.cu      answer-loss / gradient-norm / BertAdam passes over the flat parameter buffers
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace xggm {

int gemm_simt(int op, const float* A, const float* Bm, const float* bias, const float* resid,
              float* C, int M, int N, int K, int accumulate, cudaStream_t st);
int colsum(const float* g, float* out, int R, int C, int accumulate, cudaStream_t st);
bool gemm_tc_supported(int M, int N, int K);
int gemm_tc(bool a_mn, bool b_mn, const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, const __nv_bfloat16* b_hi,
            const __nv_bfloat16* b_lo, const float* bias, const float* resid, float* C, __nv_bfloat16* c_hi,
            __nv_bfloat16* c_lo, int M, int N, int K, int accumulate, int allow_split_k, int npass, cudaStream_t st,
            int lda = 0, int ldb = 0);
bool gemm_tc_ragged_ok(int M, int N, int K);
int split_planes_pitched(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long R, int C, int P, cudaStream_t st);
// one product of a grouped launch (gemm_tc_group): same shape / operand layout for every member
struct GemmProb {
    const __nv_bfloat16 *a_hi, *a_lo, *b_hi, *b_lo;
    const float* bias;
    const float* resid;
    float* C;
    __nv_bfloat16 *c_hi, *c_lo;
    int accumulate;
    int lda, ldb;   // row pitch (elements) of padded A / B planes; 0 = dense
};
int gemm_tc_group(bool a_mn, bool b_mn, const GemmProb* pr, int count, int M, int N, int K, int allow_split_k,
                  int npass, cudaStream_t st, bool kcat = false);
bool adj_tc_supported(int N, int H);
long long adj_tc_coef_elems(int B, int N);
int build_blockdiag(const float* adj, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int N, float alpha0,
                    const float* alpha_dev, float self_w, int trans, cudaStream_t st);
int adj_apply_tc(const __nv_bfloat16* c_hi, const __nv_bfloat16* c_lo, const __nv_bfloat16* x_hi,
                 const __nv_bfloat16* x_lo, float* out, __nv_bfloat16* o_hi, __nv_bfloat16* o_lo, int B, int N, int H,
                 int accumulate, int npass, cudaStream_t st, const float* resid = nullptr, const __nv_bfloat16* r_hi = nullptr,
                 const __nv_bfloat16* r_lo = nullptr);
bool adj_ln_tc_supported(int N, int H);
int adj_ln_tc(const __nv_bfloat16* c_hi, const __nv_bfloat16* c_lo, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
              const __nv_bfloat16* r_hi, const __nv_bfloat16* r_lo, const float* gamma, const float* beta, float* xhat,
              int xhat_bf16, float* rstd, float* h_out, __nv_bfloat16* o_hi, __nv_bfloat16* o_lo, int B, int N, int H, float eps,
              int npass, cudaStream_t st);
bool gram_tc_supported(int N, int H);
int gram_tc(const __nv_bfloat16* p_hi, const __nv_bfloat16* p_lo, const __nv_bfloat16* q_hi, const __nv_bfloat16* q_lo,
            float* S, int B, int N, int H, int npass, cudaStream_t st);
int adj_regen_from_s(const float* S, float* adj_out, int32_t* amax, int B, int N, int squash, cudaStream_t st);
int scale_accum(const float* S, float* out, long long n, float alpha0, const float* alpha_dev, int accumulate,
                const float* dot_ref, float* dot_out, cudaStream_t st);
int split_planes(const float* const* src, __nv_bfloat16* const* hi, __nv_bfloat16* const* lo, const long long* n,
                 int count, cudaStream_t st);
int split_planes_t(const float* const* src, __nv_bfloat16* const* hi, __nv_bfloat16* const* lo, int R, int C,
                   int count, cudaStream_t st, int pitch = 0);
void gemm_tc_set_debug(unsigned long long* dev_buf);
int gemm_prof_enable(int on);
int gemm_prof_read(double* total_ms, long long* launches, double* flops);
int layernorm_fwd(const float*, const float*, const float*, float*, float*, float*, __nv_bfloat16*, __nv_bfloat16*, int, int, float, cudaStream_t, int xhat_bf16 = 0);
int layernorm_bwd(const float*, const float*, const float*, const float*, float*, float*, float*, __nv_bfloat16*, __nv_bfloat16*, int, int, cudaStream_t, int xhat_bf16 = 0);
int gelu_ln_drop_fwd(const float*, const float*, const float*, const DropSpec&, float*, float*, float*, __nv_bfloat16*, __nv_bfloat16*, int, int, float, int, cudaStream_t, int z_bf16 = 0);
int gelu_ln_drop_bwd(const float*, const float*, const float*, const float*, const float*, const DropSpec&, float*, float*, float*, float*, __nv_bfloat16*, __nv_bfloat16*, int, int, cudaStream_t, int z_bf16 = 0);
int adj_apply(const float*, const float*, float*, __nv_bfloat16*, __nv_bfloat16*, int, int, int, float, const float*, float, bool, int, cudaStream_t);
bool adj_ln_supported(int N, int H);
int adj_ln_fwd(const float* adj, const float* P, const float* resid, const float* gamma, const float* beta, float* h,
               float* xhat, float* rstd, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int N, int H, float eps, cudaStream_t st);
int bmm_nt(const float*, const float*, float*, int, int, int, float, const float*, int, const float*, float*, cudaStream_t);
int adj_regen_fwd(const float*, float*, float*, int32_t*, int, int, int, int, cudaStream_t);
int adj_regen_bwd(const float*, const float*, const float*, const int32_t*, float*, float*, int, int, int, int, int, cudaStream_t);
int adj_regen_bwd_coeffs(const float* gadj, const float* S, const int32_t* amax, float* D, int B, int N, int squash, cudaStream_t st);
int gat_attn_fwd(const float*, const float*, const float*, float*, float*, float*, int, int, int, float, int, cudaStream_t);
int gat_attn_bwd(const float*, const float*, const float*, const float*, const float*, const float*, float*, float*, float*, int, int, int, float, int, cudaStream_t);
int gelu_fwd(const float*, float*, long long, cudaStream_t);
int gelu_bwd(const float*, const float*, float*, long long, cudaStream_t);
int mask_scale(const float*, const uint8_t*, float, float*, long long, cudaStream_t);
int avg2_drop(const float*, const float*, const uint8_t*, float, float*, long long, cudaStream_t);
int strip_diag(const float*, float*, int, int, cudaStream_t);
int triu_scatter_fwd(const float*, float*, int, int, cudaStream_t);
int triu_scatter_bwd(const float*, float*, int, int, cudaStream_t);
int edge_noise(const float*, const float*, float, float, float*, float*, int, int, cudaStream_t);
int feat_noise(const float*, const float*, float, float, float*, float*, __nv_bfloat16*, __nv_bfloat16*, int, int, int, int, cudaStream_t);
int feat_noise_philox(const float* f, uint64_t seed, uint64_t stream, const uint64_t* epoch, float sigma, float sigma2,
                      float* noisy, float* target, __nv_bfloat16* hi, __nv_bfloat16* lo, int B, int N, int H, int bcast,
                      cudaStream_t st);
int sum_nodes(const float*, float*, int, int, int, cudaStream_t);
int score_mse_fwd(const float*, const float*, float, float*, long long, cudaStream_t);
int score_mse_bwd(const float*, const float*, const float*, float, float*, long long, cudaStream_t);
int sym_kl_fwd(const float*, const float*, float*, int, int, cudaStream_t);
int sym_kl_bwd(const float*, const float*, const float*, float*, float*, int, int, cudaStream_t);
int fuse_readout_fwd(const float*, const float*, float*, int, int, int, cudaStream_t);
int fuse_readout_bwd(const float*, const float*, float*, float*, int, int, int, int, cudaStream_t);
int node_tail_fwd(const float* nodes, const float* feat, const float* target, const float* xp, float sigma, float kl_w,
                  float sm_w, float* loss, float* cat, int B, int N, int H, cudaStream_t st);
int node_tail_bwd(const float* nodes, const float* feat, const float* target, const float* cat, const float* gloss,
                  const float* gcat, float sigma, float kl_w, float sm_w, float* gnodes, float* gfeat, float* gxp,
                  float* grow, int B, int N, int H, cudaStream_t st);
int bce_logits_fwd(const float* x, const float* t, float scale, float* loss, long long n, cudaStream_t st);
int bce_logits_bwd(const float* x, const float* t, const float* gloss, float scale, float* gx, long long n, cudaStream_t st);
int grad_sumsq(const float* g, long long n, float* out, int accumulate, cudaStream_t st);
int bertadam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double b1, double b2, double eps,
                  double wd, const float* sumsq, double max_norm, const xggm_lr_schedule_t* sched, cudaStream_t st);
int dp_bertadam_step(const xggm_dp_peers_t* peers, float* m, float* v, long long n, const long long* range_lo,
                     const long long* range_hi, const double* range_lr, int n_ranges, double lr, double b1, double b2, double eps, double wd,
                     double max_norm, const xggm_lr_schedule_t* sched, float* sumsq_out, cudaStream_t st);
bool visn_tail_supported(int H, int pos_dim);
int visn_tail_fwd(const float* z, const float* boxes, const float* Wb, const float* bb, const float* g1, const float* b1,
                  const float* g2, const float* b2, const DropSpec& drop, float* out, float* xhat1, float* rstd1, float* mean2,
                  float* rstd2, int M, int H, float eps, cudaStream_t st);
int visn_tail_bwd(const float* gout, const float* xhat1, const float* rstd1, const float* boxes, const float* Wb,
                  const float* bb, const float* mean2, const float* rstd2, const float* g1, const float* g2,
                  const DropSpec& drop, float* gz, float* gt, float* gg1, float* gb1, float* gg2, float* gb2, int M, int H,
                  cudaStream_t st);
long long weight_planes_elems(int N, int K);
void weight_planes_views(void* buf, int N, int K, __nv_bfloat16** hi, __nv_bfloat16** lo, __nv_bfloat16** thi,
                         __nv_bfloat16** tlo);
int weight_planes_build(const float* const* W, void* const* bufs, const int* N, const int* K, int count, bool with_lo,
                        cudaStream_t st);
int sigmoid_fwd(const float*, float*, long long, cudaStream_t);
int sigmoid_bwd(const float*, const float*, float*, long long, cudaStream_t);
int keep_mask(uint8_t*, long long, float, uint64_t, uint64_t, const uint64_t*, cudaStream_t);


}  // namespace xggm
