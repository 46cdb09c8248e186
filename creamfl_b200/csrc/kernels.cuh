// Internal host-side entry points of the kernel translation units (shared by capi.cu).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace cfl {
// gemm_tc.cu
struct GemmParams {
  int M, N, K;
  int split_k;
  void* out;
  long long ldo;
  int out_bf16;
  void* out2;
  const float* bias;
  int act;
  float alpha;
  const void* add;
  long long ld_add;
  int add_bf16;
  const __nv_bfloat16* aux;
  long long ld_aux;
  int atomic_out;  // accumulate into the fp32 output with atomics even when split_k == 1
  long long ldo2;  // pitch of out2 (0: same as ldo)
  int tma_out;     // set by gemm_bf16: bf16 output leaves through a staged TMA store
  double* stats;   // optional [2N]: += per-column sum and sum of squares of the bf16 output (BatchNorm statistics)
  DropSpec drop;   // optional dropout of (acc*alpha + bias) BEFORE the residual add (HF BertSelfOutput / BertOutput)
};
int gemm_bf16(const void* a, long long lda, int a_mn, const void* b, long long ldb, int b_mn, GemmParams p,
              cudaStream_t stream);
int gemm_tile_n(int N, int K);
int plan_split_k(long long tiles, long long nkb, long long min_per, double epi_kb, int workers = 0);
int gemm_plan_split(int M, int N, int K);
// sim_tc.cu
size_t rowlse_workspace_bytes(int M, int N);
int rowlse_bf16(const void* Q, const void* G, const long long* labels, int M, int N, int D, float scale,
                float* score, float* lse2, void* workspace, size_t ws_bytes, cudaStream_t stream);
int softmax_emit_bf16(const void* Q, const void* G, const long long* labels, const float* lse2, int M, int N,
                      int D, float scale, void* P, long long ldp, cudaStream_t stream);
// loss_ops.cu
int pcme_fwd(const float*, const float*, int, int, const float*, const float*, float*, float*, float*, size_t,
             cudaStream_t);
int pcme_bwd(const float*, const float*, const float*, int, int, const float*, const float*, const float*,
             float*, float*, float*, float*, float*, size_t, cudaStream_t);
int moon_fwd(const float*, const float*, const float*, const long long*, int, int, float, float, float*, float*,
             float*, cudaStream_t);
int moon_bwd(const float*, const float*, const long long*, const float*, const float*, int, int, float*,
             cudaStream_t);
int mse_gather_fwd(const float*, const float*, const long long*, int, int, float*, float*, size_t, cudaStream_t);
int mse_gather_bwd(const float*, const float*, const long long*, const float*, int, int, float*, cudaStream_t);
int l2norm_fwd(const float*, int, int, float*, void*, float*, cudaStream_t);
int l2norm_bwd(const float*, const float*, const float*, int, int, float*, cudaStream_t);
int cast_f32_bf16(const float*, long long, void*, cudaStream_t);
int scale_by_scalar(float*, long long, const float*, float, cudaStream_t);
int act_bwd_f32(const float*, const float*, long long, int, void*, cudaStream_t);
__global__ void sum_finish_kernel(const float* in, int n, float scale, float* out);
// agg_ops.cu
int conw_reduce(const float* const*, const float*, int, int, int, float*, float*, cudaStream_t);
int recall_ranks(const float*, const float*, const long long*, const long long*, int, int, int, int*, void*,
                 size_t, cudaStream_t);
}  // namespace cfl

namespace cfl {
// conv_tc.cu
int conv_same_fprop(const void*, const void*, int, int, int, int, int, int, int, void*, cudaStream_t,
                    const float* bias = nullptr, const void* add = nullptr, int relu = 0, int stride = 1);
int conv_same_dgrad(const void*, const void*, int, int, int, int, int, int, int, void*, const void*, cudaStream_t,
                    int stride = 1);
int conv_same_wgrad(const void*, const void*, int, int, int, int, int, int, int, float*, cudaStream_t, int stride = 1);
// stem_tc.cu
bool stem_supported(int C, int H, int W, int R, int S, int stride, int pad, int Cout);
int stem_fprop(const float* x, const void* wt, long long ldw, int N, int H, int W, void* y, cudaStream_t stream);
int stem_wgrad(const float* x, const void* dy, int N, int H, int W, float* dw, cudaStream_t stream);
// nn_ops.cu
int bn_train_fwd(const void*, long long, int, const float*, const float*, float, float, float*, float*, double*,
                 float*, float*, float*, float*, const void*, int, int, long long*, void*, cudaStream_t,
                 void* relu_mask = nullptr);
int bn_stats_only(const void*, long long, int, double*, cudaStream_t);
int bn_fold_layers(const long long*, const long long*, int, long long, float, cudaStream_t);
int bn_eval_fwd(const void*, long long, int, const float*, const float*, float, const float*, const float*, float*,
                float*, const void*, int, void*, cudaStream_t);
int bn_train_bwd(const void*, const void*, const void*, long long, int, const float*, const float*, int, const float*,
                 const float*, double*, float*, float*, float*, void*, void*, cudaStream_t,
                 const void* relu_mask = nullptr, const void* pool_idx = nullptr, int H = 0, int W = 0);
int maxpool_fwd(const void*, int, int, int, int, void*, void*, cudaStream_t, const float* scale = nullptr,
                const float* shift = nullptr);
int maxpool_bwd(const void*, const void*, int, int, int, int, void*, cudaStream_t);
int im2col_nhwc(const void*, int, int, int, int, int, int, int, int, int, void*, cudaStream_t);
int im2col_nchw_f32(const float*, int, int, int, int, int, int, int, int, int, void*, cudaStream_t);
int col2im_nhwc(const void*, int, int, int, int, int, int, int, int, int, const void*, void*, cudaStream_t);
size_t layernorm_bwd_workspace_bytes(int);
int layernorm_fwd(const void*, const void*, const float*, const float*, float, int, int, int, void*, float*, float*,
                  cudaStream_t);
int layernorm_bwd(const void*, const void*, const void*, const float*, const float*, const float*, int, int, int,
                  void*, float*, float*, float*, void*, size_t, cudaStream_t);
int layernorm_fwd_drop(const void*, const void*, const float*, const float*, float, int, int, int, void*, float*, float*,
                       DropSpec, cudaStream_t);
int layernorm_bwd_drop(const void*, const void*, const void*, const float*, const float*, const float*, int, int, int,
                       void*, void*, float*, float*, float*, void*, size_t, DropSpec, DropSpec, cudaStream_t);
int attn_fwd_drop(const void*, const float*, int, int, int, int, void*, void*, DropSpec, cudaStream_t);
int attn_bwd_drop(const void*, const void*, const void*, int, int, int, int, void*, float*, DropSpec, cudaStream_t);
int dropout_mask(DropSpec, long long, void*, cudaStream_t);
int rng_tick(unsigned long long*, cudaStream_t);
int colsum_bf16(const void*, int, int, long long, float*, cudaStream_t);
int add_bf16(const void*, const void*, long long, void*, cudaStream_t);
int embed_fwd(const long long*, const long long*, const float*, const float*, const float*, int, int, int, void*,
              cudaStream_t);
int embed_bwd(const long long*, const long long*, const void*, int, int, int, float*, float*, float*, cudaStream_t);
int attn_fwd(const void*, const float*, int, int, int, int, void*, void*, cudaStream_t);
int attn_bwd(const void*, const void*, const void*, int, int, int, int, void*, float*, cudaStream_t);
int pie_pool_fwd(const void*, const void*, const float*, int, int, int, int, float*, void*, void*, cudaStream_t);
int pie_pool_bwd(const void*, const void*, const float*, const float*, const void*, const void*, int, int, int, int,
                 void*, void*, float*, cudaStream_t);
// optim.cu
int optimizer_step(const void*, int, const void*, int, const float*, float*, double*, float*, int*, float*, float*,
                   cudaStream_t);
int avgpool_fwd(const void*, int, int, int, float, float*, void*, cudaStream_t);
int avgpool_bwd(const void*, int, int, int, float, void*, cudaStream_t);
int ce_fwd(const float*, long long, const long long*, int, int, float, float*, float*, float*, cudaStream_t);
int relu_inplace(float*, void*, long long, cudaStream_t);
// text_ops.cu
int wemb_gather_fwd(const long long*, const float*, int, int, int, int, void*, cudaStream_t);
int wemb_scatter_bwd(const long long*, const void*, int, int, int, int, float*, cudaStream_t);
int gru_fwd(const float*, const float*, const float*, const int*, int, int, int, int, float*, float*, float*,
            cudaStream_t);
int gru_bwd(const float*, const float*, const float*, const int*, const float*, const float*, int, int, int, int, void*,
            void*, void*, cudaStream_t);
int seq_pool_fwd(const void*, const void*, const float*, const int*, int, int, int, int, int, int, float*, void*,
                 cudaStream_t);
int seq_pool_bwd(const void*, const void*, const float*, const float*, const void*, const int*, int, int, int, int, int,
                 int, void*, void*, float*, cudaStream_t);
int scale_relu_fwd(const float*, long long, float, float*, cudaStream_t);
int scale_relu_bwd(const float*, const float*, long long, float, float*, cudaStream_t);
}  // namespace cfl
