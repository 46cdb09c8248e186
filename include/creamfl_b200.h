/* creamfl_b200 - C ABI of the B200-native hot path of CreamFL.
 *
 * The reference (FLAIR-THU/CreamFL) has no FFI layer: its hot path is reached through Python factories
 * (get_model / get_criterion / losses.create) and inline torch code in the trainers.  Every entry point below
 * replaces one of those library call sites; the comment on each names the reference lines it stands in for.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host
 *   - the caller owns all memory, including workspaces (size from the matching *_workspace_bytes function)
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *   - row-major contiguous tensors; bf16 operands of tensor-core kernels must be 16-byte aligned
 *   - return value: 0 ok, -1 bad argument, -2 workspace too small, -3 CUDA error; text from creamfl_last_error()
 *   - no CPU fallback exists: on a machine without an sm_100a GPU the calls fail with -3
 */
#ifndef CREAMFL_B200_H
#define CREAMFL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CREAMFL_OK 0
#define CREAMFL_EINVAL (-1)
#define CREAMFL_EWORKSPACE (-2)
#define CREAMFL_ECUDA (-3)

/* activation codes of creamfl_gemm_bf16 */
#define CREAMFL_ACT_NONE 0
#define CREAMFL_ACT_GELU 1      /* erf GELU (HF BertIntermediate) */
#define CREAMFL_ACT_RELU 2
#define CREAMFL_ACT_TANH 3      /* PIENet w_1 (pie_model.py:30) */
#define CREAMFL_ACT_DGELU 4     /* out = acc * gelu'(aux)   (backward of GELU) */
#define CREAMFL_ACT_DRELU 5     /* out = acc * (aux > 0)    (backward of ReLU) */
#define CREAMFL_ACT_SIGMOID 6   /* PIENet residual gate (pie_model.py:63) */

const char* creamfl_last_error(void);
int creamfl_abi_version(void);

/* ---- dense contraction on tcgen05 ---------------------------------------------------------------------
 * out[M,N] = act(alpha * sum_k A(m,k) B(n,k) + bias[n] + add[m,n]).
 * a_mn/b_mn = 0: operand stored [rows, K] (K contiguous); = 1: stored [K, rows].
 * Replaces cuBLAS behind nn.Linear / HF BertModel (src/networks/models/pcme.py:31-44,
 * pie_model.py:18-19,51, image_encoder.py:30,57) and 1x1 convolutions of torchvision ResNet
 * (image_encoder.py:24) in NHWC.  split_k > 1 or accumulate != 0 adds into the existing fp32 `out` with atomics
 * (weight gradients accumulate into the parameter's grad buffer); split_k == 0 lets the library choose. */
int creamfl_gemm_bf16(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, int M, int N,
                      int K, void* out, int64_t ldo, int out_bf16, void* out_preact_bf16, const float* bias,
                      int act, float alpha, const void* add, int64_t ld_add, int add_bf16, const void* aux_bf16,
                      int64_t ld_aux, int split_k, int accumulate, void* stream);
/* the split the library chooses for an accumulating [M, N, K] product when split_k == 0 (host-only introspection:
 * whole waves of tiles * split units on the device's SMs; 148 is assumed when no device is present) */
int creamfl_plan_split_k(int M, int N, int K);
/* elementwise derivative of an activation from its OUTPUT y: kind CREAMFL_ACT_SIGMOID -> dy*y*(1-y),
 * CREAMFL_ACT_TANH -> dy*(1-y^2); fp32 in, bf16 out (the result feeds a dgrad/wgrad GEMM) */
int creamfl_act_bwd_f32(const float* dy, const float* y, int64_t n, int kind, void* out_bf16, void* stream);

/* ---- inter-modal InfoNCE (MMClientTrainer.py:193-201,301-308; ClientTrainer.py:388-401,493-502) ------
 * logits = inv_tau * Q G^T  (never materialised in fp32), CE against column labels[i], mean over rows.
 * q_bf16 [B,D], g_bf16 [N,D], labels int64 [B] (distill_dict positions), D in {64,128,192,256}.
 * Outputs: loss[1]; row_score[B] = logit_pos - logsumexp (= -CE per row); lse2[B] (log2 domain, for bwd). */
size_t creamfl_rowlse_workspace_bytes(int M, int N);
int creamfl_infonce_fwd(const void* q_bf16, const void* g_bf16, const int64_t* labels, int B, int N, int D,
                        float inv_tau, float* loss, float* row_score, float* lse2, void* workspace,
                        size_t workspace_bytes, void* stream);
/* dQ[B,D] (fp32) = gout[0] * inv_tau / B * (softmax(logits) - onehot) G */
size_t creamfl_infonce_bwd_workspace_bytes(int B, int N);
int creamfl_infonce_bwd(const void* q_bf16, const void* g_bf16, const int64_t* labels, const float* lse2, int B,
                        int N, int D, float inv_tau, const float* gout, float* dq, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ---- con_w aggregation (MMFL.py:298-335) --------------------------------------------------------------
 * score[n] = <V[n],G[n]> - log sum_j exp <V[n],G[j]>   (MMFL.py:304-307): v_bf16, g_bf16 [N,D]. */
int creamfl_conw_score(const void* v_bf16, const void* g_bf16, int N, int D, float* score, void* workspace,
                       size_t workspace_bytes, void* stream);
/* out[n,:] = sum_c softmax_c(scores[:,n])[c] * vecs[c][n,:]   (MMFL.py:311-314).  vecs_host: HOST array of C
 * device pointers to fp32 [N,D]; scores fp32 [C,N]; weights (optional) fp32 [C,N]. */
int creamfl_conw_reduce(const float* const* vecs_host, const float* scores, int C, int N, int D, float* out,
                        float* weights, void* stream);

/* ---- PCME soft-contrastive loss (src/criterions/probemb.py:185-256) -----------------------------------
 * img, txt fp32 [N,D]; shift, neg_scale: device scalars (learnable).  out3 = {loss, pos part, neg part} where
 * loss = i2t + t2i and the parts are per direction; dist [N,N] is kept for the backward. */
size_t creamfl_pcme_workspace_bytes(int N);
int creamfl_pcme_fwd(const float* img, const float* txt, int N, int D, const float* shift,
                     const float* neg_scale, float* dist, float* out3, void* workspace, size_t workspace_bytes,
                     void* stream);
int creamfl_pcme_bwd(const float* img, const float* txt, const float* dist, int N, int D, const float* shift,
                     const float* neg_scale, const float* gout, float* d_img, float* d_txt, float* d_shift,
                     float* d_neg_scale, void* workspace, size_t workspace_bytes, void* stream);

/* ---- intra-modal (MOON) contrast (MMClientTrainer.py:169-191; ClientTrainer.py:404-414) ---------------
 * row r: CE([<z,bank[idx]>, <z,zold>] * inv_tau, 0) / denom.  loss (optional) = sum of loss_rows. */
int creamfl_moon_fwd(const float* z, const float* zold, const float* bank, const int64_t* idx, int R, int D,
                     float inv_tau, float denom, float* loss_rows, float* coef, float* loss, void* stream);
int creamfl_moon_bwd(const float* zold, const float* bank, const int64_t* idx, const float* coef,
                     const float* gout, int R, int D, float* dz, void* stream);

/* ---- distillation MSE against aggregated rows (MMFL.py:296,355-378) ----------------------------------- */
size_t creamfl_mse_workspace_bytes(void);
int creamfl_mse_gather_fwd(const float* x, const float* bank, const int64_t* idx, int R, int D, float* loss,
                           void* workspace, size_t workspace_bytes, void* stream);
int creamfl_mse_gather_bwd(const float* x, const float* bank, const int64_t* idx, const float* gout, int R,
                           int D, float* dx, void* stream);

/* ---- L2 normalisation (src/utils/tensor_utils.py:25-27) and casts ------------------------------------- */
int creamfl_l2norm_fwd(const float* x, int R, int D, float* y, void* y_bf16, float* inv_norm, void* stream);
int creamfl_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, int R, int D, float* dx,
                       void* stream);
int creamfl_cast_f32_bf16(const float* x, int64_t n, void* y_bf16, void* stream);

/* ---- Recall@K ranks (src/algorithms/eval_coco.py:273-334) ---------------------------------------------
 * ranks[q] = #{g : <Q_q,G_g> > max_{g': g_lab[g'] == q_lab[q]} <Q_q,G_g'>}  (0-based best-positive rank). */
size_t creamfl_recall_workspace_bytes(int Nq);
int creamfl_recall_ranks(const float* q, const float* g, const int64_t* q_lab, const int64_t* g_lab, int Nq,
                         int Ng, int D, int32_t* ranks, void* workspace, size_t workspace_bytes, void* stream);


/* ======================================================================================================
 * Encoder towers.  Activations are NHWC bf16 (images) / [tokens, hidden] bf16 (text); parameters are fp32
 * masters with bf16 shadows for the tensor-core operands; filters are [Cout, R, S, Cin] (the memory order of
 * a torch channels_last OIHW tensor).  Replaces cuDNN / cuBLAS behind torchvision ResNet
 * (src/networks/models/image_encoder.py:24,55; src/networks/resnet_client.py:163-201) and HF BertModel
 * (src/networks/models/pcme.py:31-44).
 * ====================================================================================================== */

/* ---- convolution: implicit GEMM on tcgen05 for stride-1 "same" filters, plain GEMM for 1x1, im2col + GEMM
 * for strided filters (workspace = patch matrix).  w_pitch = elements between filter rows (>= R*S*Cin). */
size_t creamfl_conv2d_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad);
/* bn_sums (optional, 2*Cout doubles, accumulated): per-channel sum and sum of squares of the bf16 output - the
 * BatchNorm statistics of the layer that follows, produced by the GEMM epilogue where the output tile is on chip
 * (pass stats_ready = 1 to creamfl_bn_train_fwd afterwards) */
int creamfl_conv2d_fprop(const void* x_bf16, const void* w_bf16, int N, int H, int W, int Cin, int Cout, int R, int S,
                         int stride, int pad, int64_t w_pitch, void* y_bf16, double* bn_sums, void* workspace,
                         size_t workspace_bytes, void* stream);
/* Inference-mode convolution with the BatchNorm that follows folded in (torchvision ResNet under model.eval():
 * image_encoder.py:24,55 reached from MMFL.py:194-221 / MMClientTrainer.py:326-359 / the old model of
 * MMClientTrainer.py:154-167): y = [relu]( conv(x, w_folded) + bias [+ add] ), w_folded = w * gamma / sqrt(var + eps)
 * per output channel and bias = beta - mean * gamma / sqrt(var + eps) (creamfl_bn_fold_rows).  add (optional) is a bf16
 * tensor shaped like y (the residual).  Same three strategies and workspace as creamfl_conv2d_fprop. */
int creamfl_conv2d_fprop_affine(const void* x_bf16, const void* w_bf16, int N, int H, int W, int Cin, int Cout, int R,
                                int S, int stride, int pad, int64_t w_pitch, const float* bias, const void* add_bf16,
                                int relu, void* y_bf16, void* workspace, size_t workspace_bytes, void* stream);
/* Fold eval-mode BatchNorm layers into the filters of the convolutions that feed them, all layers in one launch.
 * layers: device table [n_layers, 10] int64, one entry per convolution:
 *   {w_f32 (address of the fp32 filters, [Cout, K] contiguous), K = R*S*Cin, Cout, w_out_bf16 (address), pitch of w_out
 *    in elements (>= K), gamma, beta, running_mean, running_var (addresses of the C-float vectors), bias_out (address)}
 * row_start: device [n_layers + 1] int64 prefix sums of Cout (row_start[n_layers] = total filter count).
 * w_out[c, k] = bf16( w[c, k] * s_c ), bias_out[c] = beta[c] - mean[c] * s_c,  s_c = gamma[c] / sqrt(var[c] + eps). */
int creamfl_bn_fold_layers(const int64_t* layers, const int64_t* row_start, int n_layers, int64_t total_rows, float eps,
                           void* stream);
/* dx = conv_transpose(dy, w) [+ add] ; add (optional) is a bf16 tensor shaped like dx (residual-branch gradient) */
int creamfl_conv2d_dgrad(const void* dy_bf16, const void* w_bf16, int N, int H, int W, int Cin, int Cout, int R, int S,
                         int stride, int pad, int64_t w_pitch, const void* add_bf16, void* dx_bf16, void* workspace,
                         size_t workspace_bytes, void* stream);
/* dw[Cout, R*S*Cin] (fp32, pitch R*S*Cin) += dy^T * patches(x).  If col_bf16 is non-null it is the patch matrix
 * kept from the forward pass (strided path) and x is not read. */
int creamfl_conv2d_wgrad(const void* dy_bf16, const void* x_bf16, const void* col_bf16, int N, int H, int W, int Cin,
                         int Cout, int R, int S, int stride, int pad, float* dw, void* workspace,
                         size_t workspace_bytes, void* stream);
/* 7x7/2 stem: fp32 NCHW images -> bf16 patch matrix [N*Ho*Wo, col_pitch] (column order r, s, c; zero tail) */
int creamfl_im2col_nchw_f32(const float* images, int N, int C, int H, int W, int R, int S, int stride, int pad,
                            int col_pitch, void* col_bf16, void* stream);

/* ---- fused ResNet stem: conv 7x7 / stride 2 / pad 3, 3 -> 64 channels, straight from fp32 NCHW images (torchvision
 * ResNet.conv1 reached from image_encoder.py:24,55 and resnet_client.py:164).  The patch operand of the implicit GEMM
 * is assembled in shared memory, no patch matrix is written to HBM.  Supported when the output row fits one tile
 * (image width <= 256); creamfl_stem_supported tells (callers fall back to creamfl_im2col_nchw_f32 + creamfl_gemm_bf16).
 * w_bf16: filters [64, w_pitch >= 147] bf16 in (r, s, c) column order, zeros between 147 and the pitch;
 * y [N, Ho, Wo, 64] bf16; dw [64, 147] fp32 accumulated (+= dy^T patches). */
int creamfl_stem_supported(int H, int W);
int creamfl_stem_fprop(const float* images, int N, int H, int W, const void* w_bf16, int64_t w_pitch, void* y_bf16,
                       void* stream);
int creamfl_stem_wgrad(const float* images, const void* dy_bf16, int N, int H, int W, float* dw, void* stream);

/* ---- BatchNorm2d over NHWC bf16 (P = N*H*W pixels), fused with the residual add and ReLU of the ResNet blocks.
 * train: batch statistics (fp64 accumulation in `sums`, 2*C doubles, zero on entry and on exit), running stats
 * updated with `momentum`; keeps mean/rstd for the backward.  scale/shift: C-float scratch each.
 * num_batches_tracked (optional, device int64 scalar) is incremented (nn.BatchNorm2d's buffer). */
int creamfl_bn_train_fwd(const void* x_bf16, int64_t P, int C, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, double* sums, float* mean,
                         float* rstd, float* scale, float* shift, const void* res_bf16, int relu, int stats_ready,
                         int64_t* num_batches_tracked, void* y_bf16, void* stream);
/* stand-alone statistics pass: sums[0..C) += sum_p x, sums[C..2C) += sum_p x^2 */
int creamfl_bn_stats(const void* x_bf16, int64_t P, int C, double* sums, void* stream);
int creamfl_bn_eval_fwd(const void* x_bf16, int64_t P, int C, const float* gamma, const float* beta, float eps,
                        const float* running_mean, const float* running_var, float* scale, float* shift,
                        const void* res_bf16, int relu, void* y_bf16, void* stream);
/* g = dy * gate;  dgamma += sum g*xhat, dbeta += sum g;  dx = BN'(g).  The ReLU gate is (y > 0) when y_bf16 is
 * given (needed when a residual was added before the ReLU), (gamma*xhat + beta > 0) recomputed from x when y_bf16 is
 * null and relu_from_x != 0 (saves reading y), and 1 otherwise.  g_out (optional) receives g (gradient of the
 * residual branch).  coef: 5*C floats scratch. */
int creamfl_bn_train_bwd(const void* dy_bf16, const void* y_bf16, const void* x_bf16, int64_t P, int C,
                         const float* gamma, const float* beta, int relu_from_x, const float* mean, const float* rstd,
                         double* sums, float* coef, float* dgamma, float* dbeta, void* dx_bf16, void* g_out_bf16,
                         void* stream);

/* Same pair with the ReLU gate handed from the forward to the backward pass as one bit per element instead of the
 * output tensor (needed when a residual was added before the ReLU: the gate is not a function of x alone):
 * relu_mask is P*C/8 bytes, bit i of byte t = (y[8t + i] > 0).  The backward pass then reads dy, x and the mask
 * (1/16 of y), which also lets the second read of a 14x14 layer's operands hit L2. */
int creamfl_bn_train_fwd_mask(const void* x_bf16, int64_t P, int C, const float* gamma, const float* beta, float eps,
                              float momentum, float* running_mean, float* running_var, double* sums, float* mean,
                              float* rstd, float* scale, float* shift, const void* res_bf16, int relu, int stats_ready,
                              int64_t* num_batches_tracked, void* y_bf16, void* relu_mask, void* stream);
int creamfl_bn_train_bwd_mask(const void* dy_bf16, const void* relu_mask, const void* x_bf16, int64_t P, int C,
                              const float* gamma, const float* mean, const float* rstd, double* sums, float* coef,
                              float* dgamma, float* dbeta, void* dx_bf16, void* g_out_bf16, void* stream);

/* ---- fused tail of the ResNet stem: conv -> BatchNorm -> ReLU -> maxpool 3x3/2 (torchvision resnet.py via
 * image_encoder.py:24) without ever writing the normalised 112 x 112 map.
 *   forward : creamfl_bn_train_stats (batch statistics of the raw conv output + running-stat update + scale / shift;
 *             creamfl_bn_eval_affine in inference mode), then creamfl_maxpool_affine_fwd:
 *             y = maxpool(relu(scale * x + shift)), idx = winning tap per output element;
 *   backward: creamfl_bn_pool_bwd = creamfl_bn_train_bwd with the ReLU gate recomputed from x, whose incoming gradient
 *             is the max-pooling backward of dy_pooled gathered on the fly (x is the raw conv output [N, H, W, C]). */
int creamfl_bn_train_stats(const void* x_bf16, int64_t P, int C, const float* gamma, const float* beta, float eps,
                           float momentum, float* running_mean, float* running_var, double* sums, float* mean,
                           float* rstd, float* scale, float* shift, int stats_ready, int64_t* num_batches_tracked,
                           void* stream);
int creamfl_bn_eval_affine(int C, const float* gamma, const float* beta, float eps, const float* running_mean,
                           const float* running_var, float* scale, float* shift, void* stream);
int creamfl_maxpool_affine_fwd(const void* x_bf16, const float* scale, const float* shift, int N, int H, int W, int C,
                               void* y_bf16, void* idx_u8, void* stream);
int creamfl_bn_pool_bwd(const void* dy_pooled_bf16, const void* idx_u8, const void* x_bf16, int N, int H, int W, int C,
                        const float* gamma, const float* beta, const float* mean, const float* rstd, double* sums,
                        float* coef, float* dgamma, float* dbeta, void* dx_bf16, void* stream);

/* ---- 3x3 / stride 2 / pad 1 max pooling (ResNet stem); idx: one byte per output element */
int creamfl_maxpool_fwd(const void* x_bf16, int N, int H, int W, int C, void* y_bf16, void* idx_u8, void* stream);
int creamfl_maxpool_bwd(const void* dy_bf16, const void* idx_u8, int N, int H, int W, int C, void* dx_bf16,
                        void* stream);

/* ---- LayerNorm over the last dimension of [R, D] (+ optional residual input), bf16 or fp32 activations */
size_t creamfl_layernorm_bwd_workspace_bytes(int D);
int creamfl_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, float eps, int R,
                          int D, int is_bf16, void* y, float* mean, float* rstd, void* stream);
/* dx_colsum (optional, D floats, accumulated): column sums of dx = bias gradient of the linear layer whose output
 * (plus residual) this LayerNorm normalises - saves a separate pass over dx */
int creamfl_layernorm_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* mean,
                          const float* rstd, int R, int D, int is_bf16, void* dx, float* dgamma, float* dbeta,
                          float* dx_colsum, void* workspace, size_t workspace_bytes, void* stream);

/* ---- out[n] += sum_m x[m, n]  (bias gradients) */
int creamfl_colsum_bf16(const void* x_bf16, int M, int N, int64_t ld, float* out, void* stream);
int creamfl_add_bf16(const void* a_bf16, const void* b_bf16, int64_t n, void* y_bf16, void* stream);

/* ---- BERT embeddings (word + position + token type; LayerNorm is a separate call) */
int creamfl_embed_fwd(const int64_t* ids, const int64_t* token_type, const float* word, const float* pos,
                      const float* type, int T, int L, int D, void* out_bf16, void* stream);
int creamfl_embed_bwd(const int64_t* ids, const int64_t* token_type, const void* dh_bf16, int T, int L, int D,
                      float* dword, float* dpos, float* dtype, void* stream);

/* ---- BERT self-attention, L <= 64, head dim 64: qkv [B*L, 3*H*64] bf16, mask [B, L] fp32 (1 = attend),
 * ctx [B*L, H*64] bf16, probs [B, H, L, L] bf16 */
int creamfl_attn_fwd(const void* qkv_bf16, const float* mask, int B, int L, int H, int head_dim, void* ctx_bf16,
                     void* probs_bf16, void* stream);
/* dbias (optional, 3*H*64 floats, accumulated): column sums of dqkv = bias gradient of the fused q/k/v projection */
int creamfl_attn_bwd(const void* qkv_bf16, const void* probs_bf16, const void* dctx_bf16, int B, int L, int H,
                     int head_dim, void* dqkv_bf16, float* dbias, void* stream);

/* ---- PIENet attention pooling over the P (= 49) positions of the final feature map (pie_model.py:28-40,61-67)
 * and global average pooling (image_encoder.py:56): x [B, P, C] bf16, h = tanh(x W1^T) [B, P, Hd] bf16 */
int creamfl_pie_pool_fwd(const void* x_bf16, const void* h_bf16, const float* w2, int B, int P, int C, int Hd,
                         float* attn, void* r_bf16, void* pooled_bf16, void* stream);
int creamfl_pie_pool_bwd(const void* x_bf16, const void* h_bf16, const float* w2, const float* attn,
                         const void* d_r_bf16, const void* d_pooled_bf16, int B, int P, int C, int Hd, void* dx_bf16,
                         void* dpre_bf16, float* dw2, void* stream);

/* ---- unimodal client heads (src/networks/resnet_client.py:175-201, src/algorithms/ClientTrainer.py:322-363) ------
 * global average pool of an NHWC map times `scale` ([N, P, C] bf16 -> [N, C] fp32 and/or bf16) and its backward */
int creamfl_avgpool_fwd(const void* x_bf16, int N, int P, int C, float scale, float* y, void* y_bf16, void* stream);
int creamfl_avgpool_bwd(const void* dy_bf16, int N, int P, int C, float scale, void* dx_bf16, void* stream);
/* nn.CrossEntropyLoss()(x - margin * onehot(labels), labels): loss_rows[r] (already / R), dlogits = d(loss)/dx,
 * loss (optional) = sum of loss_rows.  labels == NULL means labels[r] = r (the class-centre loss on W W^T). */
int creamfl_ce_fwd(const float* x, int64_t ldx, const int64_t* labels, int R, int C, float margin, float* loss_rows,
                   float* dlogits, float* loss, void* stream);
/* in-place ReLU clamp of a parameter and its bf16 shadow (class_fc weights, resnet_client.py:193-197) */
int creamfl_relu_inplace(float* x, void* shadow_bf16, int64_t n, void* stream);

/* ---- GRU text towers of the clients (src/networks/models/caption_encoder.py:87-116 - the multimodal client's
 * EncoderText; src/networks/language_model.py:93-130 - the unimodal text client).  The dense parts (x W_ih^T, PIENet
 * w_1 / fc, all weight gradients) go through creamfl_gemm_bf16; these entry points are the rest.
 * word embedding: out[t, 0:Dw] = bf16(table[ids[t]]), out[t, Dw:pitch] = 0 (pitch % 8 == 0: TMA row pitch);
 * backward: dtable[ids[t], :] += dx[t, 0:Dw] (dx bf16 with row pitch `pitch`).  Replaces nn.Embedding
 * (caption_encoder.py:41,90; language_model.py:40,96). */
int creamfl_wemb_gather_fwd(const int64_t* ids, const float* table, int T, int V, int Dw, int pitch, void* out_bf16,
                            void* stream);
int creamfl_wemb_scatter_bwd(const int64_t* ids, const void* dx_bf16, int T, int V, int Dw, int pitch, float* dtable,
                             void* stream);
/* recurrent part of nn.GRU(bidirectional=True, batch_first=True) over pack_padded_sequence input
 * (caption_encoder.py:93-97; language_model.py:99-103).  xproj [B, L, 2, 3H] fp32 = x W_ih^T + b_ih of both
 * directions (gate order r, z, n); w_hh [2, 3H, H], b_hh [2, 3H] fp32; lengths int32 [B]; H in {32, 64, 128}.
 * hseq [B, L, 2H] (optional) = pad_packed_sequence output (zeros past each length); hlast [B, 2H] (optional) =
 * hseq[b, len_b - 1, :] (the gather of caption_encoder.py:99-101); gates [B, L, 2, 4, H] (optional) = r, z, n and
 * W_hn h + b_hn kept for the backward.  rev_steps > 0 runs only the first rev_steps steps of the reverse
 * direction (1 is all the hlast consumer needs); 0 = the full sequence. */
int creamfl_gru_fwd(const float* xproj, const float* w_hh, const float* b_hh, const int32_t* lengths, int B, int L,
                    int H, int rev_steps, float* hseq, float* hlast, float* gates, void* stream);
/* backward through time given d(hseq) and / or d(hlast).  Outputs (bf16, zero where no step ran):
 * dxp [B*L, 2, 3H] = gradient at xproj; dgh [B*L, 2, 3H] = gradient at W_hh h + b_hh; hprev [B*L, 2, H] = the state
 * each step started from.  dW_ih = dxp^T x, dW_hh = dgh^T hprev, dx = dxp W_ih follow as creamfl_gemm_bf16 calls,
 * db_ih / db_hh as creamfl_colsum_bf16. */
int creamfl_gru_bwd(const float* gates, const float* hseq, const float* w_hh, const int32_t* lengths,
                    const float* dhseq, const float* dhlast, int B, int L, int H, int rev_steps, void* dxp_bf16,
                    void* dgh_bf16, void* hprev_bf16, void* stream);
/* PIENet attention pooling over the words of a caption with the pad mask (pie_model.py:28-40 with mask):
 * x [B, L, pitch] bf16 (C valid columns), h = tanh(x W1^T) [B, L, hpitch] bf16 (Hd valid columns), w2 [Hd];
 * attn [B, L] fp32 = softmax over p < len_b (0 at pad positions), r [B, pitch] bf16 = sum_p attn[p] x[p, :]. */
int creamfl_seq_pool_fwd(const void* x_bf16, const void* h_bf16, const float* w2, const int32_t* lengths, int B, int L,
                         int C, int pitch, int Hd, int hpitch, float* attn, void* r_bf16, void* stream);
/* dx [B, L, pitch] = attn[p] d_r;  dpre [B, L, hpitch] = gradient at the tanh pre-activation;  dw2 [Hd] accumulated */
int creamfl_seq_pool_bwd(const void* x_bf16, const void* h_bf16, const float* w2, const float* attn,
                         const void* d_r_bf16, const int32_t* lengths, int B, int L, int C, int pitch, int Hd,
                         int hpitch, void* dx_bf16, void* dpre_bf16, float* dw2, void* stream);
/* y = relu(x * scale) (language_model.py:111-112) and dx = dy * scale * (y > 0) */
int creamfl_scale_relu_fwd(const float* x, int64_t n, float scale, float* y, void* stream);
int creamfl_scale_relu_bwd(const float* dy, const float* y, int64_t n, float scale, float* dx, void* stream);

/* ---- BERT dropout (HF BertConfig defaults hidden_dropout_prob = attention_probs_dropout_prob = 0.1, constructed
 * at src/networks/models/pcme.py:31 and run under model.train() at src/algorithms/retrieval_trainer.py:187 and
 * MMFL.py:293; replaces the 37 nn.Dropout / F.dropout calls of one BertModel forward).  Masks are never stored:
 * every kernel regenerates them with Philox4x32-10 from
 *   rng   : device uint64[2] {seed, step} (creamfl_rng_tick advances `step` once per training step - inside a
 *           captured CUDA graph too, so replays draw fresh masks)
 *   site  : which dropout of the forward (0 embeddings, 1 + 3*layer attention probabilities, 2 + 3*layer attention
 *           output dense, 3 + 3*layer FFN output dense); site < 0, rng == NULL or p == 0 disable the dropout
 *   element index e inside the site's tensor; element e is kept iff the 16-bit Philox field (e & 7) of counter
 *   {e >> 3, site, step}, key seed, is >= round(p * 65536); survivors are scaled by 1 / (1 - p).
 * creamfl_dropout_mask writes that keep mask (uint8 [n]) so parity tests can feed identical masks to the oracle. */
int creamfl_rng_tick(void* rng, void* stream);
int creamfl_dropout_mask(const void* rng, int site, int64_t n, float p_drop, void* keep_u8, void* stream);
/* out[M,N] = dropout(A B^T + bias) + add   (element index row * N + column; N % 8 == 0) */
int creamfl_gemm_bf16_drop(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, int M, int N,
                           int K, void* out, int64_t ldo, int out_bf16, const float* bias, const void* add,
                           int64_t ld_add, int add_bf16, const void* rng, int site, float p_drop, void* stream);
/* y = dropout(LayerNorm(x + res))   (HF BertEmbeddings; element index row * D + column) */
int creamfl_layernorm_fwd_drop(const void* x, const void* res, const float* gamma, const float* beta, float eps, int R,
                               int D, int is_bf16, void* y, float* mean, float* rstd, const void* rng, int site,
                               float p_drop, void* stream);
/* creamfl_layernorm_bwd with (a) site_in >= 0: dy is the gradient of dropout(LayerNorm(.)) and is masked on load,
 * (b) site_out >= 0: dx_drop = dropout'(dx) is written next to dx - the gradient of the dense layer whose dropped
 * output (plus residual) this LayerNorm normalises - and dx_colsum sums dx_drop instead of dx */
int creamfl_layernorm_bwd_drop(const void* dy, const void* x, const void* res, const float* gamma, const float* mean,
                               const float* rstd, int R, int D, int is_bf16, void* dx, void* dx_drop, float* dgamma,
                               float* dbeta, float* dx_colsum, void* workspace, size_t workspace_bytes,
                               const void* rng, int site_in, int site_out, float p_drop, void* stream);
/* ctx = dropout(softmax(q k^T / 8 + mask)) v; probs keeps the un-dropped probabilities
 * (element index ((b*H + h)*L + i)*L + j) */
int creamfl_attn_fwd_drop(const void* qkv_bf16, const float* mask, int B, int L, int H, int head_dim, void* ctx_bf16,
                          void* probs_bf16, const void* rng, int site, float p_drop, void* stream);
int creamfl_attn_bwd_drop(const void* qkv_bf16, const void* probs_bf16, const void* dctx_bf16, int B, int L, int H,
                          int head_dim, void* dqkv_bf16, float* dbias, const void* rng, int site, float p_drop,
                          void* stream);

/* ---- fused optimizer step: global-norm clipping + AdamP / Adam / SGD-momentum + bf16 shadow refresh in four
 * launches for any number of tensors.  Replaces adamp.AdamP.step (third-party adamp==0.3.0, call site
 * src/algorithms/optimizers.py:24-28), clip_grad_norm_ (retrieval_trainer.py:211-214) and torch.optim.SGD
 * (ClientTrainer.py:287-288).
 *   rows    : device array of n_rows records {float* p, g, m, v; bf16* shadow; int32 len; int32 tensor} (48 bytes)
 *   tensors : device array of n_tensors records {int32 row_begin, row_end, project, clip; int64 numel} (24 bytes)
 *   hyper   : device float[9]  {lr, beta1|momentum, beta2, eps, weight_decay, delta, wd_ratio, max_norm, mode}
 *             mode 0 AdamP, 1 Adam, 2 SGD with momentum
 *   state   : device float[3]  {step, clip coefficient, gradient norm} (step advances by one per call)
 *   total_gg: device double[1], zero on entry and on exit; stats float[3*n_rows]; flag int32[n_tensors];
 *   tnorm, layer_acc float[n_tensors] - scratch */
int creamfl_optimizer_step(const void* rows, int n_rows, const void* tensors, int n_tensors, const float* hyper,
                           float* state, double* total_gg, float* stats, int32_t* flag, float* tnorm,
                           float* layer_acc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CREAMFL_B200_H */
