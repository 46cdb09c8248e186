// extern "C" surface of the encoder-tower operations (declared in include/creamfl_b200.h): argument checking and
// the choice between the three convolution strategies; no kernel code lives here.
#include "../../include/creamfl_b200.h"
#include "kernels.cuh"
#include <stdlib.h>

using namespace cfl;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

namespace {
struct ConvShape {
  int N, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo;
  long long P_out;
  int kcols;   // R*S*Cin
  int ldc;     // patch-matrix pitch
  bool is_1x1, is_same;
  bool is_strided;   // stride-2 1x1 / 3x3 with padding R/2: implicit GEMM through TMA element strides (fprop, wgrad)
};
bool strided_tma_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CREAMFL_CONV_STRIDED_TMA");      // 0: explicit patch matrix + GEMM (A/B measurements)
    on = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return on != 0;
}
ConvShape shape_of(int N, int H, int W, int Cin, int Cout, int R, int S_, int stride, int pad) {
  ConvShape c{N, H, W, Cin, Cout, R, S_, stride, pad, 0, 0, 0, 0, 0, false, false, false};
  c.Ho = (H + 2 * pad - R) / stride + 1;
  c.Wo = (W + 2 * pad - S_) / stride + 1;
  c.P_out = (long long)N * c.Ho * c.Wo;
  c.kcols = R * S_ * Cin;
  c.ldc = (int)round_up(c.kcols, 8);
  c.is_1x1 = (R == 1 && S_ == 1 && stride == 1 && pad == 0);
  c.is_same = (!c.is_1x1 && stride == 1 && (R & 1) && (S_ & 1) && pad == R / 2 && pad == S_ / 2 && Cin % 64 == 0 &&
               Cout % 64 == 0);
  c.is_strided = (stride == 2 && R == S_ && (R == 1 || R == 3) && pad == R / 2 && Cin % 64 == 0 && Cout % 64 == 0 &&
                  strided_tma_enabled());
  return c;
}
int check_shape(const char* who, const ConvShape& c) {
  if (c.N <= 0 || c.H <= 0 || c.W <= 0 || c.Cin <= 0 || c.Cout <= 0 || c.R <= 0 || c.S <= 0 || c.stride <= 0 ||
      c.pad < 0 || c.Ho <= 0 || c.Wo <= 0) {
    set_error("%s: bad convolution shape", who);
    return CFL_EINVAL;
  }
  if (c.Cin % 8 || c.Cout % 8) {
    set_error("%s: channel counts must be multiples of 8 (Cin=%d Cout=%d)", who, c.Cin, c.Cout);
    return CFL_EINVAL;
  }
  return CFL_OK;
}
int split_for(long long M, long long N, long long K) { return gemm_plan_split((int)M, (int)N, (int)K); }
}  // namespace

extern "C" {

size_t creamfl_conv2d_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S_, int stride, int pad) {
  const ConvShape c = shape_of(N, H, W, Cin, Cout, R, S_, stride, pad);
  if (c.is_1x1 || c.is_same || c.P_out <= 0) return 0;
  return (size_t)c.P_out * c.ldc * 2;
}

int creamfl_conv2d_fprop(const void* x, const void* w, int N, int H, int W, int Cin, int Cout, int R, int S_,
                         int stride, int pad, int64_t w_pitch, void* y, double* bn_sums, void* ws, size_t ws_bytes,
                         void* stream) {
  const ConvShape c = shape_of(N, H, W, Cin, Cout, R, S_, stride, pad);
  int rc = check_shape("conv2d_fprop", c);
  if (rc) return rc;
  if (!x || !w || !y) {
    set_error("conv2d_fprop: null pointer");
    return CFL_EINVAL;
  }
  if ((c.is_same || c.is_strided) && w_pitch == c.kcols) {
    if ((rc = conv_same_fprop(x, w, N, H, W, Cin, Cout, R, S_, y, S(stream), nullptr, nullptr, 0, stride))) return rc;
    return bn_sums ? bn_stats_only(y, c.P_out, Cout, bn_sums, S(stream)) : CFL_OK;
  }
  GemmParams p{};
  p.M = (int)c.P_out; p.N = Cout; p.split_k = 1;
  p.out = y; p.ldo = Cout; p.out_bf16 = 1; p.alpha = 1.0f;
  p.stats = bn_sums;
  if (c.is_1x1) {
    p.K = Cin;
    return gemm_bf16(x, Cin, 0, w, w_pitch, 0, p, S(stream));
  }
  const size_t need = (size_t)c.P_out * c.ldc * 2;
  if (!ws || ws_bytes < need) {
    set_error("conv2d_fprop: workspace %zu B < %zu B", ws_bytes, need);
    return CFL_EWORKSPACE;
  }
  if (w_pitch < c.ldc) {
    set_error("conv2d_fprop: filter pitch %lld < padded patch width %d", (long long)w_pitch, c.ldc);
    return CFL_EINVAL;
  }
  if ((rc = im2col_nhwc(x, N, H, W, Cin, R, S_, stride, pad, c.ldc, ws, S(stream)))) return rc;
  p.K = c.ldc;
  return gemm_bf16(ws, c.ldc, 0, w, w_pitch, 0, p, S(stream));
}

int creamfl_conv2d_fprop_affine(const void* x, const void* w, int N, int H, int W, int Cin, int Cout, int R, int S_,
                                int stride, int pad, int64_t w_pitch, const float* bias, const void* add, int relu,
                                void* y, void* ws, size_t ws_bytes, void* stream) {
  const ConvShape c = shape_of(N, H, W, Cin, Cout, R, S_, stride, pad);
  int rc = check_shape("conv2d_fprop_affine", c);
  if (rc) return rc;
  if (!x || !w || !y) {
    set_error("conv2d_fprop_affine: null pointer");
    return CFL_EINVAL;
  }
  if ((c.is_same || c.is_strided) && w_pitch == c.kcols)
    return conv_same_fprop(x, w, N, H, W, Cin, Cout, R, S_, y, S(stream), bias, add, relu, stride);
  GemmParams p{};
  p.M = (int)c.P_out; p.N = Cout; p.split_k = 1;
  p.out = y; p.ldo = Cout; p.out_bf16 = 1; p.alpha = 1.0f;
  p.bias = bias;
  p.add = add; p.ld_add = Cout; p.add_bf16 = 1;
  p.act = relu ? 2 : 0;
  if (c.is_1x1) {
    p.K = Cin;
    return gemm_bf16(x, Cin, 0, w, w_pitch, 0, p, S(stream));
  }
  const size_t need = (size_t)c.P_out * c.ldc * 2;
  if (!ws || ws_bytes < need) {
    set_error("conv2d_fprop_affine: workspace %zu B < %zu B", ws_bytes, need);
    return CFL_EWORKSPACE;
  }
  if (w_pitch < c.ldc) {
    set_error("conv2d_fprop_affine: filter pitch %lld < padded patch width %d", (long long)w_pitch, c.ldc);
    return CFL_EINVAL;
  }
  if ((rc = im2col_nhwc(x, N, H, W, Cin, R, S_, stride, pad, c.ldc, ws, S(stream)))) return rc;
  p.K = c.ldc;
  return gemm_bf16(ws, c.ldc, 0, w, w_pitch, 0, p, S(stream));
}

int creamfl_bn_fold_layers(const int64_t* layers, const int64_t* row_start, int n_layers, int64_t total_rows, float eps,
                           void* stream) {
  if (!layers || !row_start || n_layers <= 0 || total_rows <= 0) {
    set_error("bn_fold_layers: null pointer / empty table");
    return CFL_EINVAL;
  }
  return bn_fold_layers(reinterpret_cast<const long long*>(layers), reinterpret_cast<const long long*>(row_start),
                        n_layers, total_rows, eps, S(stream));
}

int creamfl_conv2d_dgrad(const void* dy, const void* w, int N, int H, int W, int Cin, int Cout, int R, int S_,
                         int stride, int pad, int64_t w_pitch, const void* add, void* dx, void* ws, size_t ws_bytes,
                         void* stream) {
  const ConvShape c = shape_of(N, H, W, Cin, Cout, R, S_, stride, pad);
  int rc = check_shape("conv2d_dgrad", c);
  if (rc) return rc;
  if (!dy || !w || !dx) {
    set_error("conv2d_dgrad: null pointer");
    return CFL_EINVAL;
  }
  if (c.is_same && w_pitch == c.kcols) return conv_same_dgrad(dy, w, N, H, W, Cin, Cout, R, S_, dx, add, S(stream));
  if (c.is_strided && R == 3 && !(H & 1) && !(W & 1) && w_pitch == c.kcols)
    return conv_same_dgrad(dy, w, N, H, W, Cin, Cout, R, S_, dx, add, S(stream), stride);
  GemmParams p{};
  p.M = (int)c.P_out; p.K = Cout; p.split_k = 1; p.alpha = 1.0f; p.out_bf16 = 1;
  if (c.is_1x1) {
    p.N = Cin; p.out = dx; p.ldo = Cin;
    p.add = add; p.ld_add = Cin; p.add_bf16 = 1;
    return gemm_bf16(dy, Cout, 0, w, w_pitch, 1, p, S(stream));
  }
  const size_t need = (size_t)c.P_out * c.ldc * 2;
  if (!ws || ws_bytes < need) {
    set_error("conv2d_dgrad: workspace %zu B < %zu B", ws_bytes, need);
    return CFL_EWORKSPACE;
  }
  p.N = c.kcols; p.out = ws; p.ldo = c.ldc;
  if ((rc = gemm_bf16(dy, Cout, 0, w, w_pitch, 1, p, S(stream)))) return rc;
  return col2im_nhwc(ws, N, H, W, Cin, R, S_, stride, pad, c.ldc, add, dx, S(stream));
}

int creamfl_conv2d_wgrad(const void* dy, const void* x, const void* col, int N, int H, int W, int Cin, int Cout, int R,
                         int S_, int stride, int pad, float* dw, void* ws, size_t ws_bytes, void* stream) {
  const ConvShape c = shape_of(N, H, W, Cin, Cout, R, S_, stride, pad);
  int rc = check_shape("conv2d_wgrad", c);
  if (rc) return rc;
  if (!dy || !dw || (!x && !col)) {
    set_error("conv2d_wgrad: null pointer");
    return CFL_EINVAL;
  }
  if ((c.is_same || c.is_strided) && !col) return conv_same_wgrad(dy, x, N, H, W, Cin, Cout, R, S_, dw, S(stream), stride);
  GemmParams p{};
  p.M = Cout; p.K = (int)c.P_out; p.alpha = 1.0f; p.out = dw; p.ldo = c.kcols; p.out_bf16 = 0; p.atomic_out = 1;
  if (c.is_1x1 && !col) {
    p.N = Cin;
    p.split_k = split_for(p.M, p.N, p.K);
    return gemm_bf16(dy, Cout, 1, x, Cin, 1, p, S(stream));
  }
  const void* patches = col;
  if (!patches) {
    const size_t need = (size_t)c.P_out * c.ldc * 2;
    if (!ws || ws_bytes < need) {
      set_error("conv2d_wgrad: workspace %zu B < %zu B", ws_bytes, need);
      return CFL_EWORKSPACE;
    }
    if ((rc = im2col_nhwc(x, N, H, W, Cin, R, S_, stride, pad, c.ldc, ws, S(stream)))) return rc;
    patches = ws;
  }
  p.N = c.kcols;
  p.split_k = split_for(p.M, p.N, p.K);
  return gemm_bf16(dy, Cout, 1, patches, c.ldc, 1, p, S(stream));
}

int creamfl_im2col_nchw_f32(const float* images, int N, int C, int H, int W, int R, int S_, int stride, int pad,
                            int col_pitch, void* col, void* stream) {
  if (!images || !col || N <= 0 || C <= 0 || (col_pitch & 7)) {
    set_error("im2col_nchw_f32: bad argument (pitch must be a multiple of 8)");
    return CFL_EINVAL;
  }
  return im2col_nchw_f32(images, N, C, H, W, R, S_, stride, pad, col_pitch, col, S(stream));
}

int creamfl_bn_train_fwd(const void* x, int64_t P, int C, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, double* sums, float* mean,
                         float* rstd, float* scale, float* shift, const void* res, int relu, int stats_ready,
                         int64_t* num_batches_tracked, void* y, void* stream) {
  if (!x || !gamma || !beta || !sums || !mean || !rstd || !scale || !shift || !y) {
    set_error("bn_train_fwd: null pointer");
    return CFL_EINVAL;
  }
  return bn_train_fwd(x, P, C, gamma, beta, eps, momentum, running_mean, running_var, sums, mean, rstd, scale, shift,
                      res, relu, stats_ready, reinterpret_cast<long long*>(num_batches_tracked), y, S(stream));
}

int creamfl_bn_train_fwd_mask(const void* x, int64_t P, int C, const float* gamma, const float* beta, float eps,
                              float momentum, float* running_mean, float* running_var, double* sums, float* mean,
                              float* rstd, float* scale, float* shift, const void* res, int relu, int stats_ready,
                              int64_t* num_batches_tracked, void* y, void* relu_mask, void* stream) {
  if (!x || !gamma || !beta || !sums || !mean || !rstd || !scale || !shift || !y || !relu_mask) {
    set_error("bn_train_fwd_mask: null pointer");
    return CFL_EINVAL;
  }
  return bn_train_fwd(x, P, C, gamma, beta, eps, momentum, running_mean, running_var, sums, mean, rstd, scale, shift,
                      res, relu, stats_ready, reinterpret_cast<long long*>(num_batches_tracked), y, S(stream), relu_mask);
}

int creamfl_bn_train_bwd_mask(const void* dy, const void* relu_mask, const void* x, int64_t P, int C, const float* gamma,
                              const float* mean, const float* rstd, double* sums, float* coef, float* dgamma,
                              float* dbeta, void* dx, void* g_out, void* stream) {
  if (!dy || !relu_mask || !x || !gamma || !mean || !rstd || !sums || !coef || !dx) {
    set_error("bn_train_bwd_mask: null pointer");
    return CFL_EINVAL;
  }
  return bn_train_bwd(dy, nullptr, x, P, C, gamma, nullptr, 0, mean, rstd, sums, coef, dgamma, dbeta, dx, g_out,
                      S(stream), relu_mask);
}

int creamfl_bn_stats(const void* x, int64_t P, int C, double* sums, void* stream) {
  if (!x || !sums) {
    set_error("bn_stats: null pointer");
    return CFL_EINVAL;
  }
  return bn_stats_only(x, P, C, sums, S(stream));
}

int creamfl_bn_eval_fwd(const void* x, int64_t P, int C, const float* gamma, const float* beta, float eps,
                        const float* running_mean, const float* running_var, float* scale, float* shift,
                        const void* res, int relu, void* y, void* stream) {
  if (!x || !gamma || !beta || !running_mean || !running_var || !scale || !shift || !y) {
    set_error("bn_eval_fwd: null pointer");
    return CFL_EINVAL;
  }
  return bn_eval_fwd(x, P, C, gamma, beta, eps, running_mean, running_var, scale, shift, res, relu, y, S(stream));
}

int creamfl_bn_train_bwd(const void* dy, const void* y, const void* x, int64_t P, int C, const float* gamma,
                         const float* beta, int relu_from_x, const float* mean, const float* rstd, double* sums,
                         float* coef, float* dgamma, float* dbeta, void* dx, void* g_out, void* stream) {
  if (!dy || !x || !gamma || !mean || !rstd || !sums || !coef || !dx) {
    set_error("bn_train_bwd: null pointer");
    return CFL_EINVAL;
  }
  return bn_train_bwd(dy, y, x, P, C, gamma, beta, relu_from_x, mean, rstd, sums, coef, dgamma, dbeta, dx, g_out,
                      S(stream));
}

int creamfl_bn_train_stats(const void* x, int64_t P, int C, const float* gamma, const float* beta, float eps,
                           float momentum, float* running_mean, float* running_var, double* sums, float* mean,
                           float* rstd, float* scale, float* shift, int stats_ready, int64_t* num_batches_tracked,
                           void* stream) {
  if (!x || !gamma || !beta || !sums || !mean || !rstd || !scale || !shift) {
    set_error("bn_train_stats: null pointer");
    return CFL_EINVAL;
  }
  return bn_train_fwd(x, P, C, gamma, beta, eps, momentum, running_mean, running_var, sums, mean, rstd, scale, shift,
                      nullptr, 0, stats_ready, reinterpret_cast<long long*>(num_batches_tracked), nullptr, S(stream));
}

int creamfl_bn_eval_affine(int C, const float* gamma, const float* beta, float eps, const float* running_mean,
                           const float* running_var, float* scale, float* shift, void* stream) {
  if (!gamma || !beta || !running_mean || !running_var || !scale || !shift) {
    set_error("bn_eval_affine: null pointer");
    return CFL_EINVAL;
  }
  return bn_eval_fwd(nullptr, 1, C, gamma, beta, eps, running_mean, running_var, scale, shift, nullptr, 0, nullptr,
                     S(stream));
}

int creamfl_maxpool_affine_fwd(const void* x, const float* scale, const float* shift, int N, int H, int W, int C,
                               void* y, void* idx, void* stream) {
  if (!x || !y || !scale || !shift) {
    set_error("maxpool_affine_fwd: null pointer");
    return CFL_EINVAL;
  }
  return maxpool_fwd(x, N, H, W, C, y, idx, S(stream), scale, shift);
}

int creamfl_bn_pool_bwd(const void* dy_pooled, const void* idx, const void* x, int N, int H, int W, int C,
                        const float* gamma, const float* beta, const float* mean, const float* rstd, double* sums,
                        float* coef, float* dgamma, float* dbeta, void* dx, void* stream) {
  if (!dy_pooled || !idx || !x || !gamma || !beta || !mean || !rstd || !sums || !coef || !dx) {
    set_error("bn_pool_bwd: null pointer");
    return CFL_EINVAL;
  }
  return bn_train_bwd(dy_pooled, nullptr, x, (long long)N * H * W, C, gamma, beta, 1, mean, rstd, sums, coef, dgamma,
                      dbeta, dx, nullptr, S(stream), nullptr, idx, H, W);
}

int creamfl_maxpool_fwd(const void* x, int N, int H, int W, int C, void* y, void* idx, void* stream) {
  if (!x || !y) {
    set_error("maxpool_fwd: null pointer");
    return CFL_EINVAL;
  }
  return maxpool_fwd(x, N, H, W, C, y, idx, S(stream));
}

int creamfl_maxpool_bwd(const void* dy, const void* idx, int N, int H, int W, int C, void* dx, void* stream) {
  if (!dy || !idx || !dx) {
    set_error("maxpool_bwd: null pointer");
    return CFL_EINVAL;
  }
  return maxpool_bwd(dy, idx, N, H, W, C, dx, S(stream));
}

size_t creamfl_layernorm_bwd_workspace_bytes(int D) { return D > 0 ? layernorm_bwd_workspace_bytes(D) : 0; }

int creamfl_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, float eps, int R,
                          int D, int is_bf16, void* y, float* mean, float* rstd, void* stream) {
  if (!x || !gamma || !beta || !y) {
    set_error("layernorm_fwd: null pointer");
    return CFL_EINVAL;
  }
  return layernorm_fwd(x, res, gamma, beta, eps, R, D, is_bf16, y, mean, rstd, S(stream));
}

int creamfl_layernorm_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* mean,
                          const float* rstd, int R, int D, int is_bf16, void* dx, float* dgamma, float* dbeta,
                          float* dx_colsum, void* ws, size_t ws_bytes, void* stream) {
  if (!dy || !x || !gamma || !mean || !rstd || !dx || !ws) {
    set_error("layernorm_bwd: null pointer");
    return CFL_EINVAL;
  }
  return layernorm_bwd(dy, x, res, gamma, mean, rstd, R, D, is_bf16, dx, dgamma, dbeta, dx_colsum, ws, ws_bytes,
                       S(stream));
}

int creamfl_colsum_bf16(const void* x, int M, int N, int64_t ld, float* out, void* stream) {
  if (!x || !out) {
    set_error("colsum_bf16: null pointer");
    return CFL_EINVAL;
  }
  return colsum_bf16(x, M, N, ld, out, S(stream));
}

int creamfl_add_bf16(const void* a, const void* b, int64_t n, void* y, void* stream) {
  if (!a || !b || !y) {
    set_error("add_bf16: null pointer");
    return CFL_EINVAL;
  }
  return add_bf16(a, b, n, y, S(stream));
}

int creamfl_embed_fwd(const int64_t* ids, const int64_t* tt, const float* word, const float* pos, const float* type,
                      int T, int L, int D, void* out, void* stream) {
  if (!ids || !word || !pos || !type || !out) {
    set_error("embed_fwd: null pointer");
    return CFL_EINVAL;
  }
  return embed_fwd(reinterpret_cast<const long long*>(ids), reinterpret_cast<const long long*>(tt), word, pos, type, T,
                   L, D, out, S(stream));
}

int creamfl_embed_bwd(const int64_t* ids, const int64_t* tt, const void* dh, int T, int L, int D, float* dword,
                      float* dpos, float* dtype, void* stream) {
  if (!ids || !dh || !dword || !dpos || !dtype) {
    set_error("embed_bwd: null pointer");
    return CFL_EINVAL;
  }
  return embed_bwd(reinterpret_cast<const long long*>(ids), reinterpret_cast<const long long*>(tt), dh, T, L, D, dword,
                   dpos, dtype, S(stream));
}

int creamfl_attn_fwd(const void* qkv, const float* mask, int B, int L, int H, int head_dim, void* ctx, void* probs,
                     void* stream) {
  if (!qkv || !mask || !ctx || !probs) {
    set_error("attn_fwd: null pointer");
    return CFL_EINVAL;
  }
  return attn_fwd(qkv, mask, B, L, H, head_dim, ctx, probs, S(stream));
}

int creamfl_attn_bwd(const void* qkv, const void* probs, const void* dctx, int B, int L, int H, int head_dim,
                     void* dqkv, float* dbias, void* stream) {
  if (!qkv || !probs || !dctx || !dqkv) {
    set_error("attn_bwd: null pointer");
    return CFL_EINVAL;
  }
  return attn_bwd(qkv, probs, dctx, B, L, H, head_dim, dqkv, dbias, S(stream));
}

int creamfl_pie_pool_fwd(const void* x, const void* h, const float* w2, int B, int P, int C, int Hd, float* attn,
                         void* r, void* pooled, void* stream) {
  if (!x || !h || !w2 || !attn || !r || !pooled) {
    set_error("pie_pool_fwd: null pointer");
    return CFL_EINVAL;
  }
  return pie_pool_fwd(x, h, w2, B, P, C, Hd, attn, r, pooled, S(stream));
}

int creamfl_pie_pool_bwd(const void* x, const void* h, const float* w2, const float* attn, const void* d_r,
                         const void* d_pooled, int B, int P, int C, int Hd, void* dx, void* dpre, float* dw2,
                         void* stream) {
  if (!x || !h || !w2 || !attn || !d_r || !d_pooled || !dx || !dpre || !dw2) {
    set_error("pie_pool_bwd: null pointer");
    return CFL_EINVAL;
  }
  return pie_pool_bwd(x, h, w2, attn, d_r, d_pooled, B, P, C, Hd, dx, dpre, dw2, S(stream));
}

int creamfl_avgpool_fwd(const void* x, int N, int P, int C, float scale, float* y, void* y16, void* stream) {
  if (!x || (!y && !y16)) {
    set_error("avgpool_fwd: null pointer");
    return CFL_EINVAL;
  }
  return avgpool_fwd(x, N, P, C, scale, y, y16, S(stream));
}

int creamfl_avgpool_bwd(const void* dy, int N, int P, int C, float scale, void* dx, void* stream) {
  if (!dy || !dx) {
    set_error("avgpool_bwd: null pointer");
    return CFL_EINVAL;
  }
  return avgpool_bwd(dy, N, P, C, scale, dx, S(stream));
}

int creamfl_ce_fwd(const float* x, int64_t ldx, const int64_t* labels, int R, int C, float margin, float* loss_rows,
                   float* dlogits, float* loss, void* stream) {
  if (!x || !loss_rows || !dlogits) {
    set_error("ce_fwd: null pointer");
    return CFL_EINVAL;
  }
  return ce_fwd(x, ldx, reinterpret_cast<const long long*>(labels), R, C, margin, loss_rows, dlogits, loss, S(stream));
}

int creamfl_relu_inplace(float* x, void* shadow, int64_t n, void* stream) {
  if (!x) {
    set_error("relu_inplace: null pointer");
    return CFL_EINVAL;
  }
  return relu_inplace(x, shadow, n, S(stream));
}

int creamfl_optimizer_step(const void* rows, int n_rows, const void* tensors, int n_tensors, const float* hyper,
                           float* state, double* total_gg, float* stats, int32_t* flag, float* tnorm,
                           float* layer_acc, void* stream) {
  if (!rows || !tensors || !hyper || !state || !total_gg || !stats || !flag || !tnorm || !layer_acc) {
    set_error("optimizer_step: null pointer");
    return CFL_EINVAL;
  }
  return optimizer_step(rows, n_rows, tensors, n_tensors, hyper, state, total_gg, stats, flag, tnorm, layer_acc,
                        S(stream));
}

// ---- GRU text towers (text_ops.cu)
int creamfl_wemb_gather_fwd(const int64_t* ids, const float* table, int T, int V, int Dw, int pitch, void* out_bf16,
                            void* stream) {
  if (!ids || !table || !out_bf16) {
    set_error("wemb_gather_fwd: null pointer");
    return CFL_EINVAL;
  }
  return wemb_gather_fwd(reinterpret_cast<const long long*>(ids), table, T, V, Dw, pitch, out_bf16, S(stream));
}

int creamfl_wemb_scatter_bwd(const int64_t* ids, const void* dx_bf16, int T, int V, int Dw, int pitch, float* dtable,
                             void* stream) {
  const void* dx = dx_bf16;
  if (!ids || !dx || !dtable) {
    set_error("wemb_scatter_bwd: null pointer");
    return CFL_EINVAL;
  }
  return wemb_scatter_bwd(reinterpret_cast<const long long*>(ids), dx, T, V, Dw, pitch, dtable, S(stream));
}

int creamfl_gru_fwd(const float* xproj, const float* w_hh, const float* b_hh, const int32_t* lengths, int B, int L,
                    int H, int rev_steps, float* hseq, float* hlast, float* gates, void* stream) {
  if (!xproj || !w_hh || !b_hh || !lengths || (!hseq && !hlast)) {
    set_error("gru_fwd: null pointer");
    return CFL_EINVAL;
  }
  return gru_fwd(xproj, w_hh, b_hh, lengths, B, L, H, rev_steps, hseq, hlast, gates, S(stream));
}

int creamfl_gru_bwd(const float* gates, const float* hseq, const float* w_hh, const int32_t* lengths,
                    const float* dhseq, const float* dhlast, int B, int L, int H, int rev_steps, void* dxp_bf16,
                    void* dgh_bf16, void* hprev_bf16, void* stream) {
  if (!gates || !hseq || !w_hh || !lengths || (!dhseq && !dhlast) || !dxp_bf16 || !dgh_bf16 || !hprev_bf16) {
    set_error("gru_bwd: null pointer");
    return CFL_EINVAL;
  }
  return gru_bwd(gates, hseq, w_hh, lengths, dhseq, dhlast, B, L, H, rev_steps, dxp_bf16, dgh_bf16, hprev_bf16,
                 S(stream));
}

int creamfl_seq_pool_fwd(const void* x, const void* h, const float* w2, const int32_t* lengths, int B, int L, int C,
                         int pitch, int Hd, int hpitch, float* attn, void* r_bf16, void* stream) {
  if (!x || !h || !w2 || !lengths || !attn || !r_bf16) {
    set_error("seq_pool_fwd: null pointer");
    return CFL_EINVAL;
  }
  return seq_pool_fwd(x, h, w2, lengths, B, L, C, pitch, Hd, hpitch, attn, r_bf16, S(stream));
}

int creamfl_seq_pool_bwd(const void* x, const void* h, const float* w2, const float* attn, const void* d_r,
                         const int32_t* lengths, int B, int L, int C, int pitch, int Hd, int hpitch, void* dx_bf16,
                         void* dpre_bf16, float* dw2, void* stream) {
  if (!x || !h || !w2 || !attn || !d_r || !lengths || !dx_bf16 || !dpre_bf16 || !dw2) {
    set_error("seq_pool_bwd: null pointer");
    return CFL_EINVAL;
  }
  return seq_pool_bwd(x, h, w2, attn, d_r, lengths, B, L, C, pitch, Hd, hpitch, dx_bf16, dpre_bf16, dw2, S(stream));
}

int creamfl_scale_relu_fwd(const float* x, int64_t n, float scale, float* y, void* stream) {
  if (!x || !y) {
    set_error("scale_relu_fwd: null pointer");
    return CFL_EINVAL;
  }
  return scale_relu_fwd(x, n, scale, y, S(stream));
}

int creamfl_scale_relu_bwd(const float* dy, const float* y, int64_t n, float scale, float* dx, void* stream) {
  if (!dy || !y || !dx) {
    set_error("scale_relu_bwd: null pointer");
    return CFL_EINVAL;
  }
  return scale_relu_bwd(dy, y, n, scale, dx, S(stream));
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- dropout variants
static inline DropSpec make_drop(const void* rng, int site, float p) {
  DropSpec d{};
  if (rng != nullptr && site >= 0 && p > 0.0f) {
    d.rng = reinterpret_cast<const unsigned long long*>(rng);
    d.site = site;
    d.thresh = drop_thresh16(p);
    d.scale = 1.0f / (1.0f - p);
  }
  return d;
}
static int check_p(const char* who, float p) {
  if (!(p >= 0.0f && p < 1.0f)) {
    set_error("%s: dropout probability %g outside [0, 1)", who, (double)p);
    return CFL_EINVAL;
  }
  return CFL_OK;
}

extern "C" int creamfl_gemm_bf16_drop(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, int M,
                                      int N, int K, void* out, int64_t ldo, int out_bf16, const float* bias,
                                      const void* add, int64_t ld_add, int add_bf16, const void* rng, int site,
                                      float p_drop, void* stream) {
  if (!a || !b || !out) {
    set_error("gemm_bf16_drop: null pointer");
    return CFL_EINVAL;
  }
  int rc = check_p("gemm_bf16_drop", p_drop);
  if (rc) return rc;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.split_k = 1;
  p.out = out; p.ldo = ldo; p.out_bf16 = out_bf16;
  p.bias = bias; p.act = 0; p.alpha = 1.0f;
  p.add = add; p.ld_add = ld_add; p.add_bf16 = add_bf16;
  p.drop = make_drop(rng, site, p_drop);
  return gemm_bf16(a, lda, a_mn, b, ldb, b_mn, p, S(stream));
}

extern "C" int creamfl_layernorm_fwd_drop(const void* x, const void* res, const float* gamma, const float* beta,
                                          float eps, int R, int D, int is_bf16, void* y, float* mean, float* rstd,
                                          const void* rng, int site, float p_drop, void* stream) {
  if (!x || !gamma || !beta || !y) {
    set_error("layernorm_fwd_drop: null pointer");
    return CFL_EINVAL;
  }
  int rc = check_p("layernorm_fwd_drop", p_drop);
  if (rc) return rc;
  return layernorm_fwd_drop(x, res, gamma, beta, eps, R, D, is_bf16, y, mean, rstd, make_drop(rng, site, p_drop),
                            S(stream));
}

extern "C" int creamfl_layernorm_bwd_drop(const void* dy, const void* x, const void* res, const float* gamma,
                                          const float* mean, const float* rstd, int R, int D, int is_bf16, void* dx,
                                          void* dx_drop, float* dgamma, float* dbeta, float* dx_colsum, void* ws,
                                          size_t ws_bytes, const void* rng, int site_in, int site_out, float p_drop,
                                          void* stream) {
  if (!dy || !x || !gamma || !mean || !rstd || !dx || !ws) {
    set_error("layernorm_bwd_drop: null pointer");
    return CFL_EINVAL;
  }
  int rc = check_p("layernorm_bwd_drop", p_drop);
  if (rc) return rc;
  return layernorm_bwd_drop(dy, x, res, gamma, mean, rstd, R, D, is_bf16, dx, dx_drop, dgamma, dbeta, dx_colsum, ws,
                            ws_bytes, make_drop(rng, site_in, p_drop), make_drop(rng, site_out, p_drop), S(stream));
}

extern "C" int creamfl_attn_fwd_drop(const void* qkv, const float* mask, int B, int L, int H, int head_dim, void* ctx,
                                     void* probs, const void* rng, int site, float p_drop, void* stream) {
  if (!qkv || !mask || !ctx || !probs) {
    set_error("attn_fwd_drop: null pointer");
    return CFL_EINVAL;
  }
  int rc = check_p("attn_fwd_drop", p_drop);
  if (rc) return rc;
  return attn_fwd_drop(qkv, mask, B, L, H, head_dim, ctx, probs, make_drop(rng, site, p_drop), S(stream));
}

extern "C" int creamfl_attn_bwd_drop(const void* qkv, const void* probs, const void* dctx, int B, int L, int H,
                                     int head_dim, void* dqkv, float* dbias, const void* rng, int site, float p_drop,
                                     void* stream) {
  if (!qkv || !probs || !dctx || !dqkv) {
    set_error("attn_bwd_drop: null pointer");
    return CFL_EINVAL;
  }
  int rc = check_p("attn_bwd_drop", p_drop);
  if (rc) return rc;
  return attn_bwd_drop(qkv, probs, dctx, B, L, H, head_dim, dqkv, dbias, make_drop(rng, site, p_drop), S(stream));
}

extern "C" int creamfl_dropout_mask(const void* rng, int site, int64_t n, float p_drop, void* keep_u8, void* stream) {
  int rc = check_p("dropout_mask", p_drop);
  if (rc) return rc;
  if (p_drop <= 0.0f || site < 0) {
    set_error("dropout_mask: needs p > 0 and site >= 0");
    return CFL_EINVAL;
  }
  return dropout_mask(make_drop(rng, site, p_drop), n, keep_u8, S(stream));
}

extern "C" int creamfl_rng_tick(void* rng, void* stream) {
  return rng_tick(reinterpret_cast<unsigned long long*>(rng), S(stream));
}

// ---------------------------------------------------------------------------------------------- fused ResNet stem
extern "C" int creamfl_stem_supported(int H, int W) { return stem_supported(3, H, W, 7, 7, 2, 3, 64) ? 1 : 0; }

extern "C" int creamfl_stem_fprop(const float* images, int N, int H, int W, const void* w_bf16, int64_t w_pitch,
                                  void* y_bf16, void* stream) {
  if (!images || !w_bf16 || !y_bf16) {
    set_error("stem_fprop: null pointer");
    return CFL_EINVAL;
  }
  return stem_fprop(images, w_bf16, w_pitch, N, H, W, y_bf16, S(stream));
}

extern "C" int creamfl_stem_wgrad(const float* images, const void* dy_bf16, int N, int H, int W, float* dw,
                                  void* stream) {
  if (!images || !dy_bf16 || !dw) {
    set_error("stem_wgrad: null pointer");
    return CFL_EINVAL;
  }
  return stem_wgrad(images, dy_bf16, N, H, W, dw, S(stream));
}
