// extern "C" surface of libcreamfl_b200.so (declared in include/creamfl_b200.h).  Argument checking and the
// composition of kernels into the operations the reference performs; no kernel code lives here.
#include "../../include/creamfl_b200.h"
#include "kernels.cuh"



using namespace cfl;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

extern "C" {

int creamfl_abi_version(void) { return 1; }

int creamfl_gemm_bf16(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, int M, int N,
                      int K, void* out, int64_t ldo, int out_bf16, void* out_preact_bf16, const float* bias,
                      int act, float alpha, const void* add, int64_t ld_add, int add_bf16, const void* aux_bf16,
                      int64_t ld_aux, int split_k, int accumulate, void* stream) {
  if (!a || !b || !out) {
    set_error("gemm_bf16: null pointer");
    return CFL_EINVAL;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  if (split_k == 0) split_k = accumulate ? gemm_plan_split(M, N, K) : 1;
  p.split_k = split_k;
  p.atomic_out = accumulate ? 1 : 0;
  p.out = out; p.ldo = ldo; p.out_bf16 = out_bf16;
  p.out2 = out_preact_bf16;
  p.bias = bias; p.act = act; p.alpha = alpha;
  p.add = add; p.ld_add = ld_add; p.add_bf16 = add_bf16;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(aux_bf16); p.ld_aux = ld_aux;
  return gemm_bf16(a, lda, a_mn, b, ldb, b_mn, p, S(stream));
}

int creamfl_plan_split_k(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 1;
  return gemm_plan_split(M, N, K);
}

size_t creamfl_rowlse_workspace_bytes(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return rowlse_workspace_bytes(M, N);
}

int creamfl_infonce_fwd(const void* q, const void* g, const int64_t* labels, int B, int N, int D, float inv_tau,
                        float* loss, float* row_score, float* lse2, void* ws, size_t ws_bytes, void* stream) {
  if (!labels || !row_score) {
    set_error("infonce_fwd: labels and row_score are required");
    return CFL_EINVAL;
  }
  int rc = rowlse_bf16(q, g, reinterpret_cast<const long long*>(labels), B, N, D, inv_tau, row_score, lse2, ws,
                       ws_bytes, S(stream));
  if (rc) return rc;
  if (loss) {
    sum_finish_kernel<<<1, 256, 0, S(stream)>>>(row_score, B, -1.0f / (float)B, loss);
    return check_launch("infonce_fwd/mean");
  }
  return CFL_OK;
}

size_t creamfl_infonce_bwd_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (size_t)B * (size_t)round_up(N, 64) * 2;
}

int creamfl_infonce_bwd(const void* q, const void* g, const int64_t* labels, const float* lse2, int B, int N,
                        int D, float inv_tau, const float* gout, float* dq, void* ws, size_t ws_bytes,
                        void* stream) {
  if (!lse2 || !labels || !gout || !dq) {
    set_error("infonce_bwd: null pointer");
    return CFL_EINVAL;
  }
  const long long ldp = round_up(N, 64);
  const size_t need = (size_t)B * (size_t)ldp * 2;
  if (!ws || ws_bytes < need) {
    set_error("infonce_bwd: workspace %zu B < %zu B", ws_bytes, need);
    return CFL_EWORKSPACE;
  }
  int rc = softmax_emit_bf16(q, g, reinterpret_cast<const long long*>(labels), lse2, B, N, D, inv_tau, ws, ldp,
                             S(stream));
  if (rc) return rc;
  cudaMemsetAsync(dq, 0, (size_t)B * D * sizeof(float), S(stream));
  GemmParams p{};
  p.M = B; p.N = D; p.K = N;
  p.out = dq; p.ldo = D; p.out_bf16 = 0;
  p.alpha = inv_tau / (float)B;
  // enough K splits to put every SM to work on the [B, D] output
  const int tiles = ((B + 127) / 128) * ((D + 127) / 128);
  int split = sm_count() / (tiles > 0 ? tiles : 1);
  if (split < 1) split = 1;
  p.split_k = split;
  rc = gemm_bf16(ws, ldp, 0, g, D, 1, p, S(stream));
  if (rc) return rc;
  return scale_by_scalar(dq, (long long)B * D, gout, 1.0f, S(stream));
}

int creamfl_conw_score(const void* v, const void* g, int N, int D, float* score, void* ws, size_t ws_bytes,
                       void* stream) {
  if (!score) {
    set_error("conw_score: null output");
    return CFL_EINVAL;
  }
  return rowlse_bf16(v, g, nullptr, N, N, D, 1.0f, score, nullptr, ws, ws_bytes, S(stream));
}

int creamfl_conw_reduce(const float* const* vecs_host, const float* scores, int C, int N, int D, float* out,
                        float* weights, void* stream) {
  if (!vecs_host || !scores || !out) {
    set_error("conw_reduce: null pointer");
    return CFL_EINVAL;
  }
  return conw_reduce(vecs_host, scores, C, N, D, out, weights, S(stream));
}

size_t creamfl_pcme_workspace_bytes(int N) {
  if (N <= 0) return 0;
  const size_t g = (N + 15) / 16;
  const size_t a = g * g * 2 * sizeof(float), b = (size_t)N * 2 * sizeof(float);
  return a > b ? a : b;
}

int creamfl_pcme_fwd(const float* img, const float* txt, int N, int D, const float* shift,
                     const float* neg_scale, float* dist, float* out3, void* ws, size_t ws_bytes, void* stream) {
  if (!img || !txt || !shift || !neg_scale || !dist || !out3 || !ws) {
    set_error("pcme_fwd: null pointer");
    return CFL_EINVAL;
  }
  return pcme_fwd(img, txt, N, D, shift, neg_scale, dist, out3, reinterpret_cast<float*>(ws), ws_bytes,
                  S(stream));
}

int creamfl_pcme_bwd(const float* img, const float* txt, const float* dist, int N, int D, const float* shift,
                     const float* neg_scale, const float* gout, float* d_img, float* d_txt, float* d_shift,
                     float* d_neg_scale, void* ws, size_t ws_bytes, void* stream) {
  if (!img || !txt || !dist || !gout || !d_img || !d_txt || !d_shift || !d_neg_scale || !ws) {
    set_error("pcme_bwd: null pointer");
    return CFL_EINVAL;
  }
  if (N <= 0 || D <= 0) {
    set_error("pcme_bwd: empty batch");
    return CFL_EINVAL;
  }
  return pcme_bwd(img, txt, dist, N, D, shift, neg_scale, gout, d_img, d_txt, d_shift, d_neg_scale,
                  reinterpret_cast<float*>(ws), ws_bytes, S(stream));
}

int creamfl_moon_fwd(const float* z, const float* zold, const float* bank, const int64_t* idx, int R, int D,
                     float inv_tau, float denom, float* loss_rows, float* coef, float* loss, void* stream) {
  if (!z || !zold || !bank || !idx || !loss_rows || !coef) {
    set_error("moon_fwd: null pointer");
    return CFL_EINVAL;
  }
  return moon_fwd(z, zold, bank, reinterpret_cast<const long long*>(idx), R, D, inv_tau, denom, loss_rows, coef,
                  loss, S(stream));
}

int creamfl_moon_bwd(const float* zold, const float* bank, const int64_t* idx, const float* coef,
                     const float* gout, int R, int D, float* dz, void* stream) {
  if (!zold || !bank || !idx || !coef || !gout || !dz || R <= 0) {
    set_error("moon_bwd: bad argument");
    return CFL_EINVAL;
  }
  return moon_bwd(zold, bank, reinterpret_cast<const long long*>(idx), coef, gout, R, D, dz, S(stream));
}

size_t creamfl_mse_workspace_bytes(void) { return 256 * sizeof(float); }

int creamfl_mse_gather_fwd(const float* x, const float* bank, const int64_t* idx, int R, int D, float* loss,
                           void* ws, size_t ws_bytes, void* stream) {
  if (!x || !bank || !idx || !loss || !ws) {
    set_error("mse_gather_fwd: null pointer");
    return CFL_EINVAL;
  }
  return mse_gather_fwd(x, bank, reinterpret_cast<const long long*>(idx), R, D, loss,
                        reinterpret_cast<float*>(ws), ws_bytes, S(stream));
}

int creamfl_mse_gather_bwd(const float* x, const float* bank, const int64_t* idx, const float* gout, int R,
                           int D, float* dx, void* stream) {
  if (!x || !bank || !idx || !gout || !dx || R <= 0) {
    set_error("mse_gather_bwd: bad argument");
    return CFL_EINVAL;
  }
  return mse_gather_bwd(x, bank, reinterpret_cast<const long long*>(idx), gout, R, D, dx, S(stream));
}

int creamfl_l2norm_fwd(const float* x, int R, int D, float* y, void* y_bf16, float* inv_norm, void* stream) {
  if (!x) {
    set_error("l2norm_fwd: null input");
    return CFL_EINVAL;
  }
  return l2norm_fwd(x, R, D, y, y_bf16, inv_norm, S(stream));
}

int creamfl_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, int R, int D, float* dx,
                       void* stream) {
  if (!dy || !y || !inv_norm || !dx || R <= 0) {
    set_error("l2norm_bwd: bad argument");
    return CFL_EINVAL;
  }
  return l2norm_bwd(dy, y, inv_norm, R, D, dx, S(stream));
}

int creamfl_act_bwd_f32(const float* dy, const float* y, int64_t n, int kind, void* out, void* stream) {
  if (!dy || !y || !out || n <= 0 || (kind != CREAMFL_ACT_SIGMOID && kind != CREAMFL_ACT_TANH)) {
    set_error("act_bwd_f32: bad argument");
    return CFL_EINVAL;
  }
  return act_bwd_f32(dy, y, n, kind, out, S(stream));
}

int creamfl_cast_f32_bf16(const float* x, int64_t n, void* y, void* stream) {
  if (n > 0 && (!x || !y)) {
    set_error("cast_f32_bf16: null pointer");
    return CFL_EINVAL;
  }
  return cast_f32_bf16(x, n, y, S(stream));
}

size_t creamfl_recall_workspace_bytes(int Nq) { return Nq > 0 ? (size_t)Nq * sizeof(int) : 0; }

int creamfl_recall_ranks(const float* q, const float* g, const int64_t* q_lab, const int64_t* g_lab, int Nq,
                         int Ng, int D, int32_t* ranks, void* ws, size_t ws_bytes, void* stream) {
  if (!q || !g || !q_lab || !g_lab || !ranks || !ws) {
    set_error("recall_ranks: null pointer");
    return CFL_EINVAL;
  }
  return recall_ranks(q, g, reinterpret_cast<const long long*>(q_lab), reinterpret_cast<const long long*>(g_lab),
                      Nq, Ng, D, ranks, ws, ws_bytes, S(stream));
}

}  // extern "C"
