// Small fp32 kernels of the loss side of the hot path.  They are launch/latency bound (N = batch size), so
// they stay on CUDA cores in exact fp32 - which also keeps them bit-stable run to run (fixed reduction order,
// no floating-point atomics).
//
//   pcme_*   : MCSoftContrastiveLoss fwd/bwd               (reference src/criterions/probemb.py:7-45,48-86,185-256)
//   moon_*   : intra-modal 2-way contrast vs. the old model (MMClientTrainer.py:169-191, ClientTrainer.py:404-414)
//   mse_*    : distillation MSE against aggregated rows     (MMFL.py:296,355-378)
//   l2norm_* : F.normalize(p=2, dim=-1)                     (utils/tensor_utils.py:25-27)
#include "common.cuh"

namespace cfl {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum in a fixed order; result valid in thread 0.
template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* scratch /* kThreads/32 floats */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float r = 0.0f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) r += scratch[i];
  }
  __syncthreads();
  return r;
}

__device__ __forceinline__ float softplus_f(float x) {
  // log(1 + exp(x)), stable on both tails
  return fmaxf(x, 0.0f) + log1pf(__expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// --------------------------------------------------------------------------------------------- PCME loss
// dist[i,j] = sqrt(sum_k (img[i,k]-txt[j,k])^2 + 1e-6);  l = -neg_scale*dist + shift;
// per-direction loss = sum_ij softplus(-2*m_ij*l_ij), m_ii=+1 else -1; total = i2t + t2i = 2x that sum.
constexpr int kPT = 16;  // pair tile edge

__global__ void __launch_bounds__(kPT* kPT)
pcme_fwd_kernel(const float* __restrict__ img, const float* __restrict__ txt, int N, int D,
                const float* __restrict__ shift, const float* __restrict__ neg_scale, float* __restrict__ dist,
                float* __restrict__ block_part /* [gridDim.x*gridDim.y, 2] */) {
  extern __shared__ float sm[];
  float* sa = sm;                 // [kPT][D+1]
  float* sb = sm + kPT * (D + 1); // [kPT][D+1]
  __shared__ float red[kPT * kPT / 32];
  const int ti = threadIdx.x / kPT, tj = threadIdx.x % kPT;
  const int i0 = blockIdx.y * kPT, j0 = blockIdx.x * kPT;
  for (int e = threadIdx.x; e < kPT * D; e += kPT * kPT) {
    const int r = e / D, k = e % D;
    sa[r * (D + 1) + k] = (i0 + r < N) ? img[(long long)(i0 + r) * D + k] : 0.0f;
    sb[r * (D + 1) + k] = (j0 + r < N) ? txt[(long long)(j0 + r) * D + k] : 0.0f;
  }
  __syncthreads();
  const int i = i0 + ti, j = j0 + tj;
  float acc = 0.0f;
  const float* a = sa + ti * (D + 1);
  const float* b = sb + tj * (D + 1);
  for (int k = 0; k < D; ++k) {
    const float d = a[k] - b[k];
    acc = fmaf(d, d, acc);
  }
  float pos = 0.0f, neg = 0.0f;
  if (i < N && j < N) {
    const float d = sqrtf(acc + 1e-6f);
    dist[(long long)i * N + j] = d;
    const float l = -neg_scale[0] * d + shift[0];
    if (i == j) pos = softplus_f(-2.0f * l); else neg = softplus_f(2.0f * l);
  }
  const float bp = block_sum<kPT * kPT>(pos, red);
  const float bn = block_sum<kPT * kPT>(neg, red);
  if (threadIdx.x == 0) {
    const int b_id = blockIdx.y * gridDim.x + blockIdx.x;
    block_part[2 * b_id] = bp;
    block_part[2 * b_id + 1] = bn;
  }
}

// out[0] = total loss (i2t + t2i), out[1] = per-direction positive part, out[2] = per-direction negative part
__global__ void pcme_finish_kernel(const float* __restrict__ block_part, int nblocks, float* __restrict__ out) {
  __shared__ float red[8];
  float p = 0.0f, n = 0.0f;
  for (int b = threadIdx.x; b < nblocks; b += 256) {
    p += block_part[2 * b];
    n += block_part[2 * b + 1];
  }
  const float ps = block_sum<256>(p, red);
  const float ns = block_sum<256>(n, red);
  if (threadIdx.x == 0) {
    out[0] = 2.0f * (ps + ns);
    out[1] = ps;
    out[2] = ns;
  }
}

// Backward.  Blocks [0,N): row i of d_img; blocks [N,2N): row j of d_txt.  Per-block partials of d_shift and
// d_neg_scale come from the image-row blocks only (each pair counted once) and are reduced by the finish kernel.
__global__ void __launch_bounds__(256)
pcme_bwd_kernel(const float* __restrict__ img, const float* __restrict__ txt, const float* __restrict__ dist,
                int N, int D, const float* __restrict__ shift, const float* __restrict__ neg_scale,
                const float* __restrict__ gout, float* __restrict__ d_img, float* __restrict__ d_txt,
                float* __restrict__ param_part /* [N, 2] */) {
  extern __shared__ float w[];  // [N] pair weights for this row/column
  __shared__ float red[8];
  const bool is_img = blockIdx.x < N;
  const int r = is_img ? blockIdx.x : blockIdx.x - N;
  const float s = neg_scale[0], b = shift[0], g = gout[0];
  float dsh = 0.0f, dsc = 0.0f;
  for (int o = threadIdx.x; o < N; o += 256) {
    const int i = is_img ? r : o, j = is_img ? o : r;
    const float d = dist[(long long)i * N + j];
    const float l = -s * d + b;
    const float m = (i == j) ? 1.0f : -1.0f;
    // d/dl [2 * softplus(-2 m l)] = -4 m sigmoid(-2 m l)
    const float gl = g * (-4.0f * m) * sigmoid_f(-2.0f * m * l);
    w[o] = gl * (-s) / d;
    dsh += gl;
    dsc += gl * (-d);
  }
  __syncthreads();
  const float* self = (is_img ? img : txt) + (long long)r * D;
  const float* other = is_img ? txt : img;
  float* dst = (is_img ? d_img : d_txt) + (long long)r * D;
  for (int k = threadIdx.x; k < D; k += 256) {
    const float x = self[k];
    float acc = 0.0f;
    for (int o = 0; o < N; ++o) acc = fmaf(w[o], x - other[(long long)o * D + k], acc);
    dst[k] = acc;
  }
  const float t_sh = block_sum<256>(dsh, red);
  const float t_sc = block_sum<256>(dsc, red);
  if (is_img && threadIdx.x == 0) {
    param_part[2 * r] = t_sh;
    param_part[2 * r + 1] = t_sc;
  }
}

__global__ void pcme_bwd_finish_kernel(const float* __restrict__ param_part, int N, float* __restrict__ d_shift,
                                       float* __restrict__ d_neg_scale) {
  __shared__ float red[8];
  float a = 0.0f, c = 0.0f;
  for (int i = threadIdx.x; i < N; i += 256) {
    a += param_part[2 * i];
    c += param_part[2 * i + 1];
  }
  const float sa = block_sum<256>(a, red);
  const float sc = block_sum<256>(c, red);
  if (threadIdx.x == 0) {
    d_shift[0] = sa;
    d_neg_scale[0] = sc;
  }
}

// --------------------------------------------------------------------------------------------- MOON intra
// Row r: pos = <z_r, bank[idx_r]>, neg = <z_r, zold_r>; CE([pos,neg]*inv_tau, 0) = softplus((neg-pos)*inv_tau).
// loss_rows[r] = CE_r / denom ; coef[r] = d(loss)/d(neg-pos) = inv_tau*sigmoid((neg-pos)*inv_tau)/denom.
__global__ void moon_fwd_kernel(const float* __restrict__ z, const float* __restrict__ zold,
                                const float* __restrict__ bank, const long long* __restrict__ idx, int R, int D,
                                float inv_tau, float denom, float* __restrict__ loss_rows,
                                float* __restrict__ coef) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* zr = z + (long long)r * D;
  const float* orow = zold + (long long)r * D;
  const float* prow = bank + idx[r] * D;
  float pos = 0.0f, neg = 0.0f;
  for (int k = lane; k < D; k += 32) {
    const float x = zr[k];
    pos = fmaf(x, prow[k], pos);
    neg = fmaf(x, orow[k], neg);
  }
  pos = warp_sum(pos);
  neg = warp_sum(neg);
  if (lane == 0) {
    const float t = (neg - pos) * inv_tau;
    loss_rows[r] = softplus_f(t) / denom;
    coef[r] = inv_tau * sigmoid_f(t) / denom;
  }
}

__global__ void moon_bwd_kernel(const float* __restrict__ zold, const float* __restrict__ bank,
                                const long long* __restrict__ idx, const float* __restrict__ coef,
                                const float* __restrict__ gout, int R, int D, float* __restrict__ dz) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)R * D) return;
  const int r = (int)(e / D), k = (int)(e % D);
  dz[e] = gout[0] * coef[r] * (zold[e] - bank[idx[r] * D + k]);
}

// --------------------------------------------------------------------------------------------- softmax cross-entropy
// Row r: logits z = x[r, :] - margin * onehot(label_r)   (ClientTrainer.py:346-350: "fvec - inter_distance * one_hot"),
// loss_r = logsumexp(z) - z[label_r]; out: loss_rows[r] = loss_r / R and dlogits[r, :] = (softmax(z) - onehot) / R
// (nn.CrossEntropyLoss mean reduction).  One warp per row, C <= 4096.
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const float* __restrict__ x, long long ldx, const long long* __restrict__ labels, int R, int C,
              float margin, float* __restrict__ loss_rows, float* __restrict__ dlogits) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const long long lab = labels ? labels[r] : (long long)r;
  const float* xr = x + (long long)r * ldx;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, xr[c] - (c == lab ? margin : 0.0f));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.0f;
  for (int c = lane; c < C; c += 32) s += __expf(xr[c] - (c == lab ? margin : 0.0f) - m);
  s = warp_sum(s);
  const float lse = m + __logf(s);
  const float inv_r = 1.0f / (float)R;
  for (int c = lane; c < C; c += 32) {
    const float z = xr[c] - (c == lab ? margin : 0.0f);
    dlogits[(long long)r * C + c] = (__expf(z - lse) - (c == lab ? 1.0f : 0.0f)) * inv_r;
  }
  if (lane == 0) loss_rows[r] = (lse - (xr[lab] - margin)) * inv_r;
}

// --------------------------------------------------------------------------------------------- distill MSE
// loss = mean_{r,k} (x[r,k] - bank[idx_r,k])^2 ; two-stage fixed-order reduction.
__global__ void __launch_bounds__(256)
mse_gather_fwd_kernel(const float* __restrict__ x, const float* __restrict__ bank,
                      const long long* __restrict__ idx, int R, int D, float* __restrict__ block_part) {
  __shared__ float red[8];
  float acc = 0.0f;
  const long long total = (long long)R * D;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int r = (int)(e / D), k = (int)(e % D);
    const float d = x[e] - bank[idx[r] * D + k];
    acc = fmaf(d, d, acc);
  }
  const float s = block_sum<256>(acc, red);
  if (threadIdx.x == 0) block_part[blockIdx.x] = s;
}

// out[0] = scale * sum(in[0..n))   (single block, fixed order)
__global__ void sum_finish_kernel(const float* __restrict__ in, int n, float scale, float* __restrict__ out) {
  __shared__ float red[8];
  float a = 0.0f;
  for (int i = threadIdx.x; i < n; i += 256) a += in[i];
  const float s = block_sum<256>(a, red);
  if (threadIdx.x == 0) out[0] = s * scale;
}

__global__ void mse_gather_bwd_kernel(const float* __restrict__ x, const float* __restrict__ bank,
                                      const long long* __restrict__ idx, const float* __restrict__ gout, int R,
                                      int D, float* __restrict__ dx) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)R * D;
  if (e >= total) return;
  const int r = (int)(e / D), k = (int)(e % D);
  dx[e] = gout[0] * 2.0f / (float)total * (x[e] - bank[idx[r] * D + k]);
}

// --------------------------------------------------------------------------------------------- L2 normalise
// y = x / max(||x||, 1e-12)   (F.normalize semantics); also emits a bf16 copy for the tensor-core kernels.
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, int R, int D, float* __restrict__ y,
                                  __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ inv_norm) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* xr = x + (long long)r * D;
  float ss = 0.0f;
  for (int k = lane; k < D; k += 32) ss = fmaf(xr[k], xr[k], ss);
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int k = lane; k < D; k += 32) {
    const float v = xr[k] * inv;
    if (y) y[(long long)r * D + k] = v;
    if (y_bf16) y_bf16[(long long)r * D + k] = __float2bfloat16(v);
  }
  if (lane == 0 && inv_norm) inv_norm[r] = inv;
}

// dx = (dy - y * <dy, y>) * inv_norm
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                  const float* __restrict__ inv_norm, int R, int D, float* __restrict__ dx) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* dr = dy + (long long)r * D;
  const float* yr = y + (long long)r * D;
  float dot = 0.0f;
  for (int k = lane; k < D; k += 32) dot = fmaf(dr[k], yr[k], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[r];
  for (int k = lane; k < D; k += 32) dx[(long long)r * D + k] = (dr[k] - yr[k] * dot) * inv;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ y) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(y + i) = pk;
  } else {
    for (long long j = i; j < n; ++j) y[j] = __float2bfloat16(x[j]);
  }
}

// out = dy * f'(y) from the activation OUTPUT: kind 6 sigmoid, 3 tanh
__global__ void act_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, int kind,
                                   __nv_bfloat16* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = y[i];
  out[i] = __float2bfloat16(dy[i] * (kind == 6 ? v * (1.0f - v) : 1.0f - v * v));
}

// dQ init for the InfoNCE backward is implicit (P already holds softmax - onehot); this scales fp32 rows.
__global__ void scale_by_scalar_kernel(float* __restrict__ x, long long n, const float* __restrict__ s,
                                       float c) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= s[0] * c;
}

// ============================================================================================ host side
int pcme_fwd(const float* img, const float* txt, int N, int D, const float* shift, const float* neg_scale,
             float* dist, float* out3, float* workspace, size_t ws_bytes, cudaStream_t st) {
  if (N <= 0 || D <= 0) {
    set_error("pcme_fwd: empty batch N=%d D=%d", N, D);
    return CFL_EINVAL;
  }
  const int g = (N + kPT - 1) / kPT;
  const size_t need = (size_t)g * g * 2 * sizeof(float);
  if (ws_bytes < need) {
    set_error("pcme_fwd: workspace %zu < %zu", ws_bytes, need);
    return CFL_EWORKSPACE;
  }
  const size_t smem = 2 * kPT * (D + 1) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("pcme_fwd: D=%d too large", D);
    return CFL_EINVAL;
  }
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(pcme_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pcme_fwd_kernel<<<dim3(g, g), kPT * kPT, smem, st>>>(img, txt, N, D, shift, neg_scale, dist, workspace);
  pcme_finish_kernel<<<1, 256, 0, st>>>(workspace, g * g, out3);
  return check_launch("pcme_fwd");
}

int pcme_bwd(const float* img, const float* txt, const float* dist, int N, int D, const float* shift,
             const float* neg_scale, const float* gout, float* d_img, float* d_txt, float* d_shift,
             float* d_neg_scale, float* workspace, size_t ws_bytes, cudaStream_t st) {
  const size_t need = (size_t)N * 2 * sizeof(float);
  if (ws_bytes < need) {
    set_error("pcme_bwd: workspace %zu < %zu", ws_bytes, need);
    return CFL_EWORKSPACE;
  }
  const size_t smem = (size_t)N * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("pcme_bwd: batch %d too large", N);
    return CFL_EINVAL;
  }
  pcme_bwd_kernel<<<2 * N, 256, smem, st>>>(img, txt, dist, N, D, shift, neg_scale, gout, d_img, d_txt,
                                            workspace);
  pcme_bwd_finish_kernel<<<1, 256, 0, st>>>(workspace, N, d_shift, d_neg_scale);
  return check_launch("pcme_bwd");
}

int moon_fwd(const float* z, const float* zold, const float* bank, const long long* idx, int R, int D,
             float inv_tau, float denom, float* loss_rows, float* coef, float* loss_out, cudaStream_t st) {
  if (R <= 0) {
    set_error("moon_fwd: empty batch");
    return CFL_EINVAL;
  }
  moon_fwd_kernel<<<(R + 7) / 8, 256, 0, st>>>(z, zold, bank, idx, R, D, inv_tau, denom, loss_rows, coef);
  if (loss_out) sum_finish_kernel<<<1, 256, 0, st>>>(loss_rows, R, 1.0f, loss_out);
  return check_launch("moon_fwd");
}

int moon_bwd(const float* zold, const float* bank, const long long* idx, const float* coef, const float* gout,
             int R, int D, float* dz, cudaStream_t st) {
  const long long n = (long long)R * D;
  moon_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(zold, bank, idx, coef, gout, R, D, dz);
  return check_launch("moon_bwd");
}

int ce_fwd(const float* x, long long ldx, const long long* labels, int R, int C, float margin, float* loss_rows,
           float* dlogits, float* loss_out, cudaStream_t st) {
  if (R <= 0 || C <= 0) {
    set_error("ce_fwd: empty problem");
    return CFL_EINVAL;
  }
  ce_fwd_kernel<<<(R + 7) / 8, 256, 0, st>>>(x, ldx, labels, R, C, margin, loss_rows, dlogits);
  if (loss_out) sum_finish_kernel<<<1, 256, 0, st>>>(loss_rows, R, 1.0f, loss_out);
  return check_launch("ce_fwd");
}

// x = max(x, 0) in place on the fp32 master and its bf16 shadow (resnet_client.py:193-197 clamps class_fc weights)
__global__ void relu_inplace_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ shadow, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = fmaxf(x[i], 0.0f);
  x[i] = v;
  if (shadow) shadow[i] = __float2bfloat16(v);
}

int relu_inplace(float* x, void* shadow, long long n, cudaStream_t st) {
  if (n <= 0) return CFL_OK;
  relu_inplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(shadow), n);
  return check_launch("relu_inplace");
}

int mse_gather_fwd(const float* x, const float* bank, const long long* idx, int R, int D, float* loss_out,
                   float* workspace, size_t ws_bytes, cudaStream_t st) {
  if (R <= 0) {
    set_error("mse_gather_fwd: empty batch");
    return CFL_EINVAL;
  }
  const long long n = (long long)R * D;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 256) blocks = 256;
  if (ws_bytes < blocks * sizeof(float)) {
    set_error("mse_gather_fwd: workspace too small");
    return CFL_EWORKSPACE;
  }
  mse_gather_fwd_kernel<<<blocks, 256, 0, st>>>(x, bank, idx, R, D, workspace);
  sum_finish_kernel<<<1, 256, 0, st>>>(workspace, blocks, 1.0f / (float)n, loss_out);
  return check_launch("mse_gather_fwd");
}

int mse_gather_bwd(const float* x, const float* bank, const long long* idx, const float* gout, int R, int D,
                   float* dx, cudaStream_t st) {
  const long long n = (long long)R * D;
  mse_gather_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, bank, idx, gout, R, D, dx);
  return check_launch("mse_gather_bwd");
}

int l2norm_fwd(const float* x, int R, int D, float* y, void* y_bf16, float* inv_norm, cudaStream_t st) {
  if (R <= 0) {
    set_error("l2norm_fwd: empty");
    return CFL_EINVAL;
  }
  l2norm_fwd_kernel<<<(R + 7) / 8, 256, 0, st>>>(x, R, D, y, reinterpret_cast<__nv_bfloat16*>(y_bf16), inv_norm);
  return check_launch("l2norm_fwd");
}

int l2norm_bwd(const float* dy, const float* y, const float* inv_norm, int R, int D, float* dx, cudaStream_t st) {
  l2norm_bwd_kernel<<<(R + 7) / 8, 256, 0, st>>>(dy, y, inv_norm, R, D, dx);
  return check_launch("l2norm_bwd");
}

int cast_f32_bf16(const float* x, long long n, void* y, cudaStream_t st) {
  if (n <= 0) return CFL_OK;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 7)) {
    set_error("cast_f32_bf16: misaligned");
    return CFL_EINVAL;
  }
  const long long thr = (n + 3) / 4;
  cast_f32_bf16_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, st>>>(x, n, reinterpret_cast<__nv_bfloat16*>(y));
  return check_launch("cast_f32_bf16");
}

int act_bwd_f32(const float* dy, const float* y, long long n, int kind, void* out, cudaStream_t st) {
  act_bwd_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dy, y, n, kind, reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("act_bwd_f32");
}

int scale_by_scalar(float* x, long long n, const float* s, float c, cudaStream_t st) {
  if (n <= 0) return CFL_OK;
  scale_by_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, s, c);
  return check_launch("scale_by_scalar");
}

}  // namespace cfl
