// Fused optimizer step over a table of parameter rows: global-norm gradient clipping + AdamP (or plain Adam / SGD
// with momentum) + refresh of the bf16 parameter shadow, in four launches regardless of the number of tensors.
//
// Replaces, for the server and the multimodal clients, the per-tensor Python loop of torch elementwise kernels in
// the third-party `adamp==0.3.0` package (call site src/algorithms/optimizers.py:24-28) plus
// nn.utils.clip_grad_norm_ (retrieval_trainer.py:211-214, MMClientTrainer.py:133-135,212-214); for the unimodal
// clients torch.optim.SGD(momentum 0.9, weight decay 5e-5) (ClientTrainer.py:287-288).
//
// AdamP (Heo et al., ICLR 2021; algorithm restated from the paper / package, whose source is not in the reference
// tree - parity of this kernel is pinned against the restatement in oracle/creamfl_oracle.py only):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; u = m / (sqrt(v)/sqrt(1-b2^t) + eps)
//   for tensors with more than one dimension: if max_rows |cos(g_row, p_row)| < delta / sqrt(row length)
//       u_row -= p^_row <p^_row, u_row>,  p^ = p / (|p_row| + eps)         ("channel" view)
//   else the same test and projection with the whole tensor as one row        ("layer" view)
//   p *= 1 - lr * wd * (wd_ratio if projected else 1) ; p -= lr / (1-b1^t) * u
//
// A "row" of the table is one output channel of a >1-D tensor, or a chunk of a 1-D tensor (no projection).
#include "kernels.cuh"

namespace cfl {

struct OptRow {
  float* p;
  float* g;
  float* m;
  float* v;
  __nv_bfloat16* shadow;  // may be null
  int len;
  int tensor;             // index into the per-tensor arrays
};

struct OptTensor {
  int row_begin, row_end;
  int project;   // 1: tensor has > 1 dimension (AdamP projection applies)
  int clip;      // 1: takes part in the global gradient norm
  long long numel;
};

// hyper[]: 0 lr, 1 beta1, 2 beta2, 3 eps, 4 weight_decay, 5 delta, 6 wd_ratio, 7 max_norm (<= 0: no clipping),
//          8 mode (0 AdamP, 1 Adam, 2 SGD-momentum with hyper[1] = momentum)
// state[]: 0 step (float), 1 clip coefficient, 2 total grad norm (for logging)
constexpr int kOptThreads = 128;

__device__ __forceinline__ float block_sum_128(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// rows whose length is a multiple of 4 and whose buffers are 16-byte aligned (every convolution / linear row of the
// towers) take the float4 path; the 7x7x3 stem rows and ragged heads stay scalar
__device__ __forceinline__ bool row_is_vec4(const OptRow& row, bool with_moments) {
  uintptr_t a = reinterpret_cast<uintptr_t>(row.p) | reinterpret_cast<uintptr_t>(row.g);
  if (with_moments) a |= reinterpret_cast<uintptr_t>(row.m) | reinterpret_cast<uintptr_t>(row.v);
  return (row.len & 3) == 0 && (a & 15) == 0 && (reinterpret_cast<uintptr_t>(row.shadow) & 7) == 0;
}
__device__ __forceinline__ void store_shadow4(__nv_bfloat16* dst, const float4& p) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(dst) = pk;
}

// K1: per-row <g,p>, |g|^2, |p|^2 ; global |g|^2 (fp64 atomic)
__global__ void __launch_bounds__(kOptThreads)
opt_row_stats_kernel(const OptRow* __restrict__ rows, const OptTensor* __restrict__ tensors, int n_rows,
                     float* __restrict__ stats /* [n_rows, 3] */, double* __restrict__ total_gg) {
  __shared__ float red[4];
  const int r = blockIdx.x;
  if (r >= n_rows) return;
  const OptRow row = rows[r];
  float dot = 0.f, gg = 0.f, pp = 0.f;
  if (row_is_vec4(row, false)) {   // 16-byte loads: four times the bytes in flight per thread (the pass is HBM-bound)
    const float4* g4 = reinterpret_cast<const float4*>(row.g);
    const float4* p4 = reinterpret_cast<const float4*>(row.p);
    for (int i = threadIdx.x; i < (row.len >> 2); i += kOptThreads) {
      const float4 g = g4[i], p = p4[i];
      dot = fmaf(g.x, p.x, dot); dot = fmaf(g.y, p.y, dot); dot = fmaf(g.z, p.z, dot); dot = fmaf(g.w, p.w, dot);
      gg = fmaf(g.x, g.x, gg); gg = fmaf(g.y, g.y, gg); gg = fmaf(g.z, g.z, gg); gg = fmaf(g.w, g.w, gg);
      pp = fmaf(p.x, p.x, pp); pp = fmaf(p.y, p.y, pp); pp = fmaf(p.z, p.z, pp); pp = fmaf(p.w, p.w, pp);
    }
  } else {
    for (int i = threadIdx.x; i < row.len; i += kOptThreads) {
      const float g = row.g[i], p = row.p[i];
      dot = fmaf(g, p, dot);
      gg = fmaf(g, g, gg);
      pp = fmaf(p, p, pp);
    }
  }
  dot = block_sum_128(dot, red);
  gg = block_sum_128(gg, red);
  pp = block_sum_128(pp, red);
  if (threadIdx.x == 0) {
    stats[3 * r] = dot;
    stats[3 * r + 1] = gg;
    stats[3 * r + 2] = pp;
    if (tensors[row.tensor].clip) atomicAdd(total_gg, (double)gg);
  }
}

// K2: per-tensor projection decision; block 0 also advances the step counter and derives the clip coefficient.
// flag: 0 none, 1 channel view, 2 layer view.  tnorm[t] = |p| of the whole tensor (layer view).
__global__ void __launch_bounds__(256)
opt_decide_kernel(const OptTensor* __restrict__ tensors, int n_tensors, const float* __restrict__ stats,
                  const float* __restrict__ hyper, float* __restrict__ state, double* __restrict__ total_gg,
                  int* __restrict__ flag, float* __restrict__ tnorm, float* __restrict__ layer_acc) {
  __shared__ float smax[8];
  __shared__ float sdot[8], sgg[8], spp[8];
  const int t = blockIdx.x;
  if (t >= n_tensors) return;
  const OptTensor tt = tensors[t];
  const float eps = hyper[3], delta = hyper[5];
  float mx = 0.f, dot = 0.f, gg = 0.f, pp = 0.f;
  for (int r = tt.row_begin + threadIdx.x; r < tt.row_end; r += 256) {
    const float d = stats[3 * r], g2 = stats[3 * r + 1], p2 = stats[3 * r + 2];
    mx = fmaxf(mx, fabsf(d) / (sqrtf(g2) + eps) / (sqrtf(p2) + eps));
    dot += d; gg += g2; pp += p2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    gg += __shfl_xor_sync(0xffffffffu, gg, o);
    pp += __shfl_xor_sync(0xffffffffu, pp, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smax[w] = mx; sdot[w] = dot; sgg[w] = gg; spp[w] = pp; }
  __syncthreads();
  if (threadIdx.x == 0) {
    mx = 0.f; dot = 0.f; gg = 0.f; pp = 0.f;
    for (int i = 0; i < 8; ++i) { mx = fmaxf(mx, smax[i]); dot += sdot[i]; gg += sgg[i]; pp += spp[i]; }
    int f = 0;
    if (tt.project && hyper[8] == 0.f) {
      const int n_rows = tt.row_end - tt.row_begin;
      const float row_len = (float)(tt.numel / (n_rows > 0 ? n_rows : 1));
      if (mx < delta / sqrtf(row_len)) {
        f = 1;
      } else {
        const float c = fabsf(dot) / (sqrtf(gg) + eps) / (sqrtf(pp) + eps);
        if (c < delta / sqrtf((float)tt.numel)) f = 2;
      }
    }
    flag[t] = f;
    tnorm[t] = sqrtf(pp);
    layer_acc[t] = 0.f;
    if (t == 0) {
      state[0] += 1.0f;
      const float norm = (float)sqrt(*total_gg);
      state[2] = norm;
      const float max_norm = hyper[7];
      float coef = 1.0f;
      if (max_norm > 0.f) coef = fminf(1.0f, max_norm / (norm + 1e-6f));
      state[1] = coef;
    }
  }
}

// K3: moments + update (rows of tensors with flag 0 or 1); flag 2 rows only accumulate <p^, u> for K4.
__global__ void __launch_bounds__(kOptThreads)
opt_update_kernel(const OptRow* __restrict__ rows, const OptTensor* __restrict__ tensors, int n_rows,
                  const float* __restrict__ stats,
                  const float* __restrict__ hyper, const float* __restrict__ state, const int* __restrict__ flag,
                  const float* __restrict__ tnorm, float* __restrict__ layer_acc) {
  __shared__ float red[4];
  const int r = blockIdx.x;
  if (r >= n_rows) return;
  const OptRow row = rows[r];
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], wd_ratio = hyper[6];
  const int mode = (int)hyper[8];
  const float coef = tensors[row.tensor].clip ? state[1] : 1.0f;   // clip_grad_norm_ scales model gradients only
  if (mode == 2) {  // SGD with momentum (torch semantics: g += wd*p ; buf = mom*buf + g ; p -= lr*buf)
    for (int i = threadIdx.x; i < row.len; i += kOptThreads) {
      const float p = row.p[i];
      const float g = fmaf(wd, p, row.g[i] * coef);
      const float buf = (state[0] <= 1.0f) ? g : fmaf(b1, row.m[i], g);
      row.m[i] = buf;
      const float np = fmaf(-lr, buf, p);
      row.p[i] = np;
      if (row.shadow) row.shadow[i] = __float2bfloat16(np);
    }
    return;
  }
  const float step = state[0];
  const float bc1 = 1.0f - powf(b1, step), bc2 = 1.0f - powf(b2, step);
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float step_size = lr / bc1;
  const int f = flag[row.tensor];
  const bool vec = row_is_vec4(row, true);
  float acc = 0.f;
  if (vec) {
    float4* g4 = reinterpret_cast<float4*>(row.g);
    float4* m4 = reinterpret_cast<float4*>(row.m);
    float4* v4 = reinterpret_cast<float4*>(row.v);
    float4* p4 = reinterpret_cast<float4*>(row.p);
    const float decay = wd > 0.f ? 1.0f - lr * wd : 1.0f;
    for (int i = threadIdx.x; i < (row.len >> 2); i += kOptThreads) {
      const float4 g = g4[i];
      float4 m = m4[i], v = v4[i], p = p4[i], u;
#define CFL_MOMENT(c)                                           \
      {                                                         \
        const float gc = g.c * coef;                            \
        m.c = fmaf(b1, m.c, (1.0f - b1) * gc);                  \
        v.c = fmaf(b2, v.c, (1.0f - b2) * gc * gc);             \
        u.c = m.c / (sqrtf(v.c) * inv_sqrt_bc2 + eps);          \
      }
      CFL_MOMENT(x) CFL_MOMENT(y) CFL_MOMENT(z) CFL_MOMENT(w)
#undef CFL_MOMENT
      m4[i] = m;
      v4[i] = v;
      if (f == 0) {
        p.x = fmaf(-step_size, u.x, p.x * decay); p.y = fmaf(-step_size, u.y, p.y * decay);
        p.z = fmaf(-step_size, u.z, p.z * decay); p.w = fmaf(-step_size, u.w, p.w * decay);
        p4[i] = p;
        if (row.shadow) store_shadow4(row.shadow + 4 * i, p);
      } else {
        acc = fmaf(p.x, u.x, acc); acc = fmaf(p.y, u.y, acc); acc = fmaf(p.z, u.z, acc); acc = fmaf(p.w, u.w, acc);
      }
    }
  } else {
    for (int i = threadIdx.x; i < row.len; i += kOptThreads) {
      const float g = row.g[i] * coef;
      const float m = fmaf(b1, row.m[i], (1.0f - b1) * g);
      const float v = fmaf(b2, row.v[i], (1.0f - b2) * g * g);
      row.m[i] = m;
      row.v[i] = v;
      const float u = m / (sqrtf(v) * inv_sqrt_bc2 + eps);
      if (f == 0) {
        float p = row.p[i];
        if (wd > 0.f) p *= 1.0f - lr * wd;
        p = fmaf(-step_size, u, p);
        row.p[i] = p;
        if (row.shadow) row.shadow[i] = __float2bfloat16(p);
      } else {
        acc = fmaf(row.p[i], u, acc);
      }
    }
  }
  if (f == 0) return;
  acc = block_sum_128(acc, red);
  if (f == 2) {
    if (threadIdx.x == 0) atomicAdd(layer_acc + row.tensor, acc);
    return;
  }
  // channel view: p^ = p / (|p_row| + eps) ; u -= p^ <p^, u>
  const float inv = 1.0f / (sqrtf(stats[3 * r + 2]) + eps);
  const float s = acc * inv * inv;
  if (vec) {
    const float4* m4 = reinterpret_cast<const float4*>(row.m);
    const float4* v4 = reinterpret_cast<const float4*>(row.v);
    float4* p4 = reinterpret_cast<float4*>(row.p);
    const float decay = wd > 0.f ? 1.0f - lr * wd * wd_ratio : 1.0f;
    for (int i = threadIdx.x; i < (row.len >> 2); i += kOptThreads) {
      const float4 m = m4[i], v = v4[i];
      float4 p = p4[i];
#define CFL_PROJ(c)                                                                   \
      {                                                                               \
        const float u = m.c / (sqrtf(v.c) * inv_sqrt_bc2 + eps) - p.c * s;            \
        p.c = fmaf(-step_size, u, p.c * decay);                                       \
      }
      CFL_PROJ(x) CFL_PROJ(y) CFL_PROJ(z) CFL_PROJ(w)
#undef CFL_PROJ
      p4[i] = p;
      if (row.shadow) store_shadow4(row.shadow + 4 * i, p);
    }
    return;
  }
  for (int i = threadIdx.x; i < row.len; i += kOptThreads) {
    float p = row.p[i];
    const float u = row.m[i] / (sqrtf(row.v[i]) * inv_sqrt_bc2 + eps) - p * s;
    if (wd > 0.f) p *= 1.0f - lr * wd * wd_ratio;
    p = fmaf(-step_size, u, p);
    row.p[i] = p;
    if (row.shadow) row.shadow[i] = __float2bfloat16(p);
  }
}

// K4: rows of layer-view tensors (the layer view fires for a sizeable share of the large matrices, so this runs over
// the row table like K3); block 0 clears the gradient-norm accumulator for the next step.
__global__ void __launch_bounds__(kOptThreads)
opt_layer_fix_kernel(const OptRow* __restrict__ rows, int n_rows, const float* __restrict__ hyper,
                     const float* __restrict__ state, const int* __restrict__ flag, const float* __restrict__ tnorm,
                     const float* __restrict__ layer_acc, double* __restrict__ total_gg) {
  const int r = blockIdx.x;
  if (r == 0 && threadIdx.x == 0) *total_gg = 0.0;
  if (r >= n_rows) return;
  const OptRow row = rows[r];
  if (flag[row.tensor] != 2) return;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], wd_ratio = hyper[6];
  const float step = state[0];
  const float bc1 = 1.0f - powf(b1, step), bc2 = 1.0f - powf(b2, step);
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float step_size = lr / bc1;
  const float inv = 1.0f / (tnorm[row.tensor] + eps);
  const float s = layer_acc[row.tensor] * inv * inv;
  for (int i = threadIdx.x; i < row.len; i += kOptThreads) {
    float p = row.p[i];
    const float u = row.m[i] / (sqrtf(row.v[i]) * inv_sqrt_bc2 + eps) - p * s;
    if (wd > 0.f) p *= 1.0f - lr * wd * wd_ratio;
    p = fmaf(-step_size, u, p);
    row.p[i] = p;
    if (row.shadow) row.shadow[i] = __float2bfloat16(p);
  }
}

int optimizer_step(const void* rows, int n_rows, const void* tensors, int n_tensors, const float* hyper, float* state,
                   double* total_gg, float* stats, int* flag, float* tnorm, float* layer_acc, cudaStream_t st) {
  if (n_rows <= 0 || n_tensors <= 0) {
    set_error("optimizer_step: empty table");
    return CFL_EINVAL;
  }
  const OptRow* r = reinterpret_cast<const OptRow*>(rows);
  const OptTensor* t = reinterpret_cast<const OptTensor*>(tensors);
  opt_row_stats_kernel<<<n_rows, kOptThreads, 0, st>>>(r, t, n_rows, stats, total_gg);
  opt_decide_kernel<<<n_tensors, 256, 0, st>>>(t, n_tensors, stats, hyper, state, total_gg, flag, tnorm, layer_acc);
  opt_update_kernel<<<n_rows, kOptThreads, 0, st>>>(r, t, n_rows, stats, hyper, state, flag, tnorm, layer_acc);
  opt_layer_fix_kernel<<<n_rows, kOptThreads, 0, st>>>(r, n_rows, hyper, state, flag, tnorm, layer_acc, total_gg);
  return check_launch("optimizer_step");
}

}  // namespace cfl
