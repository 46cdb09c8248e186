// GRU text towers of the clients: everything of caption_encoder.EncoderText (src/networks/models/caption_encoder.py:
// 87-116) and language_model.EncoderText (src/networks/language_model.py:93-130) that is not a dense contraction.
//
//   wemb_*      word-embedding gather (fp32 table -> bf16 rows padded to a 16-byte pitch for TMA) and its scatter-add
//               backward                                                   (nn.Embedding, caption_encoder.py:41,90)
//   gru_*       the recurrent part of the packed bidirectional GRU (nn.GRU + pack_padded_sequence,
//               caption_encoder.py:93-97): the input projection x W_ih^T + b_ih is ONE tensor-core GEMM over all tokens
//               of both directions (gemm_tc); these kernels run the sequential part h_t = f(xproj_t, W_hh h_{t-1}) with
//               the recurrent matrix held in REGISTERS for the whole sequence and the state in shared memory
//   seq_pool_*  PIENet attention pooling over the words of a caption with the pad mask of pie_model.py:31-34
//   scale_relu  `relu(out * scale)` of the unimodal text client (language_model.py:111-112)
//
// The recurrence is latency-bound (B*H*3H*2 FLOP per step, ~100 kFLOP per sequence), not bandwidth- or
// tensor-bound: the design goal is the shortest dependent chain per time step - 2 block barriers, no global
// round trip (gates are prefetched before the matvec), weights never re-read.
#include "kernels.cuh"

namespace cfl {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------ word embedding
__global__ void __launch_bounds__(256)
wemb_gather_kernel(const long long* __restrict__ ids, const float* __restrict__ table, long long total, int V, int Dw,
                   int pitch, __nv_bfloat16* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long t = i / pitch;
  const int k = (int)(i - t * pitch);
  float v = 0.0f;
  if (k < Dw) {
    long long id = ids[t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    v = table[id * Dw + k];
  }
  out[i] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
wemb_scatter_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ dx, long long total, int V, int Dw,
                    int pitch, float* __restrict__ dtable) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long t = i / Dw;
  const int k = (int)(i - t * Dw);
  long long id = ids[t];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  const float g = __bfloat162float(dx[t * pitch + k]);
  if (g != 0.0f) atomicAdd(dtable + id * Dw + k, g);
}

// ------------------------------------------------------------------------------------------------ GRU recurrence
// PyTorch gate order (r, z, n):
//   r = sigmoid(xr + W_hr h + b_hr)   z = sigmoid(xz + W_hz h + b_hz)   n = tanh(xn + r * (W_hn h + b_hn))
//   h' = (1 - z) * n + z * h
// Grid: (ceil(B / BT), 2 directions).  Block: 6H threads.  Thread (row, half) keeps half a row of W_hh[dir] (H/2
// floats) in registers; the two halves of a row are adjacent lanes and meet through one shuffle.  The state h of
// the BT sequences of the block lives in shared memory, split in two halves whose pitch (H/2 + 4 floats) puts the
// two 16-byte addresses a warp reads at once on different bank groups.
// Packed-sequence semantics (pack_padded_sequence): sequence b runs len_b steps; the reverse direction starts at
// t = len_b - 1.  rev_steps > 0 limits the reverse direction to its first rev_steps steps: the towers consume only
// the output at t = len_b - 1 (caption_encoder.py:99-101), which the reverse direction produces in its first step.
template <int H, int BT>
__global__ void __launch_bounds__(6 * H)
gru_fwd_kernel(const float* __restrict__ xproj /* [B, L, 2, 3H] */, const float* __restrict__ w_hh /* [2, 3H, H] */,
               const float* __restrict__ b_hh /* [2, 3H] */, const int* __restrict__ lengths, int B, int L,
               int rev_steps, float* __restrict__ hseq /* [B, L, 2H] */, float* __restrict__ hlast /* [B, 2H] */,
               float* __restrict__ gates /* [B, L, 2, 4, H] */) {
  constexpr int HH = H / 2;
  constexpr int HP = HH + 4;
  __shared__ __align__(16) float sh[BT][2][HP];
  __shared__ float sg[BT][3 * H];
  __shared__ int slen[BT], snst[BT];

  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * BT;
  const int tid = threadIdx.x;
  const int row = tid >> 1, half = tid & 1;

  float w[HH];
  {
    const float* wr = w_hh + ((size_t)dir * 3 * H + row) * H + half * HH;
#pragma unroll
    for (int k = 0; k < HH; k += 4) {
      const float4 v = *reinterpret_cast<const float4*>(wr + k);
      w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
    }
  }
  const float bias = b_hh[dir * 3 * H + row];
  if (tid < BT) {
    const int b = b0 + tid;
    int len = b < B ? lengths[b] : 0;
    len = len < 0 ? 0 : (len > L ? L : len);
    slen[tid] = len;
    snst[tid] = (dir == 1 && rev_steps > 0 && rev_steps < len) ? rev_steps : len;
  }
  for (int i = tid; i < BT * 2 * HP; i += 6 * H) (&sh[0][0][0])[i] = 0.0f;
  __syncthreads();
  int max_steps = 0;
#pragma unroll
  for (int i = 0; i < BT; ++i) max_steps = max(max_steps, snst[i]);

  const bool upd = tid < BT * H;
  const int ubt = upd ? tid / H : 0, uj = tid % H;
  const int ulen = slen[ubt], unst = snst[ubt];
  float hreg = 0.0f;

  for (int s = 0; s < max_steps; ++s) {
    // gate pre-activations of this step: issued before the matvec so that their latency is hidden behind it
    const bool act = upd && s < unst;
    float xr = 0.0f, xz = 0.0f, xn = 0.0f;
    int t = 0;
    if (act) {
      t = dir == 0 ? s : ulen - 1 - s;
      const float* xp = xproj + (((size_t)(b0 + ubt) * L + t) * 2 + dir) * 3 * H;
      xr = xp[uj]; xz = xp[H + uj]; xn = xp[2 * H + uj];
    }
    float acc[BT];
#pragma unroll
    for (int i = 0; i < BT; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < HH; k += 4) {
#pragma unroll
      for (int i = 0; i < BT; ++i) {
        const float4 hv = *reinterpret_cast<const float4*>(&sh[i][half][k]);
        acc[i] = fmaf(w[k], hv.x, acc[i]);
        acc[i] = fmaf(w[k + 1], hv.y, acc[i]);
        acc[i] = fmaf(w[k + 2], hv.z, acc[i]);
        acc[i] = fmaf(w[k + 3], hv.w, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < BT; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
    if (half == 0) {
#pragma unroll
      for (int i = 0; i < BT; ++i) sg[i][row] = acc[i] + bias;
    }
    __syncthreads();
    if (act) {
      const float hr = sg[ubt][uj], hz = sg[ubt][H + uj], hn = sg[ubt][2 * H + uj];
      const float r = sigmoidf_(xr + hr), z = sigmoidf_(xz + hz);
      const float n = tanhf(fmaf(r, hn, xn));
      const float hnew = fmaf(z, hreg - n, n);          // (1 - z) n + z h
      hreg = hnew;
      sh[ubt][uj / HH][uj % HH] = hnew;
      const size_t pos = (size_t)(b0 + ubt) * L + t;
      if (hseq) hseq[pos * 2 * H + dir * H + uj] = hnew;
      if (gates) {
        float* g = gates + (pos * 2 + dir) * 4 * H;
        g[uj] = r; g[H + uj] = z; g[2 * H + uj] = n; g[3 * H + uj] = hn;
      }
      if (hlast && t == ulen - 1) hlast[(size_t)(b0 + ubt) * 2 * H + dir * H + uj] = hnew;
    }
    __syncthreads();
  }
}

// Backward through time.  Thread (part, j), part in [0, 6), keeps W_hh[dir][part*H/2 .. +H/2, j] (a column piece of
// the recurrent matrix) in registers; per step the gate gradients of the BT sequences are formed by the BT*H update
// threads, broadcast through shared memory, multiplied by W_hh^T in six partial sums and folded into dh.
//   dn = dh (1 - z); dz = dh (h_prev - n); dh_prev = dh z + W_hh^T dgh
//   d(pre n) = dn (1 - n^2); d(pre z) = dz z (1 - z); d(pre r) = d(pre n) hn r (1 - r)
//   dxp = [d(pre r), d(pre z), d(pre n)]  (gradient at x W_ih^T + b_ih);  dgh = [d(pre r), d(pre z), d(pre n) r]
// dxp / dgh / hprev leave in bf16: they are the operands of the weight-gradient GEMMs (dW_ih = dxp^T x,
// dW_hh = dgh^T h_prev) and of the input-gradient GEMM (dx = dxp W_ih); entries of steps that were not run stay
// zero (the caller clears the buffers).
template <int H, int BT>
__global__ void __launch_bounds__(6 * H)
gru_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ hseq, const float* __restrict__ w_hh,
               const int* __restrict__ lengths, const float* __restrict__ dhseq, const float* __restrict__ dhlast,
               int B, int L, int rev_steps, __nv_bfloat16* __restrict__ dxp /* [B*L, 2, 3H] */,
               __nv_bfloat16* __restrict__ dgh /* [B*L, 2, 3H] */, __nv_bfloat16* __restrict__ hprev /* [B*L, 2, H] */) {
  constexpr int HH = H / 2;
  __shared__ __align__(16) float sd[BT][3 * H];
  __shared__ float sp[6][BT][H];
  __shared__ int slen[BT], snst[BT];

  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * BT;
  const int tid = threadIdx.x;
  const int part = tid / H, j = tid % H;

  float w[HH];
#pragma unroll
  for (int k = 0; k < HH; ++k) w[k] = w_hh[((size_t)dir * 3 * H + part * HH + k) * H + j];
  if (tid < BT) {
    const int b = b0 + tid;
    int len = b < B ? lengths[b] : 0;
    len = len < 0 ? 0 : (len > L ? L : len);
    slen[tid] = len;
    snst[tid] = (dir == 1 && rev_steps > 0 && rev_steps < len) ? rev_steps : len;
  }
  __syncthreads();
  int max_steps = 0;
#pragma unroll
  for (int i = 0; i < BT; ++i) max_steps = max(max_steps, snst[i]);

  const bool upd = tid < BT * H;
  const int ubt = upd ? tid / H : 0, uj = j;
  const int ulen = slen[ubt], unst = snst[ubt];
  float dh = 0.0f;

  for (int s = max_steps - 1; s >= 0; --s) {
    float zreg = 0.0f;
    if (upd) {
      float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
      if (s < unst) {
        const int t = dir == 0 ? s : ulen - 1 - s;
        const size_t pos = (size_t)(b0 + ubt) * L + t;
        float go = 0.0f;
        if (dhseq) go += dhseq[pos * 2 * H + dir * H + uj];
        if (dhlast && t == ulen - 1) go += dhlast[(size_t)(b0 + ubt) * 2 * H + dir * H + uj];
        dh += go;
        const float* g = gates + (pos * 2 + dir) * 4 * H;
        const float r = g[uj], z = g[H + uj], n = g[2 * H + uj], hn = g[3 * H + uj];
        float hp = 0.0f;
        if (s > 0) {
          const int tp = dir == 0 ? t - 1 : t + 1;
          hp = hseq[((size_t)(b0 + ubt) * L + tp) * 2 * H + dir * H + uj];
        }
        const float dn = dh * (1.0f - z);
        const float dz = dh * (hp - n);
        const float dnp = dn * (1.0f - n * n);
        const float dzp = dz * z * (1.0f - z);
        const float drp = dnp * hn * r * (1.0f - r);
        v0 = drp; v1 = dzp; v2 = dnp * r;
        zreg = z;
        __nv_bfloat16* ox = dxp + (pos * 2 + dir) * 3 * H;
        ox[uj] = __float2bfloat16_rn(drp);
        ox[H + uj] = __float2bfloat16_rn(dzp);
        ox[2 * H + uj] = __float2bfloat16_rn(dnp);
        __nv_bfloat16* og = dgh + (pos * 2 + dir) * 3 * H;
        og[uj] = __float2bfloat16_rn(v0);
        og[H + uj] = __float2bfloat16_rn(v1);
        og[2 * H + uj] = __float2bfloat16_rn(v2);
        hprev[(pos * 2 + dir) * H + uj] = __float2bfloat16_rn(hp);
      }
      sd[ubt][uj] = v0; sd[ubt][H + uj] = v1; sd[ubt][2 * H + uj] = v2;
    }
    __syncthreads();
    float acc[BT];
#pragma unroll
    for (int i = 0; i < BT; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < HH; k += 4) {
#pragma unroll
      for (int i = 0; i < BT; ++i) {
        const float4 dv = *reinterpret_cast<const float4*>(&sd[i][part * HH + k]);
        acc[i] = fmaf(w[k], dv.x, acc[i]);
        acc[i] = fmaf(w[k + 1], dv.y, acc[i]);
        acc[i] = fmaf(w[k + 2], dv.z, acc[i]);
        acc[i] = fmaf(w[k + 3], dv.w, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < BT; ++i) sp[part][i][j] = acc[i];
    __syncthreads();
    if (upd) {
      float sum = 0.0f;
#pragma unroll
      for (int q = 0; q < 6; ++q) sum += sp[q][ubt][uj];
      dh = fmaf(dh, zreg, sum);
    }
  }
}

// ------------------------------------------------------------------------------------------------ masked PIE pooling
// One block per caption.  a[p] = <h[p,:], w2> for p < len, attn = softmax over the valid positions (pad positions get
// -inf in the reference, i.e. weight 0), r = sum_p attn[p] x[p,:].
__global__ void __launch_bounds__(256)
seq_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ h,
                    const float* __restrict__ w2, const int* __restrict__ lengths, int L, int C, int pitch, int Hd,
                    int hpitch, float* __restrict__ attn, __nv_bfloat16* __restrict__ r_out) {
  extern __shared__ float sa[];
  const int b = blockIdx.x;
  const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int len = lengths[b];
  len = len < 0 ? 0 : (len > L ? L : len);
  for (int p = wp; p < L; p += 8) {
    float acc = 0.0f;
    if (p < len) {
      const __nv_bfloat16* hr = h + ((size_t)b * L + p) * hpitch;
      for (int k = lane; k < Hd; k += 32) acc = fmaf(__bfloat162float(hr[k]), w2[k], acc);
      acc = warp_sum_t(acc);
    }
    if (lane == 0) sa[p] = p < len ? acc : -INFINITY;
  }
  __syncthreads();
  if (wp == 0) {
    float m = -INFINITY;
    for (int p = lane; p < len; p += 32) m = fmaxf(m, sa[p]);
    m = warp_max_t(m);
    float s = 0.0f;
    for (int p = lane; p < len; p += 32) s += expf(sa[p] - m);
    s = warp_sum_t(s);
    for (int p = lane; p < L; p += 32) {
      const float a = p < len ? expf(sa[p] - m) / s : 0.0f;
      sa[p] = a;
      attn[(size_t)b * L + p] = a;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < pitch; k += 256) {
    float acc = 0.0f;
    if (k < C)
      for (int p = 0; p < len; ++p) acc = fmaf(sa[p], __bfloat162float(x[((size_t)b * L + p) * pitch + k]), acc);
    r_out[(size_t)b * pitch + k] = __float2bfloat16_rn(acc);
  }
}

//   dx[p,:]   = attn[p] * d_r                                  (the W1 and GRU paths are added by the dgrad GEMMs)
//   dattn[p]  = <d_r, x[p,:]> ;  da = attn * (dattn - sum attn*dattn)
//   dpre[p,:] = da[p] * w2 * (1 - h[p,:]^2) ;  dw2 += sum_p da[p] * h[p,:]
__global__ void __launch_bounds__(256)
seq_pool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ h,
                    const float* __restrict__ w2, const float* __restrict__ attn, const __nv_bfloat16* __restrict__ d_r,
                    const int* __restrict__ lengths, int L, int C, int pitch, int Hd, int hpitch,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dpre, float* __restrict__ dw2) {
  extern __shared__ float sm[];
  float* sat = sm;
  float* sda = sm + L;
  const int b = blockIdx.x;
  const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int len = lengths[b];
  len = len < 0 ? 0 : (len > L ? L : len);
  for (int p = threadIdx.x; p < L; p += 256) sat[p] = p < len ? attn[(size_t)b * L + p] : 0.0f;
  for (int p = wp; p < L; p += 8) {
    float acc = 0.0f;
    if (p < len) {
      const __nv_bfloat16* xr = x + ((size_t)b * L + p) * pitch;
      const __nv_bfloat16* dr = d_r + (size_t)b * pitch;
      for (int k = lane; k < C; k += 32) acc = fmaf(__bfloat162float(xr[k]), __bfloat162float(dr[k]), acc);
      acc = warp_sum_t(acc);
    }
    if (lane == 0) sda[p] = acc;
  }
  __syncthreads();
  if (wp == 0) {
    float s = 0.0f;
    for (int p = lane; p < L; p += 32) s = fmaf(sat[p], sda[p], s);
    s = warp_sum_t(s);
    __syncwarp();
    for (int p = lane; p < L; p += 32) sda[p] = sat[p] * (sda[p] - s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L * pitch; i += 256) {
    const int p = i / pitch, k = i - p * pitch;
    const float v = k < C ? sat[p] * __bfloat162float(d_r[(size_t)b * pitch + k]) : 0.0f;
    dx[(size_t)b * L * pitch + i] = __float2bfloat16_rn(v);
  }
  for (int k = threadIdx.x; k < hpitch; k += 256) {
    const float wv = k < Hd ? w2[k] : 0.0f;
    float acc = 0.0f;
    for (int p = 0; p < L; ++p) {
      const size_t o = ((size_t)b * L + p) * hpitch + k;
      const float hv = k < Hd ? __bfloat162float(h[o]) : 0.0f;
      const float da = sda[p];
      dpre[o] = __float2bfloat16_rn(da * wv * (1.0f - hv * hv));
      acc = fmaf(da, hv, acc);
    }
    if (k < Hd && acc != 0.0f) atomicAdd(dw2 + k, acc);
  }
}

// ------------------------------------------------------------------------------------------------ relu(x * scale)
__global__ void __launch_bounds__(256)
scale_relu_fwd_kernel(const float* __restrict__ x, long long n, float scale, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaxf(x[i] * scale, 0.0f);
}
__global__ void __launch_bounds__(256)
scale_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, float scale,
                      float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = y[i] > 0.0f ? dy[i] * scale : 0.0f;
}

constexpr int kGruBT = 4;

template <int H>
int launch_gru_fwd(const float* xproj, const float* w_hh, const float* b_hh, const int* lengths, int B, int L,
                   int rev_steps, float* hseq, float* hlast, float* gates, cudaStream_t st) {
  dim3 grid((B + kGruBT - 1) / kGruBT, 2);
  gru_fwd_kernel<H, kGruBT><<<grid, 6 * H, 0, st>>>(xproj, w_hh, b_hh, lengths, B, L, rev_steps, hseq, hlast, gates);
  return check_launch("gru_fwd");
}
template <int H>
int launch_gru_bwd(const float* gates, const float* hseq, const float* w_hh, const int* lengths, const float* dhseq,
                   const float* dhlast, int B, int L, int rev_steps, void* dxp, void* dgh, void* hprev,
                   cudaStream_t st) {
  dim3 grid((B + kGruBT - 1) / kGruBT, 2);
  gru_bwd_kernel<H, kGruBT><<<grid, 6 * H, 0, st>>>(gates, hseq, w_hh, lengths, dhseq, dhlast, B, L, rev_steps,
                                                    reinterpret_cast<__nv_bfloat16*>(dxp),
                                                    reinterpret_cast<__nv_bfloat16*>(dgh),
                                                    reinterpret_cast<__nv_bfloat16*>(hprev));
  return check_launch("gru_bwd");
}

}  // namespace

int wemb_gather_fwd(const long long* ids, const float* table, int T, int V, int Dw, int pitch, void* out,
                    cudaStream_t st) {
  if (T <= 0 || V <= 0 || Dw <= 0 || pitch < Dw) {
    set_error("wemb_gather_fwd: bad shape T=%d V=%d Dw=%d pitch=%d", T, V, Dw, pitch);
    return CFL_EINVAL;
  }
  const long long total = (long long)T * pitch;
  wemb_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ids, table, total, V, Dw, pitch,
                                                                     reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("wemb_gather_fwd");
}

int wemb_scatter_bwd(const long long* ids, const void* dx, int T, int V, int Dw, int pitch, float* dtable,
                     cudaStream_t st) {
  if (T <= 0 || V <= 0 || Dw <= 0 || pitch < Dw) {
    set_error("wemb_scatter_bwd: bad shape T=%d V=%d Dw=%d pitch=%d", T, V, Dw, pitch);
    return CFL_EINVAL;
  }
  const long long total = (long long)T * Dw;
  wemb_scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      ids, reinterpret_cast<const __nv_bfloat16*>(dx), total, V, Dw, pitch, dtable);
  return check_launch("wemb_scatter_bwd");
}

int gru_fwd(const float* xproj, const float* w_hh, const float* b_hh, const int* lengths, int B, int L, int H,
            int rev_steps, float* hseq, float* hlast, float* gates, cudaStream_t st) {
  if (B <= 0 || L <= 0) {
    set_error("gru_fwd: bad shape B=%d L=%d", B, L);
    return CFL_EINVAL;
  }
  if (hseq) cudaMemsetAsync(hseq, 0, (size_t)B * L * 2 * H * sizeof(float), st);
  if (hlast) cudaMemsetAsync(hlast, 0, (size_t)B * 2 * H * sizeof(float), st);
  switch (H) {
    case 32: return launch_gru_fwd<32>(xproj, w_hh, b_hh, lengths, B, L, rev_steps, hseq, hlast, gates, st);
    case 64: return launch_gru_fwd<64>(xproj, w_hh, b_hh, lengths, B, L, rev_steps, hseq, hlast, gates, st);
    case 128: return launch_gru_fwd<128>(xproj, w_hh, b_hh, lengths, B, L, rev_steps, hseq, hlast, gates, st);
    default:
      set_error("gru_fwd: hidden size %d not supported (32, 64 or 128 per direction: the recurrent matrix is "
                "register-resident)", H);
      return CFL_EINVAL;
  }
}

int gru_bwd(const float* gates, const float* hseq, const float* w_hh, const int* lengths, const float* dhseq,
            const float* dhlast, int B, int L, int H, int rev_steps, void* dxp, void* dgh, void* hprev,
            cudaStream_t st) {
  if (B <= 0 || L <= 0 || !gates || !hseq || !dxp || !dgh || !hprev) {
    set_error("gru_bwd: bad argument");
    return CFL_EINVAL;
  }
  const size_t rows = (size_t)B * L * 2;
  cudaMemsetAsync(dxp, 0, rows * 3 * H * 2, st);
  cudaMemsetAsync(dgh, 0, rows * 3 * H * 2, st);
  cudaMemsetAsync(hprev, 0, rows * H * 2, st);
  switch (H) {
    case 32: return launch_gru_bwd<32>(gates, hseq, w_hh, lengths, dhseq, dhlast, B, L, rev_steps, dxp, dgh, hprev, st);
    case 64: return launch_gru_bwd<64>(gates, hseq, w_hh, lengths, dhseq, dhlast, B, L, rev_steps, dxp, dgh, hprev, st);
    case 128:
      return launch_gru_bwd<128>(gates, hseq, w_hh, lengths, dhseq, dhlast, B, L, rev_steps, dxp, dgh, hprev, st);
    default:
      set_error("gru_bwd: hidden size %d not supported (32, 64 or 128 per direction)", H);
      return CFL_EINVAL;
  }
}

int seq_pool_fwd(const void* x, const void* h, const float* w2, const int* lengths, int B, int L, int C, int pitch,
                 int Hd, int hpitch, float* attn, void* r, cudaStream_t st) {
  if (B <= 0 || L <= 0 || L > 4096 || C <= 0 || pitch < C || Hd <= 0 || hpitch < Hd) {
    set_error("seq_pool_fwd: bad shape B=%d L=%d C=%d pitch=%d Hd=%d hpitch=%d", B, L, C, pitch, Hd, hpitch);
    return CFL_EINVAL;
  }
  seq_pool_fwd_kernel<<<B, 256, (size_t)L * sizeof(float), st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(h), w2, lengths, L, C, pitch,
      Hd, hpitch, attn, reinterpret_cast<__nv_bfloat16*>(r));
  return check_launch("seq_pool_fwd");
}

int seq_pool_bwd(const void* x, const void* h, const float* w2, const float* attn, const void* d_r, const int* lengths,
                 int B, int L, int C, int pitch, int Hd, int hpitch, void* dx, void* dpre, float* dw2,
                 cudaStream_t st) {
  if (B <= 0 || L <= 0 || L > 4096 || C <= 0 || pitch < C || Hd <= 0 || hpitch < Hd) {
    set_error("seq_pool_bwd: bad shape");
    return CFL_EINVAL;
  }
  seq_pool_bwd_kernel<<<B, 256, (size_t)2 * L * sizeof(float), st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(h), w2, attn,
      reinterpret_cast<const __nv_bfloat16*>(d_r), lengths, L, C, pitch, Hd, hpitch,
      reinterpret_cast<__nv_bfloat16*>(dx), reinterpret_cast<__nv_bfloat16*>(dpre), dw2);
  return check_launch("seq_pool_bwd");
}

int scale_relu_fwd(const float* x, long long n, float scale, float* y, cudaStream_t st) {
  if (n <= 0) return CFL_OK;
  scale_relu_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, scale, y);
  return check_launch("scale_relu_fwd");
}

int scale_relu_bwd(const float* dy, const float* y, long long n, float scale, float* dx, cudaStream_t st) {
  if (n <= 0) return CFL_OK;
  scale_relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dy, y, n, scale, dx);
  return check_launch("scale_relu_bwd");
}

}  // namespace cfl
