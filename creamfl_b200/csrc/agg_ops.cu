// Server-side kernels that are bound by HBM rather than by the tensor pipe.
//
//   conw_reduce   : softmax over clients of the contrastive scores, then the weighted sum of client
//                   representations                        (reference src/algorithms/MMFL.py:311-314 / 328-331)
//   recall_ranks  : rank of the best positive for every query = #gallery items scoring strictly higher
//                   (reference src/algorithms/eval_coco.py:273-334 computes the same rank with an fp64 matmul,
//                   a full sort per row and a Python search per positive)
#include "common.cuh"

namespace cfl {

constexpr int kMaxClients = 32;

struct ClientPtrs {
  const float* v[kMaxClients];
};

// One thread handles 4 consecutive features of one public row: 16-byte loads of every client's row, weights
// recomputed per thread from the [C, N] score matrix (C <= 32 scalars, L1/L2 resident).  CT > 0: client count known at
// compile time (1..8, the configurations of the reference) - weights and the C row vectors live in registers and all
// C 16-byte loads are in flight before the first FMA; CT == 0: any C <= 32 (weights in local memory).
template <int CT>
__global__ void __launch_bounds__(256)
conw_reduce_kernel(ClientPtrs vecs, const float* __restrict__ scores /* [C, N] */, int C_rt, int N, int D,
                   float* __restrict__ out /* [N, D] */, float* __restrict__ weights /* [C, N] or null */) {
  const int C = CT > 0 ? CT : C_rt;
  const int d4 = D >> 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * d4) return;
  const int n = (int)(t / d4), k4 = (int)(t % d4);
  constexpr int W = CT > 0 ? CT : kMaxClients;
  float w[W];
  float4 v[CT > 0 ? CT : 1];
  if (CT > 0) {
#pragma unroll
    for (int c = 0; c < CT; ++c) v[c] = __ldg(reinterpret_cast<const float4*>(vecs.v[c] + (long long)n * D) + k4);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < W; ++c) {
    if (c < C) {
      w[c] = __ldg(scores + (long long)c * N + n);
      mx = fmaxf(mx, w[c]);
    }
  }
  float den = 0.0f;
#pragma unroll
  for (int c = 0; c < W; ++c) {
    if (c < C) {
      w[c] = __expf(w[c] - mx);
      den += w[c];
    }
  }
  const float inv = 1.0f / den;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < W; ++c) {
    if (c < C) {
      const float wc = w[c] * inv;
      const float4 x = CT > 0 ? v[CT > 0 ? c : 0]
                              : __ldg(reinterpret_cast<const float4*>(vecs.v[c] + (long long)n * D) + k4);
      acc.x = fmaf(wc, x.x, acc.x);
      acc.y = fmaf(wc, x.y, acc.y);
      acc.z = fmaf(wc, x.z, acc.z);
      acc.w = fmaf(wc, x.w, acc.w);
      if (weights != nullptr && k4 == 0) weights[(long long)c * N + n] = wc;
    }
  }
  reinterpret_cast<float4*>(out + (long long)n * D)[k4] = acc;
}

int conw_reduce(const float* const* vecs_host, const float* scores, int C, int N, int D, float* out,
                float* weights, cudaStream_t st) {
  if (C <= 0 || C > kMaxClients) {
    set_error("conw_reduce: client count %d outside [1, %d]", C, kMaxClients);
    return CFL_EINVAL;
  }
  if (N <= 0 || D <= 0 || (D & 3)) {
    set_error("conw_reduce: bad shape N=%d D=%d (D %% 4 == 0 required)", N, D);
    return CFL_EINVAL;
  }
  ClientPtrs p{};
  for (int c = 0; c < C; ++c) {
    if (vecs_host[c] == nullptr || (reinterpret_cast<uintptr_t>(vecs_host[c]) & 15)) {
      set_error("conw_reduce: client %d representation pointer null or not 16-byte aligned", c);
      return CFL_EINVAL;
    }
    p.v[c] = vecs_host[c];
  }
  const long long threads = (long long)N * (D >> 2);
  const unsigned grid = (unsigned)((threads + 255) / 256);
  switch (C) {
#define CFL_CONW(K) case K: conw_reduce_kernel<K><<<grid, 256, 0, st>>>(p, scores, C, N, D, out, weights); break;
    CFL_CONW(1) CFL_CONW(2) CFL_CONW(3) CFL_CONW(4) CFL_CONW(5) CFL_CONW(6) CFL_CONW(7) CFL_CONW(8)
#undef CFL_CONW
    default: conw_reduce_kernel<0><<<grid, 256, 0, st>>>(p, scores, C, N, D, out, weights);
  }
  return check_launch("conw_reduce");
}

// --------------------------------------------------------------------------------------------- Recall@K
// 64x64 similarity tile per block, 256 threads, 4x4 register tile per thread, fp32 FMA in a fixed k order so
// that pass 1 (best positive similarity) and pass 2 (count of strictly larger similarities) see bit-identical
// values for the same (query, gallery) pair.
template <int PASS>
__global__ void __launch_bounds__(256)
recall_tile_kernel(const float* __restrict__ Q, const float* __restrict__ G, const long long* __restrict__ q_lab,
                   const long long* __restrict__ g_lab, int Nq, int Ng, int D, int* __restrict__ best_bits,
                   int* __restrict__ ranks) {
  __shared__ float sq[16][64 + 4];
  __shared__ float sg[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int q0 = blockIdx.y * 64, g0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  for (int k0 = 0; k0 < D; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, k = e & 15;
      sq[k][r] = (q0 + r < Nq && k0 + k < D) ? Q[(long long)(q0 + r) * D + k0 + k] : 0.0f;
      sg[k][r] = (g0 + r < Ng && k0 + k < D) ? G[(long long)(g0 + r) * D + k0 + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = sq[k][ty * 4 + i];
        b[i] = sg[k][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    const bool q_ok = q < Nq;
    if (PASS == 1) {
      if (!q_ok) continue;
      const long long ql = q_lab[q];
      float best = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = g0 + tx * 4 + j;
        if (g < Ng && g_lab[g] == ql) best = fmaxf(best, acc[i][j]);
      }
      if (best > -INFINITY) {
        // order-preserving float -> int map so that atomicMax on ints orders floats
        int bits = __float_as_int(best);
        bits = bits >= 0 ? bits : bits ^ 0x7fffffff;
        atomicMax(best_bits + q, bits);
      }
    } else {
      int bits = q_ok ? best_bits[q] : 0;
      // no positive in the gallery (the reference's evaluator raises there): every gallery item outranks the missing
      // positive, rank = Ng - never a silent perfect hit
      const bool none = bits == (int)0x80000000;
      bits = bits >= 0 ? bits : bits ^ 0x7fffffff;
      const float best = __int_as_float(bits);
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = g0 + tx * 4 + j;
        if (q_ok && g < Ng && (none || acc[i][j] > best)) ++cnt;
      }
      // 16 threads (tx) share a query row inside a half-warp: reduce before the atomic
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (tx == 0 && cnt) atomicAdd(ranks + q, cnt);
    }
  }
}

__global__ void fill_int_kernel(int* p, long long n, int v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ranks[q] (int32, 0-based) ; workspace holds Nq ints.
int recall_ranks(const float* Q, const float* G, const long long* q_lab, const long long* g_lab, int Nq, int Ng,
                 int D, int* ranks, void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (Nq <= 0 || Ng <= 0 || D <= 0) {
    set_error("recall_ranks: empty problem");
    return CFL_EINVAL;
  }
  if (ws_bytes < (size_t)Nq * sizeof(int)) {
    set_error("recall_ranks: workspace %zu < %zu", ws_bytes, (size_t)Nq * sizeof(int));
    return CFL_EWORKSPACE;
  }
  int* best = reinterpret_cast<int*>(workspace);
  fill_int_kernel<<<(Nq + 255) / 256, 256, 0, st>>>(best, Nq, (int)0x80000000);
  fill_int_kernel<<<(Nq + 255) / 256, 256, 0, st>>>(ranks, Nq, 0);
  dim3 grid((Ng + 63) / 64, (Nq + 63) / 64);
  recall_tile_kernel<1><<<grid, 256, 0, st>>>(Q, G, q_lab, g_lab, Nq, Ng, D, best, ranks);
  recall_tile_kernel<2><<<grid, 256, 0, st>>>(Q, G, q_lab, g_lab, Nq, Ng, D, best, ranks);
  return check_launch("recall_ranks");
}

}  // namespace cfl
