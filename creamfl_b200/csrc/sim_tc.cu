// Similarity-matrix kernels on tcgen05: row-wise log-sum-exp of  X = scale * Q G^T  without ever writing X.
//
// One kernel family serves three reference call sites:
//   * inter-modal InfoNCE forward  CE(Q G^T / 0.5, d_idx)       (MMClientTrainer.py:193-201, ClientTrainer.py:388-401)
//   * its backward                 P = softmax(X) - onehot       (autograd of the same lines), P is emitted in bf16 and
//                                  contracted with G by gemm_tc (dQ = P G, split-K)
//   * con_w scoring                s[n] = X[n,n] - log sum_j exp X[n,j]   with Q = client reps, M = N_pub
//                                  (MMFL.py:302-307 / 319-324; the reference builds the 50000x50000 matrix on the CPU)
//
// Work decomposition: a unit = (block of MT*128 query rows) x (chunk of consecutive 64-row tiles of G).
// The query block stays resident in shared memory for the whole unit (MT*128 x D bf16), G tiles are streamed
// through a TMA ring; every G tile is reused by MT MMAs (MT = 2 halves the L2->SM traffic of the compute-bound
// con_w case).  S tiles (128 x 64 fp32) are double-buffered in TMEM; 4*MT epilogue warps own one accumulator row
// per thread, so the running (max, sum) needs no cross-thread traffic at all.  Partial (max, sum) pairs per
// (chunk, row) are merged by lse_combine_kernel, which also evaluates the positive logit <Q_i, G_label_i>.
#include "common.cuh"
#include "ptx.cuh"

namespace cfl {

constexpr int kSimBN = 64;   // G rows per tile
constexpr int kSimKC = 64;   // bf16 elements per 128-byte swizzled row

struct SimParams {
  int M, N, D;            // Q [M, D], G [N, D]
  int m_blocks, n_chunks; // units = m_blocks * n_chunks
  int tiles_per_chunk;    // G tiles (of 64 rows) per chunk
  float scale_log2;       // inv_tau * log2(e)
  float2* partial;        // MODE 0: [n_chunks, M] (running max, running sum) in the log2 domain
  const float* lse2;      // MODE 1: [M] log2-domain LSE
  const long long* labels;// MODE 1: [M] positive column (or null)
  __nv_bfloat16* P;       // MODE 1: [M, ldp] softmax(X) - onehot(label)
  long long ldp;
};

template <int MT>
struct SimCfg {
  static constexpr int kStages = (MT == 2) ? 3 : 4;
  static constexpr int kABytes = MT * 128 * 256 * 2;       // resident query block (D <= 256)
  static constexpr int kBStageBytes = kSimBN * 256 * 2;    // one G tile
  static constexpr int kSmemBytes = kABytes + kStages * kBStageBytes + 1024 + 256;
  static constexpr int kThreads = 128 + 128 * MT;
  static constexpr uint32_t kTmemCols = (MT == 2) ? 256 : 128;
};

template <int MT, int MODE>
__global__ void __launch_bounds__(SimCfg<MT>::kThreads, 1)
sim_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, SimParams p) {
  using Cfg = SimCfg<MT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + Cfg::kStages * Cfg::kBStageBytes);
  uint64_t* full = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty = full + Cfg::kStages;       // [kStages]  MMA -> TMA
  uint64_t* a_full = empty + Cfg::kStages;     // [1]
  uint64_t* a_empty = a_full + 1;              // [1]
  uint64_t* tfull = a_empty + 1;               // [MT*2]     MMA -> epilogue
  uint64_t* tempty = tfull + MT * 2;           // [MT*2]     epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + MT * 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int dc = p.D / kSimKC;  // 128-byte k-chunks per row
  const int units = p.m_blocks * p.n_chunks;
  const int n_tiles_total = (p.N + kSimBN - 1) / kSimBN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < MT * 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ui = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
        const int mb = u % p.m_blocks;
        const int ch = u / p.m_blocks;
        const int t0 = ch * p.tiles_per_chunk;
        const int t1 = min(n_tiles_total, t0 + p.tiles_per_chunk);
        mbar_wait(a_empty, (ui & 1) ^ 1);
        mbar_arrive_expect_tx(a_full, MT * 128 * p.D * 2);
        for (int t = 0; t < MT; ++t)
          for (int c = 0; c < dc; ++c)
            tma_load_2d(&tmQ, a_full, sA + (t * 4 + c) * 16384, c * kSimKC, (mb * MT + t) * 128);
        for (int nt = t0; nt < t1; ++nt) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], kSimBN * p.D * 2);
          uint8_t* sb = sB + stage * Cfg::kBStageBytes;
          for (int c = 0; c < dc; ++c) tma_load_2d(&tmG, &full[stage], sb + c * 8192, c * kSimKC, nt * kSimBN);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(1, 128, kSimBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ui = 0;
      uint32_t it = 0;  // global tile counter (TMEM buffer phases)
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
        const int ch = u / p.m_blocks;
        const int t0 = ch * p.tiles_per_chunk;
        const int t1 = min(n_tiles_total, t0 + p.tiles_per_chunk);
        mbar_wait(a_full, ui & 1);
        tc_fence_after();
        for (int nt = t0; nt < t1; ++nt, ++it) {
          const uint32_t buf = it & 1;
          const uint32_t bphase = (it >> 1) & 1;
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sb = smem_u32(sB + stage * Cfg::kBStageBytes);
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            mbar_wait(&tempty[t * 2 + buf], bphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (t * 2 + buf) * kSimBN;
            const uint32_t sa = smem_u32(sA + t * 4 * 16384);
            for (int c = 0; c < dc; ++c) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = make_smem_desc(sa + c * 16384 + k * 32, 16, 1024);
                const uint64_t db = make_smem_desc(sb + c * 8192 + k * 32, 16, 1024);
                umma_f16_ss(tmem_d, da, db, idesc, (c > 0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit(&tfull[t * 2 + buf]);
          }
          umma_commit(&empty[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(a_empty);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: thread == accumulator row
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int mb = u % p.m_blocks;
      const int ch = u / p.m_blocks;
      const int t0 = ch * p.tiles_per_chunk;
      const int t1 = min(n_tiles_total, t0 + p.tiles_per_chunk);
      const int row = (mb * MT + t) * 128 + q * 32 + lane;
      float run_m = -INFINITY, run_l = 0.0f;
      float my_lse2 = 0.0f;
      long long my_label = -1;
      if (MODE == 1 && row < p.M) {
        my_lse2 = p.lse2[row];
        if (p.labels) my_label = p.labels[row];
      }
      for (int nt = t0; nt < t1; ++nt, ++it) {
        const uint32_t buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        mbar_wait(&tfull[t * 2 + buf], bphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (t * 2 + buf) * kSimBN;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        // the accumulator is in registers: hand the TMEM buffer back before doing the math
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[t * 2 + buf]);

        const int n0 = nt * kSimBN;
        const int valid = min(kSimBN, p.N - n0);
        if (MODE == 0) {
          float x[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            x[j] = __uint_as_float(v0[j]) * p.scale_log2;
            x[32 + j] = __uint_as_float(v1[j]) * p.scale_log2;
          }
          if (valid < kSimBN) {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (j >= valid) x[j] = -INFINITY;
          }
          float tm = x[0];
#pragma unroll
          for (int j = 1; j < 64; ++j) tm = fmaxf(tm, x[j]);
          const float new_m = fmaxf(run_m, tm);
          float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            s0 += ex2_approx(x[j] - new_m);
            s1 += ex2_approx(x[j + 1] - new_m);
            s2 += ex2_approx(x[j + 2] - new_m);
            s3 += ex2_approx(x[j + 3] - new_m);
          }
          run_l = run_l * ex2_approx(run_m - new_m) + ((s0 + s1) + (s2 + s3));
          run_m = new_m;
        } else {
          if (row < p.M) {
            __nv_bfloat16* prow = p.P + (long long)row * p.ldp + n0;
            if (valid == kSimBN) {
#pragma unroll
              for (int j = 0; j < 64; j += 8) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float s = __uint_as_float(j + i < 32 ? v0[(j + i) & 31] : v1[(j + i) & 31]);
                  e[i] = ex2_approx(s * p.scale_log2 - my_lse2);
                  if ((long long)(n0 + j + i) == my_label) e[i] -= 1.0f;
                }
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(e[0], e[1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(e[2], e[3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(e[4], e[5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(e[6], e[7]);
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2);
                pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(prow + j) = pk;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                if (j < valid) {
                  const float s = __uint_as_float(j < 32 ? v0[j & 31] : v1[j & 31]);
                  float e = ex2_approx(s * p.scale_log2 - my_lse2);
                  if ((long long)(n0 + j) == my_label) e -= 1.0f;
                  prow[j] = __float2bfloat16(e);
                }
              }
            }
          }
        }
      }
      if (MODE == 0 && row < p.M) p.partial[(long long)ch * p.M + row] = make_float2(run_m, run_l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// Merge per-chunk (max, sum) pairs; evaluate the positive logit; emit per-row outputs.
//   lse2[i]   = log2-domain LSE               (kept for the backward pass)
//   score[i]  = pos_i - lse_i                 (natural log domain; con_w's diag(log_prob), -CE per row)
// One warp per row.
__global__ void lse_combine_kernel(const float2* __restrict__ partial, int n_chunks, int M, int D,
                                   const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ G,
                                   const long long* __restrict__ labels, float scale, float* __restrict__ lse2,
                                   float* __restrict__ score) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float m = -INFINITY;
  for (int c = lane; c < n_chunks; c += 32) m = fmaxf(m, partial[(long long)c * M + row].x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float l = 0.0f;
  for (int c = lane; c < n_chunks; c += 32) {
    const float2 pr = partial[(long long)c * M + row];
    l += pr.y * exp2f(pr.x - m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  const float l2 = m + log2f(l);
  const long long lab = labels ? labels[row] : (long long)row;
  float dot = 0.0f;
  const __nv_bfloat16* qr = Q + (long long)row * D;
  const __nv_bfloat16* gr = G + lab * D;
  for (int k = lane * 2; k < D; k += 64) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(qr + k);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(gr + k);
    dot += __bfloat162float(a.x) * __bfloat162float(b.x) + __bfloat162float(a.y) * __bfloat162float(b.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (lane == 0) {
    if (lse2) lse2[row] = l2;
    if (score) score[row] = dot * scale - l2 * 0.6931471805599453f;
  }
}

// Choose the decomposition: rows per block (MT) and the number of G chunks so that the unit count fills the
// SMs in whole waves.
static void plan_units(int M, int N, int* mt, int* m_blocks, int* n_chunks, int* tiles_per_chunk) {
  const int sms = sm_count();
  *mt = (M > 128) ? 2 : 1;
  *m_blocks = (M + 128 * (*mt) - 1) / (128 * (*mt));
  const int n_tiles = (N + kSimBN - 1) / kSimBN;
  int best_c = 1;
  double best_eff = -1.0;
  const int max_c = n_tiles < 4 * sms ? n_tiles : 4 * sms;
  for (int c = 1; c <= max_c; ++c) {
    const int per = (n_tiles + c - 1) / c;
    const int cc = (n_tiles + per - 1) / per;  // chunks actually non-empty
    const long long units = (long long)(*m_blocks) * cc;
    const long long waves = (units + sms - 1) / sms;
    // cost ~ waves * (tiles per chunk + fixed per-unit overhead of ~4 tiles for the A reload / drain)
    const double cost = (double)waves * (per + 4.0);
    const double ideal = (double)(*m_blocks) * n_tiles / sms;
    const double eff = ideal / cost;
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best_c = cc;
    }
  }
  const int per = (n_tiles + best_c - 1) / best_c;
  *tiles_per_chunk = per;
  *n_chunks = (n_tiles + per - 1) / per;
}

size_t rowlse_workspace_bytes(int M, int N) {
  int mt, mb, nc, tpc;
  plan_units(M, N, &mt, &mb, &nc, &tpc);
  return (size_t)nc * (size_t)M * sizeof(float2);
}

template <int MT, int MODE>
static int launch_sim(const CUtensorMap& tq, const CUtensorMap& tg, const SimParams& p, cudaStream_t stream) {
  using Cfg = SimCfg<MT>;
  auto kern = sim_tc_kernel<MT, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("sim_tc: cudaFuncSetAttribute(%d B smem): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return CFL_ECUDA;
    }
    attr_set = true;
  }
  const int units = p.m_blocks * p.n_chunks;
  const int grid = units < sm_count() ? units : sm_count();
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tq, tg, p);
  return check_launch("sim_tc_kernel");
}

static int check_sim_args(const char* who, const void* Q, const void* G, int M, int N, int D) {
  if (M <= 0 || N <= 0) {
    set_error("%s: empty problem M=%d N=%d", who, M, N);
    return CFL_EINVAL;
  }
  if (D % 64 != 0 || D < 64 || D > 256) {
    set_error("%s: feature dim %d unsupported (need 64, 128, 192 or 256)", who, D);
    return CFL_EINVAL;
  }
  if (!Q || !G) {
    set_error("%s: null operand", who);
    return CFL_EINVAL;
  }
  return CFL_OK;
}

// score[i] = scale*<Q_i, G_label_i> - log sum_j exp(scale*<Q_i, G_j>);  lse2 (optional) feeds softmax_emit.
int rowlse_bf16(const void* Q, const void* G, const long long* labels, int M, int N, int D, float scale,
                float* score, float* lse2, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  int rc = check_sim_args("rowlse", Q, G, M, N, D);
  if (rc) return rc;
  SimParams p{};
  int mt;
  plan_units(M, N, &mt, &p.m_blocks, &p.n_chunks, &p.tiles_per_chunk);
  const size_t need = (size_t)p.n_chunks * (size_t)M * sizeof(float2);
  if (ws_bytes < need || workspace == nullptr) {
    set_error("rowlse: workspace %zu B < %zu B", ws_bytes, need);
    return CFL_EWORKSPACE;
  }
  p.M = M; p.N = N; p.D = D;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.partial = reinterpret_cast<float2*>(workspace);
  CUtensorMap tq, tg;
  if ((rc = make_tmap_2d(&tq, Q, 2, M, D, D, kSimKC, 128))) return rc;
  if ((rc = make_tmap_2d(&tg, G, 2, N, D, D, kSimKC, kSimBN))) return rc;
  rc = (mt == 2) ? launch_sim<2, 0>(tq, tg, p, stream) : launch_sim<1, 0>(tq, tg, p, stream);
  if (rc) return rc;
  const int wpb = 8;
  lse_combine_kernel<<<(M + wpb - 1) / wpb, wpb * 32, 0, stream>>>(
      p.partial, p.n_chunks, M, D, reinterpret_cast<const __nv_bfloat16*>(Q),
      reinterpret_cast<const __nv_bfloat16*>(G), labels, scale, lse2, score);
  return check_launch("lse_combine_kernel");
}

// P[i, j] = exp(scale*<Q_i,G_j> - lse_i) - [j == label_i]   (bf16, row pitch ldp >= N, ldp % 8 == 0)
int softmax_emit_bf16(const void* Q, const void* G, const long long* labels, const float* lse2, int M, int N,
                      int D, float scale, void* P, long long ldp, cudaStream_t stream) {
  int rc = check_sim_args("softmax_emit", Q, G, M, N, D);
  if (rc) return rc;
  if (ldp < N || (ldp % 8) != 0 || (reinterpret_cast<uintptr_t>(P) & 15)) {
    set_error("softmax_emit: P pitch %lld must be >= N=%d, a multiple of 8, base 16-byte aligned", ldp, N);
    return CFL_EINVAL;
  }
  SimParams p{};
  int mt;
  plan_units(M, N, &mt, &p.m_blocks, &p.n_chunks, &p.tiles_per_chunk);
  p.M = M; p.N = N; p.D = D;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2;
  p.labels = labels;
  p.P = reinterpret_cast<__nv_bfloat16*>(P);
  p.ldp = ldp;
  CUtensorMap tq, tg;
  if ((rc = make_tmap_2d(&tq, Q, 2, M, D, D, kSimKC, 128))) return rc;
  if ((rc = make_tmap_2d(&tg, G, 2, N, D, D, kSimKC, kSimBN))) return rc;
  return (mt == 2) ? launch_sim<2, 1>(tq, tg, p, stream) : launch_sim<1, 1>(tq, tg, p, stream);
}

}  // namespace cfl
