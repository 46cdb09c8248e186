// Similarity-matrix kernels on tcgen05: row-wise log-sum-exp of  X = scale * Q G^T  without ever writing X.
//
// One kernel family serves three reference call sites:
//   * inter-modal InfoNCE forward  CE(Q G^T / 0.5, d_idx)       (MMClientTrainer.py:193-201, ClientTrainer.py:388-401)
//   * its backward                 P = softmax(X) - onehot       (autograd of the same lines), P is emitted in bf16 and
//                                  contracted with G by gemm_tc (dQ = P G, split-K)
//   * con_w scoring                s[n] = X[n,n] - log sum_j exp X[n,j]   with Q = client reps, M = N_pub
//                                  (MMFL.py:302-307 / 319-324; the reference builds the 50000x50000 matrix on the CPU)
//
// Work decomposition: a unit = (block of 128 query rows) x (chunk of consecutive 256-row tiles of G).  The query
// block stays resident in shared memory for the whole unit (128 x D bf16); G is streamed through a TMA ring in
// (256 rows x 64 columns) stages.  Every S tile is ONE accumulator of 128 x 256 fp32 built by D/16 tcgen05.mma of
// shape 128x256x16 - with N = 256 the shared-memory operand traffic is 96 B/clk per SM, below the 128 B/clk port
// (N = 64 needs 192 B/clk and leaves the tensor pipe half idle).  Two S tiles fit in TMEM (512 columns), so the
// epilogue of tile i overlaps the MMAs of tile i+1.  Eight epilogue warps: thread = accumulator row x column half,
// running (max, sum) per thread, no cross-thread traffic.  Partial (max, sum) pairs per (chunk, half, row) are
// merged by lse_combine_kernel, which also evaluates the positive logit <Q_i, G_label_i>.
#include "common.cuh"
#include "ptx.cuh"

namespace cfl {

constexpr int kSimBN = 256;  // G rows per tile (= UMMA N)
constexpr int kSimKC = 64;   // bf16 elements per 128-byte swizzled row
constexpr int kSimThreads = 384;
constexpr int kSimStages = 4;

struct SimParams {
  int M, N, D;            // Q [M, D], G [N, D]
  int m_blocks, n_chunks; // units = m_blocks * n_chunks
  int tiles_per_chunk;    // G tiles (of 256 rows) per chunk
  float scale_log2;       // inv_tau * log2(e)
  float2* partial;        // MODE 0: [n_chunks * 2, M] (running max, running sum) in the log2 domain
  const float* lse2;      // MODE 1: [M] log2-domain LSE
  const long long* labels;// MODE 1: [M] positive column (or null)
  __nv_bfloat16* P;       // MODE 1: [M, ldp] softmax(X) - onehot(label)
  long long ldp;
};

struct SimCfg {
  static constexpr int kABytes = 128 * 256 * 2;            // resident query block (D <= 256)
  static constexpr int kBStageBytes = kSimBN * kSimKC * 2; // one (256 x 64) slice of a G tile
  static constexpr int kSmemBytes = kABytes + kSimStages * kBStageBytes + 1024 + 256;
  static constexpr uint32_t kTmemCols = 512;
};

template <int MODE>
__global__ void __launch_bounds__(kSimThreads, 1)
sim_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, SimParams p) {
  using Cfg = SimCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kSimStages * Cfg::kBStageBytes);
  uint64_t* full = bars;                       // [kSimStages]  TMA -> MMA
  uint64_t* empty = full + kSimStages;         // [kSimStages]  MMA -> TMA
  uint64_t* a_full = empty + kSimStages;       // [1]
  uint64_t* a_empty = a_full + 1;              // [1]
  uint64_t* tfull = a_empty + 1;               // [2]           MMA -> epilogue
  uint64_t* tempty = tfull + 2;                // [2]           epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

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
    for (int i = 0; i < kSimStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
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
        mbar_arrive_expect_tx(a_full, 128 * p.D * 2);
        for (int c = 0; c < dc; ++c) tma_load_2d(&tmQ, a_full, sA + c * 16384, c * kSimKC, mb * 128);
        for (int nt = t0; nt < t1; ++nt) {
          for (int c = 0; c < dc; ++c) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], Cfg::kBStageBytes);
            tma_load_2d(&tmG, &full[stage], sB + stage * Cfg::kBStageBytes, c * kSimKC, nt * kSimBN);
            if (++stage == kSimStages) {
              stage = 0;
              phase ^= 1;
            }
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
        const uint32_t sa = smem_u32(sA);
        for (int nt = t0; nt < t1; ++nt, ++it) {
          const uint32_t buf = it & 1;
          const uint32_t bphase = (it >> 1) & 1;
          mbar_wait(&tempty[buf], bphase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * kSimBN;
          for (int c = 0; c < dc; ++c) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sb = smem_u32(sB + stage * Cfg::kBStageBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_smem_desc(sa + c * 16384 + k * 32, 16, 1024);
              const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024);
              umma_f16_ss(tmem_d, da, db, idesc, (c > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty[stage]);
            if (++stage == kSimStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit(&tfull[buf]);
        }
        umma_commit(a_empty);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: thread = (accumulator row, column half)
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int mb = u % p.m_blocks;
      const int ch = u / p.m_blocks;
      const int t0 = ch * p.tiles_per_chunk;
      const int t1 = min(n_tiles_total, t0 + p.tiles_per_chunk);
      const int row = mb * 128 + q * 32 + lane;
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
        mbar_wait(&tfull[buf], bphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kSimBN + half * 128;
        const int nbase = nt * kSimBN + half * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          const int n0 = nbase + c * 32;
          const int valid = min(32, p.N - n0);   // may be <= 0 for the ragged last tile
          if (MODE == 0) {
            if (valid <= 0) continue;
            float tm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (valid == 32 || j < valid) tm = fmaxf(tm, __uint_as_float(v[j]));
            const float new_m = fmaxf(run_m, tm * p.scale_log2);     // scale > 0: max commutes with the scaling
            if (new_m > run_m) {
              run_l *= ex2_approx(run_m - new_m);
              run_m = new_m;
            }
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
            if (valid == 32) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                s0 += ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2, -run_m));
                s1 += ex2_approx(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -run_m));
                s2 += ex2_approx(fmaf(__uint_as_float(v[j + 2]), p.scale_log2, -run_m));
                s3 += ex2_approx(fmaf(__uint_as_float(v[j + 3]), p.scale_log2, -run_m));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) s0 += ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2, -run_m));
            }
            run_l += (s0 + s1) + (s2 + s3);
          } else {
            if (row < p.M && valid > 0) {
              __nv_bfloat16* prow = p.P + (long long)row * p.ldp + n0;
              if (valid == 32) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  float e[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    e[i] = ex2_approx(fmaf(__uint_as_float(v[j + i]), p.scale_log2, -my_lse2));
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
                for (int j = 0; j < 32; ++j) {
                  if (j < valid) {
                    float e = ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2, -my_lse2));
                    if ((long long)(n0 + j) == my_label) e -= 1.0f;
                    prow[j] = __float2bfloat16(e);
                  }
                }
              }
            }
          }
        }
        // accumulator consumed: hand the TMEM buffer back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
      }
      if (MODE == 0 && row < p.M)
        p.partial[((long long)ch * 2 + half) * p.M + row] = make_float2(run_m, run_l);
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

// Choose the number of G chunks so that the unit count fills the SMs in whole waves.
static void plan_units(int M, int N, int* m_blocks, int* n_chunks, int* tiles_per_chunk) {
  const int sms = sm_count();
  *m_blocks = (M + 127) / 128;
  const int n_tiles = (N + kSimBN - 1) / kSimBN;
  int best_c = 1;
  double best_eff = -1.0;
  const int max_c = n_tiles < 4 * sms ? n_tiles : 4 * sms;
  for (int c = 1; c <= max_c; ++c) {
    const int per = (n_tiles + c - 1) / c;
    const int cc = (n_tiles + per - 1) / per;  // chunks actually non-empty
    const long long units = (long long)(*m_blocks) * cc;
    const long long waves = (units + sms - 1) / sms;
    // cost ~ waves * (tiles per chunk + fixed per-unit overhead of ~1 tile for the A reload / drain)
    const double cost = (double)waves * (per + 1.0);
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
  int mb, nc, tpc;
  plan_units(M, N, &mb, &nc, &tpc);
  return (size_t)nc * 2 * (size_t)M * sizeof(float2);
}

template <int MODE>
static int launch_sim(const CUtensorMap& tq, const CUtensorMap& tg, const SimParams& p, cudaStream_t stream) {
  using Cfg = SimCfg;
  auto kern = sim_tc_kernel<MODE>;
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
  kern<<<grid, kSimThreads, Cfg::kSmemBytes, stream>>>(tq, tg, p);
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
  if (!(scale > 0.0f)) {
    set_error("rowlse: scale must be positive");
    return CFL_EINVAL;
  }
  SimParams p{};
  plan_units(M, N, &p.m_blocks, &p.n_chunks, &p.tiles_per_chunk);
  const size_t need = (size_t)p.n_chunks * 2 * (size_t)M * sizeof(float2);
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
  rc = launch_sim<0>(tq, tg, p, stream);
  if (rc) return rc;
  const int wpb = 8;
  lse_combine_kernel<<<(M + wpb - 1) / wpb, wpb * 32, 0, stream>>>(
      p.partial, p.n_chunks * 2, M, D, reinterpret_cast<const __nv_bfloat16*>(Q),
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
  plan_units(M, N, &p.m_blocks, &p.n_chunks, &p.tiles_per_chunk);
  p.M = M; p.N = N; p.D = D;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2;
  p.labels = labels;
  p.P = reinterpret_cast<__nv_bfloat16*>(P);
  p.ldp = ldp;
  CUtensorMap tq, tg;
  if ((rc = make_tmap_2d(&tq, Q, 2, M, D, D, kSimKC, 128))) return rc;
  if ((rc = make_tmap_2d(&tg, G, 2, N, D, D, kSimKC, kSimBN))) return rc;
  return launch_sim<1>(tq, tg, p, stream);
}

}  // namespace cfl
