// Dense bf16 GEMM on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
//   D[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
//
// Both operands are plain row-major 2-D arrays in HBM and either may be "K-major" (stored [rows, K]) or
// "MN-major" (stored [K, rows]); the same kernel therefore serves the three GEMMs of a linear layer without any
// transposition pass:   fwd  Y = X W^T      (A=X  K-major,  B=W  K-major)
//                       dgrad dX = dY W     (A=dY K-major,  B=W  MN-major)
//                       wgrad dW = dY^T X   (A=dY MN-major, B=X  MN-major, split-K + fp32 atomics)
// It replaces the cuBLAS calls reached from the reference through nn.Linear / HF BertModel
// (reference: src/networks/models/pcme.py:31-44, pie_model.py:18-19,51, image_encoder.py:30,57) and the 1x1
// convolutions of torchvision ResNet in NHWC (image_encoder.py:24).
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0   : TMA producer    (one elected lane)            smem ring of kStages x {A 128x64, B BNx64} bf16
//   warp 1   : MMA issuer      (one elected lane)            tcgen05.mma 128 x BN x 16, fp32 accum in TMEM
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue        (thread = one accumulator row) tcgen05.ld -> bias/act/residual -> global
// TMEM holds two accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "kernels.cuh"
#include "ptx.cuh"
#include <string.h>
#include <stdlib.h>

namespace cfl {


constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 384;   // 4 control warps + 8 epilogue warps

// ADD_TMA: the bf16 residual operand of the epilogue (`add`) is prefetched tile by tile into shared memory by the
// TMA producer (two buffers), so its DRAM latency hides behind the previous tiles instead of stalling the epilogue.
template <int BN, bool ADD_TMA = false, bool CTA2 = false>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBRows = CTA2 ? BN / 2 : BN;          // CTA pair: each CTA stages half of the B tile
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = kBM * BN * 2;            // bf16 output tile staged for the TMA store
  static constexpr int kOutBufs = (BN == 256) ? 1 : 2;      // staging buffers (BN = 256: 64 KB, single)
  static constexpr int kAddBufs = ADD_TMA ? 2 : 0;
  // BatchNorm statistics: every epilogue thread owns one (row group, column pair) slot of fp64 accumulators
  // (sum, sum of squares): [512 / BN row groups][BN] doubles x 2 = 8 KB, no shared-memory atomics
  static constexpr int kStatBytes = ADD_TMA ? 0 : 2 * 512 * 8;
  static constexpr int kFixedBytes = (kOutBufs + kAddBufs) * kOutBytes + 1024 /*align*/ + 256 /*barriers*/ + kStatBytes;
  // single CTA: the round-1 stage counts; CTA pair: smaller stages, as many as fit (at most 8)
  static constexpr int kPairStages = (227 * 1024 - kFixedBytes) / kStageBytes;
  static constexpr int kStages = CTA2 ? (kPairStages > 8 ? 8 : kPairStages)
                                      : (ADD_TMA ? 3 : ((BN == 256) ? 3 : (BN == 128 ? 4 : 6)));
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
  static constexpr uint32_t kTmemCols = 2 * BN;              // two accumulator buffers (power of two)
};

// erf with |error| <= 1.5e-7 (Abramowitz-Stegun 7.1.26): one ex2, one rcp, 6 FMA - the epilogue must not outlast
// the MMAs of the next tile.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float y = 1.0f - p * t * ex2_approx(-1.4426950408889634f * ax * ax);
  return copysignf(y, x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erf_fast(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * ex2_approx(-0.72134752044448170f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// CTA2: the kernel runs as clusters of two CTAs on the two SMs of a TPC; a pair owns a 256 x BN output tile, the
// leader (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256 for both, each CTA loads its own 128 rows of A
// and HALF of the B tile (L2 -> shared-memory operand traffic per FLOP drops by 1/4 .. 1/3), and each CTA runs the
// epilogue of its own 128 accumulator rows.
template <int BN, bool A_MN, bool B_MN, bool ADD_TMA, bool STATS, bool CTA2>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, GemmParams p) {
  using Cfg = GemmCfg<BN, ADD_TMA, CTA2>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sout = smem + Cfg::kStages * Cfg::kStageBytes;    // [2][kOutBytes], 1024-aligned (stage sizes are)
  uint8_t* sadd = sout + Cfg::kOutBufs * Cfg::kOutBytes;    // [kAddBufs][kOutBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sadd + Cfg::kAddBufs * Cfg::kOutBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* dfull = tempty + 2;                              // [2] residual tile landed
  uint64_t* dempty = dfull + 2;                              // [2] residual tile consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);
  double* s_csum = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [RG][BN] (only if kStatBytes)
  double* s_csq = s_csum + 512;                                                         // [RG][BN], RG = 512 / BN

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work units: (M block, N block, K split); a CTA pair takes M blocks of 256 rows, rank r of the pair the r-th half
  const int cta_rank = CTA2 ? (int)cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int wid = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // worker (CTA or CTA pair) index
  const int wstride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_m = CTA2 ? (p.M + 2 * kBM - 1) / (2 * kBM) : (p.M + kBM - 1) / kBM;
  const int num_n = (p.N + BN - 1) / BN;
  const int nkb = (p.K + kBK - 1) / kBK;
  const int kb_per = (nkb + p.split_k - 1) / p.split_k;
  const int units = num_m * num_n * p.split_k;
  auto row_block = [&](int u) { return CTA2 ? 2 * (u % num_m) + cta_rank : u % num_m; };   // 128-row block of this CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmC);
    if (ADD_TMA) tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], CTA2 ? 16 : 8);     // pair: the epilogue warps of BOTH CTAs release the leader's buffer
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CTA2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot); else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();     // pair: the peer's barriers are initialised too
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = wid; u < units; u += wstride, ++it) {
        const int m_blk = row_block(u);
        const int n_blk = (u / num_m) % num_n;
        const int ks = u / (num_m * num_n);
        const int kb0 = ks * kb_per;
        const int kb1 = min(nkb, kb0 + kb_per);
        const int b_row0 = n_blk * BN + (CTA2 ? cta_rank * (BN / 2) : 0);     // this CTA's slice of the B tile
        if (ADD_TMA) {
          const int db = it & 1;
          mbar_wait(&dempty[db], ((it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&dfull[db], Cfg::kOutBytes);
#pragma unroll
          for (int r = 0; r < BN / 64; ++r)
            tma_load_2d(&tmD, &dfull[db], sadd + db * Cfg::kOutBytes + r * (kBM * 128), n_blk * BN + r * 64,
                        m_blk * kBM);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          // pair: both CTAs' bytes are counted on the LEADER's barrier (the MMA issuer waits there)
          if (!CTA2) mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
          else if (leader) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
          auto load = [&](const CUtensorMap* m, void* dst, int c0, int c1) {
            if (CTA2) tma_load_2d_pair(m, &full[stage], dst, c0, c1); else tma_load_2d(m, &full[stage], dst, c0, c1);
          };
          if constexpr (!A_MN) {
            load(&tmA, sa, kb * kBK, m_blk * kBM);
          } else {
#pragma unroll
            for (int j = 0; j < kBM / 64; ++j) load(&tmA, sa + j * 8192, m_blk * kBM + j * 64, kb * kBK);
          }
          if constexpr (!B_MN) {
            load(&tmB, sb, kb * kBK, b_row0);
          } else {
#pragma unroll
            for (int j = 0; j < Cfg::kBRows / 64; ++j) load(&tmB, sb + j * 8192, b_row0 + j * 64, kb * kBK);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------ MMA issuer (pair: the leader CTA only)
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(1, CTA2 ? 2 * kBM : kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = wid; u < units; u += wstride, ++it) {
        const int ks = u / (num_m * num_n);
        const int kb0 = ks * kb_per;
        const int kb1 = min(nkb, kb0 + kb_per);
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        mbar_wait(&tempty[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
            if (CTA2) umma_f16_ss_pair(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_f16_ss(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (CTA2) umma_commit_pair(&empty[stage]); else umma_commit(&empty[stage]);   // pair: frees both CTAs' slots
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (CTA2) umma_commit_pair(&tfull[buf]); else umma_commit(&tfull[buf]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: 8 warps, thread = row, warp set = column half
    constexpr int HC = BN / 2;             // columns per thread
    const int q = warp & 3;                // TMEM lane group of this warp
    const int half = (warp - 4) >> 2;      // column half
    const int rloc = q * 32 + lane;
    const bool issuer = (warp == 4 && lane == 0);
    // fused BatchNorm statistics of the bf16 output (per-column sum / sum of squares): accumulated per CTA in shared
    // memory over consecutive tiles of the same column block, flushed with one fp64 atomic per column
    const bool do_stats = STATS && !ADD_TMA && p.stats != nullptr && p.tma_out;
    const int etid = threadIdx.x - 128;
    int acc_n_blk = -1;
    constexpr int kStatRG = 512 / BN;      // row groups of the column-sum pass below
    if (do_stats) {
      for (int e = etid; e < 512; e += 256) { s_csum[e] = 0.0; s_csq[e] = 0.0; }
      asm volatile("bar.sync 3, 256;" ::: "memory");
    }
    // dropout of the linear output (HF BertSelfOutput / BertOutput: dense -> dropout -> + residual): the mask is a
    // function of (seed, step, site, row * N + column) and is regenerated by the backward kernels (philox.cuh)
    const bool do_drop = p.drop.rng != nullptr;
    const unsigned long long drop_seed = do_drop ? p.drop.rng[0] : 0ull;
    const uint32_t drop_step = do_drop ? (uint32_t)p.drop.rng[1] : 0u;
    int it = 0;
    for (int u = wid; u < units; u += wstride, ++it) {
      const int m_blk = row_block(u);
      const int n_blk = (u / num_m) % num_n;
      const int ks = u / (num_m * num_n);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const bool has_k = ks * kb_per < nkb;
      uint8_t* stile = sout + (Cfg::kOutBufs == 2 ? buf : 0) * Cfg::kOutBytes;
      if (p.tma_out) {
        // the TMA store that last read this staging buffer must have drained
        if (issuer) tma_store_wait_read<Cfg::kOutBufs - 1>();
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (do_stats && acc_n_blk != n_blk) {
        if (acc_n_blk >= 0) {
          if (etid < BN && acc_n_blk * BN + etid < p.N) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int g = 0; g < kStatRG; ++g) { a += s_csum[g * BN + etid]; b += s_csq[g * BN + etid]; }
            atomicAdd(p.stats + acc_n_blk * BN + etid, a);
            atomicAdd(p.stats + p.N + acc_n_blk * BN + etid, b);
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          for (int e = etid; e < 512; e += 256) { s_csum[e] = 0.0; s_csq[e] = 0.0; }
          asm volatile("bar.sync 3, 256;" ::: "memory");
        }
        acc_n_blk = n_blk;
      }
      const int row = m_blk * kBM + rloc;
      // residual operand (bf16): fetched one 32-column chunk ahead of its use, the first chunk before the
      // accumulator is even ready, so the global-load latency hides behind the MMAs / the previous chunk
      const bool add_fast = !ADD_TMA && p.add != nullptr && p.add_bf16 && ks == 0 && row < p.M && has_k &&
                            ((p.ld_add & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.add) & 15) == 0);
      const __nv_bfloat16* add_row =
          add_fast ? reinterpret_cast<const __nv_bfloat16*>(p.add) + (long long)row * p.ld_add + n_blk * BN + half * HC
                   : nullptr;
      uint4 add_cur[4], add_nxt[4];
      bool cur_ok = false, nxt_ok = false;
      if (add_fast && n_blk * BN + half * HC + 32 <= p.N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) add_cur[j] = __ldg(reinterpret_cast<const uint4*>(add_row) + j);
        cur_ok = true;
      }
      if (ADD_TMA) mbar_wait(&dfull[buf], bphase);
      mbar_wait(&tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + half * HC;
#pragma unroll 1
      for (int c = 0; c < HC / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        nxt_ok = false;
        if (add_fast && c + 1 < HC / 32 && n_blk * BN + half * HC + (c + 2) * 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) add_nxt[j] = __ldg(reinterpret_cast<const uint4*>(add_row + (c + 1) * 32) + j);
          nxt_ok = true;
        }
        tmem_ld_wait();
        const int ctile = half * HC + c * 32;          // column offset inside the tile
        const int col0 = n_blk * BN + ctile;
        const bool live = row < p.M && col0 < p.N && has_k;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
        const bool full_chunk = (col0 + 32 <= p.N);
        if (live) {
          if (p.bias != nullptr && ks == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full_chunk || col0 + j < p.N) f[j] += __ldg(p.bias + col0 + j);
          }
          if (do_drop) {
            const unsigned long long e8 = ((unsigned long long)row * (unsigned long long)p.N + col0) >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (full_chunk || col0 + 8 * j < p.N) {
                const uint32_t keep = drop_keep8(drop_seed, drop_step, (uint32_t)p.drop.site, e8 + j, p.drop.thresh);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[8 * j + i] = ((keep >> i) & 1u) ? f[8 * j + i] * p.drop.scale : 0.0f;
              }
            }
          }
          if (ADD_TMA) {
            // residual tile staged by the producer: same 128-byte-row swizzled layout as the output staging
            const uint8_t* dbase = sadd + buf * Cfg::kOutBytes + (ctile >> 6) * (kBM * 128);
            const int du = (ctile & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 pk = lds128(smem_u32(dbase) + sw128_offset(rloc, du + j));
              const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                f[8 * j + 2 * i] += t2.x;
                f[8 * j + 2 * i + 1] += t2.y;
              }
            }
          } else if (p.add != nullptr && ks == 0) {
            if (cur_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t w[4] = {add_cur[j].x, add_cur[j].y, add_cur[j].z, add_cur[j].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                  f[8 * j + 2 * i] += t2.x;
                  f[8 * j + 2 * i + 1] += t2.y;
                }
              }
            } else if (p.add_bf16) {
              const __nv_bfloat16* ar = reinterpret_cast<const __nv_bfloat16*>(p.add) + (long long)row * p.ld_add + col0;
              if (full_chunk && ((reinterpret_cast<uintptr_t>(ar) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  const uint4 pk = __ldg(reinterpret_cast<const uint4*>(ar + j));
                  const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                    f[j + 2 * i] += t2.x;
                    f[j + 2 * i + 1] += t2.y;
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (full_chunk || col0 + j < p.N) f[j] += __bfloat162float(ar[j]);
              }
            } else {
              const float* ar = reinterpret_cast<const float*>(p.add) + (long long)row * p.ld_add + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full_chunk || col0 + j < p.N) f[j] += ar[j];
            }
          }
          if (p.out2 != nullptr) {
            __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(p.out2) + (long long)row * p.ldo2 + col0;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(o2) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                *reinterpret_cast<uint4*>(o2 + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                                               pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full_chunk || col0 + j < p.N) o2[j] = __float2bfloat16(f[j]);
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          } else if (p.act == 3) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = tanh_fast(f[j]);
          } else if (p.act == 6) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __frcp_rn(1.0f + ex2_approx(-1.4426950408889634f * f[j]));
          } else if (p.act == 4 || p.act == 5) {
            const __nv_bfloat16* xr = p.aux + (long long)row * p.ld_aux + col0;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(xr) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 pk = __ldg(reinterpret_cast<const uint4*>(xr + j));
                const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                  f[j + 2 * i] *= (p.act == 4) ? dgelu_erf(t2.x) : (t2.x > 0.0f ? 1.0f : 0.0f);
                  f[j + 2 * i + 1] *= (p.act == 4) ? dgelu_erf(t2.y) : (t2.y > 0.0f ? 1.0f : 0.0f);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (full_chunk || col0 + j < p.N) {
                  const float x = __bfloat162float(xr[j]);
                  f[j] *= (p.act == 4) ? dgelu_erf(x) : (x > 0.0f ? 1.0f : 0.0f);
                }
              }
            }
          }
        }
        if (p.tma_out) {
          // stage the bf16 row chunk: 128-byte rows, 16-byte units XOR-swizzled like the TMA store map expects
          const int region = ctile >> 6;                      // staged regions are 64 columns (128 bytes) wide
          const int unit0 = (ctile & 63) >> 3;                // first 16-byte unit of this chunk inside the row
          uint8_t* rbase = stile + region * (kBM * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 pk = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                        pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
            sts128(smem_u32(rbase) + sw128_offset(rloc, unit0 + j), pk);
          }
        } else if (live) {
          if (p.split_k > 1 || p.atomic_out) {
            float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)       // red.global.add.v4.f32: a quarter of the atomic instructions
                atomicAdd(reinterpret_cast<float4*>(o + j), make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full_chunk || col0 + j < p.N) atomicAdd(o + j, f[j]);
            }
          } else if (p.out_bf16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col0;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                                              pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = __float2bfloat16(f[j]);
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = f[j];
            }
          }
        }
        cur_ok = nxt_ok;
#pragma unroll
        for (int j = 0; j < 4; ++j) add_cur[j] = add_nxt[j];
      }
      // accumulator is consumed: hand the TMEM buffer (and the residual buffer) back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_leader(&tempty[buf]); else mbar_arrive(&tempty[buf]);
        if (ADD_TMA) mbar_arrive(&dempty[buf]);
      }
      if (p.tma_out) {
        fence_proxy_async_smem();                         // generic-proxy smem writes -> visible to the TMA engine
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (issuer && has_k) {
#pragma unroll
          for (int r = 0; r < BN / 64; ++r)
            tma_store_2d(&tmC, stile + r * (kBM * 128), n_blk * BN + r * 64, m_blk * kBM);
          tma_store_commit();
        }
        if (do_stats) {
          // column sums of the staged bf16 tile (rows / columns outside the problem were staged as zeros): a thread
          // owns one column pair and a group of rows; a warp reads 128 contiguous bytes of one row per instruction
          constexpr int CP = BN / 2, RG = 256 / CP, RPT = kBM / RG;
          const int cp = etid % CP, rg = etid / CP;
          const int col = cp * 2;
          const uint8_t* cbase = stile + (col >> 6) * (kBM * 128) + (col & 7) * 2;
          const int unit = (col & 63) >> 3;
          const uint32_t cbase_u32 = smem_u32(cbase);
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
          for (int r = rg * RPT; r < (rg + 1) * RPT; ++r) {
            const uint32_t raw = lds32(cbase_u32 + sw128_offset(r, unit));
            const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
            s0 += v.x; s1 += v.y;
            q0 = fmaf(v.x, v.x, q0); q1 = fmaf(v.y, v.y, q1);
          }
          // this thread's own fp64 slots (row group rg, columns col, col + 1): plain read-modify-write, no atomics
          const uint32_t ps = smem_u32(s_csum + rg * BN + col), pq = smem_u32(s_csq + rg * BN + col);
          sts_f64(ps, lds_f64(ps) + (double)s0);
          sts_f64(ps + 8, lds_f64(ps + 8) + (double)s1);
          sts_f64(pq, lds_f64(pq) + (double)q0);
          sts_f64(pq + 8, lds_f64(pq + 8) + (double)q1);
        }
      }
    }
    if (do_stats) {
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (acc_n_blk >= 0 && etid < BN && acc_n_blk * BN + etid < p.N) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int g = 0; g < kStatRG; ++g) { a += s_csum[g * BN + etid]; b += s_csq[g * BN + etid]; }
        atomicAdd(p.stats + acc_n_blk * BN + etid, a);
        atomicAdd(p.stats + p.N + acc_n_blk * BN + etid, b);
      }
    }
    // the staging buffers must have been READ before the CTA (and its shared memory) goes away; the global writes of
    // the bulk stores complete with the grid (kernel-boundary semantics), no need to sit on them here
    if (p.tma_out && issuer) tma_store_wait_read<0>();
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();     // pair: no CTA leaves while its peer may still signal it
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base); else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, bool A_MN, bool B_MN, bool ADD_TMA = false, bool STATS = false, bool CTA2 = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td,
                       const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, ADD_TMA, CTA2>;
  static_assert(Cfg::kStages >= 2 && Cfg::kSmemBytes <= 227 * 1024, "shared-memory budget");
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, ADD_TMA, STATS, CTA2>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute(%d B smem): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return CFL_ECUDA;
    }
    attr_set = true;
  }
  const int num_n = (p.N + BN - 1) / BN;
  if (!CTA2) {
    const int num_m = (p.M + kBM - 1) / kBM;
    const int units = num_m * num_n * p.split_k;
    const int grid = units < sm_count() ? units : sm_count();
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tc, td, p);
  } else {
    // one cluster of two CTAs (the two SMs of a TPC) per 256-row work unit, persistent over units
    const int num_m = (p.M + 2 * kBM - 1) / (2 * kBM);
    const int units = num_m * num_n * p.split_k;
    const int pairs = units < sm_count() / 2 ? units : sm_count() / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, td, p);
    if (e != cudaSuccess) {
      set_error("gemm_tc (CTA pair): cudaLaunchKernelEx: %s", cudaGetErrorString(e));
      return CFL_ECUDA;
    }
  }
  return check_launch("gemm_tc_kernel");
}

// CTA-pair (tcgen05 cta_group::2) eligibility of an [M, N, K] problem with tile width BN.
bool gemm_use_pair(int M, int N, int K, int a_mn, int b_mn, int BN) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("CREAMFL_GEMM_2CTA");
    enabled = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  (void)N; (void)K; (void)a_mn;
  if (!enabled || (sm_count() & 1)) return false;
  if (M < 2 * kBM) return false;
  if (b_mn && BN < 128) return false;          // MN-major B travels in 64-column blocks: half a tile needs >= 64
  return true;
}

// Tile width the launcher picks for an [M, N, K] problem (see gemm_bf16 below).
int gemm_tile_n(int N, int K) {
  return (N <= 64) ? 64 : ((((N >= 256 && N % 256 == 0) || N >= 1024) && K >= 512) ? 256 : 128);
}

// Split-K planner of the accumulating (weight-gradient) GEMMs.  A launch runs `tiles * split` units on sm_count()
// persistent CTAs, i.e. ceil(units / SMs) waves, each costing the unit's k-blocks plus its atomic epilogue
// (~4 k-block times, calibrated on the ResNet101 / BERT shapes with scripts/sweep_wgrad.py): pick the split with the
// smallest waves * (k-blocks per unit + epilogue).  The earlier "two waves of 128-wide tiles" rule produced e.g.
// 152 units on 148 SMs for the 23 + 22 1x1 convolutions of ResNet101's layer 3 (a second wave for 4 units).
int plan_split_k(long long tiles, long long nkb, long long min_per, double epi_kb) {
  const long long sms = sm_count();
  long long max_split = nkb / (min_per > 0 ? min_per : 1);
  if (max_split < 1) max_split = 1;
  if (max_split > 4 * sms) max_split = 4 * sms;
  int best = 1;
  double best_cost = 1e300;
  for (long long s = 1; s <= max_split; ++s) {
    const long long per = (nkb + s - 1) / s;
    if ((nkb + per - 1) / per != s) continue;          // not a distinct partition
    const long long waves = (tiles * s + sms - 1) / sms;
    const double cost = (double)waves * ((double)per + epi_kb);
    if (cost < best_cost * (1.0 - 1e-9)) {
      best_cost = cost;
      best = (int)s;
    }
  }
  return best;
}

int gemm_plan_split(int M, int N, int K) {
  const int BN = gemm_tile_n(N, K);
  const long long tiles = ((M + kBM - 1LL) / kBM) * ((N + BN - 1LL) / BN);
  return plan_split_k(tiles, (K + kBK - 1LL) / kBK, 4, 4.0);
}

// Host entry shared by the C ABI wrappers.  a/b are bf16.  a_mn / b_mn select the storage order (see top).
int gemm_bf16(const void* a, long long lda, int a_mn, const void* b, long long ldb, int b_mn, GemmParams p,
              cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) {
    set_error("gemm_bf16: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
    return CFL_EINVAL;
  }
  const int nkb = (p.K + kBK - 1) / kBK;
  if (p.split_k < 1) p.split_k = 1;
  if (p.split_k > nkb) p.split_k = nkb;
  {
    // every split must own at least one k-block
    const int per = (nkb + p.split_k - 1) / p.split_k;
    p.split_k = (nkb + per - 1) / per;
  }
  if ((p.split_k > 1 || p.atomic_out) && (p.out_bf16 || p.act != 0 || p.out2 != nullptr)) {
    set_error("gemm_bf16: split-K requires fp32 output without activation");
    return CFL_EINVAL;
  }
  if (p.drop.rng != nullptr && (p.split_k > 1 || p.atomic_out || (p.N & 7))) {
    set_error("gemm_bf16: dropout needs split_k == 1, a plain output and N %% 8 == 0 (N=%d)", p.N);
    return CFL_EINVAL;
  }
  if ((p.act == 4 || p.act == 5) && p.aux == nullptr) {
    set_error("gemm_bf16: act %d needs aux", p.act);
    return CFL_EINVAL;
  }
  // Narrow outputs waste MMA columns: pick the tile width from N.  128x256 tiles read 96 B/clk of operands from
  // shared memory per SM (128x128: 128 B/clk, the port limit), so wide outputs use them whenever N fills them.
  // (short-K problems are HBM-bound: they keep the 128-wide configuration with double-buffered output staging)
  int BN = gemm_tile_n(p.N, p.K);
  // residual operand through TMA: needs the 128-wide configuration (two residual + two output staging buffers)
  const bool add_tma = p.add != nullptr && p.add_bf16 && p.split_k == 1 && !a_mn && p.N > 64 &&
                       (reinterpret_cast<uintptr_t>(p.add) & 15) == 0 && ((p.ld_add * 2) & 15) == 0;
  if (add_tma) BN = 128;
  CUtensorMap ta, tb, tc, td;
  memset(&tc, 0, sizeof(tc));
  memset(&td, 0, sizeof(td));
  int rc;
  // bf16 outputs that TMA can address leave through a staged, fully coalesced TMA store
  p.tma_out = (p.out_bf16 && p.split_k == 1 && !p.atomic_out && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
               ((p.ldo * 2) & 15) == 0) ? 1 : 0;
  if (p.out2 != nullptr && p.ldo2 == 0) p.ldo2 = p.ldo;
  if (p.tma_out && (rc = make_tmap_2d(&tc, p.out, 2, p.M, p.N, p.ldo, 64, kBM))) return rc;
  if (add_tma && (rc = make_tmap_2d(&td, p.add, 2, p.M, p.N, p.ld_add, 64, kBM))) return rc;
  if (!a_mn)
    rc = make_tmap_2d(&ta, a, 2, p.M, p.K, lda, kBK, kBM);
  else
    rc = make_tmap_2d(&ta, a, 2, p.K, p.M, lda, 64, kBK);
  if (rc) return rc;
  const bool pair = gemm_use_pair(p.M, p.N, p.K, a_mn, b_mn, BN);
  if (!b_mn)
    rc = make_tmap_2d(&tb, b, 2, p.N, p.K, ldb, kBK, pair ? BN / 2 : BN);   // <= 256 rows per box; pair: half a tile
  else
    rc = make_tmap_2d(&tb, b, 2, p.K, p.N, ldb, 64, kBK);
  if (rc) return rc;

#define CFL_DISPATCH2(BNV, P2)                                                                          \
  do {                                                                                                  \
    if (!a_mn && !b_mn) return launch_gemm<BNV, false, false, false, false, P2>(ta, tb, tc, td, p, stream);  \
    if (!a_mn && b_mn) return launch_gemm<BNV, false, true, false, false, P2>(ta, tb, tc, td, p, stream);    \
    if (a_mn && !b_mn) return launch_gemm<BNV, true, false, false, false, P2>(ta, tb, tc, td, p, stream);    \
    return launch_gemm<BNV, true, true, false, false, P2>(ta, tb, tc, td, p, stream);                        \
  } while (0)
#define CFL_DISPATCH(BNV)            \
  do {                               \
    if (pair) CFL_DISPATCH2(BNV, true); \
    CFL_DISPATCH2(BNV, false);       \
  } while (0)
  if (p.stats != nullptr && (!p.tma_out || add_tma || p.act != 0 || p.ldo != p.N)) {
    // statistics cannot ride in the epilogue for this configuration: run the GEMM, then the stand-alone pass
    double* stats = p.stats;
    p.stats = nullptr;
    if (!p.out_bf16 || p.ldo != p.N) {
      set_error("gemm_bf16: BatchNorm statistics need a contiguous bf16 output");
      return CFL_EINVAL;
    }
    rc = gemm_bf16(a, lda, a_mn, b, ldb, b_mn, p, stream);
    if (rc) return rc;
    return bn_stats_only(p.out, p.M, p.N, stats, stream);
  }
  if (add_tma) {
    if (pair) {
      if (b_mn) return launch_gemm<128, false, true, true, false, true>(ta, tb, tc, td, p, stream);
      return launch_gemm<128, false, false, true, false, true>(ta, tb, tc, td, p, stream);
    }
    if (b_mn) return launch_gemm<128, false, true, true>(ta, tb, tc, td, p, stream);
    return launch_gemm<128, false, false, true>(ta, tb, tc, td, p, stream);
  }
  if (p.stats != nullptr) {
    if (a_mn || b_mn) {
      set_error("gemm_bf16: fused statistics are implemented for K-major operands (forward convolutions)");
      return CFL_EINVAL;
    }
    if (pair) {
      if (BN == 64) return launch_gemm<64, false, false, false, true, true>(ta, tb, tc, td, p, stream);
      if (BN == 256) return launch_gemm<256, false, false, false, true, true>(ta, tb, tc, td, p, stream);
      return launch_gemm<128, false, false, false, true, true>(ta, tb, tc, td, p, stream);
    }
    if (BN == 64) return launch_gemm<64, false, false, false, true>(ta, tb, tc, td, p, stream);
    if (BN == 256) return launch_gemm<256, false, false, false, true>(ta, tb, tc, td, p, stream);
    return launch_gemm<128, false, false, false, true>(ta, tb, tc, td, p, stream);
  }
  if (BN == 64) CFL_DISPATCH(64);
  if (BN == 256) CFL_DISPATCH(256);
  CFL_DISPATCH(128);
#undef CFL_DISPATCH
#undef CFL_DISPATCH2
}

}  // namespace cfl
