// Kernel template of the tcgen05 GEMM (see gemm_tc.cu for the overview); instantiated from gemm_tc.cu (general
// epilogue) and gemm_tc_lean.cu (specialised epilogues).
#pragma once
#include "kernels.cuh"
#include "ptx.cuh"
#include <string.h>
#include <stdlib.h>

namespace cfl {


constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 384;   // 4 control warps + 8 epilogue warps

// Epilogue flavours (compile-time): the general one decides everything at run time (bias / activation / dropout /
// residual / second output / fp32 or bf16 / atomics); the specialised ones serve the shapes that dominate a training
// step with a fraction of the instructions (the epilogue warps, not the tensor pipe, bounded those launches:
// profiles/r02_ncu_gemm_cases.md).
constexpr int kEpiGeneral = 0;
constexpr int kEpiStore = 1;    // bf16 tile -> swizzled staging -> TMA store; alpha = 1, optional bias, optional TMA residual,
                                // optional BatchNorm statistics; N % 32 == 0
constexpr int kEpiStoreGelu = 3;   // kEpiStore + erf-GELU of (acc + bias); optional second output = the pre-activation
constexpr int kEpiStoreDgelu = 4;  // kEpiStore + multiply by GELU'(aux) (aux = bf16 pre-activation of the forward pass)
constexpr int kEpiAtomic = 2;   // fp32 red.global.add.v4 (split-K weight gradients); alpha = 1, nothing else; N % 32 == 0

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
// bf16 outputs: Phi(x) = 1 / (1 + 2^(x q(x^2))) with q a quadratic in x^2 fitted to atanh(erf(x / sqrt 2)) / x
// (max |error| of Phi and of x Phi over the reals: 3.1e-5, an order of magnitude below the bf16 rounding of the
// result; relative accuracy holds in the tails because the sigmoid form has no cancellation).  9 instructions, 2 MUFU
// instead of 17 / 2: the FFN1 epilogue, not the MMAs, bounded that launch.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float phi_cdf_fast(float x) {
  const float u = fminf(x * x, 54.0f);                       // q(u) is monotone up to u = 54; beyond it Phi is 0 / 1
  const float qv = fmaf(fmaf(9.844209789e-04f, u, -1.065445952e-01f), u, -2.301466763e+00f);
  return rcp_approx(1.0f + ex2_approx(x * qv));
}
__device__ __forceinline__ float gelu_fast(float x) { return x * phi_cdf_fast(x); }
__device__ __forceinline__ float dgelu_fast(float x) {
  const float pdf = ex2_approx(-0.72134752044448170f * x * x);
  return fmaf(0.3989422804014327f * x, pdf, phi_cdf_fast(x));
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
template <int BN, bool A_MN, bool B_MN, bool ADD_TMA, bool STATS, bool CTA2, int EPI = kEpiGeneral>
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
   if constexpr (EPI == kEpiStore || EPI == kEpiStoreGelu || EPI == kEpiStoreDgelu) {
    // ------------------------------------------------------------ specialised epilogue: bf16 tile -> TMA store
    // (alpha = 1, N % 32 == 0, optional bias / TMA-staged residual / BatchNorm statistics; checked by the host)
    constexpr int HC = BN / 2;             // columns per thread
    constexpr int NCH = HC / 32;           // 32-column chunks per thread
    const int q = warp & 3;                // TMEM lane group of this warp
    const int half = (warp - 4) >> 2;      // column half
    const int rloc = q * 32 + lane;
    const uint32_t x7 = rloc & 7;
    const bool issuer = (warp == 4 && lane == 0);
    const bool do_stats = STATS && !ADD_TMA && p.stats != nullptr;
    const int etid = threadIdx.x - 128;
    int acc_n_blk = -1;
    constexpr int kStatRG = 512 / BN;      // row groups of the column-sum pass
    // column-sum pass geometry: a thread owns one column pair and RPT consecutive rows of the staged tile
    constexpr int CP = BN / 2, RPT = kBM / (256 / CP);
    const int s_col = (etid % CP) * 2, s_rg = etid / CP;
    const uint32_t s_unit = (s_col & 63) >> 3;
    const uint32_t s_base = (s_col >> 6) * (kBM * 128) + (s_col & 7) * 2 + s_rg * RPT * 128;
    if (do_stats) {
      for (int e = etid; e < 512; e += 256) { s_csum[e] = 0.0; s_csq[e] = 0.0; }
      asm volatile("bar.sync 3, 256;" ::: "memory");
    }
    auto flush_stats = [&](int nb) {
      if (etid < BN && nb * BN + etid < p.N) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int g = 0; g < kStatRG; ++g) { a += s_csum[g * BN + etid]; b += s_csq[g * BN + etid]; }
        atomicAdd(p.stats + nb * BN + etid, a);
        atomicAdd(p.stats + p.N + nb * BN + etid, b);
      }
    };
    int it = 0;
    for (int u = wid; u < units; u += wstride, ++it) {
      const int m_blk = row_block(u);
      const int n_blk = (u / num_m) % num_n;
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const uint32_t stile = smem_u32(sout + (Cfg::kOutBufs == 2 ? buf : 0) * Cfg::kOutBytes);
      // the TMA store that last read this staging buffer must have drained
      if (issuer) tma_store_wait_read<Cfg::kOutBufs - 1>();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (do_stats && acc_n_blk != n_blk) {
        if (acc_n_blk >= 0) {
          flush_stats(acc_n_blk);
          asm volatile("bar.sync 3, 256;" ::: "memory");
          for (int e = etid; e < 512; e += 256) { s_csum[e] = 0.0; s_csq[e] = 0.0; }
          asm volatile("bar.sync 3, 256;" ::: "memory");
        }
        acc_n_blk = n_blk;
      }
      // dGELU: this thread's 32 pre-activations of a chunk (64 contiguous bytes) are fetched one chunk ahead, the
      // first chunk before the accumulator is ready
      const int row = m_blk * kBM + rloc;
      const bool row_ok = row < p.M;
      const __nv_bfloat16* aux_row = nullptr;
      uint4 aux_v[2][4];
      if constexpr (EPI == kEpiStoreDgelu) {
        aux_row = p.aux + (long long)row * p.ld_aux + n_blk * BN + half * HC;
        if (row_ok && n_blk * BN + half * HC < p.N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) aux_v[0][j] = __ldg(reinterpret_cast<const uint4*>(aux_row) + j);
        }
      }
      if (ADD_TMA) mbar_wait(&dfull[buf], bphase);
      mbar_wait(&tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + half * HC;
      uint32_t v[2][32];
      tmem_ld_32x32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tmem_ld_wait();
        if (c + 1 < NCH) {
          tmem_ld_32x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);     // in flight while chunk c is processed
          if constexpr (EPI == kEpiStoreDgelu) {
            if (row_ok && n_blk * BN + half * HC + (c + 1) * 32 < p.N) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                aux_v[(c + 1) & 1][j] = __ldg(reinterpret_cast<const uint4*>(aux_row + (c + 1) * 32) + j);
            }
          }
        } else {
          // the accumulator now lives in registers: hand the TMEM buffer back before the arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CTA2) mbar_arrive_leader(&tempty[buf]); else mbar_arrive(&tempty[buf]); }
        }
        const uint32_t(&w)[32] = v[c & 1];
        const int ctile = half * HC + c * 32;          // column offset inside the tile
        const int col0 = n_blk * BN + ctile;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(w[j]);
        if (p.bias != nullptr && col0 < p.N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
            f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
          }
        }
        if constexpr (EPI == kEpiStoreGelu) {
          if (p.out2 != nullptr && row_ok && col0 < p.N) {      // pre-activation for the backward pass
            uint4* o2 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + (long long)row * p.ldo2 + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o2[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
        }
        if constexpr (EPI == kEpiStoreDgelu) {
          if (row_ok && col0 < p.N) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 pk = aux_v[c & 1][j];
              const uint32_t r4[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                f[8 * j + 2 * i] *= dgelu_fast(__uint_as_float(r4[i] << 16));
                f[8 * j + 2 * i + 1] *= dgelu_fast(__uint_as_float(r4[i] & 0xffff0000u));
              }
            }
          }
        }
        const uint32_t region = static_cast<uint32_t>(ctile >> 6) * (kBM * 128) + rloc * 128u;
        const uint32_t unit0 = (ctile & 63) >> 3;
        if (ADD_TMA) {
          // residual tile staged by the producer: same 128-byte-row swizzled layout as the output staging
          const uint32_t dbase = smem_u32(sadd + buf * Cfg::kOutBytes) + region;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 pk = lds128(dbase + (((unit0 + j) ^ x7) << 4));
            const uint32_t r4[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              f[8 * j + 2 * i] += __uint_as_float(r4[i] << 16);
              f[8 * j + 2 * i + 1] += __uint_as_float(r4[i] & 0xffff0000u);
            }
          }
        }
        if (EPI == kEpiStore && p.act == 2) {      // ReLU after the residual (eval-mode conv + folded BatchNorm)
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 pk = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                      pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
          sts128(stile + region + (((unit0 + j) ^ x7) << 4), pk);
        }
      }
      if (ADD_TMA) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&dempty[buf]);
      }
      fence_proxy_async_smem();                         // generic-proxy smem writes -> visible to the TMA engine
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (issuer) {
#pragma unroll
        for (int r = 0; r < BN / 64; ++r)
          tma_store_2d(&tmC, sout + (Cfg::kOutBufs == 2 ? buf : 0) * Cfg::kOutBytes + r * (kBM * 128), n_blk * BN + r * 64,
                       m_blk * kBM);
        tma_store_commit();
      }
      if (do_stats) {
        // column sums of the staged bf16 tile (rows / columns outside the problem were staged as zeros); row
        // r0 + 8 i + j sits at byte (8 i + j) * 128 + ((unit ^ j) << 4) of its region: eight base addresses, the rest
        // are immediate offsets
        uint32_t xo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) xo[j] = stile + s_base + ((s_unit ^ j) << 4);
        float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f}, q0[2] = {0.f, 0.f}, q1[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < RPT / 8; ++i) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t raw = lds32(xo[j] + (8 * i + j) * 128);       // ptxas folds the constant into the LDS offset
            const float a = __uint_as_float(raw << 16), b = __uint_as_float(raw & 0xffff0000u);
            s0[j & 1] += a; s1[j & 1] += b;
            q0[j & 1] = fmaf(a, a, q0[j & 1]); q1[j & 1] = fmaf(b, b, q1[j & 1]);
          }
        }
        // this thread's own fp64 slots (row group, columns s_col, s_col + 1): plain read-modify-write, no atomics
        const uint32_t ps = smem_u32(s_csum + s_rg * BN + s_col), pq = smem_u32(s_csq + s_rg * BN + s_col);
        sts_f64(ps, lds_f64(ps) + (double)(s0[0] + s0[1]));
        sts_f64(ps + 8, lds_f64(ps + 8) + (double)(s1[0] + s1[1]));
        sts_f64(pq, lds_f64(pq) + (double)(q0[0] + q0[1]));
        sts_f64(pq + 8, lds_f64(pq + 8) + (double)(q1[0] + q1[1]));
      }
    }
    if (do_stats) {
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (acc_n_blk >= 0) flush_stats(acc_n_blk);
    }
    if (issuer) tma_store_wait_read<0>();
   } else if constexpr (EPI == kEpiAtomic) {
    // ------------------------------------------------------------ specialised epilogue: fp32 reds (weight gradients)
    constexpr int HC = BN / 2;
    constexpr int NCH = HC / 32;
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int rloc = q * 32 + lane;
    int it = 0;
    for (int u = wid; u < units; u += wstride, ++it) {
      const int m_blk = row_block(u);
      const int n_blk = (u / num_m) % num_n;
      const int ks = u / (num_m * num_n);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const bool has_k = ks * kb_per < nkb;
      const int row = m_blk * kBM + rloc;
      mbar_wait(&tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + half * HC;
      float* orow = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + n_blk * BN + half * HC;
      uint32_t v[2][32];
      tmem_ld_32x32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tmem_ld_wait();
        if (c + 1 < NCH) {
          tmem_ld_32x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CTA2) mbar_arrive_leader(&tempty[buf]); else mbar_arrive(&tempty[buf]); }
        }
        const uint32_t(&w)[32] = v[c & 1];
        const int col0 = n_blk * BN + half * HC + c * 32;
        if (row < p.M && col0 < p.N && has_k) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)       // red.global.add.v4.f32
            atomicAdd(reinterpret_cast<float4*>(orow + c * 32 + j),
                      make_float4(__uint_as_float(w[j]), __uint_as_float(w[j + 1]), __uint_as_float(w[j + 2]),
                                  __uint_as_float(w[j + 3])));
        }
      }
    }
   } else {
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
            for (int j = 0; j < 32; ++j) f[j] = p.out_bf16 ? gelu_fast(f[j]) : gelu_erf(f[j]);
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
                  f[j + 2 * i] *= (p.act == 4) ? (p.out_bf16 ? dgelu_fast(t2.x) : dgelu_erf(t2.x)) : (t2.x > 0.0f ? 1.0f : 0.0f);
                  f[j + 2 * i + 1] *= (p.act == 4) ? (p.out_bf16 ? dgelu_fast(t2.y) : dgelu_erf(t2.y)) : (t2.y > 0.0f ? 1.0f : 0.0f);
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
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();     // pair: no CTA leaves while its peer may still signal it
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base); else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, bool A_MN, bool B_MN, bool ADD_TMA = false, bool STATS = false, bool CTA2 = false,
          int EPI = kEpiGeneral>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td,
                       const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, ADD_TMA, CTA2>;
  static_assert(Cfg::kStages >= 2 && Cfg::kSmemBytes <= 227 * 1024, "shared-memory budget");
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, ADD_TMA, STATS, CTA2, EPI>;
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

// gemm_tc_lean.cu: instantiations with the specialised epilogues
int launch_gemm_store(int BN, bool b_mn, bool add_tma, bool stats, bool pair, const CUtensorMap& ta,
                      const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td, const GemmParams& p,
                      cudaStream_t stream);
int launch_gemm_gelu(bool dgelu, bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                     const CUtensorMap& td, const GemmParams& p, cudaStream_t stream);      // 256-wide tiles
int launch_gemm_atomic(int BN, bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                       const CUtensorMap& td, const GemmParams& p, cudaStream_t stream);
bool gemm_lean_enabled();
}  // namespace cfl
