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

#include "gemm_tc_kernel.cuh"

namespace cfl {

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

// CREAMFL_GEMM_LEAN=0 routes every launch through the general epilogue (A/B switch for measurements and tests).
bool gemm_lean_enabled() {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("CREAMFL_GEMM_LEAN");
    enabled = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return enabled != 0;
}

// Tile width the launcher picks for an [M, N, K] problem (see gemm_bf16 below).
int gemm_tile_n(int N, int K) {
  static int min_k = -1;      // smallest K served by 256-wide tiles (CREAMFL_GEMM_BN256_MINK overrides: measurements)
  if (min_k < 0) {
    const char* e = getenv("CREAMFL_GEMM_BN256_MINK");
    min_k = e ? atoi(e) : 512;
  }
  return (N <= 64) ? 64 : ((((N >= 256 && N % 256 == 0) || N >= 1024) && K >= min_k) ? 256 : 128);
}

// Split-K planner of the accumulating (weight-gradient) GEMMs.  A launch runs `tiles * split` units on sm_count()
// persistent CTAs, i.e. ceil(units / SMs) waves, each costing the unit's k-blocks plus its atomic epilogue
// (~4 k-block times, calibrated on the ResNet101 / BERT shapes with scripts/sweep_wgrad.py): pick the split with the
// smallest waves * (k-blocks per unit + epilogue).  The earlier "two waves of 128-wide tiles" rule produced e.g.
// 152 units on 148 SMs for the 23 + 22 1x1 convolutions of ResNet101's layer 3 (a second wave for 4 units).
int plan_split_k(long long tiles, long long nkb, long long min_per, double epi_kb, int workers) {
  const long long sms = workers > 0 ? workers : sm_count();
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
  // specialised epilogues (gemm_tc_lean.cu) for the launches that dominate a training step
  const bool lean_base = gemm_lean_enabled() && p.alpha == 1.0f && p.drop.rng == nullptr && p.out2 == nullptr &&
                         (p.N & 31) == 0;
  const bool lean_common = lean_base && p.act == 0;
  if (lean_base && (p.act == 0 || (p.act == 2 && p.stats == nullptr)) && p.tma_out && !a_mn && (p.add == nullptr || add_tma) && (p.stats == nullptr || !b_mn) &&
      (p.bias == nullptr || (p.stats == nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)))
    return launch_gemm_store(BN, b_mn != 0, add_tma, p.stats != nullptr, pair, ta, tb, tc, td, p, stream);
  // GELU (forward, K-major B, optional pre-activation output) / dGELU (data gradient, MN-major B) on 256-wide tiles
  if (gemm_lean_enabled() && BN == 256 && p.tma_out && !a_mn && p.alpha == 1.0f && p.drop.rng == nullptr &&
      p.add == nullptr && p.stats == nullptr && (p.N & 31) == 0 &&
      (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
    if (p.act == 1 && !b_mn && (p.out2 == nullptr || ((reinterpret_cast<uintptr_t>(p.out2) & 15) == 0 && (p.ldo2 & 7) == 0)))
      return launch_gemm_gelu(false, pair, ta, tb, tc, td, p, stream);
    if (p.act == 4 && b_mn && p.out2 == nullptr && p.bias == nullptr && (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0 &&
        (p.ld_aux & 7) == 0)
      return launch_gemm_gelu(true, pair, ta, tb, tc, td, p, stream);
  }
  if (lean_common && a_mn && b_mn && !p.out_bf16 && (p.split_k > 1 || p.atomic_out) && p.bias == nullptr &&
      p.add == nullptr && (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
    return launch_gemm_atomic(BN, pair, ta, tb, tc, td, p, stream);
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
