// Instantiations of the tcgen05 GEMM (gemm_tc_kernel.cuh) with the specialised epilogues:
//   kEpiStore  - forward / data-gradient GEMMs whose bf16 output leaves through the staged TMA store (1x1 convolutions
//                with fused BatchNorm statistics, data gradients with the TMA-staged residual, plain linears + bias)
//   kEpiAtomic - split-K weight gradients (fp32 red.global.add.v4)
// The general epilogue decides bias / activation / dropout / residual / output type per element group at run time
// and spent ~9 instructions per output element on the 8 epilogue warps; these paths spend ~1-2 (+3.5 for the
// statistics), which is what bounded the short-K launches of a training step (profiles/r02_ncu_gemm_cases.md).
#include "gemm_tc_kernel.cuh"

namespace cfl {

template <int BN, bool B_MN, bool ADD_TMA, bool STATS>
static int store_pair(bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                      const CUtensorMap& td, const GemmParams& p, cudaStream_t stream) {
  if (pair) return launch_gemm<BN, false, B_MN, ADD_TMA, STATS, true, kEpiStore>(ta, tb, tc, td, p, stream);
  return launch_gemm<BN, false, B_MN, ADD_TMA, STATS, false, kEpiStore>(ta, tb, tc, td, p, stream);
}

int launch_gemm_store(int BN, bool b_mn, bool add_tma, bool stats, bool pair, const CUtensorMap& ta,
                      const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& td, const GemmParams& p,
                      cudaStream_t stream) {
  if (add_tma) {          // the residual configuration is 128 wide (two residual + two output staging buffers)
    if (b_mn) return store_pair<128, true, true, false>(pair, ta, tb, tc, td, p, stream);
    return store_pair<128, false, true, false>(pair, ta, tb, tc, td, p, stream);
  }
  if (stats) {            // fused statistics: forward convolutions (K-major operands)
    if (BN == 64) return store_pair<64, false, false, true>(pair, ta, tb, tc, td, p, stream);
    if (BN == 256) return store_pair<256, false, false, true>(pair, ta, tb, tc, td, p, stream);
    return store_pair<128, false, false, true>(pair, ta, tb, tc, td, p, stream);
  }
  if (b_mn) {
    if (BN == 64) return store_pair<64, true, false, false>(pair, ta, tb, tc, td, p, stream);
    if (BN == 256) return store_pair<256, true, false, false>(pair, ta, tb, tc, td, p, stream);
    return store_pair<128, true, false, false>(pair, ta, tb, tc, td, p, stream);
  }
  if (BN == 64) return store_pair<64, false, false, false>(pair, ta, tb, tc, td, p, stream);
  if (BN == 256) return store_pair<256, false, false, false>(pair, ta, tb, tc, td, p, stream);
  return store_pair<128, false, false, false>(pair, ta, tb, tc, td, p, stream);
}

// BERT FFN: forward GELU (K-major weights, + pre-activation output) and the data gradient through it (MN-major
// weights, multiply by GELU'(pre-activation)); 256-wide tiles (N = 3072, K >= 512)
int launch_gemm_gelu(bool dgelu, bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                     const CUtensorMap& td, const GemmParams& p, cudaStream_t stream) {
  if (dgelu) {
    if (pair) return launch_gemm<256, false, true, false, false, true, kEpiStoreDgelu>(ta, tb, tc, td, p, stream);
    return launch_gemm<256, false, true, false, false, false, kEpiStoreDgelu>(ta, tb, tc, td, p, stream);
  }
  if (pair) return launch_gemm<256, false, false, false, false, true, kEpiStoreGelu>(ta, tb, tc, td, p, stream);
  return launch_gemm<256, false, false, false, false, false, kEpiStoreGelu>(ta, tb, tc, td, p, stream);
}

template <int BN>
static int atomic_pair(bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                       const CUtensorMap& td, const GemmParams& p, cudaStream_t stream) {
  if (pair) return launch_gemm<BN, true, true, false, false, true, kEpiAtomic>(ta, tb, tc, td, p, stream);
  return launch_gemm<BN, true, true, false, false, false, kEpiAtomic>(ta, tb, tc, td, p, stream);
}

int launch_gemm_atomic(int BN, bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                       const CUtensorMap& td, const GemmParams& p, cudaStream_t stream) {
  if (BN == 64) return atomic_pair<64>(pair, ta, tb, tc, td, p, stream);
  if (BN == 256) return atomic_pair<256>(pair, ta, tb, tc, td, p, stream);
  return atomic_pair<128>(pair, ta, tb, tc, td, p, stream);
}

}  // namespace cfl
