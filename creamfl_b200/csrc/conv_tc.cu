// Convolution as implicit GEMM on tcgen05 (NHWC bf16 activations, [Cout, R, S, Cin] bf16 filters, fp32 accumulate).
//
// Replaces the cuDNN fprop / dgrad / wgrad calls reached from the reference through torchvision ResNet101
// (src/networks/models/image_encoder.py:24,55) and the client ResNet18 (src/networks/resnet_client.py:164-172).
//
// Stride-1 "same" convolutions (every 3x3 of ResNet except three) never materialise an im2col matrix: the A
// operand of the GEMM is fetched by TMA as a 4-D box (64 channels x bw x bh x bn pixels) of the activation
// tensor, shifted by the filter tap; out-of-image pixels are zero-filled by the TMA unit, which is exactly the
// zero padding of the convolution.
//
//   MODE 0  fprop : Y[p, co]  = sum_{tap, ci} X[p + tap, ci] * Wt[co, tap, ci]      M = pixels, N = Cout, K = taps*Cin
//   MODE 1  dgrad : dX[p, ci] = sum_{tap, co} dY[p - tap, co] * Wt[co, tap, ci]     M = pixels, N = Cin,  K = taps*Cout
//   MODE 2  wgrad : dW[co, tap, ci] += sum_p dY[p, co] * X[p + tap, ci]             M = Cout,   N = Cin,  K = pixels
//
// Same warp-specialised pipeline as gemm_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue, double-buffered accumulators in TMEM, persistent over output tiles.  Strided convolutions and
// the 7x7 stem go through an explicit im2col buffer + gemm_tc (they are 4 % of the FLOPs).
#include "kernels.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace cfl {

struct ConvParams {
  int N, H, W;           // OUTPUT extent (== input extent for stride 1, same padding)
  int stride;            // fprop / wgrad: input pixel = output pixel * stride + tap - pad (TMA element strides sample
                         // the input box)
  int up;                // dgrad of a stride-2 convolution (1: stride 1): N, H, W describe dY; dX is [N, 2H, 2W, Cin] and
                         // splits into 4 parity classes (h % 2, w % 2), each a dense stride-1 problem over the taps
                         // r = (parity + pad) mod 2, +2, ... that reach it: dX[2h'+ph, 2w'+pw] += dY[h' + dh, w' + dw] W[r, s]
                         // with dh = (ph + pad - r) / 2.  A work unit = (pixel tile of dY, class, channel block).
  int Cin, Cout;
  int R, S, pad_h, pad_w;
  int bw, bh, bn;        // pixel box of one tile (product 128 for fprop/dgrad, 64 for wgrad)
  int tiles_w, tiles_h, tiles_n;
  int split_k;           // wgrad only
  void* out;             // bf16 [N,H,W,Cout|Cin] (fprop/dgrad) or fp32 [Cout, R*S*Cin] accumulated (wgrad)
  const __nv_bfloat16* add;  // optional bf16 tensor added to the fprop/dgrad output (same layout as out)
  const float* bias;         // optional per-output-channel bias (fprop: folded eval-mode BatchNorm shift)
  int relu;                  // fprop: ReLU after bias / residual
};

constexpr int kCBM = 128;
constexpr int kCBK = 64;

// CTA2: two CTAs of a cluster (the two SMs of a TPC) own one 256 x BN tile (tcgen05 cta_group::2): each CTA stages
// its own 128 rows of A and HALF of the B tile, so the L2 -> shared-memory traffic per output element halves against
// the 128 x 128 single-CTA tile.  The 3x3 convolutions at 14 x 14 ran at the L2 throughput cap (462 MB per launch at
// 11.9 TB/s, profiles/r02_ncu_gemm_cases.md), not at the tensor pipe.
template <int BN, bool CTA2 = false>
struct ConvCfg {
  static constexpr int kABytes = kCBM * kCBK * 2;
  static constexpr int kBRows = CTA2 ? BN / 2 : BN;
  static constexpr int kBBytes = kBRows * kCBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = CTA2 ? ((BN == 256) ? 6 : 8) : ((BN == 128) ? 6 : 8);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static constexpr uint32_t kTmemCols = (BN == 256) ? 512 : ((BN == 128) ? 256 : 128);
};

// taps of parity class (ph, pw) of a stride-2 data gradient: r = r0 + 2 i (i < nr), s = s0 + 2 j (j < ns)
struct ClassTaps {
  int ph, pw, r0, nr, s0, ns;
};
__device__ __forceinline__ ClassTaps class_taps(const ConvParams& p, int pc) {
  ClassTaps c;
  c.ph = pc >> 1; c.pw = pc & 1;
  c.r0 = (c.ph + p.pad_h) & 1; c.nr = (p.R - c.r0 + 1) >> 1;
  c.s0 = (c.pw + p.pad_w) & 1; c.ns = (p.S - c.s0 + 1) >> 1;
  return c;
}

template <int BN, int MODE, bool CTA2>
__global__ void __launch_bounds__(256, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvParams p) {
  using Cfg = ConvCfg<BN, CTA2>;
  const int cta_rank = CTA2 ? (int)cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int wid = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // worker (CTA or CTA pair) index
  const int wstride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.R * p.S;
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;

  // GEMM view of this mode
  const int n_extent = (MODE == 0) ? p.Cout : p.Cin;                 // N of the GEMM
  const int num_n = (n_extent + BN - 1) / BN;
  int num_m, nkb, kb_per, units;
  // pair: num_m counts 256-row blocks; rank r of the pair owns the 128-row block 2 * m + r
  constexpr int kRowsPerUnit = CTA2 ? 2 * kCBM : kCBM;
  if (MODE == 2) {
    num_m = (p.Cout + kRowsPerUnit - 1) / kRowsPerUnit;
    nkb = pix_tiles;
    kb_per = (nkb + p.split_k - 1) / p.split_k;
    units = num_m * num_n * taps * p.split_k;
  } else {
    num_m = CTA2 ? (pix_tiles + 1) / 2 : pix_tiles;
    const int cred = (MODE == 0) ? p.Cin : p.Cout;
    nkb = taps * (cred / kCBK);
    kb_per = nkb;
    units = num_m * num_n * ((MODE == 1 && p.up == 2) ? 4 : 1);
  }
  const bool upx = (MODE == 1 && p.up == 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], CTA2 ? 8 : 4);      // pair: the epilogue warps of BOTH CTAs release the leader's buffer
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
      // pair: both CTAs' bytes are counted on the LEADER's barrier (the MMA issuer waits there)
      auto expect = [&](int st) {
        if (!CTA2) mbar_arrive_expect_tx(&full[st], Cfg::kStageBytes);
        else if (leader) mbar_arrive_expect_tx(&full[st], 2 * Cfg::kStageBytes);
      };
      auto load4 = [&](const CUtensorMap* m, int st, void* dst, int c0, int c1, int c2, int c3) {
        if (CTA2) tma_load_4d_pair(m, &full[st], dst, c0, c1, c2, c3); else tma_load_4d(m, &full[st], dst, c0, c1, c2, c3);
      };
      auto load2 = [&](const CUtensorMap* m, int st, void* dst, int c0, int c1) {
        if (CTA2) tma_load_2d_pair(m, &full[st], dst, c0, c1); else tma_load_2d(m, &full[st], dst, c0, c1);
      };
      const int b_off = CTA2 ? cta_rank * (BN / 2) : 0;      // this CTA's slice of the B tile
      for (int u = wid; u < units; u += wstride) {
        if (MODE != 2) {
          const int m_blk = CTA2 ? 2 * (u % num_m) + cta_rank : u % num_m;
          const int rest = u / num_m;
          const int n_blk = upx ? rest >> 2 : rest;
          const ClassTaps ct = class_taps(p, upx ? rest & 3 : 0);
          const int w0 = (m_blk % p.tiles_w) * p.bw;
          const int h0 = ((m_blk / p.tiles_w) % p.tiles_h) * p.bh;
          const int n0 = (m_blk / (p.tiles_w * p.tiles_h)) * p.bn;
          const int cred = (MODE == 0) ? p.Cin : p.Cout;
          const int cchunks = cred / kCBK;
          const int nkb_u = upx ? ct.nr * ct.ns * cchunks : nkb;
          for (int kb = 0; kb < nkb_u; ++kb) {
            const int c0 = (kb % cchunks) * kCBK;
            int tap = kb / cchunks, r, s, dh, dw;
            if (upx) {
              r = ct.r0 + 2 * (tap / ct.ns); s = ct.s0 + 2 * (tap % ct.ns);
              tap = r * p.S + s;
              dh = (ct.ph + p.pad_h - r) >> 1; dw = (ct.pw + p.pad_w - s) >> 1;
            } else {
              r = tap / p.S; s = tap % p.S;
              dh = (MODE == 0) ? (r - p.pad_h) : (p.pad_h - r);
              dw = (MODE == 0) ? (s - p.pad_w) : (p.pad_w - s);
            }
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            expect(stage);
            load4(&tmA, stage, sa, c0, w0 * p.stride + dw, h0 * p.stride + dh, n0);
            if (MODE == 0) {
              load2(&tmB, stage, sb, tap * p.Cin + c0, n_blk * BN + b_off);
            } else {
#pragma unroll
              for (int j = 0; j < Cfg::kBRows / 64; ++j)
                load2(&tmB, stage, sb + j * 8192, tap * p.Cin + n_blk * BN + b_off + j * 64, c0);
            }
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        } else {
          int t = u;
          const int m_blk = CTA2 ? 2 * (t % num_m) + cta_rank : t % num_m; t /= num_m;
          const int n_blk = t % num_n; t /= num_n;
          const int tap = t % taps;
          const int ks = t / taps;
          const int r = tap / p.S, s = tap % p.S;
          const int dh = r - p.pad_h, dw = s - p.pad_w;
          const int kb0 = ks * kb_per;
          const int kb1 = min(nkb, kb0 + kb_per);
          for (int kb = kb0; kb < kb1; ++kb) {
            const int w0 = (kb % p.tiles_w) * p.bw;
            const int h0 = ((kb / p.tiles_w) % p.tiles_h) * p.bh;
            const int n0 = (kb / (p.tiles_w * p.tiles_h)) * p.bn;
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kABytes;
            expect(stage);
#pragma unroll
            for (int j = 0; j < kCBM / 64; ++j)
              load4(&tmA, stage, sa + j * 8192, m_blk * kCBM + j * 64, w0, h0, n0);
#pragma unroll
            for (int j = 0; j < Cfg::kBRows / 64; ++j)
              load4(&tmB, stage, sb + j * 8192, n_blk * BN + b_off + j * 64, w0 * p.stride + dw, h0 * p.stride + dh, n0);
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------ MMA issuer (pair: the leader CTA only)
    if (elect_one()) {
      constexpr bool A_MN = (MODE == 2);
      constexpr bool B_MN = (MODE != 0);
      constexpr uint32_t idesc = make_idesc(1, CTA2 ? 2 * kCBM : kCBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = wid; u < units; u += wstride, ++it) {
        int kb0 = 0, kb1 = nkb;
        if (MODE == 2) {
          const int ks = u / (num_m * num_n * taps);
          kb0 = ks * kb_per;
          kb1 = min(nkb, kb0 + kb_per);
        } else if (upx) {
          const ClassTaps ct = class_taps(p, (u / num_m) & 3);
          kb1 = ct.nr * ct.ns * (p.Cout / kCBK);
        }
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
          for (int k = 0; k < kCBK / 16; ++k) {
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
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;
    int it = 0;
    for (int u = wid; u < units; u += wstride, ++it) {
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const int rloc = q * 32 + lane;
      bool row_ok;
      long long row_off;   // element offset of this thread's output row
      int n_blk;
      bool has_k = true;
      if (MODE != 2) {
        const int m_blk = CTA2 ? 2 * (u % num_m) + cta_rank : u % num_m;
        const int rest = u / num_m;
        n_blk = upx ? rest >> 2 : rest;
        const int w = (m_blk % p.tiles_w) * p.bw + rloc % p.bw;
        const int h = ((m_blk / p.tiles_w) % p.tiles_h) * p.bh + (rloc / p.bw) % p.bh;
        const int n = (m_blk / (p.tiles_w * p.tiles_h)) * p.bn + rloc / (p.bw * p.bh);
        row_ok = (w < p.W) && (h < p.H) && (n < p.N);
        if (upx) {       // this class's pixel of the up-sampled gradient map
          const int ph = (rest & 3) >> 1, pw = rest & 1;
          row_off = (((long long)n * (2 * p.H) + 2 * h + ph) * (2 * p.W) + 2 * w + pw) * n_extent;
        } else {
          row_off = (((long long)n * p.H + h) * p.W + w) * n_extent;
        }
      } else {
        int t = u;
        const int m_blk = CTA2 ? 2 * (t % num_m) + cta_rank : t % num_m; t /= num_m;
        n_blk = t % num_n; t /= num_n;
        const int tap = t % taps;
        const int ks = t / taps;
        has_k = ks * kb_per < nkb;
        const int row = m_blk * kCBM + rloc;
        row_ok = row < p.Cout;
        row_off = (long long)row * taps * p.Cin + (long long)tap * p.Cin;
      }
      mbar_wait(&tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row_ok && has_k && col0 < n_extent) {
          if (MODE == 2) {
            float* o = reinterpret_cast<float*>(p.out) + row_off + col0;
            if (col0 + 32 <= n_extent && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                atomicAdd(reinterpret_cast<float4*>(o + j),
                          make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                      __uint_as_float(v[j + 3])));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < n_extent) atomicAdd(o + j, __uint_as_float(v[j]));
            }
          } else {
            // channel counts are multiples of 64 (check_conv): every 32-column chunk inside the extent is complete,
            // and this thread's 32 outputs are 64 contiguous, 16-byte aligned bytes
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + col0;
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
              }
            }
            if (p.add != nullptr) {
              const uint4* ar = reinterpret_cast<const uint4*>(p.add + row_off + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 pk = __ldg(ar + j);
                const uint32_t r4[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  f[8 * j + 2 * i] += __uint_as_float(r4[i] << 16);
                  f[8 * j + 2 * i + 1] += __uint_as_float(r4[i] & 0xffff0000u);
                }
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(f[j + 0], f[j + 1]);
              __nv_bfloat162 h1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
              __nv_bfloat162 h3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&h0);
              pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2);
              pk.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(o + j) = pk;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CTA2) mbar_arrive_leader(&tempty[buf]); else mbar_arrive(&tempty[buf]); }
    }
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();     // pair: no CTA leaves while its peer may still signal it
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base); else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN, int MODE, bool CTA2 = false>
static int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const ConvParams& p, int units,
                       cudaStream_t stream) {
  using Cfg = ConvCfg<BN, CTA2>;
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "shared-memory budget");
  auto kern = conv_tc_kernel<BN, MODE, CTA2>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("conv_tc: cudaFuncSetAttribute(%d B smem): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return CFL_ECUDA;
    }
    attr_set = true;
  }
  if (!CTA2) {
    const int grid = units < sm_count() ? units : sm_count();
    kern<<<grid, 256, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  } else {
    // `units` counts 256-row work units: one cluster of two CTAs each, persistent over units
    const int pairs = units < sm_count() / 2 ? units : sm_count() / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
    if (e != cudaSuccess) {
      set_error("conv_tc (CTA pair): cudaLaunchKernelEx: %s", cudaGetErrorString(e));
      return CFL_ECUDA;
    }
  }
  return check_launch("conv_tc_kernel");
}

// CTA pairs serve the wide layers (CREAMFL_CONV_2CTA=0 switches them off: A/B measurements)
static bool conv_pair_enabled() {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("CREAMFL_CONV_2CTA");
    enabled = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return enabled != 0 && (sm_count() & 1) == 0;
}

// Largest power of two dividing x, capped.
static int pow2_divisor(int x, int cap) {
  int d = 1;
  while (d * 2 <= cap && x % (d * 2) == 0) d *= 2;
  return d;
}

// Pixel box (bw x bh x bn) with `rows` pixels that tiles an H x W map without remainder in h and w.
static void plan_box(int H, int W, int rows, int* bw, int* bh, int* bn) {
  *bw = pow2_divisor(W, rows < 16 ? rows : 16);
  *bh = pow2_divisor(H, rows / *bw);
  *bn = rows / (*bw * *bh);
}

static int check_conv(const char* who, int N, int H, int W, int Cin, int Cout, int R, int S) {
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) {
    set_error("%s: empty problem", who);
    return CFL_EINVAL;
  }
  if ((R & 1) == 0 || (S & 1) == 0) {
    set_error("%s: implicit path needs odd filter sizes (got %dx%d)", who, R, S);
    return CFL_EINVAL;
  }
  if (Cin % 64 || Cout % 64) {
    set_error("%s: implicit path needs Cin, Cout multiples of 64 (got %d, %d)", who, Cin, Cout);
    return CFL_EINVAL;
  }
  return CFL_OK;
}

// Y = conv(X, Wt), stride 1, padding (R/2, S/2).  X [N,H,W,Cin], Wt [Cout,R,S,Cin], Y [N,H,W,Cout], all bf16.
int conv_same_fprop(const void* x, const void* wt, int N, int H, int W, int Cin, int Cout, int R, int S, void* y,
                    cudaStream_t stream, const float* bias, const void* add, int relu, int stride) {
  int rc = check_conv("conv_fprop", N, H, W, Cin, Cout, R, S);
  if (rc) return rc;
  ConvParams p{};
  // padding R/2: Ho = (H + 2 (R/2) - R) / stride + 1; the pixel tiling and the epilogue live in OUTPUT coordinates
  const int Ho = (H + 2 * (R / 2) - R) / stride + 1, Wo = (W + 2 * (S / 2) - S) / stride + 1;
  p.N = N; p.H = Ho; p.W = Wo; p.stride = stride;
  p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S; p.pad_h = R / 2; p.pad_w = S / 2;
  plan_box(Ho, Wo, kCBM, &p.bw, &p.bh, &p.bn);
  p.tiles_w = Wo / p.bw; p.tiles_h = Ho / p.bh; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.split_k = 1;
  p.out = y;
  p.bias = bias;
  p.add = reinterpret_cast<const __nv_bfloat16*>(add);
  p.relu = relu;
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const bool pair = conv_pair_enabled() && Cout % 128 == 0 && pix_tiles >= 2;
  const int BN = pair ? (Cout % 256 == 0 ? 256 : 128) : ((Cout <= 64) ? 64 : 128);
  CUtensorMap ta, tb;
  // strided: the box spans (b - 1) * stride + 1 input pixels per dimension and keeps every stride-th one
  if ((rc = make_tmap_nhwc(&ta, x, N, H, W, Cin, 64, (p.bw - 1) * stride + 1, (p.bh - 1) * stride + 1, p.bn, stride)))
    return rc;
  if ((rc = make_tmap_2d(&tb, wt, 2, Cout, (uint64_t)R * S * Cin, (uint64_t)R * S * Cin, 64, pair ? BN / 2 : BN)))
    return rc;
  if (pair) {
    const int units = ((pix_tiles + 1) / 2) * (Cout / BN);
    return BN == 256 ? launch_conv<256, 0, true>(ta, tb, p, units, stream)
                     : launch_conv<128, 0, true>(ta, tb, p, units, stream);
  }
  const int units = pix_tiles * ((Cout + BN - 1) / BN);
  return BN == 64 ? launch_conv<64, 0>(ta, tb, p, units, stream) : launch_conv<128, 0>(ta, tb, p, units, stream);
}

// dX = conv_transpose(dY, Wt).  dY [N,H,W,Cout], Wt [Cout,R,S,Cin], dX [N,H,W,Cin], all bf16.
int conv_same_dgrad(const void* dy, const void* wt, int N, int H, int W, int Cin, int Cout, int R, int S, void* dx,
                    const void* add, cudaStream_t stream, int stride) {
  int rc = check_conv("conv_dgrad", N, H, W, Cin, Cout, R, S);
  if (rc) return rc;
  if (stride != 1 && (stride != 2 || (H & 1) || (W & 1) || R != 3 || S != 3)) {
    set_error("conv_dgrad: implicit path serves stride 1, or 3x3 stride 2 on even maps (got %dx%d stride %d, %dx%d)", R, S,
              stride, H, W);
    return CFL_EINVAL;
  }
  ConvParams p{};
  if (stride == 2) {     // tile dY; every tile serves the four parity classes of dX
    H /= 2;
    W /= 2;
  }
  p.N = N; p.H = H; p.W = W; p.stride = 1; p.up = stride;
  p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S; p.pad_h = R / 2; p.pad_w = S / 2;
  plan_box(H, W, kCBM, &p.bw, &p.bh, &p.bn);
  p.tiles_w = W / p.bw; p.tiles_h = H / p.bh; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.split_k = 1;
  p.out = dx;
  p.add = reinterpret_cast<const __nv_bfloat16*>(add);
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const bool pair = conv_pair_enabled() && Cin % 128 == 0 && pix_tiles >= 2;
  const int BN = pair ? (Cin % 256 == 0 ? 256 : 128) : ((Cin <= 64) ? 64 : 128);
  CUtensorMap ta, tb;
  if ((rc = make_tmap_nhwc(&ta, dy, N, H, W, Cout, 64, p.bw, p.bh, p.bn, 1))) return rc;
  if ((rc = make_tmap_2d(&tb, wt, 2, Cout, (uint64_t)R * S * Cin, (uint64_t)R * S * Cin, 64, 64))) return rc;
  const int classes = stride == 2 ? 4 : 1;
  if (pair) {
    const int units = ((pix_tiles + 1) / 2) * (Cin / BN) * classes;
    return BN == 256 ? launch_conv<256, 1, true>(ta, tb, p, units, stream)
                     : launch_conv<128, 1, true>(ta, tb, p, units, stream);
  }
  const int units = pix_tiles * ((Cin + BN - 1) / BN) * classes;
  return BN == 64 ? launch_conv<64, 1>(ta, tb, p, units, stream) : launch_conv<128, 1>(ta, tb, p, units, stream);
}

// dW[Cout,R,S,Cin] (fp32) += dY^T * shifted X.
int conv_same_wgrad(const void* dy, const void* x, int N, int H, int W, int Cin, int Cout, int R, int S, float* dw,
                    cudaStream_t stream, int stride) {
  int rc = check_conv("conv_wgrad", N, H, W, Cin, Cout, R, S);
  if (rc) return rc;
  ConvParams p{};
  const int Ho = (H + 2 * (R / 2) - R) / stride + 1, Wo = (W + 2 * (S / 2) - S) / stride + 1;
  p.N = N; p.H = Ho; p.W = Wo; p.stride = stride;
  p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S; p.pad_h = R / 2; p.pad_w = S / 2;
  plan_box(Ho, Wo, 64, &p.bw, &p.bh, &p.bn);
  p.tiles_w = Wo / p.bw; p.tiles_h = Ho / p.bh; p.tiles_n = (N + p.bn - 1) / p.bn;
  p.out = dw;
  const bool pair = conv_pair_enabled() && Cout % 256 == 0 && Cin % 128 == 0;
  const int BN = pair ? (Cin % 256 == 0 ? 256 : 128) : ((Cin <= 64) ? 64 : 128);
  const int num_m = pair ? Cout / 256 : (Cout + kCBM - 1) / kCBM, num_n = (Cin + BN - 1) / BN;
  const int base_units = num_m * num_n * R * S;
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  // whole waves of units (see plan_split_k in gemm_tc.cu); at least 8 pixel tiles per split
  const int split = plan_split_k(base_units, pix_tiles, 8, 4.0, pair ? sm_count() / 2 : 0);
  p.split_k = split;
  CUtensorMap ta, tb;
  if ((rc = make_tmap_nhwc(&ta, dy, N, Ho, Wo, Cout, 64, p.bw, p.bh, p.bn, 1))) return rc;
  if ((rc = make_tmap_nhwc(&tb, x, N, H, W, Cin, 64, (p.bw - 1) * stride + 1, (p.bh - 1) * stride + 1, p.bn, stride)))
    return rc;
  const int units = base_units * split;
  if (pair)
    return BN == 256 ? launch_conv<256, 2, true>(ta, tb, p, units, stream)
                     : launch_conv<128, 2, true>(ta, tb, p, units, stream);
  return BN == 64 ? launch_conv<64, 2>(ta, tb, p, units, stream) : launch_conv<128, 2>(ta, tb, p, units, stream);
}

}  // namespace cfl
