// ResNet stem convolution (7x7, stride 2, pad 3, 3 -> 64 channels) straight from the loader's fp32 NCHW images, as an
// implicit GEMM on tcgen05 whose patch operand is BUILT IN SHARED MEMORY - no patch matrix ever reaches HBM.
//
// Replaces torchvision ResNet.conv1 reached from src/networks/models/image_encoder.py:24,55 (server ResNet101 /
// client ResNet18) and src/networks/resnet_client.py:164.  The previous path materialised a [B*112*112, 152] bf16
// patch matrix (0.49 GB at batch 128: 0.37 ms to write, 0.21 ms GEMM re-reading it, 0.15 ms weight gradient re-reading
// it again); per forward this kernel moves the algorithmic bytes only (77 MB of images in, 205 MB of bf16 map out).
//
//   MODE 0  fprop : Y[p, co] = sum_k P[p, k] Wt[co, k]        p = output pixel, k = (r, s, c) tap index (147 real)
//   MODE 2  wgrad : dW[co, k] += sum_p dY[p, co] P[p, k]
//
// One tile = one output row (n, ho): up to 128 output pixels (Wo <= 128, i.e. images up to 256 wide).
// Warp roles (512 threads, persistent over tiles):
//   warps 0-9  (320 threads) builders: cp.async (16-byte chunks) the 7 input rows x 3 channels of the NEXT tile into a staging buffer
//                            while assembling the current tile's patch operand [128 pixels x 160 taps] bf16 directly
//                            in the 128-byte-swizzled layout the MMA descriptors read (thread = one 16-byte unit column,
//                            tap offsets held in registers)
//   warp 10    MMA issuer:   fprop 10 x tcgen05.mma 128x64x16 per tile (K = 160, taps 147..159 are zeros);
//                            wgrad 2 x 8 MMAs per tile accumulating dW^T [taps x co] in TMEM across all tiles of the CTA
//   warp 11    TMEM allocator
//   warps 12-15 epilogue:    fprop: TMEM -> bf16 -> 128 contiguous bytes per pixel; wgrad: one fp32 atomic flush per CTA
#include "kernels.cuh"
#include "ptx.cuh"
#include <string.h>

namespace cfl {

namespace {
constexpr int kSC = 3, kSR = 7, kSS = 7, kSStride = 2, kSPad = 3;
constexpr int kTaps = kSC * kSR * kSS;     // 147
constexpr int kUnits = 20;                 // 16-byte units per pixel row: 160 taps (147 real + zero tail)
constexpr int kCout = 64;
constexpr int kBuilders = 320;             // warps 0-9 prefetch the next tile's rows and assemble the current patch operand
constexpr int kStemThreads = 512;
constexpr int kBlockBytes = 128 * 128;     // one 64-tap block of the patch operand: 128 pixel rows x 128 B

struct StemParams {
  const float* x;        // [N, 3, H, W] fp32
  int N, H, W, Ho, Wo;
  void* out;             // fprop: bf16 [N, Ho, Wo, 64]; wgrad: fp32 [64, 147] accumulated
  int tiles;             // N * Ho
};

template <int MODE>
struct StemCfg {
  static constexpr int kABlocks = (MODE == 0) ? 3 : 4;          // wgrad reads taps 128..255 as a second M = 128 operand
  static constexpr int kABytes = kABlocks * kBlockBytes;
  static constexpr int kWBytes = (MODE == 0) ? 3 * 64 * 128 : 0;  // fprop: filters [64 x 192] bf16, K-major
  static constexpr int kDBytes = (MODE == 2) ? 128 * 128 : 0;     // wgrad: dY tile [128 pixels x 64 co]
  static constexpr int kRowFloats = kSC * kSR * (256 + 8);   // row pitch W + 8: image at [4, 4 + W), zero borders
  static constexpr int kSmemBytes = 2 * kABytes + kWBytes + 2 * kDBytes + 2 * kRowFloats * 4 + 1024 + 256;
};

__device__ __forceinline__ void cp_async_16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;    // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void builder_sync() { asm volatile("bar.sync 1, 320;" ::: "memory"); }
}  // namespace

template <int MODE>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmD, StemParams p) {
  using Cfg = StemCfg<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // [2][kABytes]
  uint8_t* sW = sA + 2 * Cfg::kABytes;                  // fprop filters
  uint8_t* sD = sW + Cfg::kWBytes;                      // wgrad dY tiles [2][16 KB]
  float* srow = reinterpret_cast<float*>(sD + 2 * Cfg::kDBytes);   // [2][kRowFloats]
  uint64_t* bars = reinterpret_cast<uint64_t*>(srow + 2 * Cfg::kRowFloats);
  uint64_t* full = bars;          // [2] patch operand of stage s assembled
  uint64_t* empty = bars + 2;     // [2] MMAs reading stage s retired
  uint64_t* tfull = bars + 4;     // [2] accumulator ready
  uint64_t* tempty = bars + 6;    // [2] accumulator drained
  uint64_t* wfull = bars + 8;     // filters landed (fprop)
  uint64_t* dfull = bars + 9;     // [2] dY tile landed (wgrad)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wp = p.W + 8;            // staged row pitch (floats): 16-byte aligned rows, image columns at [4, 4 + W)
  const int n_my = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  if (warp == 10 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
      mbar_init(&dfull[i], 1);
    }
    mbar_init(wfull, 1);
    fence_mbar_init();
    if (MODE == 0) tma_prefetch_desc(&tmW);
    if (MODE == 2) tma_prefetch_desc(&tmD);
  }
  if (warp == 11) tmem_alloc<128>(tmem_slot);
  if (MODE == 2) {   // taps 160..255 of both stages are never assembled: they must read as zeros
    for (int e = threadIdx.x; e < 2 * Cfg::kABytes / 16; e += kStemThreads)
      reinterpret_cast<uint4*>(sA)[e] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  // zero borders of the staged rows (left / right padding of the convolution): written once, never overwritten
  for (int e = threadIdx.x; e < 2 * kSC * kSR * 8; e += kStemThreads) {
    const int row = e >> 3, b = e & 7;
    srow[(row / (kSC * kSR)) * Cfg::kRowFloats + (row % (kSC * kSR)) * Wp + (b < 4 ? b : p.W + b)] = 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x < kBuilders) {
    // ------------------------------------------------------------ builders
    const int tid = threadIdx.x;
    // assembly: builder warp w owns the 16-byte unit columns w and w + 10 of the patch operand (tap offsets of both in
    // registers); its lanes take CONSECUTIVE pixels, so the staged-row reads of a tap are a stride-2 sweep (2-way bank
    // conflict at worst) and the swizzled 16-byte stores of 8 consecutive pixels land in 8 distinct slots
    int koff[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = 8 * (warp + 10 * h) + j;
        const int tap = k / kSC, c = k - tap * kSC, r = tap / kSS, sx = tap - r * kSS;
        koff[h][j] = (k < kTaps) ? (c * kSR + r) * Wp + sx + 1 : -1;   // input column 2 px + sx - 3 lives at 4 + column
      }
    const int w4 = p.W >> 2;                            // 16-byte chunks per image row
    auto prefetch = [&](int t, float* dst) {
      const int n = t / p.Ho, ho = t - n * p.Ho;
      const int h0 = ho * kSStride - kSPad;
      for (int e = tid; e < kSC * kSR * w4; e += kBuilders) {
        const int cr = e / w4, ch = e - cr * w4;
        const int c = cr / kSR, h = h0 + (cr - c * kSR);
        const bool hv = h >= 0 && h < p.H;
        const float* src = p.x + (((long long)n * kSC + c) * p.H + (hv ? h : 0)) * p.W + 4 * ch;
        cp_async_16(dst + cr * Wp + 4 + 4 * ch, src, hv);
      }
      cp_async_commit();
    };
    if (n_my > 0) prefetch(blockIdx.x, srow);
    for (int i = 0; i < n_my; ++i) {
      const int t = blockIdx.x + i * gridDim.x;
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      if (i + 1 < n_my) {
        prefetch(t + gridDim.x, srow + ((i + 1) & 1) * Cfg::kRowFloats);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      builder_sync();                                   // every builder's rows of tile i are in shared memory
      mbar_wait(&empty[s], ph ^ 1);                     // MMAs of tile i - 2 no longer read stage s
      if (MODE == 2 && tid == 0) {                      // dY rows of this tile: one 128 x 64 box
        const int n = t / p.Ho, ho = t - n * p.Ho;
        mbar_arrive_expect_tx(&dfull[s], 128 * 128);
        tma_load_2d(&tmD, &dfull[s], sD + s * Cfg::kDBytes, 0, (n * p.Ho + ho) * p.Wo);
      }
      const uint32_t rows = smem_u32(srow + s * Cfg::kRowFloats);      // explicit shared-space accesses (ptx.cuh)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int u = warp + 10 * h;
        const uint32_t blk = smem_u32(sA + s * Cfg::kABytes + (u >> 3) * kBlockBytes);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int px = q * 32 + lane;
          const bool live = px < p.Wo;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            f[j] = (live && koff[h][j] >= 0) ? lds_f32(rows + 4 * (koff[h][j] + kSStride * px)) : 0.0f;
          uint4 pk;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2);
          pk.w = *reinterpret_cast<uint32_t*>(&h3);
          sts128(blk + sw128_offset(px, u & 7), pk);
        }
      }
      fence_proxy_async_smem();                         // generic-proxy writes -> visible to the tensor core
      builder_sync();
      if (tid == 0) mbar_arrive(&full[s]);
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      if (MODE == 0) {
        mbar_arrive_expect_tx(wfull, Cfg::kWBytes);
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) tma_load_2d(&tmW, wfull, sW + kb * 8192, kb * 64, 0);
        mbar_wait(wfull, 0);
      }
      constexpr uint32_t idesc = (MODE == 0) ? make_idesc(1, 128, kCout, 0, 0) : make_idesc(1, 128, kCout, 1, 1);
      for (int i = 0; i < n_my; ++i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const uint32_t a0 = smem_u32(sA + s * Cfg::kABytes);
        if (MODE == 0) {
          mbar_wait(&tempty[s], ph ^ 1);
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t w0 = smem_u32(sW);
#pragma unroll
          for (int kk = 0; kk < 10; ++kk) {             // K = 160 = 10 x 16
            const int kb = kk >> 2, k = kk & 3;
            umma_f16_ss(tmem_base + s * kCout, make_smem_desc(a0 + kb * kBlockBytes + k * 32, 16, 1024),
                        make_smem_desc(w0 + kb * 8192 + k * 32, 16, 1024), idesc, kk > 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          umma_commit(&tfull[s]);
        } else {
          mbar_wait(&full[s], ph);
          mbar_wait(&dfull[s], ph);
          tc_fence_after();
          const uint32_t d0 = smem_u32(sD + s * Cfg::kDBytes);
#pragma unroll
          for (int half = 0; half < 2; ++half) {        // taps [0, 128) and [128, 256): two M = 128 operands
#pragma unroll
            for (int k = 0; k < 8; ++k) {               // K = 128 pixels = 8 x 16
              umma_f16_ss(tmem_base + half * kCout,
                          make_smem_desc(a0 + half * 2 * kBlockBytes + k * 2048, kBlockBytes, 1024),
                          make_smem_desc(d0 + k * 2048, 8192, 1024), idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[s]);
          if (i == n_my - 1) umma_commit(&tfull[0]);
        }
      }
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------ epilogue (thread = accumulator row)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (MODE == 0) {
      for (int i = 0; i < n_my; ++i) {
        const int t = blockIdx.x + i * gridDim.x;
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(&tfull[s], ph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + s * kCout;
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)t * p.Wo + row) * kCout;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (row < p.Wo) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]));
              __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
              __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
              pk.x = *reinterpret_cast<uint32_t*>(&h0);
              pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2);
              pk.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(o + c * 32 + j) = pk;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[s]);
      }
    } else if (n_my > 0) {
      // dW^T accumulated over all tiles of this CTA: row = tap (two halves of 128), column = output channel
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      float* dw = reinterpret_cast<float*>(p.out);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int tap = half * 128 + row;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * kCout;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (tap < kTaps) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dw + (long long)(c * 32 + j) * kTaps + tap, __uint_as_float(v[j]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 11) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

template <int MODE>
static int launch_stem(const CUtensorMap& tw, const CUtensorMap& td, const StemParams& p, cudaStream_t stream) {
  using Cfg = StemCfg<MODE>;
  auto kern = stem_tc_kernel<MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("stem_tc: cudaFuncSetAttribute(%d B smem): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return CFL_ECUDA;
    }
    attr_set = true;
  }
  const int grid = p.tiles < sm_count() ? p.tiles : sm_count();
  kern<<<grid, kStemThreads, Cfg::kSmemBytes, stream>>>(tw, td, p);
  return check_launch("stem_tc_kernel");
}

bool stem_supported(int C, int H, int W, int R, int S, int stride, int pad, int Cout) {
  if (C != kSC || R != kSR || S != kSS || stride != kSStride || pad != kSPad || Cout != kCout) return false;
  const int Wo = (W + 2 * pad - S) / stride + 1;
  return W >= 8 && H >= 7 && W <= 256 && (W & 3) == 0 && Wo <= 128;   // 16-byte row chunks
}

// Y [N, Ho, Wo, 64] bf16 = conv7x7/2(images fp32 NCHW, Wt [64, >= 147] bf16 with row pitch ldw in (r, s, c) order)
int stem_fprop(const float* x, const void* wt, long long ldw, int N, int H, int W, void* y, cudaStream_t stream) {
  if (!stem_supported(3, H, W, 7, 7, 2, 3, 64) || N <= 0) {
    set_error("stem_fprop: unsupported shape N=%d H=%d W=%d", N, H, W);
    return CFL_EINVAL;
  }
  StemParams p{};
  p.x = x; p.N = N; p.H = H; p.W = W;
  p.Ho = (H + 2 * kSPad - kSR) / kSStride + 1;
  p.Wo = (W + 2 * kSPad - kSS) / kSStride + 1;
  p.out = y;
  p.tiles = N * p.Ho;
  CUtensorMap tw, td;
  memset(&td, 0, sizeof(td));
  // filters [64 rows, ldw columns]: taps 147.. are zeros up to the pitch, beyond it TMA's out-of-bounds fill
  int rc = make_tmap_2d(&tw, wt, 2, kCout, (uint64_t)ldw, (uint64_t)ldw, 64, 64);
  if (rc) return rc;
  return launch_stem<0>(tw, td, p, stream);
}

// dW [64, 147] fp32 += dY^T patches(images)
int stem_wgrad(const float* x, const void* dy, int N, int H, int W, float* dw, cudaStream_t stream) {
  if (!stem_supported(3, H, W, 7, 7, 2, 3, 64) || N <= 0) {
    set_error("stem_wgrad: unsupported shape N=%d H=%d W=%d", N, H, W);
    return CFL_EINVAL;
  }
  StemParams p{};
  p.x = x; p.N = N; p.H = H; p.W = W;
  p.Ho = (H + 2 * kSPad - kSR) / kSStride + 1;
  p.Wo = (W + 2 * kSPad - kSS) / kSStride + 1;
  p.out = dw;
  p.tiles = N * p.Ho;
  CUtensorMap tw, td;
  memset(&tw, 0, sizeof(tw));
  int rc = make_tmap_2d(&td, dy, 2, (uint64_t)N * p.Ho * p.Wo, kCout, kCout, 64, 128);
  if (rc) return rc;
  return launch_stem<2>(tw, td, p, stream);
}

}  // namespace cfl
