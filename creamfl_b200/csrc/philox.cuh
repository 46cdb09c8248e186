// Counter-based dropout masks (Philox4x32-10, Salmon et al. SC'11 - the generator behind torch's CUDA dropout).
//
// The reference trains HF BertModel with hidden_dropout_prob = attention_probs_dropout_prob = 0.1
// (src/networks/models/pcme.py:31 + BertConfig defaults, model.train() at src/algorithms/retrieval_trainer.py:187,
// MMFL.py:293): 37 dropout sites per forward.  Here a mask is never stored: every kernel that applies or
// back-propagates a dropout regenerates it from (seed, step, site, element index).
//
//   rng      device uint64[2] = {seed, step}; `step` is advanced once per training step by rng_tick_kernel, so a
//            captured CUDA graph of the step draws fresh masks at every replay
//   site     which of the 37 dropouts (0 = embeddings, 1 + 3*layer = attention probabilities,
//            2 + 3*layer = attention output dense, 3 + 3*layer = FFN output dense)
//   element  linear index inside the site's tensor; one Philox call serves the 8 elements of block e >> 3
//            (4 x 32 random bits = 8 x 16-bit fields), element e keeps its value iff field[e & 7] >= thresh16
//   thresh16 = round(p * 65536)  (p = 0.1 -> 6554: keep probability 0.899994), survivors are scaled by 1 / (1 - p)
#pragma once
#include <stdint.h>

namespace cfl {

struct DropSpec {
  const unsigned long long* rng;  // null: no dropout
  int site;
  uint32_t thresh;                // 16-bit threshold
  float scale;                    // 1 / (1 - p)
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Keep bits of the 8 elements of block e8 (bit i = element 8*e8 + i survives).
__host__ __device__ __forceinline__ uint32_t drop_keep8(unsigned long long seed, uint32_t step, uint32_t site,
                                                        unsigned long long e8, uint32_t thresh) {
  uint32_t r[4];
  philox4x32_10((uint32_t)e8, (uint32_t)(e8 >> 32), site, step, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bits |= ((r[i] & 0xFFFFu) >= thresh ? 1u : 0u) << (2 * i);
    bits |= ((r[i] >> 16) >= thresh ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}

inline uint32_t drop_thresh16(float p) {
  const double t = (double)p * 65536.0;
  return (uint32_t)(t + 0.5);
}

}  // namespace cfl
