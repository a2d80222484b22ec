// Counter-based dropout masks (training mode; tf.nn.dropout at utils/layers.py:19,30,32).
//
// TF's Philox stream cannot be reproduced outside TF, so parity with dropout on is statistical by
// construction.  What matters here is that the SAME mask bit can be recomputed
// anywhere it is needed -- forward and backward kernels, other ranks -- from (seed, coordinates) alone,
// with no mask tensors in memory.  keep(seed, a, b) is a 32-bit mix of three words compared against a
// 24-bit threshold: element kept iff u < keep_prob with u uniform in [0,1) (TF: floor(keep + u) == 1
// iff u >= 1 - keep; same distribution).  tests/test_gpu_dropout.py holds the numpy replica used to
// feed identical masks to the oracle.
#pragma once
#include <stdint.h>

namespace han {

__host__ __device__ __forceinline__ uint32_t mix3(uint32_t seed, uint32_t a, uint32_t b) {
  uint32_t h = seed ^ 0x9E3779B9u;
  h = (h ^ a) * 0x85EBCA6Bu;
  h ^= h >> 13;
  h = (h ^ b) * 0xC2B2AE35u;
  h ^= h >> 16;
  h *= 0x27D4EB2Fu;
  h ^= h >> 15;
  return h;
}

// threshold = round(keep_prob * 2^24); kept iff the top 24 bits of the mix are below it
__host__ __device__ __forceinline__ bool keep24(uint32_t seed, uint32_t a, uint32_t b, uint32_t threshold) {
  return (mix3(seed, a, b) >> 8) < threshold;
}

// one stream per (purpose, meta-path, head): purposes 1 = input features, 2 = projected features,
// 3 = attention coefficients
__host__ __device__ __forceinline__ uint32_t stream_seed(uint32_t seed, uint32_t purpose, uint32_t metapath,
                                                         uint32_t head) {
  return mix3(seed, purpose * 0x01000193u + metapath, head + 0x7F4A7C15u);
}

}  // namespace han
