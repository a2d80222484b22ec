// Gradient of the projection w.r.t. its INPUT -- what a stacked attention layer (models/gat.py:48-57)
// hands to the layer below.  For one meta-path, every head k projected its own dropped copy of the input
// (utils/layers.py:18-20), S_k = (X * m_k / keep) W_k, so
//
//     dX[n][f] (+)= sum_k  m_k(n,f)/keep * sum_h dS[n][kH+h] * W[f][kH+h]
//
// with the forward's per-head masks regenerated from (seed, meta-path, node, feature) -- han_rng.cuh, the
// same bits sgemm_nn_drop_kernel used.  Without dropout (seed_ptr == NULL) every mask is 1 and this is the
// plain dS W^T.  The inputs of a stacked layer are the previous layer's concatenated heads (F = K*H, 64 in
// every configuration the reference ships), so the kernel keeps a 64-feature slab of W in shared memory and
// streams node rows through it: 16 lanes share a node (dS row broadcast from L1), each lane owns the
// features fq, fq+16, fq+32, fq+48 of the slab (conflict-free 128-bit shared loads, coalesced stores).
#include "han_common.cuh"
#include "han_rng.cuh"

namespace han {

constexpr int kDxThreads = 256;
constexpr int kDxFeat = 64;      // features per CTA slab
constexpr int kDxMaxD = 128;

__global__ void __launch_bounds__(kDxThreads)
project_dx_kernel(const float* __restrict__ dS, int64_t n, int D, int K, int H, const float* __restrict__ W,
                  int64_t ldw, int64_t F, float* __restrict__ dX, int64_t ldx, int accumulate,
                  const uint32_t* __restrict__ seed_ptr, uint32_t thr, float inv_keep, uint32_t metapath,
                  int64_t row0) {
  extern __shared__ __align__(16) float Ws[];      // [kDxFeat][D + 4]
  const int LD = D + 4;
  const int tid = threadIdx.x;
  const int64_t f0 = (int64_t)blockIdx.y * kDxFeat;
  for (int i = tid; i < kDxFeat * (D / 4); i += kDxThreads) {
    const int f = i / (D / 4), c = i % (D / 4);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f0 + f < F) w = ldg4(W + (f0 + f) * ldw + 4 * c);
    *reinterpret_cast<float4*>(Ws + f * LD + 4 * c) = w;
  }
  __syncthreads();
  const int fq = tid % 16, rl = tid / 16;
  const bool drop = seed_ptr != nullptr;
  const uint32_t sseed = drop ? stream_seed(*seed_ptr, 1u, metapath, 0u) : 0u;
  for (int64_t r = (int64_t)blockIdx.x * 16 + rl; r < n; r += (int64_t)gridDim.x * 16) {
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t bits[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bits[j] = 0xFFFFFFFFu;
      if (drop) {
        // identical to head_mask_bits(mix3(sseed, node, f), K, thr) in project_drop.cu
        const uint32_t base = mix3(sseed, (uint32_t)(r + row0), (uint32_t)(f0 + j * 16 + fq));
        uint32_t b = 0;
        for (int k = 0; k < K; ++k) {
          const uint32_t h = (base ^ (0x632BE5ABu * (uint32_t)(k + 1))) * 0x9E3779B1u;
          b |= ((h >> 8) < thr ? 1u : 0u) << k;
        }
        bits[j] = b;
      }
    }
    const float* ds = dS + r * D;
    for (int k = 0; k < K; ++k) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int h = 0; h < H; h += 4) {
        const float4 g = ldg4(ds + k * H + h);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(Ws + (j * 16 + fq) * LD + k * H + h);
          acc[j] = fmaf(g.x, w.x, acc[j]);
          acc[j] = fmaf(g.y, w.y, acc[j]);
          acc[j] = fmaf(g.z, w.z, acc[j]);
          acc[j] = fmaf(g.w, w.w, acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((bits[j] >> k) & 1u) tot[j] += acc[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t f = f0 + j * 16 + fq;
      if (f < F) {
        float* o = dX + r * ldx + f;
        const float v = tot[j] * inv_keep;
        *o = accumulate ? *o + v : v;
      }
    }
  }
}

}  // namespace han

using namespace han;

extern "C" int han_project_dx(const float* dS, int64_t n, int K, int H, const float* W, int64_t ldw, int64_t F,
                              float* dX, int64_t ldx, int accumulate, const void* seed_ptr, float in_keep,
                              int metapath, int64_t row0, han_stream_t stream) {
  HAN_REQUIRE(dS && W && dX, "null pointer");
  const int D = K * H;
  HAN_REQUIRE(n > 0 && F > 0 && K >= 1 && K <= 32 && H % 4 == 0 && D <= kDxMaxD, "n, F > 0; K <= 32; H % 4 == 0; K*H <= 128");
  HAN_REQUIRE(ldw % 4 == 0 && ((uintptr_t)W % 16) == 0 && ((uintptr_t)dS % 16) == 0, "16-byte aligned W rows and dS");
  HAN_REQUIRE(in_keep > 0.f && in_keep <= 1.f, "0 < in_keep <= 1");
  const bool drop = seed_ptr != nullptr && in_keep < 1.f;
  const size_t smem = (size_t)kDxFeat * (D + 4) * sizeof(float);
  int64_t gx = ceil_div64(n, 16);
  if (gx > (int64_t)kNumSMs * 8) gx = (int64_t)kNumSMs * 8;
  dim3 grid((unsigned)gx, (unsigned)ceil_div64(F, kDxFeat));
  const uint32_t thr = (uint32_t)(in_keep * 16777216.f + 0.5f);
  const float inv_keep = drop ? 1.f / ((float)thr / 16777216.f) : 1.f;     // as project_drop.cu: unbiased for the quantised keep
  project_dx_kernel<<<grid, kDxThreads, smem, as_stream(stream)>>>(
      dS, n, D, K, H, W, ldw, F, dX, ldx, accumulate, drop ? reinterpret_cast<const uint32_t*>(seed_ptr) : nullptr,
      thr, inv_keep, (uint32_t)metapath, row0);
  return check_launch(__func__);
}
