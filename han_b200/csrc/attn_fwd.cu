// Shape queries of the node-attention kernels and the per-edge coefficient kernel (return_coef,
// utils/layers.py:43-44).  The fused edge-softmax-aggregate itself (K-B) lives in attn_stream.cu.
#include "han_common.cuh"

namespace han {

// per-edge coefficients (return_coef, utils/layers.py:43-44): alpha[e][k] recomputed from m, rinv
template <int K, int H>
__global__ void __launch_bounds__(256)
attn_coefs_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  int64_t n_dst, const float* __restrict__ T, const float* __restrict__ a2,
                  const float* __restrict__ b2, const float* __restrict__ R,
                  const float* __restrict__ ew, float* __restrict__ alpha) {
  constexpr int D = K * H;
  constexpr int TS = D;
  constexpr int RS = ((D + 3 * K + 3) / 4) * 4;
  constexpr int SLOTS = 32 / K;
  const int lane = threadIdx.x & 31;
  const int head = lane % K, slot = lane / K;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_dst) return;
  const int64_t start = indptr[row], end = indptr[row + 1];
  const float f1b = R[row * RS + D + head] + __ldg(b2 + head);
  const float lse = R[row * RS + D + K + head];
  float a2v[H];
#pragma unroll
  for (int h = 0; h < H; ++h) a2v[h] = __ldg(a2 + head * H + h);
  for (int64_t e = start + slot; e < end; e += SLOTS) {
    const int col = indices[e];
    float sv[H];                          // f2_j = S_j a2 + b2 (utils/layers.py:24), from the table row
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 t = ldg4(T + (int64_t)col * TS + head * H + 4 * q);
      sv[4 * q] = t.x; sv[4 * q + 1] = t.y; sv[4 * q + 2] = t.z; sv[4 * q + 3] = t.w;
    }
    const float lgt = f1b + score_dot<H>(sv, a2v);
    const float wt = ew ? __ldg(ew + e) : 1.f;
    alpha[e * K + head] = __expf(leaky(lgt * wt) - lse);
  }
}

template <int K, int H>
static int launch_coefs(const int64_t* indptr, const int32_t* indices, int64_t n_dst, const float* T,
                        const float* a2, const float* b2, const float* R, const float* ew, float* alpha,
                        cudaStream_t st) {
  unsigned grid = (unsigned)ceil_div64(n_dst, 8);
  attn_coefs_kernel<K, H><<<grid, 256, 0, st>>>(indptr, indices, n_dst, T, a2, b2, R, ew, alpha);
  return check_launch("han_attn_coefs");
}

}  // namespace han

using namespace han;

// (K,H) instantiations: K a power of two <= 32 (lanes per edge), H a multiple of 4 (float4 lanes)
#define HAN_FOR_SHAPES(X) X(8, 8) X(4, 8) X(2, 8) X(1, 8) X(8, 4) X(4, 4) X(1, 4) X(8, 16) X(4, 16) X(1, 16) X(16, 4) X(16, 8)

extern "C" {

int han_attn_shape_supported(int K, int H) {
#define X(k, h) if (K == k && H == h) return 1;
  HAN_FOR_SHAPES(X)
#undef X
  return 0;
}

int han_table_stride(int K, int H) { return K * H; }
int han_record_stride(int K, int H) { return ((K * H + 3 * K + 3) / 4) * 4; }

int han_attn_coefs(const int64_t* indptr, const int32_t* indices, int64_t n_dst, const float* T, const float* a2,
                   const float* b2, const float* R, int K, int H, const float* edge_w, float* alpha,
                   han_stream_t stream) {
  HAN_REQUIRE(indptr && T && a2 && b2 && R && alpha, "null pointer");
  HAN_REQUIRE(n_dst > 0, "n_dst > 0 required");
#define X(k, h)         \
  if (K == k && H == h) \
    return launch_coefs<k, h>(indptr, indices, n_dst, T, a2, b2, R, edge_w, alpha, as_stream(stream));
  HAN_FOR_SHAPES(X)
#undef X
  return fail_arg(__func__, "unsupported (K,H); see han_attn_shape_supported");
}

}  // extern "C"
