// One-launch training update: L2 on every variable + Adam with TF1 semantics
// (models/base_gattn.py:12-24).  All trainable variables of the model live in one flat FP32 buffer
// (views with 256-byte aligned offsets; the padding stays zero), so the whole update is a single
// grid-stride pass and can sit at the end of a captured CUDA graph: the step counter is a device word.
//
//   g' = g + l2 * p                      (d/dp of l2 * sum(p^2)/2, :14-16)
//   m  = b1 m + (1-b1) g' ;  v = b2 v + (1-b2) g'^2
//   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) ;  p -= lr_t * m / (sqrt(v) + eps)     (tf.train.AdamOptimizer)
#include "han_common.cuh"

namespace han {

__global__ void __launch_bounds__(256)
adam_l2_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
               float* __restrict__ v, int64_t n4, const int* __restrict__ step_ptr, float lr, float beta1,
               float beta2, float eps, float l2) {
  __shared__ float lr_t_s;
  if (threadIdx.x == 0) {
    const double t = (double)*step_ptr;
    lr_t_s = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  }
  __syncthreads();
  const float lr_t = lr_t_s;
  const float c1 = 1.f - beta1, c2 = 1.f - beta2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define HAN_ADAM_LANE(c)                                  \
  {                                                       \
    const float gr = fmaf(l2, pp.c, gg.c);                \
    mm.c = fmaf(beta1, mm.c, c1 * gr);                    \
    vv.c = fmaf(beta2, vv.c, c2 * gr * gr);               \
    pp.c -= lr_t * (mm.c / (sqrtf(vv.c) + eps));          \
  }
    HAN_ADAM_LANE(x) HAN_ADAM_LANE(y) HAN_ADAM_LANE(z) HAN_ADAM_LANE(w)
#undef HAN_ADAM_LANE
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

}  // namespace han

using namespace han;

extern "C" int han_adam_l2_step(float* p, const float* g, float* m, float* v, int64_t n, const int* step_ptr,
                                float lr, float beta1, float beta2, float eps, float l2_coef,
                                han_stream_t stream) {
  HAN_REQUIRE(p && g && m && v && step_ptr, "null pointer");
  HAN_REQUIRE(n > 0 && n % 4 == 0, "n > 0 and a multiple of 4 (flat buffer with padded views)");
  HAN_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, "16-byte aligned buffers");
  const int64_t n4 = n / 4;
  int64_t blocks = ceil_div64(n4, 256);
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  adam_l2_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n4, step_ptr, lr, beta1, beta2, eps,
                                                                 l2_coef);
  return check_launch(__func__);
}
