// Shared helpers for libhan_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "han_b200.h"

namespace han {

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr int kReduceBlocks = 148 * 4;  // fixed grid for deterministic two-stage reductions
constexpr float kLeakySlope = 0.2f;   // tf.nn.leaky_relu default (utils/layers.py:27)

extern thread_local char g_last_error[512];

inline int fail_arg(const char* fn, const char* what) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: invalid argument: %s", fn, what);
  return -1;
}

inline int check_launch(const char* fn) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: CUDA error %d (%s)", fn, (int)e,
             cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define HAN_REQUIRE(cond, what)                  \
  do {                                           \
    if (!(cond)) return han::fail_arg(__func__, what); \
  } while (0)

// Opt-in to > 48 KB of dynamic shared memory.  The attribute is per DEVICE (context), so remember it per device
// (bit d of a process-wide word), thread-safe; a process that touches several GPUs sets it on each.
#define HAN_SMEM_ATTR_ONCE(kernel, bytes)                                                              \
  do {                                                                                                 \
    static std::atomic<uint64_t> han_done_{0};                                                         \
    int han_dev_ = 0;                                                                                  \
    cudaGetDevice(&han_dev_);                                                                          \
    const uint64_t han_bit_ = 1ull << (han_dev_ & 63);                                                 \
    if (!(han_done_.load(std::memory_order_acquire) & han_bit_)) {                                     \
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));         \
      han_done_.fetch_or(han_bit_, std::memory_order_release);                                         \
    }                                                                                                  \
  } while (0)

inline cudaStream_t as_stream(han_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming (read-once) 128-bit load that does not allocate in L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream_f32(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ int ldg_stream_i32(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : kLeakySlope * x; }

// exp(x) as one FMUL + one MUFU.EX2: ex2.approx.ftz (relative error 2^-22 like __expf, which spends four more
// instructions on denormal inputs / results; results below 2^-126 flush to zero -- softmax terms that small vanish anyway)
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

// S_j a2 over one head (utils/layers.py:24 without its bias), in the association every kernel uses -- even and odd
// elements as two FMA chains (the forward gather runs them as one packed FFMA2 chain), added at the end -- so that the
// forward, the backward and han_attn_coefs see bit-identical logits: logit = (f1 + b2) + score_dot(S_j, a2).
template <int H>
__device__ __forceinline__ float score_dot(const float* v, const float* a) {
  float x = v[0] * a[0], y = v[1] * a[1];
#pragma unroll
  for (int i = 1; i < H / 2; ++i) {
    x = fmaf(v[2 * i], a[2 * i], x);
    y = fmaf(v[2 * i + 1], a[2 * i + 1], y);
  }
  return x + y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace han
