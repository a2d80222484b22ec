// K-B: fused CSR edge-softmax-aggregate (forward).  Replaces the dense N x N chain of
// utils/layers.py:26-35,46 (f1 + f2^T, leaky_relu, + bias_mat, softmax, coefs @ seq_fts, bias_add,
// activation) for all K heads of one meta-path in a single pass over the edges.
//
// Mapping: one warp per destination row.  A lane is (slot, head): head = lane % K owns the H
// floats of that head, slot = lane / K selects which of the 32/K edges processed together it
// works on.  Per edge a K-lane group reads one contiguous node-table row [S_j (D) | f2_j (K)]
// (288 B for K=H=8): 128-bit loads for S, one 4-byte load for f2, so a group touches 9 full
// sectors and nothing else.  UNROLL edges per slot are loaded before any is consumed, giving
// (32/K)*UNROLL gathered rows in flight per warp (HBM latency hiding is the whole game here:
// ~0.5 FLOP/B).  Softmax is the online (running max / running sum) form, merged across slots with
// shuffles at the end of the row.  HBM-bound: 292 B/edge + 360 B/row algorithmic.
#include "han_common.cuh"

namespace han {

template <int K, int H, int UNROLL>
__global__ void __launch_bounds__(256)
attn_fwd_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                int64_t n_dst, const float* __restrict__ T, float* __restrict__ R,
                const float* __restrict__ bias, int act, float* __restrict__ out, int64_t out_stride,
                float* __restrict__ vsave, const float* __restrict__ colmean) {
  constexpr int D = K * H;
  constexpr int TS = ((D + K + 3) / 4) * 4;
  constexpr int RS = ((D + 3 * K + 3) / 4) * 4;
  constexpr int SLOTS = 32 / K;
  constexpr int HV = H / 4;
  const int lane = threadIdx.x & 31;
  const int head = lane % K;
  const int slot = lane / K;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_dst) return;

  const int64_t start = indptr[row], end = indptr[row + 1];
  const float f1v = R[row * RS + D + head];
  float m = -INFINITY, l = 0.f;
  float acc[H];
#pragma unroll
  for (int h = 0; h < H; ++h) acc[h] = 0.f;

  for (int64_t base = start; base < end; base += 32) {
    const int cnt = (int)min((int64_t)32, end - base);
    const int my_col = (lane < cnt) ? ldg_stream_i32(indices + base + lane) : 0;
    for (int t = 0; t < cnt; t += SLOTS * UNROLL) {
      float4 v[UNROLL][HV];
      float e[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int ei = t + u * SLOTS + slot;
        const int col = __shfl_sync(0xffffffffu, my_col, ei & 31);
        if (ei < cnt) {
          const float* rowp = T + (int64_t)col * TS;
          e[u] = __ldg(rowp + D + head);
#pragma unroll
          for (int q = 0; q < HV; ++q) v[u][q] = ldg4(rowp + head * H + 4 * q);
        } else {
          e[u] = -INFINITY;
#pragma unroll
          for (int q = 0; q < HV; ++q) v[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float mnew = m;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        e[u] = (e[u] == -INFINITY) ? -INFINITY : leaky(f1v + e[u]);
        mnew = fmaxf(mnew, e[u]);
      }
      if (mnew != -INFINITY) {  // at least one valid edge seen by this lane so far
        const float sc = __expf(m - mnew);  // m = -inf -> 0
        l *= sc;
#pragma unroll
        for (int h = 0; h < H; ++h) acc[h] *= sc;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const float p = __expf(e[u] - mnew);  // invalid edge: exp(-inf) = 0
          l += p;
#pragma unroll
          for (int q = 0; q < HV; ++q) {
            acc[4 * q + 0] = fmaf(p, v[u][q].x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(p, v[u][q].y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(p, v[u][q].z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(p, v[u][q].w, acc[4 * q + 3]);
          }
        }
        m = mnew;
      }
    }
  }

  // merge the SLOTS partial softmax states of each head
#pragma unroll
  for (int off = K; off < 32; off <<= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, m, off);
    const float lo = __shfl_xor_sync(0xffffffffu, l, off);
    const float mn = fmaxf(m, mo);
    const float s0 = (m == -INFINITY) ? 0.f : __expf(m - mn);
    const float s1 = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
    l = l * s0 + lo * s1;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float ao = __shfl_xor_sync(0xffffffffu, acc[h], off);
      acc[h] = acc[h] * s0 + ao * s1;
    }
    m = mn;
  }

  if (slot == 0) {
    float lse;
    if (end > start) {
      const float rinv = 1.f / l;
      lse = m + __logf(l);
#pragma unroll
      for (int h = 0; h < H; ++h) acc[h] *= rinv;
    } else {
      // row without any edge: the dense path degenerates to uniform 1/N over ALL nodes
      // (SURVEY.md section 0.6a); colmean = mean_j S_j when the builder flagged such rows.
      lse = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) acc[h] = colmean ? colmean[head * H + h] : 0.f;
    }
    R[row * RS + D + K + head] = lse;
    float* vp = vsave + row * D + head * H;
    float* op = out + row * out_stride + head * H;
#pragma unroll
    for (int q = 0; q < HV; ++q) {
      float4 a = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
      *reinterpret_cast<float4*>(vp + 4 * q) = a;
      const float4 b = ldg4(bias + head * H + 4 * q);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      if (act == HAN_ACT_ELU) {
        a.x = a.x > 0.f ? a.x : expm1f(a.x);
        a.y = a.y > 0.f ? a.y : expm1f(a.y);
        a.z = a.z > 0.f ? a.z : expm1f(a.z);
        a.w = a.w > 0.f ? a.w : expm1f(a.w);
      }
      *reinterpret_cast<float4*>(op + 4 * q) = a;
    }
  }
}

// per-edge coefficients (return_coef, utils/layers.py:43-44): alpha[e][k] recomputed from m, rinv
template <int K, int H>
__global__ void __launch_bounds__(256)
attn_coefs_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  int64_t n_dst, const float* __restrict__ T, const float* __restrict__ R,
                  float* __restrict__ alpha) {
  constexpr int D = K * H;
  constexpr int TS = ((D + K + 3) / 4) * 4;
  constexpr int RS = ((D + 3 * K + 3) / 4) * 4;
  constexpr int SLOTS = 32 / K;
  const int lane = threadIdx.x & 31;
  const int head = lane % K, slot = lane / K;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_dst) return;
  const int64_t start = indptr[row], end = indptr[row + 1];
  const float f1v = R[row * RS + D + head];
  const float lse = R[row * RS + D + K + head];
  for (int64_t e = start + slot; e < end; e += SLOTS) {
    const int col = indices[e];
    const float f2 = __ldg(T + (int64_t)col * TS + D + head);
    alpha[e * K + head] = __expf(leaky(f1v + f2) - lse);
  }
}

template <int K, int H>
static int launch_fwd(const int64_t* indptr, const int32_t* indices, int64_t n_dst, const float* T,
                      float* R, const float* bias, int act, float* out, int64_t out_stride,
                      float* vsave, const float* colmean, cudaStream_t st) {
  constexpr int UNROLL = (K >= 8) ? 4 : (K >= 4 ? 2 : 1);
  const int warps = 8;
  unsigned grid = (unsigned)ceil_div64(n_dst, warps);
  attn_fwd_kernel<K, H, UNROLL><<<grid, warps * 32, 0, st>>>(indptr, indices, n_dst, T, R, bias, act,
                                                           out, out_stride, vsave, colmean);
  return check_launch("han_attn_fwd");
}

template <int K, int H>
static int launch_coefs(const int64_t* indptr, const int32_t* indices, int64_t n_dst, const float* T,
                        const float* R, float* alpha, cudaStream_t st) {
  unsigned grid = (unsigned)ceil_div64(n_dst, 8);
  attn_coefs_kernel<K, H><<<grid, 256, 0, st>>>(indptr, indices, n_dst, T, R, alpha);
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

int han_table_stride(int K, int H) { return ((K * H + K + 3) / 4) * 4; }
int han_record_stride(int K, int H) { return ((K * H + 3 * K + 3) / 4) * 4; }

int han_attn_fwd(const int64_t* indptr, const int32_t* indices, int64_t n_dst, const float* T,
                 float* R, const float* bias, int K, int H, int act, float* out, int64_t out_stride,
                 float* vsave, const float* colmean, han_stream_t stream) {
  HAN_REQUIRE(indptr && T && R && bias && out && vsave, "null pointer");
  HAN_REQUIRE(n_dst > 0, "n_dst > 0 required");
  HAN_REQUIRE(act == HAN_ACT_ELU || act == HAN_ACT_IDENTITY, "activation");
  HAN_REQUIRE(out_stride >= (int64_t)K * H && out_stride % 4 == 0, "out_stride");
  HAN_REQUIRE(((uintptr_t)T % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)vsave % 16 == 0) &&
              ((uintptr_t)bias % 16 == 0), "16-byte alignment");
#define X(k, h)            \
  if (K == k && H == h)    \
    return launch_fwd<k, h>(indptr, indices, n_dst, T, R, bias, act, out, out_stride, vsave, colmean, as_stream(stream));
  HAN_FOR_SHAPES(X)
#undef X
  return fail_arg(__func__, "unsupported (K,H); see han_attn_shape_supported");
}

int han_attn_coefs(const int64_t* indptr, const int32_t* indices, int64_t n_dst, const float* T,
                   const float* R, int K, int H, float* alpha, han_stream_t stream) {
  HAN_REQUIRE(indptr && T && R && alpha, "null pointer");
  HAN_REQUIRE(n_dst > 0, "n_dst > 0 required");
#define X(k, h)         \
  if (K == k && H == h) \
    return launch_coefs<k, h>(indptr, indices, n_dst, T, R, alpha, as_stream(stream));
  HAN_FOR_SHAPES(X)
#undef X
  return fail_arg(__func__, "unsupported (K,H); see han_attn_shape_supported");
}

}  // extern "C"
