// K-D: backward of the fused edge-softmax-aggregate (autodiff of utils/layers.py:26-35,46 as
// TF would build it via models/base_gattn.py:22, restricted to edges).
//
//   dV_i   = dO_i * act'(V_i + bias)                     delta_i = <dV_i, V_i> per head
//   alpha  = exp(leaky(f1_i + f2_j) - m_i) * rinv_i      (recomputed, never stored)
//   dl_ij  = alpha_ij (dV_i . S_j - delta_i) * leaky'(f1_i + f2_j)
//   dS_j   = sum_i alpha_ij dV_i ;  df2_j = sum_i dl_ij   (by source: transposed structure)
//   df1_i  = sum_j dl_ij = <dV_i, V'_i> - delta_i c_i     (ROW-LOCAL: the forward kept V' and c, see attn_stream.cu)
//
// Design: ONE gather pass, by source (attn_stream.cu).  Everything an edge needs from its destination row
// lives in one contiguous row record R_i = [dV | f1 | lse | delta] (352 B for K=H=8), so the by-source
// kernel gathers exactly one record per edge and keeps S_j / f2_j in registers.  No per-edge array, no
// by-destination pass, no atomics: the backward is deterministic.  This file holds the row-local kernels around
// that pass: prep (dV, delta, df1 -> records), finish (dS_tot, parameter-gradient partials).
#include "han_common.cuh"
#include "han_rng.cuh"

namespace han {

// multimem.st: one store to a multicast address lands in the same offset of every rank's buffer (NVLS)
__device__ __forceinline__ void mc_store4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mc_store1(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// local rows -> multicast rows (for producers without a fused multicast epilogue)
__global__ void multicast_copy_kernel(const float4* __restrict__ src, float* __restrict__ dst_mc, int64_t n_vec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x)
    mc_store4(dst_mc + 4 * i, src[i]);
}

// ---- prep: row-local ---------------------------------------------------------------------------
template <int K, int H>
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const float* __restrict__ dout, int64_t dout_stride, const float* __restrict__ out,
                     int64_t out_stride, const float* __restrict__ vsave, float* __restrict__ R,
                     int64_t n_dst, int act, float* __restrict__ dbias_partial, float* __restrict__ Rmc,
                     int64_t r_row0, const float* __restrict__ vsave2, const float* __restrict__ csave,
                     float* __restrict__ df1) {
  constexpr int D = K * H;
  constexpr int RS = ((D + 3 * K + 3) / 4) * 4;
  // thread = (row, head); a block covers 256/K consecutive rows; grid-stride over row tiles
  const int head = threadIdx.x % K;
  const int rsub = threadIdx.x / K;
  constexpr int ROWS = 256 / K;
  float colsum[H];
#pragma unroll
  for (int h = 0; h < H; ++h) colsum[h] = 0.f;
  for (int64_t r0 = (int64_t)blockIdx.x * ROWS; r0 < n_dst; r0 += (int64_t)gridDim.x * ROWS) {
    const int64_t row = r0 + rsub;
    if (row < n_dst) {
      float delta = 0.f, dv2 = 0.f;
#pragma unroll
      for (int q = 0; q < H / 4; ++q) {
        float4 g = ldg4_stream(dout + row * dout_stride + head * H + 4 * q);
        if (act == HAN_ACT_ELU) {
          // TF EluGrad: outputs < 0 ? grad * (outputs + 1) : grad
          const float4 o = ldg4_stream(out + row * out_stride + head * H + 4 * q);
          g.x = o.x < 0.f ? g.x * (o.x + 1.f) : g.x;
          g.y = o.y < 0.f ? g.y * (o.y + 1.f) : g.y;
          g.z = o.z < 0.f ? g.z * (o.z + 1.f) : g.z;
          g.w = o.w < 0.f ? g.w * (o.w + 1.f) : g.w;
        }
        const float4 v = ldg4_stream(vsave + row * D + head * H + 4 * q);
        delta += g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w;
        if (df1 != nullptr) {
          const float4 v2 = ldg4_stream(vsave2 + row * D + head * H + 4 * q);
          dv2 += g.x * v2.x + g.y * v2.y + g.z * v2.z + g.w * v2.w;
        }
        if (Rmc != nullptr)   // sharded: the record goes to every rank's copy with one multicast store
          mc_store4(Rmc + (r_row0 + row) * RS + head * H + 4 * q, g);
        else
          *reinterpret_cast<float4*>(R + row * RS + head * H + 4 * q) = g;
        colsum[4 * q] += g.x; colsum[4 * q + 1] += g.y; colsum[4 * q + 2] += g.z; colsum[4 * q + 3] += g.w;
      }
      // df1_i = sum_j dl_ij = <dV_i, V'_i> - delta_i c_i : row-local thanks to the forward's second aggregate
      if (df1 != nullptr) df1[row * K + head] = dv2 - delta * csave[row * K + head];
      if (Rmc != nullptr) {
        float* rp = Rmc + (r_row0 + row) * RS + D + head;
        mc_store1(rp, R[row * RS + D + head]);              // f1  (written by the projection, local)
        mc_store1(rp + K, R[row * RS + D + K + head]);      // lse (written by the forward, local)
        mc_store1(rp + 2 * K, delta);
      } else {
        R[row * RS + D + 2 * K + head] = delta;
      }
    }
  }
  // block reduce of column sums over the ROWS sub-rows -> dbias_partial[block][D]
  __shared__ float red[256 * H];
#pragma unroll
  for (int h = 0; h < H; ++h) red[(rsub * K + head) * H + h] = colsum[h];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float s = 0.f;
    for (int r = 0; r < ROWS; ++r) s += red[r * D + c];
    dbias_partial[(int64_t)blockIdx.x * D + c] = s;
  }
}

// ---- finish: row-local; dS_tot = dS_agg + df1 a1^T + df2 a2^T and parameter-gradient partials ----
template <int K, int H>
__global__ void __launch_bounds__(256)
attn_bwd_finish_kernel(const float* __restrict__ T, int64_t n, const float* __restrict__ a1,
                       const float* __restrict__ a2, const float* __restrict__ df1,
                       const float* __restrict__ df2, float* __restrict__ dS, float* __restrict__ part,
                       const uint32_t* __restrict__ seed_ptr, uint32_t thr,
                       float inv_keep, uint32_t metapath, int64_t row0) {
  constexpr int D = K * H;
  constexpr int TS = D;
  constexpr int ROWS = 256 / K;
  const int head = threadIdx.x % K;
  const int rsub = threadIdx.x / K;
  // training-mode dropout of the projected features (layers.py:31-32): dS_agg is the gradient w.r.t. the
  // DROPPED S, so it goes through the same mask; a1/a2 gradients need the un-dropped S (what the table holds)
  const uint32_t sseed = thr ? stream_seed(*seed_ptr, 2u, metapath, 0u) : 0u;
  float a1v[H], a2v[H], da1[H], da2[H];
  float db1 = 0.f, db2 = 0.f;
#pragma unroll
  for (int h = 0; h < H; ++h) {
    a1v[h] = a1[head * H + h];
    a2v[h] = a2[head * H + h];
    da1[h] = 0.f;
    da2[h] = 0.f;
  }
  for (int64_t r0 = (int64_t)blockIdx.x * ROWS; r0 < n; r0 += (int64_t)gridDim.x * ROWS) {
    const int64_t row = r0 + rsub;
    if (row < n) {
      const float g1 = df1[row * K + head], g2 = df2[row * K + head];
      db1 += g1;
      db2 += g2;
#pragma unroll
      for (int q = 0; q < H / 4; ++q) {
        const float4 s = ldg4_stream(T + row * TS + head * H + 4 * q);
        float4 d = *reinterpret_cast<const float4*>(dS + row * D + head * H + 4 * q);
        if (thr) {
          const uint32_t node = (uint32_t)(row + row0);
          const int dd = head * H + 4 * q;
          d.x = keep24(sseed, node, (uint32_t)(dd + 0), thr) ? d.x * inv_keep : 0.f;
          d.y = keep24(sseed, node, (uint32_t)(dd + 1), thr) ? d.y * inv_keep : 0.f;
          d.z = keep24(sseed, node, (uint32_t)(dd + 2), thr) ? d.z * inv_keep : 0.f;
          d.w = keep24(sseed, node, (uint32_t)(dd + 3), thr) ? d.w * inv_keep : 0.f;
        }
        d.x += g1 * a1v[4 * q] + g2 * a2v[4 * q];
        d.y += g1 * a1v[4 * q + 1] + g2 * a2v[4 * q + 1];
        d.z += g1 * a1v[4 * q + 2] + g2 * a2v[4 * q + 2];
        d.w += g1 * a1v[4 * q + 3] + g2 * a2v[4 * q + 3];
        *reinterpret_cast<float4*>(dS + row * D + head * H + 4 * q) = d;
        da1[4 * q] += s.x * g1; da1[4 * q + 1] += s.y * g1; da1[4 * q + 2] += s.z * g1; da1[4 * q + 3] += s.w * g1;
        da2[4 * q] += s.x * g2; da2[4 * q + 1] += s.y * g2; da2[4 * q + 2] += s.z * g2; da2[4 * q + 3] += s.w * g2;
      }
    }
  }
  // block reduce over rsub -> part[block][ da1 (D) | da2 (D) | db1 (K) | db2 (K) ]
  constexpr int W = 2 * H + 2;
  __shared__ float red[256 * W];
  float* mine = red + (size_t)(rsub * K + head) * W;
#pragma unroll
  for (int h = 0; h < H; ++h) {
    mine[h] = da1[h];
    mine[H + h] = da2[h];
  }
  mine[2 * H] = db1;
  mine[2 * H + 1] = db2;
  __syncthreads();
  float* dst = part + (int64_t)blockIdx.x * (2 * D + 2 * K);
  for (int c = threadIdx.x; c < 2 * D + 2 * K; c += 256) {
    int hd, w;
    if (c < D) { hd = c / H; w = c % H; }
    else if (c < 2 * D) { hd = (c - D) / H; w = H + (c - D) % H; }
    else if (c < 2 * D + K) { hd = c - 2 * D; w = 2 * H; }
    else { hd = c - 2 * D - K; w = 2 * H + 1; }
    float s = 0.f;
    for (int r = 0; r < ROWS; ++r) s += red[(size_t)(r * K + hd) * W + w];
    dst[c] = s;
  }
}

// outv[c] = sum_b part[b][c], deterministic: 32 columns x 8 row groups per CTA; row group j sums blocks j, j+8, ...
// with 4 independent accumulators (the loads pipeline instead of one dependent chain of nblocks L2 round trips), the 8
// group sums are then added in a fixed order.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, int nblocks, int64_t cols, float* __restrict__ outv) {
  __shared__ float sm[8][33];
  const int lc = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + lc;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < cols) {
    int b = j;
    for (; b + 24 < nblocks; b += 32) {
      s0 += part[(int64_t)b * cols + c];
      s1 += part[(int64_t)(b + 8) * cols + c];
      s2 += part[(int64_t)(b + 16) * cols + c];
      s3 += part[(int64_t)(b + 24) * cols + c];
    }
    for (; b < nblocks; b += 8) s0 += part[(int64_t)b * cols + c];
  }
  sm[j][lc] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (j == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][lc];
    outv[c] = s;
  }
}

}  // namespace han

using namespace han;

#define HAN_FOR_SHAPES(X) X(8, 8) X(4, 8) X(2, 8) X(1, 8) X(8, 4) X(4, 4) X(1, 4) X(8, 16) X(4, 16) X(1, 16) X(16, 4) X(16, 8)

extern "C" {

int han_reduce_blocks(void) { return kReduceBlocks; }

int han_multicast_copy(const float* src, float* dst_mc, int64_t n_floats, han_stream_t stream) {
  HAN_REQUIRE(src && dst_mc, "null pointer");
  HAN_REQUIRE(n_floats > 0 && n_floats % 4 == 0, "n_floats must be a positive multiple of 4");
  HAN_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst_mc % 16 == 0), "16-byte alignment");
  // a small grid is enough to saturate NVLink egress and leaves the SMs to the compute kernels
  multicast_copy_kernel<<<kNumSMs / 4, 512, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(src), dst_mc,
                                                                   n_floats / 4);
  return check_launch(__func__);
}

int han_attn_bwd_prep(const float* dout, int64_t dout_stride, const float* out, int64_t out_stride,
                      const float* vsave, float* R, int64_t n_dst, int K, int H, int act,
                      float* dbias_partial, float* R_mc, int64_t r_row0, const float* vsave2, const float* csave,
                      float* df1, han_stream_t stream) {
  HAN_REQUIRE(dout && out && vsave && R && dbias_partial, "null pointer");
  HAN_REQUIRE(n_dst > 0 && r_row0 >= 0, "n_dst > 0 required");
  HAN_REQUIRE(dout_stride % 4 == 0 && out_stride % 4 == 0, "strides must be multiples of 4 floats");
  HAN_REQUIRE((uintptr_t)R_mc % 16 == 0, "multicast records must be 16-byte aligned");
  HAN_REQUIRE(!df1 || (vsave2 && csave && (uintptr_t)vsave2 % 16 == 0), "df1 needs vsave2 and csave");
#define X(k, h)                                                                                        \
  if (K == k && H == h) {                                                                              \
    attn_bwd_prep_kernel<k, h><<<kReduceBlocks, 256, 0, as_stream(stream)>>>(                          \
        dout, dout_stride, out, out_stride, vsave, R, n_dst, act, dbias_partial, R_mc, r_row0, vsave2,   \
        csave, df1);                                                                                   \
    return check_launch(__func__);                                                                     \
  }
  HAN_FOR_SHAPES(X)
#undef X
  return fail_arg(__func__, "unsupported (K,H)");
}

int han_attn_bwd_finish(const float* T, int64_t n, int K, int H, const float* a1, const float* a2,
                        const float* df1, const float* df2, float* dS, float* part,
                        const uint32_t* seed_ptr, float in_keep, int metapath, int64_t row0,
                        han_stream_t stream) {
  HAN_REQUIRE(T && a1 && a2 && df1 && df2 && dS && part, "null pointer");
  HAN_REQUIRE(n > 0, "n > 0 required");
  HAN_REQUIRE(in_keep > 0.f && in_keep <= 1.f && (in_keep == 1.f || seed_ptr), "in_keep in (0,1]; dropout needs seed_ptr");
  const uint32_t thr = in_keep < 1.f ? (uint32_t)(in_keep * 16777216.f + 0.5f) : 0u;
  const float inv_keep = in_keep < 1.f ? 1.f / ((float)thr / 16777216.f) : 1.f;
#define X(k, h)                                                                                        \
  if (K == k && H == h) {                                                                              \
    attn_bwd_finish_kernel<k, h><<<kReduceBlocks, 256, 0, as_stream(stream)>>>(                        \
        T, n, a1, a2, df1, df2, dS, part, seed_ptr, thr, inv_keep, (uint32_t)metapath, row0);          \
    return check_launch(__func__);                                                                     \
  }
  HAN_FOR_SHAPES(X)
#undef X
  return fail_arg(__func__, "unsupported (K,H)");
}

int han_reduce_partials(const float* part, int nblocks, int64_t cols, float* outv, han_stream_t stream) {
  HAN_REQUIRE(part && outv, "null pointer");
  HAN_REQUIRE(nblocks > 0 && cols > 0, "sizes");
  reduce_partials_kernel<<<(unsigned)ceil_div64(cols, 32), 256, 0, as_stream(stream)>>>(part, nblocks, cols, outv);
  return check_launch(__func__);
}

}  // extern "C"
