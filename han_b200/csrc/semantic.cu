// K-C / K-F: semantic-level attention (utils/layers.py:152-159) forward and backward.
//
//   v = tanh(Z w + b)  (N*P x A)      s = v . u  (N x P)      beta = softmax_P(s)  PER NODE (:156)
//   out[n] = sum_p beta[n,p] Z[n,p]   (:159)
//
// Forward: one CTA owns a tile of 128 (node, meta-path) rows; w (D x A) stays in shared memory for
// the whole CTA, the Z tile is staged once and used twice (for Z w and for the weighted sum), the
// tanh / u-dot / softmax / weighted sum are all fused behind the register-tiled FP32 contraction so
// Z is read from HBM exactly once and v is written once (kept for the backward).
// Backward: one persistent CTA per SM slot; dv is formed in shared memory in place of v, dZ = dv w^T
// and the dw = Z^T dv partial products run from the same staged tiles; dw/db/du partials are
// reduced deterministically (fixed grid, two stages).
#include "han_common.cuh"

namespace han {

constexpr int kSemThreads = 256;

// FP32-grade contractions on the tensor pipe: every operand is split x = hi + lo with hi = x truncated to
// TF32 (10 explicit mantissa bits) and lo = x - hi (exact in FP32; the tensor core reads its top 19 bits),
// and acc += lo_a*hi_b + hi_a*lo_b + hi_a*hi_b in FP32 (small terms first).  The dropped lo*lo term is
// <= 2^-22 relative.  Warp-level m16n8k8 fragments (PTX ISA "mma.m16n8k8 .tf32"): with gid = lane/4 and
// tig = lane%4,  A: a0=(gid,tig) a1=(gid+8,tig) a2=(gid,tig+4) a3=(gid+8,tig+4);  B: b0=(k=tig,n=gid)
// b1=(k=tig+4,n=gid);  C: c0=(gid,2tig) c1=(gid,2tig+1) c2=(gid+8,2tig) c3=(gid+8,2tig+1).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                           const uint32_t (&bhi)[2], const uint32_t (&blo)[2]) {
  mma_tf32(c, alo, bhi);
  mma_tf32(c, ahi, blo);
  mma_tf32(c, ahi, bhi);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int D, int A>
struct SemFwdCfg {
  static constexpr int TM = 128;            // rows (n,p) per tile
  static constexpr int TX = A / 8;          // column groups (8 cols each: 4 at tx*4, 4 at A/2+tx*4)
  static constexpr int TY = kSemThreads / TX;
  static constexpr int RT = TM / TY;        // rows per thread
  static constexpr int ZLD = D + 4;            // A-fragment reads (gid*ZLD + tig) hit 32 distinct banks
  // tensor-pipe path: w kept pre-split (hi | lo) with a leading dimension = 8 mod 32 so the B-fragment
  // reads (tig*WLD + gid) are conflict-free; 8 warps as 4 (rows) x 2 (columns), warp tile 32 x A/2
  static constexpr bool MMA = (D % 8 == 0) && (A % 16 == 0);
  static constexpr int WLD = MMA ? A + 8 : A;
  static constexpr size_t w_floats = MMA ? (size_t)2 * D * WLD : (size_t)D * A;
  static constexpr size_t smem_floats = w_floats + 2 * A + (size_t)TM * ZLD + 2 * TM + TM;
};

template <int D, int A>
__global__ void __launch_bounds__(kSemThreads)
semantic_fwd_kernel(const float* __restrict__ Z, int64_t n, int P, const float* __restrict__ w,
                    const float* __restrict__ b, const float* __restrict__ u, int mode,
                    float* __restrict__ out, float* __restrict__ beta, float* __restrict__ vsave,
                    float* __restrict__ scores) {
  using C = SemFwdCfg<D, A>;
  constexpr int TM = C::TM, TX = C::TX, TY = C::TY, RT = C::RT, ZLD = C::ZLD;
  static_assert(A % 8 == 0 && TX >= 4 && TX <= 16 && (TX & (TX - 1)) == 0, "A in {32,64,128}");
  static_assert(D % 4 == 0, "D % 4");
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                 // [D][A]   (tensor-pipe path: hi [D][WLD] | lo [D][WLD])
  float* bs = ws + C::w_floats;     // [A]
  float* us = bs + A;               // [A]
  float* Zs = us + A;               // [TM][ZLD]
  float* ss = Zs + TM * ZLD;        // [2][TM] scores (one half per column-half of the warp grid)
  float* bts = ss + 2 * TM;         // [TM] beta

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int nodes_per_tile = TM / P;
  const int rows_per_tile = nodes_per_tile * P;
  const int64_t n_tiles = ceil_div64(n, nodes_per_tile);

  if constexpr (C::MMA) {
    for (int i = tid; i < D * A; i += kSemThreads) {
      uint32_t hi, lo;
      split_tf32(w[i], hi, lo);
      ws[(i / A) * C::WLD + i % A] = __uint_as_float(hi);
      ws[D * C::WLD + (i / A) * C::WLD + i % A] = __uint_as_float(lo);
    }
  } else {
    for (int i = tid; i < D * A / 4; i += kSemThreads)
      reinterpret_cast<float4*>(ws)[i] = ldg4(w + 4 * i);
  }
  for (int i = tid; i < A; i += kSemThreads) {
    bs[i] = b[i];
    us[i] = u[i];
  }

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t node0 = tile * nodes_per_tile;
    const int64_t row0 = node0 * P;
    const int64_t rows_here = min((int64_t)rows_per_tile, n * P - row0);
    __syncthreads();  // previous tile fully consumed (and w/b/u staged on the first pass)
    // stage Z tile: rows are contiguous in memory ([n][P][D] row-major)
    for (int i = tid; i < TM * (D / 4); i += kSemThreads) {
      const int r = i / (D / 4), c = i % (D / 4);
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows_here) z = ldg4_stream(Z + (row0 + r) * D + 4 * c);
      *reinterpret_cast<float4*>(Zs + r * ZLD + 4 * c) = z;
    }
    __syncthreads();

    if constexpr (C::MMA) {
      constexpr int WLD = C::WLD, NT = A / 16;
      const int warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
      const int wm = warp & 3, wn = warp >> 2;
      const uint32_t* whi = reinterpret_cast<const uint32_t*>(ws);
      const uint32_t* wlo = whi + D * WLD;
      float acc[2][NT][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
#pragma unroll
      for (int k0 = 0; k0 < D; k0 += 8) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const float* zp = Zs + (wm * 32 + mt * 16 + gid) * ZLD + k0 + tig;
          split_tf32(zp[0], ahi[mt][0], alo[mt][0]);
          split_tf32(zp[8 * ZLD], ahi[mt][1], alo[mt][1]);
          split_tf32(zp[4], ahi[mt][2], alo[mt][2]);
          split_tf32(zp[8 * ZLD + 4], ahi[mt][3], alo[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int o = (k0 + tig) * WLD + wn * (A / 2) + nt * 8 + gid;
          const uint32_t bhi[2] = {whi[o], whi[o + 4 * WLD]};
          const uint32_t blo[2] = {wlo[o], wlo[o + 4 * WLD]};
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_3xtf32(acc[mt][nt], ahi[mt], alo[mt], bhi, blo);
        }
      }
      // epilogue: tanh, store of v, dot with u over this warp's A/2 columns (4 lanes share a row)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int rl = wm * 32 + mt * 16 + gid + 8 * h;
          float part = 0.f;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int col = wn * (A / 2) + nt * 8 + 2 * tig;
            const float v0 = tanhf(acc[mt][nt][2 * h] + bs[col]);
            const float v1 = tanhf(acc[mt][nt][2 * h + 1] + bs[col + 1]);
            part = fmaf(v0, us[col], part);
            part = fmaf(v1, us[col + 1], part);
            if (vsave != nullptr && rl < rows_here)
              *reinterpret_cast<float2*>(vsave + (row0 + rl) * A + col) = make_float2(v0, v1);
          }
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          if (tig == 0) ss[wn * TM + rl] = part;
        }
      __syncthreads();
      for (int r = tid; r < TM; r += kSemThreads) ss[r] += ss[TM + r];
    } else {
      float acc[RT][8];
  #pragma unroll
      for (int r = 0; r < RT; ++r)
  #pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

  #pragma unroll 2
      for (int d = 0; d < D; d += 4) {
        float4 a[RT];
  #pragma unroll
        for (int r = 0; r < RT; ++r) a[r] = *reinterpret_cast<const float4*>(Zs + (r * TY + ty) * ZLD + d);
  #pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          const float4 w0 = *reinterpret_cast<const float4*>(ws + (d + dd) * A + tx * 4);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + (d + dd) * A + A / 2 + tx * 4);
  #pragma unroll
          for (int r = 0; r < RT; ++r) {
            const float av = dd == 0 ? a[r].x : (dd == 1 ? a[r].y : (dd == 2 ? a[r].z : a[r].w));
            acc[r][0] = fmaf(av, w0.x, acc[r][0]);
            acc[r][1] = fmaf(av, w0.y, acc[r][1]);
            acc[r][2] = fmaf(av, w0.z, acc[r][2]);
            acc[r][3] = fmaf(av, w0.w, acc[r][3]);
            acc[r][4] = fmaf(av, w1.x, acc[r][4]);
            acc[r][5] = fmaf(av, w1.y, acc[r][5]);
            acc[r][6] = fmaf(av, w1.z, acc[r][6]);
            acc[r][7] = fmaf(av, w1.w, acc[r][7]);
          }
        }
      }
      // epilogue: tanh, optional store of v, dot with u, reduce over the TX lanes of a row
  #pragma unroll
      for (int r = 0; r < RT; ++r) {
        const int rl = r * TY + ty;
        float part = 0.f;
        float vv[8];
  #pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int col = (c < 4) ? (tx * 4 + c) : (A / 2 + tx * 4 + (c - 4));
          vv[c] = tanhf(acc[r][c] + bs[col]);
          part = fmaf(vv[c], us[col], part);
        }
        if (vsave != nullptr && rl < rows_here) {
          float* vp = vsave + (row0 + rl) * A;
          *reinterpret_cast<float4*>(vp + tx * 4) = make_float4(vv[0], vv[1], vv[2], vv[3]);
          *reinterpret_cast<float4*>(vp + A / 2 + tx * 4) = make_float4(vv[4], vv[5], vv[6], vv[7]);
        }
  #pragma unroll
        for (int o = TX / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (tx == 0) ss[rl] = part;
      }
    }
    __syncthreads();
    if (scores != nullptr)
      for (int r = tid; r < rows_here; r += kSemThreads) scores[row0 + r] = ss[r];
    if (mode == HAN_SEM_REFERENCE) {
      // per-node softmax over the P meta-paths (utils/layers.py:156)
      const int nodes_here = (int)(rows_here / P);
      for (int nl = tid; nl < nodes_here; nl += kSemThreads) {
        float mx = -INFINITY;
        for (int p = 0; p < P; ++p) mx = fmaxf(mx, ss[nl * P + p]);
        float sum = 0.f;
        for (int p = 0; p < P; ++p) {
          const float e = expf(ss[nl * P + p] - mx);
          bts[nl * P + p] = e;
          sum += e;
        }
        const float inv = 1.f / sum;
        for (int p = 0; p < P; ++p) {
          const float bt = bts[nl * P + p] * inv;
          bts[nl * P + p] = bt;
          beta[(node0 + nl) * P + p] = bt;
        }
      }
      __syncthreads();
      // out[n] = sum_p beta[n,p] Z[n,p]  (:159), Z from the staged tile
      for (int i = tid; i < nodes_here * (D / 4); i += kSemThreads) {
        const int nl = i / (D / 4), c = i % (D / 4);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < P; ++p) {
          const float bt = bts[nl * P + p];
          const float4 z = *reinterpret_cast<const float4*>(Zs + (nl * P + p) * ZLD + 4 * c);
          o.x = fmaf(bt, z.x, o.x); o.y = fmaf(bt, z.y, o.y); o.z = fmaf(bt, z.z, o.z); o.w = fmaf(bt, z.w, o.w);
        }
        *reinterpret_cast<float4*>(out + (node0 + nl) * D + 4 * c) = o;
      }
    }
  }
}

// paper mode second phase: out[n] = sum_p beta_p Z[n,p] with one global beta (han.pdf Eq. 9)
__global__ void semantic_combine_kernel(const float* __restrict__ Z, int64_t n, int P, int D,
                                        const float* __restrict__ beta_vec, float* __restrict__ out,
                                        float* __restrict__ beta) {
  const int64_t total = n * (D / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = i / (D / 4);
    const int c = (int)(i % (D / 4));
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < P; ++p) {
      const float bt = beta_vec[p];
      const float4 z = ldg4_stream(Z + (node * P + p) * D + 4 * c);
      o.x = fmaf(bt, z.x, o.x); o.y = fmaf(bt, z.y, o.y); o.z = fmaf(bt, z.z, o.z); o.w = fmaf(bt, z.w, o.w);
      if (c == 0 && beta != nullptr) beta[node * P + p] = bt;
    }
    *reinterpret_cast<float4*>(out + node * D + 4 * c) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int D, int A>
struct SemBwdCfg {
  static constexpr int TM = 64;
  // tensor-pipe path (3xTF32 mma.sync): dZ = dv wT as 4 x 2 warps of 16 x D/2; dw += Z^T dv with the
  // (D/16) x (A/8) accumulator tiles dealt to the 8 warps in runs of TPW along A.  Leading dimensions:
  // VLD = 4 mod 32 (dv read as a row-major A operand), WLD and ZLD = 8 mod 32 (wT read as a B operand, Z
  // read as a column-major A operand); dv as the B operand of the dw product is then 2-way conflicted.
  static constexpr bool MMA = (D % 16 == 0) && (A % 16 == 0);
  static constexpr int DW_TILES = (D / 16) * (A / 8);
  static constexpr int TPW = (DW_TILES + 7) / 8;
  static constexpr int ZLD = MMA ? D + 8 : D + 4;
  static constexpr int VLD = A + 4;
  static constexpr int WLD = MMA ? D + 8 : D + 4;   // wT [A][WLD]
  static constexpr int DZ_MT = ((TM / 4) * (D / 4) + kSemThreads - 1) / kSemThreads;   // 4x4 micro-tiles
  static constexpr int DW_MT = ((D / 4) * (A / 8) + kSemThreads - 1) / kSemThreads;    // 4x8 micro-tiles
  static constexpr int RS = kSemThreads / A;   // row groups of the column pass (thread = column x row group)
  static constexpr size_t smem_floats =
      (size_t)A * WLD + A + (size_t)TM * ZLD + (size_t)TM * VLD + (size_t)TM * ZLD /*dout rows*/ + 3 * TM;
  static constexpr size_t part_floats = (size_t)D * A + 2 * A;  // per block: dw | db | du
};

template <int D, int A>
__global__ void __launch_bounds__(kSemThreads)
semantic_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ Z,
                    const float* __restrict__ beta, const float* __restrict__ vsave, int64_t n, int P,
                    const float* __restrict__ w, const float* __restrict__ u, int mode,
                    const float* __restrict__ dsbar, float* __restrict__ dZ, float* __restrict__ part,
                    float* const* __restrict__ dz_tab, int64_t dz_stride) {
  using C = SemBwdCfg<D, A>;
  // Destination of the dZ row of (node, meta-path p): dZ[node][p][:] locally, or -- tile-sharded multi-GPU runs --
  // dz_tab[p] + node * dz_stride: a peer-mapped address inside the GPU that owns meta-path p, so the re-sharding
  // all-to-all of dZ is this kernel's own stores over NVLink.
  auto dz_row = [&](int64_t row) -> float* {
    if (dz_tab == nullptr) return dZ + row * D;
    const int64_t node = row / P;
    return dz_tab[row - node * P] + node * dz_stride;
  };
  constexpr int TM = C::TM, ZLD = C::ZLD, VLD = C::VLD, WLD = C::WLD;
  extern __shared__ __align__(16) float smem[];
  float* wT = smem;                  // [A][WLD]   wT[a][d] = w[d][a]
  float* us = wT + A * WLD;          // [A]
  float* Zs = us + A;                // [TM][ZLD]
  float* Vs = Zs + TM * ZLD;         // [TM][VLD]  v, then dv in place
  float* Gs = Vs + TM * VLD;         // [TM][ZLD]  dout row of the node of each (n,p) row
  float* gs = Gs + TM * ZLD;         // [TM] g = <dout, Z>
  float* dss = gs + TM;              // [TM] ds
  float* bts = dss + TM;             // [TM] beta

  const int tid = threadIdx.x;
  const int nodes_per_tile = TM / P;
  const int rows_per_tile = nodes_per_tile * P;
  const int64_t n_tiles = ceil_div64(n, nodes_per_tile);

  for (int i = tid; i < D * A; i += kSemThreads) {
    const int d = i / A, a = i % A;
    wT[a * WLD + d] = w[i];
  }
  for (int i = tid; i < A; i += kSemThreads) us[i] = u[i];

  constexpr int kDwFfma = C::MMA ? 1 : C::DW_MT;
  float dw_acc[kDwFfma][4][8];       // FFMA path: 4x8 micro-tiles
#pragma unroll
  for (int t = 0; t < kDwFfma; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) dw_acc[t][i][j] = 0.f;
  constexpr int kDwMma = C::MMA ? C::TPW : 1;
  float dw_frag[kDwMma][4];          // tensor-pipe path: m16n8 accumulator fragments
#pragma unroll
  for (int t = 0; t < kDwMma; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) dw_frag[t][c] = 0.f;
  const int warp = tid >> 5, gid = (tid & 31) >> 2, tig = tid & 3;
  static_assert(kSemThreads % A == 0 && TM % C::RS == 0, "column pass: A divides the CTA, RS divides TM");
  float db_acc = 0.f, du_acc = 0.f;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t node0 = tile * nodes_per_tile;
    const int64_t row0 = node0 * P;
    const int rows_here = (int)min((int64_t)rows_per_tile, n * P - row0);
    __syncthreads();
    for (int i = tid; i < TM * (D / 4); i += kSemThreads) {
      const int r = i / (D / 4), c = i % (D / 4);
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f), g = z;
      if (r < rows_here) {
        z = ldg4_stream(Z + (row0 + r) * D + 4 * c);
        g = ldg4(dout + (node0 + r / P) * D + 4 * c);
      }
      *reinterpret_cast<float4*>(Zs + r * ZLD + 4 * c) = z;
      *reinterpret_cast<float4*>(Gs + r * ZLD + 4 * c) = g;
    }
    for (int i = tid; i < TM * (A / 4); i += kSemThreads) {
      const int r = i / (A / 4), c = i % (A / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows_here) v = ldg4_stream(vsave + (row0 + r) * A + 4 * c);
      *reinterpret_cast<float4*>(Vs + r * VLD + 4 * c) = v;
    }
    for (int r = tid; r < TM; r += kSemThreads) bts[r] = (r < rows_here) ? beta[row0 + r] : 0.f;
    __syncthreads();
    // g[r] = <dout[n], Z[r]> : 4 threads per row
    {
      const int r = tid / 4, q = tid % 4;
      float s = 0.f;
      if (r < TM)
        for (int d = q; d < D; d += 4) s = fmaf(Gs[r * ZLD + d], Zs[r * ZLD + d], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0 && r < TM) gs[r] = s;
    }
    __syncthreads();
    if (tid < TM) {
      float ds = 0.f;
      if (tid < rows_here) {
        const int p = tid % P;
        if (mode == HAN_SEM_REFERENCE) {
          const int nl = tid / P;
          float dot = 0.f;
          for (int q = 0; q < P; ++q) dot = fmaf(bts[nl * P + q], gs[nl * P + q], dot);
          ds = bts[tid] * (gs[tid] - dot);
        } else {
          ds = dsbar[p];  // paper mode: d s_bar_p / N, identical for every node
        }
      }
      dss[tid] = ds;
    }
    __syncthreads();
    // dv = ds * u * (1 - v^2) in place; du += ds * v; db += dv   (thread = column a, rows rg, rg+RS, ...)
    {
      const int a = tid % A, rg = tid / A;
      const float ua = us[a];
      float du = 0.f, db = 0.f;
#pragma unroll 4
      for (int r = rg; r < TM; r += C::RS) {
        const float v = Vs[r * VLD + a];
        const float ds = dss[r];
        const float dv = ds * ua * (1.f - v * v);
        du = fmaf(ds, v, du);
        db += dv;
        Vs[r * VLD + a] = dv;
      }
      du_acc += du;
      db_acc += db;
    }
    __syncthreads();
    if constexpr (C::MMA) {
      // dZ[r][d] = beta[r] dout[n][d] + sum_a dv[r][a] w[d][a]
      {
        static_assert(C::MMA ? (A / 8) % C::TPW == 0 : true, "dw tile runs stay inside one 16-row band");
        constexpr int NT = D / 16;
        const int wm = warp & 3, wn = warp >> 2;
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[nt][c] = 0.f;
#pragma unroll 4
        for (int k0 = 0; k0 < A; k0 += 8) {
          uint32_t ahi[4], alo[4];
          const float* vp = Vs + (wm * 16 + gid) * VLD + k0 + tig;
          split_tf32(vp[0], ahi[0], alo[0]);
          split_tf32(vp[8 * VLD], ahi[1], alo[1]);
          split_tf32(vp[4], ahi[2], alo[2]);
          split_tf32(vp[8 * VLD + 4], ahi[3], alo[3]);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const float* wp = wT + (k0 + tig) * WLD + wn * (D / 2) + nt * 8 + gid;
            uint32_t bhi[2], blo[2];
            split_tf32(wp[0], bhi[0], blo[0]);
            split_tf32(wp[4 * WLD], bhi[1], blo[1]);
            mma_3xtf32(acc[nt], ahi, alo, bhi, blo);
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = wm * 16 + gid + 8 * h;
          if (r < rows_here) {
            const float bt = bts[r];
            float* zrow = dz_row(row0 + r);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const int d = wn * (D / 2) + nt * 8 + 2 * tig;
              const float2 gg = *reinterpret_cast<const float2*>(Gs + r * ZLD + d);
              *reinterpret_cast<float2*>(zrow + d) =
                  make_float2(fmaf(bt, gg.x, acc[nt][2 * h]), fmaf(bt, gg.y, acc[nt][2 * h + 1]));
            }
          }
        }
      }
      // dw[d][a] += sum_r Z[r][d] dv[r][a], accumulated across this CTA's tiles
      if (warp * C::TPW < C::DW_TILES) {
        const int tile0 = warp * C::TPW;
        const int m0 = (tile0 / (A / 8)) * 16, n0 = (tile0 % (A / 8)) * 8;
#pragma unroll 2
        for (int k0 = 0; k0 < TM; k0 += 8) {
          uint32_t ahi[4], alo[4];
          const float* zp = Zs + (k0 + tig) * ZLD + m0 + gid;
          split_tf32(zp[0], ahi[0], alo[0]);
          split_tf32(zp[8], ahi[1], alo[1]);
          split_tf32(zp[4 * ZLD], ahi[2], alo[2]);
          split_tf32(zp[4 * ZLD + 8], ahi[3], alo[3]);
#pragma unroll
          for (int j = 0; j < C::TPW; ++j) {
            const float* vp = Vs + (k0 + tig) * VLD + n0 + j * 8 + gid;
            uint32_t bhi[2], blo[2];
            split_tf32(vp[0], bhi[0], blo[0]);
            split_tf32(vp[4 * VLD], bhi[1], blo[1]);
            mma_3xtf32(dw_frag[j], ahi, alo, bhi, blo);
          }
        }
      }
    } else {
      // dZ[r][d] = beta[r] dout[n][d] + sum_a dv[r][a] w[d][a]   (4x4 micro-tiles)
  #pragma unroll
      for (int t = 0; t < C::DZ_MT; ++t) {
        const int mt = tid + t * kSemThreads;
        if (mt < (TM / 4) * (D / 4)) {
          const int dg = mt % (D / 4), rg = mt / (D / 4);
          float acc[4][4];
  #pragma unroll
          for (int i = 0; i < 4; ++i)
  #pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
          for (int a = 0; a < A; a += 4) {
            float4 dv[4], wv[4];
  #pragma unroll
            for (int i = 0; i < 4; ++i) dv[i] = *reinterpret_cast<const float4*>(Vs + (rg * 4 + i) * VLD + a);
  #pragma unroll
            for (int k = 0; k < 4; ++k) wv[k] = *reinterpret_cast<const float4*>(wT + (a + k) * WLD + dg * 4);
  #pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float d0 = dv[i].x, d1 = dv[i].y, d2 = dv[i].z, d3 = dv[i].w;
              acc[i][0] += d0 * wv[0].x + d1 * wv[1].x + d2 * wv[2].x + d3 * wv[3].x;
              acc[i][1] += d0 * wv[0].y + d1 * wv[1].y + d2 * wv[2].y + d3 * wv[3].y;
              acc[i][2] += d0 * wv[0].z + d1 * wv[1].z + d2 * wv[2].z + d3 * wv[3].z;
              acc[i][3] += d0 * wv[0].w + d1 * wv[1].w + d2 * wv[2].w + d3 * wv[3].w;
            }
          }
  #pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rg * 4 + i;
            if (r < rows_here) {
              const float bt = bts[r];
              const float4 g = *reinterpret_cast<const float4*>(Gs + r * ZLD + dg * 4);
              *reinterpret_cast<float4*>(dz_row(row0 + r) + dg * 4) =
                  make_float4(fmaf(bt, g.x, acc[i][0]), fmaf(bt, g.y, acc[i][1]),
                              fmaf(bt, g.z, acc[i][2]), fmaf(bt, g.w, acc[i][3]));
            }
          }
        }
      }
      // dw[d][a] += sum_r Z[r][d] dv[r][a]   (4x8 micro-tiles, accumulated across this CTA's tiles)
  #pragma unroll
      for (int t = 0; t < C::DW_MT; ++t) {
        const int mt = tid + t * kSemThreads;
        if (mt < (D / 4) * (A / 8)) {
          const int ag = mt % (A / 8), dg = mt / (A / 8);
          for (int r = 0; r < TM; ++r) {
            const float4 z = *reinterpret_cast<const float4*>(Zs + r * ZLD + dg * 4);
            const float4 v0 = *reinterpret_cast<const float4*>(Vs + r * VLD + ag * 4);
            const float4 v1 = *reinterpret_cast<const float4*>(Vs + r * VLD + A / 2 + ag * 4);
            const float zz[4] = {z.x, z.y, z.z, z.w};
  #pragma unroll
            for (int i = 0; i < 4; ++i) {
              dw_acc[t][i][0] = fmaf(zz[i], v0.x, dw_acc[t][i][0]);
              dw_acc[t][i][1] = fmaf(zz[i], v0.y, dw_acc[t][i][1]);
              dw_acc[t][i][2] = fmaf(zz[i], v0.z, dw_acc[t][i][2]);
              dw_acc[t][i][3] = fmaf(zz[i], v0.w, dw_acc[t][i][3]);
              dw_acc[t][i][4] = fmaf(zz[i], v1.x, dw_acc[t][i][4]);
              dw_acc[t][i][5] = fmaf(zz[i], v1.y, dw_acc[t][i][5]);
              dw_acc[t][i][6] = fmaf(zz[i], v1.z, dw_acc[t][i][6]);
              dw_acc[t][i][7] = fmaf(zz[i], v1.w, dw_acc[t][i][7]);
            }
          }
        }
      }
    }
  }
  // per-block partials: [dw (D*A) | db (A) | du (A)]
  float* my = part + (size_t)blockIdx.x * C::part_floats;
  if constexpr (C::MMA) {
    if (warp * C::TPW < C::DW_TILES) {
      const int tile0 = warp * C::TPW;
      const int m0 = (tile0 / (A / 8)) * 16, n0 = (tile0 % (A / 8)) * 8;
#pragma unroll
      for (int j = 0; j < C::TPW; ++j) {
        float* cp = my + (size_t)(m0 + gid) * A + n0 + j * 8 + 2 * tig;
        *reinterpret_cast<float2*>(cp) = make_float2(dw_frag[j][0], dw_frag[j][1]);
        *reinterpret_cast<float2*>(cp + 8 * A) = make_float2(dw_frag[j][2], dw_frag[j][3]);
      }
    }
  } else {
  #pragma unroll
    for (int t = 0; t < C::DW_MT; ++t) {
      const int mt = tid + t * kSemThreads;
      if (mt < (D / 4) * (A / 8)) {
        const int ag = mt % (A / 8), dg = mt / (A / 8);
  #pragma unroll
        for (int i = 0; i < 4; ++i) {
          float* rowp = my + (size_t)(dg * 4 + i) * A;
          *reinterpret_cast<float4*>(rowp + ag * 4) =
              make_float4(dw_acc[t][i][0], dw_acc[t][i][1], dw_acc[t][i][2], dw_acc[t][i][3]);
          *reinterpret_cast<float4*>(rowp + A / 2 + ag * 4) =
              make_float4(dw_acc[t][i][4], dw_acc[t][i][5], dw_acc[t][i][6], dw_acc[t][i][7]);
        }
      }
    }
  }
  // db / du: the RS row groups of a column are summed in a fixed order through shared memory
  __syncthreads();
  Vs[tid] = db_acc;
  Vs[kSemThreads + tid] = du_acc;
  __syncthreads();
  if (tid < A) {
    float db = 0.f, du = 0.f;
    for (int rg = 0; rg < C::RS; ++rg) {
      db += Vs[rg * A + tid];
      du += Vs[kSemThreads + rg * A + tid];
    }
    my[(size_t)D * A + tid] = db;
    my[(size_t)D * A + A + tid] = du;
  }
}

__global__ void sem_reduce_kernel(const float* __restrict__ part, int nblocks, int64_t cols,
                                  float* __restrict__ dw, int64_t ndw, float* __restrict__ db,
                                  float* __restrict__ du, int A) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += part[(int64_t)b * cols + c];
  if (c < ndw) dw[c] = s;
  else if (c < ndw + A) db[c - ndw] = s;
  else du[c - ndw - A] = s;
}

constexpr int kSemBwdBlocks = kNumSMs * 2;

template <int D, int A>
static int launch_sem_fwd(const float* Z, int64_t n, int P, const float* w, const float* b,
                          const float* u, int mode, float* out, float* beta, float* vsave,
                          float* scores, cudaStream_t st) {
  using C = SemFwdCfg<D, A>;
  size_t smem = C::smem_floats * sizeof(float);
  HAN_SMEM_ATTR_ONCE((semantic_fwd_kernel<D, A>), smem);
  int64_t n_tiles = ceil_div64(n, C::TM / P);
  unsigned grid = (unsigned)(n_tiles < (int64_t)kNumSMs * 2 ? n_tiles : (int64_t)kNumSMs * 2);
  semantic_fwd_kernel<D, A><<<grid, kSemThreads, smem, st>>>(Z, n, P, w, b, u, mode, out, beta, vsave, scores);
  return check_launch("han_semantic_fwd");
}

template <int D, int A>
static int launch_sem_bwd(const float* dout, const float* Z, const float* beta, const float* vsave,
                          int64_t n, int P, const float* w, const float* u, int mode,
                          const float* dsbar, float* dZ, float* dw, float* db, float* du, void* ws,
                          size_t ws_bytes, float* const* dz_tab, int64_t dz_stride, cudaStream_t st) {
  using C = SemBwdCfg<D, A>;
  size_t smem = C::smem_floats * sizeof(float);
  if (ws_bytes < (size_t)kSemBwdBlocks * C::part_floats * sizeof(float))
    return fail_arg("han_semantic_bwd", "workspace too small");
  HAN_SMEM_ATTR_ONCE((semantic_bwd_kernel<D, A>), smem);
  float* part = reinterpret_cast<float*>(ws);
  semantic_bwd_kernel<D, A><<<kSemBwdBlocks, kSemThreads, smem, st>>>(dout, Z, beta, vsave, n, P, w, u,
                                                                    mode, dsbar, dZ, part, dz_tab, dz_stride);
  int64_t cols = (int64_t)C::part_floats;
  sem_reduce_kernel<<<(unsigned)ceil_div64(cols, 128), 128, 0, st>>>(part, kSemBwdBlocks, cols, dw,
                                                                   (int64_t)D * A, db, du, A);
  return check_launch("han_semantic_bwd");
}

}  // namespace han

using namespace han;

#define HAN_FOR_SEM(X) X(64, 128) X(64, 64) X(32, 128) X(32, 64) X(32, 32) X(16, 32) X(8, 32) X(128, 128)

extern "C" {

int han_semantic_shape_supported(int D, int A) {
#define X(d, a) if (D == d && A == a) return 1;
  HAN_FOR_SEM(X)
#undef X
  return 0;
}

int han_semantic_fwd(const float* Z, int64_t n, int P, int D, int A, const float* w, const float* b,
                     const float* u, int mode, float* out, float* beta, float* vsave, float* scores,
                     han_stream_t stream) {
  HAN_REQUIRE(Z && w && b && u, "null pointer");
  HAN_REQUIRE(n > 0 && P > 0 && P <= 64, "n > 0, 1 <= P <= 64");
  HAN_REQUIRE(mode == HAN_SEM_REFERENCE || mode == HAN_SEM_PAPER, "mode");
  HAN_REQUIRE(mode == HAN_SEM_PAPER || (out && beta), "reference mode needs out and beta");
  HAN_REQUIRE(mode == HAN_SEM_REFERENCE || scores, "paper mode needs scores");
#define X(d, a)         \
  if (D == d && A == a) \
    return launch_sem_fwd<d, a>(Z, n, P, w, b, u, mode, out, beta, vsave, scores, as_stream(stream));
  HAN_FOR_SEM(X)
#undef X
  return fail_arg(__func__, "unsupported (D,A); see han_semantic_shape_supported");
}

int han_semantic_combine(const float* Z, int64_t n, int P, int D, const float* beta_vec, float* out,
                         float* beta, han_stream_t stream) {
  HAN_REQUIRE(Z && beta_vec && out, "null pointer");
  HAN_REQUIRE(n > 0 && P > 0 && D % 4 == 0, "sizes");
  int64_t gb = ceil_div64(n * (D / 4), 256);
  unsigned grid = (unsigned)(gb < (int64_t)kNumSMs * 16 ? gb : (int64_t)kNumSMs * 16);
  semantic_combine_kernel<<<grid, 256, 0, as_stream(stream)>>>(Z, n, P, D, beta_vec, out, beta);
  return check_launch(__func__);
}

size_t han_semantic_bwd_workspace_bytes(int P, int D, int A) {
  (void)P;
  return (size_t)kSemBwdBlocks * ((size_t)D * A + 2 * A) * sizeof(float);
}

int han_semantic_bwd(const float* dout, const float* Z, const float* beta, const float* vsave,
                     int64_t n, int P, int D, int A, const float* w, const float* u, int mode,
                     const float* dsbar, float* dZ, float* dw, float* db, float* du, void* ws,
                     size_t ws_bytes, float* const* dz_tab, int64_t dz_stride, han_stream_t stream) {
  HAN_REQUIRE(dout && Z && beta && vsave && w && u && (dZ || dz_tab) && dw && db && du && ws, "null pointer");
  HAN_REQUIRE(!dz_tab || (dz_stride >= D && dz_stride % 4 == 0), "dz_stride");
  HAN_REQUIRE(n > 0 && P > 0 && P <= 64, "n > 0, 1 <= P <= 64");
  HAN_REQUIRE(mode == HAN_SEM_REFERENCE || dsbar, "paper mode needs dsbar");
#define X(d, a)         \
  if (D == d && A == a) \
    return launch_sem_bwd<d, a>(dout, Z, beta, vsave, n, P, w, u, mode, dsbar, dZ, dw, db, du, ws, ws_bytes, dz_tab, dz_stride, as_stream(stream));
  HAN_FOR_SEM(X)
#undef X
  return fail_arg(__func__, "unsupported (D,A); see han_semantic_shape_supported");
}

}  // extern "C"
