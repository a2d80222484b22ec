// K-A / K-E (FP32 CUDA-core mode): the X.W projection of utils/layers.py:20 for all K heads of G
// meta-paths at once, the f1/f2 attention dot-products of :23-24, and dW = X^T dS for the backward.
// mode 0 here is the exact-FP32 FFMA path (register-tiled, shared-memory double buffered); the
// tcgen05 tensor-core path lives in project_tc.cu and is selected with mode >= 1.
#include "han_common.cuh"

namespace han {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int kGemmThreads = 256;

// C[M x BN-tile] = A[M x Kd] * B[Kd x N]; A row-major lda, B row-major ldb, C row-major ldc.
// blockIdx.x = row tile, blockIdx.y = column tile.  Column tile `ct` is written at
// C + ct * c_tile_stride (lets one launch fill G separate node tables).
__global__ void __launch_bounds__(kGemmThreads)
sgemm_nn_kernel(const float* __restrict__ A, int64_t M, int64_t Kd, int64_t lda,
                const float* __restrict__ B, int64_t N, int64_t ldb, float* __restrict__ C, int64_t ldc,
                int64_t c_tile_stride) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;       // 16 col groups x 16 row groups; thread tile 8 x 4
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  float* Cout = C + (int64_t)blockIdx.y * c_tile_stride;

  // global->register staging: A tile 128x16 (8 per thread), B tile 16x64 (4 per thread)
  const int a_r = tid / 16, a_k = tid % 16;     // rows a_r + 16*i
  const int b_k = tid / 64, b_c = tid % 64;     // rows b_k + 4*i
  float ra[8], rb[4];
  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = m0 + a_r + 16 * i, k = k0 + a_k;
      ra[i] = (r < M && k < Kd) ? __ldg(A + r * lda + k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t k = k0 + b_k + 4 * i, c = n0 + b_c;
      rb[i] = (k < Kd && c < N) ? __ldg(B + k * ldb + c) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][a_k][a_r + 16 * i] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[buf][b_k + 4 * i][b_c] = rb[i];
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t nk = ceil_div64(Kd, BK);
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r < M) {
      const int64_t c = (int64_t)tx * 4;  // column inside this tile's output
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + c + j < N) Cout[r * ldc + c + j] = acc[i][j];
    }
  }
}

// f1 = S a1 + b1 -> R[:, D+k]   (utils/layers.py:23); thread = (row, head).  f2 = S a2 + b2 (:24) is NOT stored: the
// gather kernels recompute it from the table row they fetch anyway, which keeps the table rows at D floats.
template <int K, int H>
__global__ void __launch_bounds__(256)
attn_scores_kernel(const float* __restrict__ T, float* __restrict__ R, int64_t n, const float* __restrict__ a1,
                   const float* __restrict__ b1) {
  constexpr int D = K * H;
  constexpr int TS = D;
  constexpr int RS = ((D + 3 * K + 3) / 4) * 4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = idx / K;
  const int head = (int)(idx % K);
  if (row >= n) return;
  float s1 = b1[head];
#pragma unroll
  for (int q = 0; q < H / 4; ++q) {
    const float4 s = *reinterpret_cast<const float4*>(T + row * TS + head * H + 4 * q);
    const float4 x1 = ldg4(a1 + head * H + 4 * q);
    s1 += s.x * x1.x + s.y * x1.y + s.z * x1.z + s.w * x1.w;
  }
  R[row * RS + D + head] = s1;
}

// C_part[split][F x Dn] = A^T[F x rows] * G[rows x Dn] over this split's row range.
// A = X [n][lda]; G = dS_g [n][Dn]; blockIdx = (f tile, group g, split)
__global__ void __launch_bounds__(kGemmThreads)
sgemm_tn_splitk_kernel(const float* __restrict__ A, int64_t n, int64_t F, int64_t lda,
                       const float* __restrict__ G, int64_t Dn, int64_t g_stride, int64_t rows_per_split,
                       float* __restrict__ part, int64_t NC, int ctiles) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t f0 = (int64_t)blockIdx.x * BM;
  const int g = blockIdx.y / ctiles;
  const int64_t c0 = (int64_t)(blockIdx.y % ctiles) * BN;   // column tile inside group g
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const float* Gg = G + (int64_t)g * g_stride;

  const int a_k = tid / 128, a_f = tid % 128;   // A tile 16 x 128: rows a_k + 2*i
  const int b_k = tid / 64, b_c = tid % 64;     // B tile 16 x 64 : rows b_k + 4*i
  float ra[8], rb[4];
  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = k0 + a_k + 2 * i, f = f0 + a_f;
      ra[i] = (r < r_end && f < F) ? __ldg(A + r * lda + f) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = k0 + b_k + 4 * i;
      rb[i] = (r < r_end && c0 + b_c < Dn) ? __ldg(Gg + r * Dn + c0 + b_c) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][a_k + 2 * i][a_f] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[buf][b_k + 4 * i][b_c] = rb[i];
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t nk = ceil_div64(max((int64_t)0, r_end - r_begin), BK);
  if (nk > 0) {
    load_tiles(r_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < nk) load_tiles(r_begin + (kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }
  float* P = part + (int64_t)blockIdx.z * F * NC;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t f = f0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (f < F) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t c = c0 + (int64_t)tx * 4 + j;
        if (c < Dn) P[f * NC + (int64_t)g * Dn + c] = acc[i][j];
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t elems,
                                     float* __restrict__ outv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= elems) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(int64_t)k * elems + i];
  outv[i] = s;
}

static int pick_splits(int64_t n, int64_t F, int G, int D) {
  int64_t tiles = ceil_div64(F, BM) * G * ceil_div64(D, BN);
  int64_t want = (kNumSMs * 4 + tiles - 1) / tiles;
  int64_t maxs = ceil_div64(n, 4 * BK);
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}

}  // namespace han

using namespace han;

#define HAN_FOR_SHAPES(X) X(8, 8) X(4, 8) X(2, 8) X(1, 8) X(8, 4) X(4, 4) X(1, 4) X(8, 16) X(4, 16) X(1, 16) X(16, 4) X(16, 8)

extern "C" {

int han_project_fwd(const float* X, int64_t n, int64_t F, int64_t ldx, const float* W, int G, int K,
                    int H, const float* a1, const float* b1, float* T, float* R, int mode, han_stream_t stream) {
  HAN_REQUIRE(X && W && a1 && b1 && T && R, "null pointer");
  HAN_REQUIRE(n > 0 && F > 0 && G > 0 && ldx >= F, "sizes");
  HAN_REQUIRE(han_attn_shape_supported(K, H), "unsupported (K,H)");
  HAN_REQUIRE(mode == 0, "only mode 0 (fp32 FFMA) is built into this library version");
  const int D = K * H;
  const int TS = han_table_stride(K, H), RS = han_record_stride(K, H);
  cudaStream_t st = as_stream(stream);
  // one launch fills all G node tables when D is a whole number of 64-column tiles
  if (D % BN == 0) {
    // column tile ct belongs to group ct / (D/BN), offset (ct % (D/BN)) * BN inside the row
    if (D == BN) {
      dim3 grid((unsigned)ceil_div64(n, BM), (unsigned)G);
      sgemm_nn_kernel<<<grid, kGemmThreads, 0, st>>>(X, n, F, ldx, W, (int64_t)G * D, (int64_t)G * D, T, TS,
                                                    (int64_t)n * TS);
    } else {
      for (int g = 0; g < G; ++g) {
        dim3 grid((unsigned)ceil_div64(n, BM), (unsigned)(D / BN));
        sgemm_nn_kernel<<<grid, kGemmThreads, 0, st>>>(X, n, F, ldx, W + (int64_t)g * D, D, (int64_t)G * D,
                                                      T + (int64_t)g * n * TS, TS, BN);
      }
    }
  } else {
    for (int g = 0; g < G; ++g) {
      dim3 grid((unsigned)ceil_div64(n, BM), (unsigned)ceil_div64(D, BN));
      sgemm_nn_kernel<<<grid, kGemmThreads, 0, st>>>(X, n, F, ldx, W + (int64_t)g * D, D, (int64_t)G * D,
                                                    T + (int64_t)g * n * TS, TS, BN);
    }
  }
  int rc = check_launch(__func__);
  if (rc) return rc;
  for (int g = 0; g < G; ++g) {
    float* Tg = T + (int64_t)g * n * TS;
    float* Rg = R + (int64_t)g * n * RS;
    unsigned grid = (unsigned)ceil_div64(n * K, 256);
#define X_(k, h)                                                                                   \
  if (K == k && H == h)                                                                            \
    attn_scores_kernel<k, h><<<grid, 256, 0, st>>>(Tg, Rg, n, a1 + (int64_t)g * D, b1 + (int64_t)g * K);
    HAN_FOR_SHAPES(X_)
#undef X_
  }
  return check_launch(__func__);
}

size_t han_project_bwd_workspace_bytes(int64_t n, int64_t F, int G, int D) {
  return (size_t)pick_splits(n, F, G, D) * (size_t)F * G * D * sizeof(float);
}

int han_project_bwd(const float* X, int64_t n, int64_t F, int64_t ldx, const float* dS, int G, int D,
                    float* dW, void* ws, size_t ws_bytes, int mode, han_stream_t stream) {
  HAN_REQUIRE(X && dS && dW && ws, "null pointer");
  HAN_REQUIRE(n > 0 && F > 0 && G > 0 && D > 0 && ldx >= F, "sizes");
  HAN_REQUIRE(mode == 0, "only mode 0 (fp32 FFMA) is built into this library version");
  const int splits = pick_splits(n, F, G, D);
  HAN_REQUIRE(ws_bytes >= (size_t)splits * F * G * D * sizeof(float), "workspace too small");
  cudaStream_t st = as_stream(stream);
  int64_t rows_per_split = ceil_div64(ceil_div64(n, splits), BK) * BK;
  const int ctiles = (int)ceil_div64(D, BN);
  dim3 grid((unsigned)ceil_div64(F, BM), (unsigned)(G * ctiles), (unsigned)splits);
  float* part = reinterpret_cast<float*>(ws);
  sgemm_tn_splitk_kernel<<<grid, kGemmThreads, 0, st>>>(X, n, F, ldx, dS, D, (int64_t)n * D, rows_per_split,
                                                       part, (int64_t)G * D, ctiles);
  int64_t elems = F * (int64_t)G * D;
  splitk_reduce_kernel<<<(unsigned)ceil_div64(elems, 256), 256, 0, st>>>(part, splits, elems, dW);
  return check_launch(__func__);
}

}  // extern "C"
