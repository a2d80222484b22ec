// Training-mode projection with the reference's feed-forward dropout (utils/layers.py:18-19 and :31-32):
//   * every head k of every meta-path g sees its OWN dropped copy of the input,  S_k = (X * m_k / keep) W_k
//     (each attn_head call draws a fresh mask), so the shared-A tensor-core GEMM does not apply; this is
//     the exact-FP32 FFMA kernel with the per-head mask bits generated while the X tile is staged;
//   * f1, f2 are taken from the UN-dropped S (:23-24), then S itself is dropped for the aggregation (:31-32);
//     the un-dropped S is kept (S_keep) because da1/da2 need it in the backward.
// Masks are pure functions of (seed, meta-path, head, node, feature) -- han_rng.cuh -- so the dW kernel
// below regenerates exactly the masks of the forward.
#include "han_common.cuh"
#include "han_rng.cuh"

namespace han {

constexpr int DBM = 128, DBN = 64, DBK = 16;
constexpr int kDropThreads = 256;

struct DropIn {
  const uint32_t* seed_ptr;
  uint32_t thr;        // keep * 2^24
  float inv_keep;
  uint32_t metapath;
  int64_t row0;        // global id of local row 0
};

// bit k of the result: head k keeps input element (node, f)
__device__ __forceinline__ uint32_t head_mask_bits(uint32_t base, int K, uint32_t thr) {
  uint32_t bits = 0;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    const uint32_t h = (base ^ (0x632BE5ABu * (uint32_t)(k + 1))) * 0x9E3779B1u;
    bits |= ((h >> 8) < thr ? 1u : 0u) << k;
  }
  return bits;
}

// C[M x 64-col tile] = sum_f (A[m][f] * mask_head(m,f)/keep) * B[f][c]  for ONE meta-path (N = D columns)
__global__ void __launch_bounds__(kDropThreads)
sgemm_nn_drop_kernel(const float* __restrict__ A, int64_t M, int64_t Kd, int64_t lda, const float* __restrict__ B,
                     int64_t N, int64_t ldb, float* __restrict__ C, int64_t ldc, int K, int H, DropIn dr) {
  __shared__ __align__(16) float As[2][DBK][DBM + 4];
  __shared__ __align__(16) float Bs[2][DBK][DBN + 4];
  __shared__ __align__(4) uint8_t Ms[2][DBK][DBM + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.x * DBM;
  const int64_t n0 = (int64_t)blockIdx.y * DBN;
  const uint32_t sseed = stream_seed(*dr.seed_ptr, 1u, dr.metapath, 0u);
  const int head = (int)(((n0 + tx * 4) / H) % K);

  const int a_r = tid / 16, a_k = tid % 16;
  const int b_k = tid / 64, b_c = tid % 64;
  float ra[8], rb[4];
  uint32_t rm[8];
  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = m0 + a_r + 16 * i, k = k0 + a_k;
      const bool ok = r < M && k < Kd;
      ra[i] = ok ? __ldg(A + r * lda + k) : 0.f;
      rm[i] = ok ? head_mask_bits(mix3(sseed, (uint32_t)(r + dr.row0), (uint32_t)k), K, dr.thr) : 0u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t k = k0 + b_k + 4 * i, c = n0 + b_c;
      rb[i] = (k < Kd && c < N) ? __ldg(B + k * ldb + c) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      As[buf][a_k][a_r + 16 * i] = ra[i];
      Ms[buf][a_k][a_r + 16 * i] = (uint8_t)rm[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[buf][b_k + 4 * i][b_c] = rb[i];
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t nk = ceil_div64(Kd, DBK);
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < nk) load_tiles((kt + 1) * DBK);
#pragma unroll
    for (int k = 0; k < DBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const uint32_t mk0 = *reinterpret_cast<const uint32_t*>(&Ms[buf][k][ty * 4]);
      const uint32_t mk1 = *reinterpret_cast<const uint32_t*>(&Ms[buf][k][64 + ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t byte = ((i < 4 ? mk0 : mk1) >> (8 * (i & 3))) & 0xFFu;
        av[i] = ((byte >> head) & 1u) ? av[i] * dr.inv_keep : 0.f;
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
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + tx * 4 + j < N) C[r * ldc + n0 + tx * 4 + j] = acc[i][j];
    }
  }
}

// f1 from the un-dropped S (layers.py:23).  The table keeps the UN-dropped S: f2 (:24) is recomputed from it by the
// gather kernels, which also apply the feature mask of :31-32 to the rows they fetch (same counter-based bits).
template <int K, int H>
__global__ void __launch_bounds__(256)
scores_drop_kernel(const float* __restrict__ T, float* __restrict__ R, int64_t n, const float* __restrict__ a1,
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

// dW partial[split][F x D] = (X * mask_head / keep)^T dS_g over this split's rows, ONE meta-path
__global__ void __launch_bounds__(kDropThreads)
sgemm_tn_drop_kernel(const float* __restrict__ A, int64_t n, int64_t F, int64_t lda, const float* __restrict__ G,
                     int64_t Dn, int64_t rows_per_split, float* __restrict__ part, int K, int H, DropIn dr) {
  __shared__ __align__(16) float As[2][DBK][DBM + 4];
  __shared__ __align__(16) float Bs[2][DBK][DBN + 4];
  __shared__ __align__(4) uint8_t Ms[2][DBK][DBM + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t f0 = (int64_t)blockIdx.x * DBM;
  const int64_t c0 = (int64_t)blockIdx.y * DBN;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const uint32_t sseed = stream_seed(*dr.seed_ptr, 1u, dr.metapath, 0u);
  const int head = (int)(((c0 + tx * 4) / H) % K);

  const int a_k = tid / 128, a_f = tid % 128;
  const int b_k = tid / 64, b_c = tid % 64;
  float ra[8], rb[4];
  uint32_t rm[8];
  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t r = k0 + a_k + 2 * i, f = f0 + a_f;
      const bool ok = r < r_end && f < F;
      ra[i] = ok ? __ldg(A + r * lda + f) : 0.f;
      rm[i] = ok ? head_mask_bits(mix3(sseed, (uint32_t)(r + dr.row0), (uint32_t)f), K, dr.thr) : 0u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = k0 + b_k + 4 * i;
      rb[i] = (r < r_end && c0 + b_c < Dn) ? __ldg(G + r * Dn + c0 + b_c) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      As[buf][a_k + 2 * i][a_f] = ra[i];
      Ms[buf][a_k + 2 * i][a_f] = (uint8_t)rm[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[buf][b_k + 4 * i][b_c] = rb[i];
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int64_t nk = ceil_div64(max((int64_t)0, r_end - r_begin), DBK);
  if (nk > 0) {
    load_tiles(r_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < nk) load_tiles(r_begin + (kt + 1) * DBK);
#pragma unroll
    for (int k = 0; k < DBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const uint32_t mk0 = *reinterpret_cast<const uint32_t*>(&Ms[buf][k][ty * 4]);
      const uint32_t mk1 = *reinterpret_cast<const uint32_t*>(&Ms[buf][k][64 + ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t byte = ((i < 4 ? mk0 : mk1) >> (8 * (i & 3))) & 0xFFu;
        av[i] = ((byte >> head) & 1u) ? av[i] * dr.inv_keep : 0.f;
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }
  float* P = part + (int64_t)blockIdx.z * F * Dn;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t f = f0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (f < F) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t c = c0 + (int64_t)tx * 4 + j;
        if (c < Dn) P[f * Dn + c] = acc[i][j];
      }
    }
  }
}

// dW[:, col0 : col0 + Dn] = sum over splits
__global__ void drop_reduce_kernel(const float* __restrict__ part, int splits, int64_t F, int64_t Dn,
                                   float* __restrict__ dW, int64_t ldw, int64_t col0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * Dn) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(int64_t)k * F * Dn + i];
  dW[(i / Dn) * ldw + col0 + (i % Dn)] = s;
}

static int drop_splits(int64_t n, int64_t F, int D) {
  int64_t tiles = ceil_div64(F, DBM) * ceil_div64(D, DBN);
  int64_t want = (kNumSMs * 4 + tiles - 1) / tiles;
  int64_t maxs = ceil_div64(n, 4 * DBK);
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}

static DropIn make_dropin(const uint32_t* seed_ptr, float keep, int metapath, int64_t row0) {
  DropIn d;
  d.seed_ptr = seed_ptr;
  d.thr = (uint32_t)(keep * 16777216.f + 0.5f);
  d.inv_keep = 1.f / ((float)d.thr / 16777216.f);
  d.metapath = (uint32_t)metapath;
  d.row0 = row0;
  return d;
}

}  // namespace han

using namespace han;

#define HAN_FOR_SHAPES(X) X(8, 8) X(4, 8) X(2, 8) X(1, 8) X(8, 4) X(4, 4) X(1, 4) X(8, 16) X(4, 16) X(1, 16)

extern "C" {

int han_project_fwd_drop(const float* X, int64_t n, int64_t F, int64_t ldx, const float* W, int64_t ldw, int G, int K,
                         int H,
                         const float* a1, const float* b1, float* T, float* R,
                         const uint32_t* seed_ptr, float in_keep, int metapath0, int64_t row0,
                         han_stream_t stream) {
  HAN_REQUIRE(X && W && a1 && b1 && T && R && seed_ptr, "null pointer");
  HAN_REQUIRE(n > 0 && F > 0 && G > 0 && ldx >= F && ldw >= (int64_t)G * K * H, "sizes");
  HAN_REQUIRE(in_keep > 0.f && in_keep < 1.f, "in_keep in (0,1): use han_project_fwd when dropout is off");
  HAN_REQUIRE(K <= 8 && han_attn_shape_supported(K, H), "dropout projection: K <= 8 and a supported (K,H)");
  const int D = K * H;
  const int TS = han_table_stride(K, H), RS = han_record_stride(K, H);
  cudaStream_t st = as_stream(stream);
  for (int g = 0; g < G; ++g) {
    DropIn dr = make_dropin(seed_ptr, in_keep, metapath0 + g, row0);
    float* Tg = T + (int64_t)g * n * TS;
    float* Rg = R + (int64_t)g * n * RS;
    dim3 grid((unsigned)ceil_div64(n, DBM), (unsigned)ceil_div64(D, DBN));
    sgemm_nn_drop_kernel<<<grid, kDropThreads, 0, st>>>(X, n, F, ldx, W + (int64_t)g * D, D, ldw, Tg, TS, K, H, dr);
    unsigned sgrid = (unsigned)ceil_div64(n * K, 256);
#define X_(k, h)                                                                                             \
  if (K == k && H == h)                                                                                      \
    scores_drop_kernel<k, h><<<sgrid, 256, 0, st>>>(Tg, Rg, n, a1 + (int64_t)g * D, b1 + (int64_t)g * K);
    HAN_FOR_SHAPES(X_)
#undef X_
  }
  return check_launch(__func__);
}

size_t han_project_bwd_drop_workspace_bytes(int64_t n, int64_t F, int D) {
  return (size_t)drop_splits(n, F, D) * (size_t)F * D * sizeof(float);
}

int han_project_bwd_drop(const float* X, int64_t n, int64_t F, int64_t ldx, const float* dS, int G, int K, int H,
                         float* dW, int64_t ldw, void* ws, size_t ws_bytes, const uint32_t* seed_ptr, float in_keep,
                         int metapath0, int64_t row0, han_stream_t stream) {
  HAN_REQUIRE(X && dS && dW && ws && seed_ptr, "null pointer");
  HAN_REQUIRE(n > 0 && F > 0 && G > 0 && ldx >= F && K <= 8, "sizes");
  HAN_REQUIRE(in_keep > 0.f && in_keep < 1.f, "in_keep in (0,1)");
  const int D = K * H;
  const int splits = drop_splits(n, F, D);
  HAN_REQUIRE(ws_bytes >= (size_t)splits * F * D * sizeof(float), "workspace too small");
  cudaStream_t st = as_stream(stream);
  const int64_t rows_per_split = ceil_div64(ceil_div64(n, splits), DBK) * DBK;
  float* part = reinterpret_cast<float*>(ws);
  for (int g = 0; g < G; ++g) {
    DropIn dr = make_dropin(seed_ptr, in_keep, metapath0 + g, row0);
    dim3 grid((unsigned)ceil_div64(F, DBM), (unsigned)ceil_div64(D, DBN), (unsigned)splits);
    sgemm_tn_drop_kernel<<<grid, kDropThreads, 0, st>>>(X, n, F, ldx, dS + (int64_t)g * n * D, D, rows_per_split, part,
                                                       K, H, dr);
    drop_reduce_kernel<<<(unsigned)ceil_div64(F * (int64_t)D, 256), 256, 0, st>>>(part, splits, F, D, dW, ldw,
                                                                                 (int64_t)g * D);
  }
  return check_launch(__func__);
}

}  // extern "C"
