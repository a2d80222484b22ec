// The classifier and the training loss of the step (models/gat.py:66-72 `tf.layers.dense(final_embed, nb_classes)`;
// models/base_gattn.py:41-48 `masked_softmax_cross_entropy`) as hand-written kernels, so that the timed step launches
// no stock-library GEMM / softmax kernels.  Exact-FP32 FFMA: these are skinny products (n x D by D x C with D = 64 and
// C = 3..349) that are bound by reading final_embed once and writing the logits once.
//
//   han_dense_fwd : Y = X W + b                                   one pass over X, W resident in shared memory
//   han_dense_bwd : dX = dY W^T ; dW = X^T dY ; db = sum dY       persistent CTAs, dW / db in registers across tiles,
//                                                                 deterministic two-stage reduce (han_reduce_partials)
//   han_masked_ce : loss = sum_i mask_i xent_i / sum mask ; dlogits = (softmax * sum(labels) - labels) mask_i / sum mask
//                   (tf.nn.softmax_cross_entropy_with_logits: labels need not be one-hot; no gradient to them)
#include "han_common.cuh"

namespace han {

constexpr int DN_ROWS = 64;        // rows per tile
constexpr int DN_THREADS = 256;
constexpr int DN_MAXC = 384;
constexpr int DN_MAXD = 64;

// ---- forward: thread = 4 rows x columns tx, tx+16, ... ------------------------------------------------------------
__global__ void __launch_bounds__(DN_THREADS)
dense_fwd_kernel(const float* __restrict__ X, int64_t n, int D, int64_t ldx, const float* __restrict__ W, int C,
                 const float* __restrict__ b, float* __restrict__ Y) {
  extern __shared__ __align__(16) float smem[];
  const int CP = C | 1;                         // odd leading dimension: column-strided reads are conflict-free
  float* Ws = smem;                             // [D][CP]
  float* Xs = Ws + (size_t)D * CP;              // [DN_ROWS][D + 1]
  float* bsm = Xs + (size_t)DN_ROWS * (D + 1);  // [C]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int i = tid; i < D * C; i += DN_THREADS) Ws[(i / C) * CP + (i % C)] = W[i];
  for (int i = tid; i < C; i += DN_THREADS) bsm[i] = b[i];
  const int64_t n_tiles = ceil_div64(n, DN_ROWS);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * DN_ROWS;
    __syncthreads();
    for (int i = tid; i < DN_ROWS * D; i += DN_THREADS) {
      const int r = i / D, d = i % D;
      Xs[r * (D + 1) + d] = (r0 + r < n) ? ldg_stream_f32(X + (r0 + r) * ldx + d) : 0.f;
    }
    __syncthreads();
    for (int cb = 0; cb < C; cb += 64) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int d = 0; d < D; ++d) {
        float xv[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = Xs[(4 * ty + i) * (D + 1) + d];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = cb + tx + 16 * j;
          wv[j] = (c < C) ? Ws[d * CP + c] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + 4 * ty + i;
        if (r < n) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = cb + tx + 16 * j;
            if (c < C) Y[r * C + c] = acc[i][j] + bsm[c];
          }
        }
      }
    }
  }
}

// ---- backward ---------------------------------------------------------------------------------------------------------
// Per tile: dX rows (thread = 4 rows x d = tx, tx+16, ...) and the dW / db contributions (thread = (d, c = cq + 4 j)).
template <int CT>   // CT = ceil(C / 4): dW columns per thread
__global__ void __launch_bounds__(DN_THREADS)
dense_bwd_kernel(const float* __restrict__ X, int64_t n, int D, int64_t ldx, const float* __restrict__ W, int C,
                 const float* __restrict__ dY, const float* __restrict__ scale, float* __restrict__ dX,
                 float* __restrict__ part) {
  extern __shared__ __align__(16) float smem[];
  const int CP = C | 1;
  float* Ws = smem;                               // [D][CP]
  float* Xs = Ws + (size_t)D * CP;                // [DN_ROWS][D + 1]
  float* Gs = Xs + (size_t)DN_ROWS * (D + 1);     // [DN_ROWS][CP]   dY tile (times the upstream scalar)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int i = tid; i < D * C; i += DN_THREADS) Ws[(i / C) * CP + (i % C)] = W[i];
  const float sc = scale ? *scale : 1.f;
  const int dd = tid >> 2, cq = tid & 3;          // dW: this thread's feature d (D <= 64) and column residue
  float dw[CT];
#pragma unroll
  for (int j = 0; j < CT; ++j) dw[j] = 0.f;
  float dbv = 0.f, dbv2 = 0.f;                    // db: thread t sums column t (and t + 256)
  const int64_t n_tiles = ceil_div64(n, DN_ROWS);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * DN_ROWS;
    __syncthreads();
    for (int i = tid; i < DN_ROWS * D; i += DN_THREADS) {
      const int r = i / D, d = i % D;
      Xs[r * (D + 1) + d] = (r0 + r < n) ? ldg_stream_f32(X + (r0 + r) * ldx + d) : 0.f;
    }
    for (int i = tid; i < DN_ROWS * C; i += DN_THREADS) {
      const int r = i / C, c = i % C;
      Gs[r * CP + c] = (r0 + r < n) ? sc * ldg_stream_f32(dY + (r0 + r) * C + c) : 0.f;
    }
    __syncthreads();
    // dX = dY W^T
    if (dX != nullptr) {
      for (int db0 = 0; db0 < D; db0 += 64) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int c = 0; c < C; ++c) {
          float gv[4], wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) gv[i] = Gs[(4 * ty + i) * CP + c];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = db0 + tx + 16 * j;
            wv[j] = (d < D) ? Ws[d * CP + c] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], wv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t r = r0 + 4 * ty + i;
          if (r < n) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int d = db0 + tx + 16 * j;
              if (d < D) dX[r * D + d] = acc[i][j];
            }
          }
        }
      }
    }
    // dW += X^T dY, db += column sums
    if (dd < D) {
      for (int r = 0; r < DN_ROWS; ++r) {
        const float xv = Xs[r * (D + 1) + dd];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          const int c = cq + 4 * j;
          if (c < C) dw[j] = fmaf(xv, Gs[r * CP + c], dw[j]);
        }
      }
    }
    if (tid < C) {
      for (int r = 0; r < DN_ROWS; ++r) dbv += Gs[r * CP + tid];
    }
    if (tid + DN_THREADS < C) {
      for (int r = 0; r < DN_ROWS; ++r) dbv2 += Gs[r * CP + tid + DN_THREADS];
    }
  }
  // per-CTA partials: [dW (D*C) | db (C)]
  float* my = part + (size_t)blockIdx.x * ((size_t)D * C + C);
  if (dd < D) {
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      const int c = cq + 4 * j;
      if (c < C) my[(size_t)dd * C + c] = dw[j];
    }
  }
  if (tid < C) my[(size_t)D * C + tid] = dbv;
  if (tid + DN_THREADS < C) my[(size_t)D * C + tid + DN_THREADS] = dbv2;
}

// ---- few classes (C <= 8, the node-classification case: 3 for ACM/IMDB, 4 for DBLP) ---------------------------------------
// The tiled kernels above spend 13/16 of their threads on padding when C = 3 and stage X through scalar shared-memory
// stores (0.8 ms for a 0.5 GB read on the 2M-node config).  Here LPR = D/4 lanes own one row: each lane loads ONE
// float4 of it (a fully coalesced 16 B x 32 lanes request), keeps its 4 x C slice of W in registers, and a butterfly
// over the LPR lanes finishes the C dot products.  Persistent grid; a warp handles 32 / LPR rows per iteration.
template <int LPR, int CMAX>
__global__ void __launch_bounds__(256)
dense_fwd_small_kernel(const float* __restrict__ X, int64_t n, int64_t ldx, const float* __restrict__ W, int C,
                       const float* __restrict__ b, float* __restrict__ Y) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
  float w[4][CMAX];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < CMAX; ++c) w[i][c] = (c < C) ? __ldg(W + (size_t)(4 * sub + i) * C + c) : 0.f;
  const float bias = (sub < C) ? __ldg(b + sub) : 0.f;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = warp0 * RPW; r0 < n; r0 += n_warps * RPW) {
    const int64_t r = r0 + rsel;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n) x = ldg4_stream(X + r * ldx + 4 * sub);
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = fmaf(x.x, w[0][c], fmaf(x.y, w[1][c], fmaf(x.z, w[2][c], x.w * w[3][c])));
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int c = 0; c < CMAX; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    float y = 0.f;            // lane `sub` of the row keeps column `sub` (no dynamic register indexing)
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (sub == c) y = acc[c];
    if (r < n && sub < C) Y[r * C + sub] = y + bias;
  }
}

// backward of the same: dX row = dY row W^T (each lane its 4 features), dW / db accumulated in registers over the
// whole grid-stride loop, then one deterministic CTA reduction into this CTA's partial [dW (D*C) | db (C)].
template <int LPR, int CMAX>
__global__ void __launch_bounds__(256)
dense_bwd_small_kernel(const float* __restrict__ X, int64_t n, int64_t ldx, const float* __restrict__ W, int C,
                       const float* __restrict__ dY, const float* __restrict__ scale, float* __restrict__ dX,
                       float* __restrict__ part) {
  constexpr int RPW = 32 / LPR;
  constexpr int D = 4 * LPR;
  __shared__ float red[8][RPW][D + 1][CMAX];      // 8 warps x row groups: <= 8 * 2 * 65 * 8 * 4 B = 33 KB
  __shared__ float redb[8][RPW][CMAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LPR, rsel = lane / LPR;
  float w[4][CMAX];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < CMAX; ++c) w[i][c] = (c < C) ? __ldg(W + (size_t)(4 * sub + i) * C + c) : 0.f;
  const float sc = scale ? *scale : 1.f;
  float dw[4][CMAX], dbv[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    dbv[c] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) dw[i][c] = 0.f;
  }
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = warp0 * RPW; r0 < n; r0 += n_warps * RPW) {
    const int64_t r = r0 + rsel;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    float g[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) g[c] = 0.f;
    if (r < n) {
      x = ldg4_stream(X + r * ldx + 4 * sub);
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) g[c] = sc * __ldg(dY + r * C + c);      // the row's C gradients: a broadcast within its lanes
    }
    if (dX != nullptr && r < n) {
      float4 o;
      o.x = o.y = o.z = o.w = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        o.x = fmaf(g[c], w[0][c], o.x);
        o.y = fmaf(g[c], w[1][c], o.y);
        o.z = fmaf(g[c], w[2][c], o.z);
        o.w = fmaf(g[c], w[3][c], o.w);
      }
      *reinterpret_cast<float4*>(dX + r * D + 4 * sub) = o;
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      dw[0][c] = fmaf(x.x, g[c], dw[0][c]);
      dw[1][c] = fmaf(x.y, g[c], dw[1][c]);
      dw[2][c] = fmaf(x.z, g[c], dw[2][c]);
      dw[3][c] = fmaf(x.w, g[c], dw[3][c]);
      dbv[c] += g[c];
    }
  }
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red[warp][rsel][4 * sub + i][c] = dw[i][c];
    if (sub == 0) redb[warp][rsel][c] = dbv[c];
  }
  __syncthreads();
  float* my = part + (size_t)blockIdx.x * ((size_t)D * C + C);
  for (int i = threadIdx.x; i < D * C + C; i += blockDim.x) {
    float t = 0.f;
    if (i < D * C) {
      const int d = i / C, c = i % C;
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int q = 0; q < RPW; ++q) t += red[k][q][d][c];
    } else {
      const int c = i - D * C;
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int q = 0; q < RPW; ++q) t += redb[k][q][c];
    }
    my[i] = t;
  }
}

// ---- masked softmax cross-entropy: warp per row --------------------------------------------------------------------
__global__ void __launch_bounds__(256)
masked_ce_kernel(const float* __restrict__ logits, const float* __restrict__ labels, const float* __restrict__ mask,
                 const float* __restrict__ mask_total, int64_t n, int C, float* __restrict__ loss_part,
                 float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_total = 1.f / *mask_total;
  float loss = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n; row += (int64_t)gridDim.x * 8) {
    const float* lg = logits + row * C;
    const float* lb = labels + row * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f, ls = 0.f, dot = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float l = lg[c], y = lb[c];
      se += expf(l - mx);
      ls += y;
      dot = fmaf(y, l, dot);
    }
    se = warp_sum(se);
    ls = warp_sum(ls);
    dot = warp_sum(dot);
    const float lse = mx + logf(se);
    const float wgt = mask[row] * inv_total;
    loss += (lse * ls - dot) * wgt;                       // -sum_c y_c log_softmax_c
    if (dlogits != nullptr) {
      const float rinv = 1.f / se;
      for (int c = lane; c < C; c += 32) dlogits[row * C + c] = (expf(lg[c] - mx) * rinv * ls - lb[c]) * wgt;
    }
  }
  // lanes hold identical sums; one value per warp, 8 warps per CTA summed in a fixed order
  __shared__ float ws[8];
  if (lane == 0) ws[warp] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += ws[k];
    loss_part[blockIdx.x] = s;
  }
}

// small C (the usual 3..8 classes): one THREAD per row -- a row is C contiguous floats, thousands of rows are in flight
// per SM, nothing to shuffle.  (The warp-per-row kernel above serialises ~850 dependent row iterations per warp on the
// 2M-node config: 1.6 ms for 190 MB of traffic.)
template <int CMAX>
__global__ void __launch_bounds__(256)
masked_ce_rows_kernel(const float* __restrict__ logits, const float* __restrict__ labels, const float* __restrict__ mask,
                      const float* __restrict__ mask_total, int64_t n, int C, float* __restrict__ loss_part,
                      float* __restrict__ dlogits) {
  const float inv_total = 1.f / *mask_total;
  float loss = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x) {
    float lg[CMAX], lb[CMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        lg[c] = ldg_stream_f32(logits + row * C + c);
        lb[c] = ldg_stream_f32(labels + row * C + c);
        mx = fmaxf(mx, lg[c]);
      }
    }
    float se = 0.f, ls = 0.f, dot = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        lg[c] = expf(lg[c] - mx);          // keep the exponential (reused by dlogits)
        se += lg[c];
        ls += lb[c];
      }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) dot = fmaf(lb[c], __ldg(logits + row * C + c) - mx, dot);      // raw logits again (L1 hit): exact
    const float wgt = mask[row] * inv_total;
    loss += (logf(se) * ls - dot) * wgt;                   // -sum_c y_c log_softmax_c  (mx cancels)
    if (dlogits != nullptr) {
      const float rinv = 1.f / se;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) dlogits[row * C + c] = (lg[c] * rinv * ls - lb[c]) * wgt;
    }
  }
  // deterministic block sum
  __shared__ float ws[8];
  loss = warp_sum(loss);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += ws[k];
    loss_part[blockIdx.x] = t;
  }
}

static size_t dense_fwd_smem(int D, int C) { return ((size_t)D * (C | 1) + (size_t)DN_ROWS * (D + 1) + C) * sizeof(float); }
static size_t dense_bwd_smem(int D, int C) {
  return ((size_t)D * (C | 1) + (size_t)DN_ROWS * (D + 1) + (size_t)DN_ROWS * (C | 1)) * sizeof(float);
}

}  // namespace han

using namespace han;

extern "C" {

// persistent grid of the dense / CE kernels: 4 CTAs per SM so that (with the small tiles of C <= 32) several CTAs per SM
// overlap one another's tile loads; also the number of per-CTA partials the backward and the loss write
int han_dense_blocks(void) { return kNumSMs * 4; }

int han_dense_fwd(const float* X, int64_t n, int D, int64_t ldx, const float* W, int C, const float* b, float* Y,
                  han_stream_t stream) {
  HAN_REQUIRE(X && W && b && Y, "null pointer");
  HAN_REQUIRE(n > 0 && D > 0 && D <= DN_MAXD && C > 0 && C <= DN_MAXC && ldx >= D, "sizes: D <= 64, C <= 384");
  if (C <= 8 && (D == 64 || D == 32) && ldx % 4 == 0 && (uintptr_t)X % 16 == 0) {
    const unsigned g = (unsigned)han_dense_blocks();
    if (D == 64) dense_fwd_small_kernel<16, 8><<<g, 256, 0, as_stream(stream)>>>(X, n, ldx, W, C, b, Y);
    else dense_fwd_small_kernel<8, 8><<<g, 256, 0, as_stream(stream)>>>(X, n, ldx, W, C, b, Y);
    return check_launch(__func__);
  }
  const size_t smem = dense_fwd_smem(D, C);
  HAN_SMEM_ATTR_ONCE(dense_fwd_kernel, dense_fwd_smem(DN_MAXD, DN_MAXC));
  const int64_t tiles = ceil_div64(n, DN_ROWS);
  const int64_t cap = (int64_t)kNumSMs * 8;
  const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
  dense_fwd_kernel<<<grid, DN_THREADS, smem, as_stream(stream)>>>(X, n, D, ldx, W, C, b, Y);
  return check_launch(__func__);
}

/* part: [han_dense_blocks()][D*C + C] per-CTA partials of dW | db (reduce with han_reduce_partials).
 * scale (nullable): device scalar multiplied into dY (the upstream gradient of a scalar loss). dX nullable. */
int han_dense_bwd(const float* X, int64_t n, int D, int64_t ldx, const float* W, int C, const float* dY,
                  const float* scale, float* dX, float* part, han_stream_t stream) {
  HAN_REQUIRE(X && W && dY && part, "null pointer");
  HAN_REQUIRE(n > 0 && D > 0 && D <= DN_MAXD && C > 0 && C <= DN_MAXC && ldx >= D, "sizes: D <= 64, C <= 384");
  const size_t smem = dense_bwd_smem(D, C);
  cudaStream_t st = as_stream(stream);
  const unsigned grid = (unsigned)han_dense_blocks();     // every CTA writes its partial (zeros if it has no tile)
  if (C <= 8 && (D == 64 || D == 32) && ldx % 4 == 0 && (uintptr_t)X % 16 == 0 && (!dX || (uintptr_t)dX % 16 == 0)) {
    if (D == 64) dense_bwd_small_kernel<16, 8><<<grid, 256, 0, st>>>(X, n, ldx, W, C, dY, scale, dX, part);
    else dense_bwd_small_kernel<8, 8><<<grid, 256, 0, st>>>(X, n, ldx, W, C, dY, scale, dX, part);
    return check_launch(__func__);
  }
  const int ct = (C + 3) / 4;
#define LAUNCH(CT)                                                                                             \
  {                                                                                                            \
    HAN_SMEM_ATTR_ONCE(dense_bwd_kernel<CT>, dense_bwd_smem(DN_MAXD, CT * 4 > DN_MAXC ? DN_MAXC : CT * 4));    \
    dense_bwd_kernel<CT><<<grid, DN_THREADS, smem, st>>>(X, n, D, ldx, W, C, dY, scale, dX, part);             \
  }
  if (ct <= 4) LAUNCH(4)
  else if (ct <= 16) LAUNCH(16)
  else if (ct <= 48) LAUNCH(48)
  else LAUNCH(96)
#undef LAUNCH
  return check_launch(__func__);
}

/* loss_part: [han_dense_blocks()] per-CTA partial sums of the loss; dlogits nullable (evaluation). */
int han_masked_ce(const float* logits, const float* labels, const float* mask, const float* mask_total, int64_t n,
                  int C, float* loss_part, float* dlogits, han_stream_t stream) {
  HAN_REQUIRE(logits && labels && mask && mask_total && loss_part, "null pointer");
  HAN_REQUIRE(n > 0 && C > 0, "sizes");
  if (C <= 8)
    masked_ce_rows_kernel<8><<<(unsigned)han_dense_blocks(), 256, 0, as_stream(stream)>>>(logits, labels, mask, mask_total, n,
                                                                                         C, loss_part, dlogits);
  else if (C <= 32)
    masked_ce_rows_kernel<32><<<(unsigned)han_dense_blocks(), 256, 0, as_stream(stream)>>>(logits, labels, mask, mask_total, n,
                                                                                          C, loss_part, dlogits);
  else
    masked_ce_kernel<<<(unsigned)han_dense_blocks(), 256, 0, as_stream(stream)>>>(logits, labels, mask, mask_total, n, C,
                                                                                 loss_part, dlogits);
  return check_launch(__func__);
}

}  // extern "C"
