// K-0: device-side graph builder.  Replaces utils/process.py:14-25 (adj_to_bias): instead of a
// dense fp64 N x N bias matrix that is re-fed every step, the mask is turned into CSR once, on the
// device, bit-exact against np.nonzero(bias == 0); plus the transposed structure the backward needs.
#include "han_common.cuh"

namespace han {

thread_local char g_last_error[512] = {0};

// ---------------------------------------------------------------------------------------------
// dense -> mask predicate
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool is_edge(T v, bool diag, int kind, bool* bad) {
  if (kind == HAN_DENSE_ADJ) {
    // mt = I @ (adj + I) is exactly adj + I in fp64 (process.py:18-20, nhood=1); edge iff > 0 (:23)
    return ((double)v + (diag ? 1.0 : 0.0)) > 0.0;
  }
  if (kind == HAN_DENSE_POSITIVE) return v > (T)0;  // an already-formed mt (process.py:23)
  // bias == 0 <=> mt == 1 (process.py:25); anything else must be a mask value (<= -1e8)
  bool e = (v == (T)0);
  *bad = !e && !(v <= (T)-1e8);
  return e;
}

template <typename T>
__global__ void dense_row_counts_kernel(const T* __restrict__ dense, int kind, int64_t n, int64_t ld,
                                        int32_t* __restrict__ row_counts, int32_t* bad_count) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int nbad = 0;
  for (int64_t row = warp; row < n; row += nwarps) {
    const T* r = dense + row * ld;
    int cnt = 0;
    for (int64_t c0 = 0; c0 < n; c0 += 32) {
      int64_t c = c0 + lane;
      bool bad = false;
      bool e = (c < n) && is_edge<T>(r[c], c == row, kind, &bad);
      nbad += (c < n && bad) ? 1 : 0;
      cnt += __popc(__ballot_sync(0xffffffffu, e));
    }
    if (lane == 0) row_counts[row] = cnt;
  }
  if (bad_count != nullptr && nbad) atomicAdd(bad_count, nbad);
}

template <typename T>
__global__ void dense_fill_kernel(const T* __restrict__ dense, int kind, int64_t n, int64_t ld,
                                  const int64_t* __restrict__ indptr, int32_t* __restrict__ indices) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp; row < n; row += nwarps) {
    const T* r = dense + row * ld;
    int64_t pos = indptr[row];
    for (int64_t c0 = 0; c0 < n; c0 += 32) {
      int64_t c = c0 + lane;
      bool bad;
      bool e = (c < n) && is_edge<T>(r[c], c == row, kind, &bad);
      unsigned b = __ballot_sync(0xffffffffu, e);
      if (e) indices[pos + __popc(b & ((1u << lane) - 1u))] = (int32_t)c;
      pos += __popc(b);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan int32 -> int64, three stream-ordered kernels (no host sync)
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* total) {
  __shared__ int64_t warp_tot[kScanThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  int64_t off = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < kScanThreads / 32; ++i) {
    if (i < w) off += warp_tot[i];
    tot += warp_tot[i];
  }
  __syncthreads();
  *total = tot;
  return off + inc - v;
}

__global__ void scan_chunk_sums_kernel(const int32_t* __restrict__ counts, int64_t n,
                                       int64_t* __restrict__ chunk_sums) {
  int64_t base = (int64_t)blockIdx.x * kScanChunk;
  int64_t s = 0;
  for (int i = threadIdx.x; i < kScanChunk; i += kScanThreads) {
    int64_t idx = base + i;
    if (idx < n) s += counts[idx];
  }
  int64_t tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) chunk_sums[blockIdx.x] = tot;
}

__global__ void scan_chunk_offsets_kernel(int64_t* chunk_sums, int64_t nchunks) {
  // single block; sequential over tiles of kScanThreads
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t b = 0; b < nchunks; b += kScanThreads) {
    int64_t i = b + threadIdx.x;
    int64_t v = (i < nchunks) ? chunk_sums[i] : 0;
    int64_t tot;
    int64_t ex = block_exclusive_scan(v, &tot);
    int64_t carry = carry_s;
    if (i < nchunks) chunk_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
}

__global__ void scan_apply_kernel(const int32_t* __restrict__ counts, int64_t n,
                                  const int64_t* __restrict__ chunk_offsets,
                                  int64_t* __restrict__ indptr) {
  int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    v[i] = (idx < n) ? counts[idx] : 0;
    s += v[i];
  }
  int64_t tot;
  int64_t ex = block_exclusive_scan(s, &tot) + chunk_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    if (idx < n) indptr[idx] = ex;
    ex += v[i];
    if (idx == n - 1) indptr[n] = ex;
  }
}

// ---------------------------------------------------------------------------------------------
// transpose: count -> scan -> atomic fill -> per-segment sort (restores a deterministic order)
// ---------------------------------------------------------------------------------------------
__global__ void col_count_kernel(const int32_t* __restrict__ indices, int64_t nnz,
                                 int32_t* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) atomicAdd(&counts[indices[i]], 1);
}

// One pass places the edges whose column lies in [c_lo, c_hi).  The host runs the passes over column blocks whose slice
// of t_indices (and perm) fits in L2: the scattered 4-byte stores of a pass then complete their lines in L2 and leave
// as full-line write-backs, instead of 100 M partial-sector evictions over a 400 MB array (2M-node graph: the fill was
// 8.3 of the transposition's 12.7 ms).  Re-reading the column array once per pass is a coalesced 400 MB stream.
__global__ void transpose_fill_kernel(int64_t n_rows, const int64_t* __restrict__ indptr,
                                      const int32_t* __restrict__ indices,
                                      const int64_t* __restrict__ t_indptr, int32_t* cursor,
                                      int32_t* __restrict__ t_indices, int32_t* __restrict__ perm, int32_t c_lo,
                                      int32_t c_hi) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp; row < n_rows; row += nwarps) {
    int64_t s = indptr[row], e = indptr[row + 1];
    for (int64_t k = s + lane; k < e; k += 32) {
      int32_t c = indices[k];
      if (c < c_lo || c >= c_hi) continue;
      int64_t pos = t_indptr[c] + atomicAdd(&cursor[c], 1);
      t_indices[pos] = (int32_t)row;
      if (perm != nullptr) perm[pos] = (int32_t)k;
    }
  }
}

// all-ascending ("flip") bitonic network: valid for any length L with virtual +inf padding
template <bool HAS_VAL>
__device__ __forceinline__ void cmpx(int32_t* keys, int32_t* vals, int64_t i, int64_t l) {
  int32_t a = keys[i], b = keys[l];
  if (a > b) {
    keys[i] = b;
    keys[l] = a;
    if (HAS_VAL) {
      int32_t t = vals[i];
      vals[i] = vals[l];
      vals[l] = t;
    }
  }
}

// sorts keys[0..L) (with vals) using `nthreads` cooperating threads whose id is `tid`;
// `sync` is __syncwarp or __syncthreads via the template flag.
template <bool HAS_VAL, bool BLOCK>
__device__ __forceinline__ void bitonic_sort(int32_t* keys, int32_t* vals, int64_t L, int tid,
                                             int nthreads) {
  int64_t Lpad = 1;
  while (Lpad < L) Lpad <<= 1;
  for (int64_t k = 2; k <= Lpad; k <<= 1) {
    for (int64_t t = tid; t < (Lpad >> 1); t += nthreads) {  // flip step
      int64_t i = (t / (k >> 1)) * k + (t % (k >> 1));
      int64_t l = i ^ (k - 1);
      if (l < L) cmpx<HAS_VAL>(keys, vals, i, l);
    }
    if (BLOCK) __syncthreads(); else __syncwarp();
    for (int64_t j = k >> 2; j > 0; j >>= 1) {
      for (int64_t t = tid; t < (Lpad >> 1); t += nthreads) {
        int64_t i = (t / j) * (j << 1) + (t % j);
        int64_t l = i + j;
        if (l < L) cmpx<HAS_VAL>(keys, vals, i, l);
      }
      if (BLOCK) __syncthreads(); else __syncwarp();
    }
  }
}

constexpr int kWarpSortMax = 128;     // per-warp smem segment
constexpr int kBlockSortMax = 4096;   // per-CTA smem segment

// Bitonic sort of up to 64 (key, value) pairs held two per lane (element r * 32 + lane in register r), ascending:
// 20 shuffle stages + 1 in-thread stage, no shared memory.  Padding keys are INT32_MAX.
template <bool HAS_VAL>
__device__ __forceinline__ void warp_sort64(int32_t& k0, int32_t& k1, int32_t& v0, int32_t& v1, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {            // k == 64: the partner is this lane's other register; ascending
        if (k0 > k1) {
          const int32_t t = k0; k0 = k1; k1 = t;
          if (HAS_VAL) { const int32_t u = v0; v0 = v1; v1 = u; }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          int32_t& kk = r ? k1 : k0;
          int32_t& vv = r ? v1 : v0;
          const int32_t ok = __shfl_xor_sync(0xffffffffu, kk, j);
          const int32_t ov = HAS_VAL ? __shfl_xor_sync(0xffffffffu, vv, j) : 0;
          const bool up = ((r * 32 + lane) & k) == 0;
          const bool lower = (lane & j) == 0;
          const bool take = (up == lower) ? (ok < kk) : (ok > kk);
          if (take) {
            kk = ok;
            if (HAS_VAL) vv = ov;
          }
        }
      }
    }
  }
}

// one warp per segment, L <= kWarpSortMax; longer segments are appended to `long_list`
template <bool HAS_VAL>
__global__ void seg_sort_warp_kernel(int64_t nseg, const int64_t* __restrict__ segptr,
                                     int32_t* keys, int32_t* vals, int32_t* long_list,
                                     int32_t* long_count, int32_t* dup_count) {
  __shared__ int32_t sk[8][kWarpSortMax];
  __shared__ int32_t sv[8][kWarpSortMax];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * 8 + w;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t seg = warp; seg < nseg; seg += nwarps) {
    int64_t s = segptr[seg];
    int64_t L = segptr[seg + 1] - s;
    if (L > kWarpSortMax) {
      if (lane == 0) long_list[atomicAdd(long_count, 1)] = (int32_t)seg;
      continue;
    }
    if (L < 2) continue;
    if (L <= 64) {      // registers only (the common case: by-source segments of a degree-50 graph)
      int32_t k0 = lane < L ? keys[s + lane] : INT32_MAX, k1 = 32 + lane < L ? keys[s + 32 + lane] : INT32_MAX;
      int32_t v0 = 0, v1 = 0;
      if (HAS_VAL) {
        v0 = lane < L ? vals[s + lane] : 0;
        v1 = 32 + lane < L ? vals[s + 32 + lane] : 0;
      }
      warp_sort64<HAS_VAL>(k0, k1, v0, v1, lane);
      if (lane < L) keys[s + lane] = k0;
      if (32 + lane < L) keys[s + 32 + lane] = k1;
      if (HAS_VAL) {
        if (lane < L) vals[s + lane] = v0;
        if (32 + lane < L) vals[s + 32 + lane] = v1;
      }
      if (dup_count != nullptr) {      // equal neighbours after sorting (element 32 r + lane against its predecessor)
        const int32_t p0 = __shfl_up_sync(0xffffffffu, k0, 1), l31 = __shfl_sync(0xffffffffu, k0, 31);
        const int32_t p1 = __shfl_up_sync(0xffffffffu, k1, 1);
        int d = (lane > 0 && lane < L && p0 == k0) + (32 + lane < L && (lane > 0 ? p1 : l31) == k1);
        d = __reduce_add_sync(0xffffffffu, d);
        if (d && lane == 0) atomicAdd(dup_count, d);
      }
      continue;
    }
    bool sorted = true;
    for (int i = lane; i < L; i += 32) {
      sk[w][i] = keys[s + i];
      if (HAS_VAL) sv[w][i] = vals[s + i];
      if (i > 0 && keys[s + i - 1] > keys[s + i]) sorted = false;
    }
    sorted = __all_sync(0xffffffffu, sorted);
    __syncwarp();
    if (!sorted) {
      bitonic_sort<HAS_VAL, false>(sk[w], sv[w], L, lane, 32);
      for (int i = lane; i < L; i += 32) {
        keys[s + i] = sk[w][i];
        if (HAS_VAL) vals[s + i] = sv[w][i];
      }
    }
    if (dup_count != nullptr) {
      int d = 0;
      for (int i = lane + 1; i < L; i += 32) d += (sk[w][i] == sk[w][i - 1]);
      if (d) atomicAdd(dup_count, d);
    }
    __syncwarp();
  }
}

// persistent CTAs over the list of long segments; smem sort when it fits, global otherwise
template <bool HAS_VAL>
__global__ void seg_sort_block_kernel(const int64_t* __restrict__ segptr, int32_t* keys,
                                      int32_t* vals, const int32_t* __restrict__ long_list,
                                      const int32_t* __restrict__ long_count, int32_t* dup_count) {
  extern __shared__ int32_t smem[];
  int32_t* sk = smem;
  int32_t* sv = smem + kBlockSortMax;
  const int nlong = *long_count;
  for (int li = blockIdx.x; li < nlong; li += gridDim.x) {
    int64_t seg = long_list[li];
    int64_t s = segptr[seg];
    int64_t L = segptr[seg + 1] - s;
    if (L <= kBlockSortMax) {
      for (int i = threadIdx.x; i < L; i += blockDim.x) {
        sk[i] = keys[s + i];
        if (HAS_VAL) sv[i] = vals[s + i];
      }
      __syncthreads();
      bitonic_sort<HAS_VAL, true>(sk, sv, L, threadIdx.x, blockDim.x);
      for (int i = threadIdx.x; i < L; i += blockDim.x) {
        keys[s + i] = sk[i];
        if (HAS_VAL) vals[s + i] = sv[i];
      }
      __syncthreads();
    } else {
      __syncthreads();
      bitonic_sort<HAS_VAL, true>(keys + s, vals + s, L, threadIdx.x, blockDim.x);
    }
    if (dup_count != nullptr) {
      int d = 0;
      for (int64_t i = threadIdx.x + 1; i < L; i += blockDim.x) d += (keys[s + i] == keys[s + i - 1]);
      if (d) atomicAdd(dup_count, d);
    }
    __syncthreads();
  }
}

static int launch_scan(const int32_t* counts, int64_t n, int64_t* indptr, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  int64_t nchunks = ceil_div64(n, kScanChunk);
  if (nchunks < 1) nchunks = 1;
  if (ws_bytes < (size_t)nchunks * sizeof(int64_t)) return fail_arg("han_scan_counts", "workspace too small");
  int64_t* chunk = reinterpret_cast<int64_t*>(ws);
  scan_chunk_sums_kernel<<<(unsigned)nchunks, kScanThreads, 0, st>>>(counts, n, chunk);
  scan_chunk_offsets_kernel<<<1, kScanThreads, 0, st>>>(chunk, nchunks);
  scan_apply_kernel<<<(unsigned)nchunks, kScanThreads, 0, st>>>(counts, n, chunk, indptr);
  return check_launch("han_scan_counts");
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

template <bool HAS_VAL>
static int launch_seg_sort(int64_t nseg, const int64_t* segptr, int32_t* keys, int32_t* vals,
                           int32_t* long_list, int32_t* long_count, int32_t* dup_count,
                           cudaStream_t st) {
  cudaMemsetAsync(long_count, 0, sizeof(int32_t), st);
  unsigned grid = (unsigned)((nseg + 7) / 8);
  if (grid > 148u * 32u) grid = 148u * 32u;
  if (grid < 1) grid = 1;
  seg_sort_warp_kernel<HAS_VAL><<<grid, 256, 0, st>>>(nseg, segptr, keys, vals, long_list, long_count,
                                                      dup_count);
  size_t smem = 2 * kBlockSortMax * sizeof(int32_t);
  seg_sort_block_kernel<HAS_VAL><<<kNumSMs * 2, 512, smem, st>>>(segptr, keys, vals, long_list,
                                                                 long_count, dup_count);
  return check_launch("han_seg_sort");
}

}  // namespace han

using namespace han;

extern "C" {

int han_version(void) { return 100; }
const char* han_last_error(void) { return g_last_error; }

int han_dense_row_counts(const void* dense, int dtype, int kind, int64_t n, int64_t ld,
                         int32_t* row_counts, int32_t* bad_count, han_stream_t stream) {
  HAN_REQUIRE(dense && row_counts, "null pointer");
  HAN_REQUIRE(n > 0 && ld >= n, "n > 0 and ld >= n required");
  HAN_REQUIRE(kind == HAN_DENSE_ADJ || kind == HAN_DENSE_BIAS || kind == HAN_DENSE_POSITIVE, "kind");
  HAN_REQUIRE(dtype == HAN_F32 || dtype == HAN_F64, "dtype");
  cudaStream_t st = as_stream(stream);
  if (bad_count) cudaMemsetAsync(bad_count, 0, sizeof(int32_t), st);
  unsigned grid = (unsigned)ceil_div64(n, 8);
  if (grid > 148u * 16u) grid = 148u * 16u;
  if (dtype == HAN_F32)
    dense_row_counts_kernel<float><<<grid, 256, 0, st>>>((const float*)dense, kind, n, ld, row_counts, bad_count);
  else
    dense_row_counts_kernel<double><<<grid, 256, 0, st>>>((const double*)dense, kind, n, ld, row_counts, bad_count);
  return check_launch(__func__);
}

size_t han_scan_workspace_bytes(int64_t n) {
  int64_t nchunks = ceil_div64(n > 0 ? n : 1, kScanChunk);
  return align256((size_t)nchunks * sizeof(int64_t));
}

int han_scan_counts(const int32_t* counts, int64_t n, int64_t* indptr, void* ws, size_t ws_bytes,
                    han_stream_t stream) {
  HAN_REQUIRE(counts && indptr && ws, "null pointer");
  HAN_REQUIRE(n > 0, "n > 0 required");
  return launch_scan(counts, n, indptr, ws, ws_bytes, as_stream(stream));
}

int han_dense_fill_indices(const void* dense, int dtype, int kind, int64_t n, int64_t ld,
                           const int64_t* indptr, int32_t* indices, han_stream_t stream) {
  HAN_REQUIRE(dense && indptr && indices, "null pointer");
  HAN_REQUIRE(n > 0 && ld >= n, "n > 0 and ld >= n required");
  HAN_REQUIRE(n < ((int64_t)1 << 31), "n must fit int32 column ids");
  cudaStream_t st = as_stream(stream);
  unsigned grid = (unsigned)ceil_div64(n, 8);
  if (grid > 148u * 16u) grid = 148u * 16u;
  if (dtype == HAN_F32)
    dense_fill_kernel<float><<<grid, 256, 0, st>>>((const float*)dense, kind, n, ld, indptr, indices);
  else if (dtype == HAN_F64)
    dense_fill_kernel<double><<<grid, 256, 0, st>>>((const double*)dense, kind, n, ld, indptr, indices);
  else
    return fail_arg(__func__, "dtype");
  return check_launch(__func__);
}

// workspace layout: [counts/cursor int32 n_cols][scan chunks][long_list int32 n_cols][long_count]
size_t han_transpose_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t nnz) {
  (void)n_rows; (void)nnz;
  return align256((size_t)n_cols * 4) + han_scan_workspace_bytes(n_cols) + align256((size_t)n_cols * 4) + 256;
}

int han_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr,
                      const int32_t* indices, int64_t* t_indptr, int32_t* t_indices, int32_t* perm,
                      void* ws, size_t ws_bytes, han_stream_t stream) {
  HAN_REQUIRE(indptr && t_indptr && ws, "null pointer");
  HAN_REQUIRE(n_rows > 0 && n_cols > 0 && nnz >= 0, "sizes");
  HAN_REQUIRE(nnz < ((int64_t)1 << 31) && n_rows < ((int64_t)1 << 31), "nnz and n_rows must fit int32");
  HAN_REQUIRE(ws_bytes >= han_transpose_workspace_bytes(n_rows, n_cols, nnz), "workspace too small");
  cudaStream_t st = as_stream(stream);
  char* p = reinterpret_cast<char*>(ws);
  int32_t* counts = reinterpret_cast<int32_t*>(p);
  p += align256((size_t)n_cols * 4);
  void* scan_ws = p;
  size_t scan_bytes = han_scan_workspace_bytes(n_cols);
  p += scan_bytes;
  int32_t* long_list = reinterpret_cast<int32_t*>(p);
  p += align256((size_t)n_cols * 4);
  int32_t* long_count = reinterpret_cast<int32_t*>(p);

  cudaMemsetAsync(counts, 0, (size_t)n_cols * 4, st);
  if (nnz > 0) {
    unsigned g = (unsigned)ceil_div64(nnz, 256);
    if (g > 148u * 32u) g = 148u * 32u;
    col_count_kernel<<<g, 256, 0, st>>>(indices, nnz, counts);
  }
  int rc = launch_scan(counts, n_cols, t_indptr, scan_ws, scan_bytes, st);
  if (rc) return rc;
  if (nnz == 0) return 0;
  cudaMemsetAsync(counts, 0, (size_t)n_cols * 4, st);  // reuse as cursor
  unsigned g = (unsigned)ceil_div64(n_rows, 8);
  if (g > 148u * 32u) g = 148u * 32u;
  {
    // column blocks of ~48 MB of output each (assuming columns of similar weight; a skewed graph only loses locality)
    const int64_t bytes = nnz * (perm ? 8 : 4);
    int64_t passes = (bytes + (48ll << 20) - 1) / (48ll << 20);
    if (passes < 1) passes = 1;
    if (passes > 64) passes = 64;
    const int64_t per = (n_cols + passes - 1) / passes;
    for (int64_t b = 0; b < passes; ++b) {
      const int64_t lo = b * per, hi = (lo + per < n_cols) ? lo + per : n_cols;
      if (lo >= hi) break;
      transpose_fill_kernel<<<g, 256, 0, st>>>(n_rows, indptr, indices, t_indptr, counts, t_indices, perm, (int32_t)lo,
                                               (int32_t)hi);
    }
  }
  rc = check_launch(__func__);
  if (rc) return rc;
  if (perm == nullptr) return launch_seg_sort<false>(n_cols, t_indptr, t_indices, nullptr, long_list, long_count, nullptr, st);
  return launch_seg_sort<true>(n_cols, t_indptr, t_indices, perm, long_list, long_count, nullptr, st);
}

// ws-free variant for CSR rows: uses payload==NULL -> keys only.  Needs a long-segment list, so
// the caller passes `scratch` = int32[n_rows + 64] via payload-independent argument below.
int han_csr_sort_rows(int64_t n_rows, const int64_t* indptr, int32_t* indices, int32_t* payload,
                      int32_t* scratch, int32_t* dup_count, han_stream_t stream) {
  HAN_REQUIRE(indptr && indices && scratch, "null pointer");
  HAN_REQUIRE(n_rows > 0, "n_rows > 0 required");
  cudaStream_t st = as_stream(stream);
  if (dup_count) cudaMemsetAsync(dup_count, 0, sizeof(int32_t), st);
  int32_t* long_count = scratch;
  int32_t* long_list = scratch + 64;
  if (payload)
    return launch_seg_sort<true>(n_rows, indptr, indices, payload, long_list, long_count, dup_count, st);
  return launch_seg_sort<false>(n_rows, indptr, indices, nullptr, long_list, long_count, dup_count, st);
}

}  // extern "C"
