// K-A on 5th-generation tensor cores: T[g] = [ X W_g | f2 ], f1 -> R[g], for up to 4 meta-paths
// (N = G*64 <= 256 accumulator columns) in one pass over X.  Replaces utils/layers.py:20,23,24.
//
//   operands  : TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) -> shared memory, 2-stage mbarrier ring
//   math      : tcgen05.mma.cta_group::1.kind::tf32, M=128 x N=G*64 x K=8 per instruction,
//               accumulators in TMEM (128 lanes x 256 columns fp32)
//   precision : there is no FP32 MMA.  mode 1 (3xTF32): X and W are split hi + lo (hi = fp32 with the
//               13 low mantissa bits cleared, lo = x - hi exactly) and D += Xhi Whi + Xlo Whi + Xhi Wlo,
//               dropping only the lo*lo term (2^-22 relative) -> FP32-grade.  mode 2 (2xTF32): X is
//               exactly representable in tf32 (0/1 bag-of-words features of ACM/DBLP/IMDB), only W is
//               split.  mode 3: plain TF32.
//   epilogue  : tcgen05.ld (one accumulator row per thread) -> f1 = S a1 + b1 fused (f2 = S a2 + b2 is recomputed by the gather kernels from the table row)
//               here, rows written straight into the node table / row record layout.
//
// Warp roles (192 threads, 1 CTA per SM): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
// (one elected lane), warps 2-5 = X hi/lo split (mode 1) and epilogue (warp w owns TMEM lanes
// 32*(w%4) .. +31).
#include <cuda.h>

#include "han_common.cuh"

namespace han {

constexpr int TC_BM = 128;       // rows of X per CTA
constexpr int TC_BK = 32;        // fp32 K elements per stage = 128 B = one swizzle span
constexpr int TC_STAGES = 2;
constexpr int TC_MAXN = 256;
constexpr int TC_THREADS = 192;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 4;      // 16 KB
constexpr uint32_t TC_B_BYTES = TC_MAXN * TC_BK * 4;    // 32 KB (N <= 256 rows)
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // A_hi, A_lo, B_hi, B_lo
constexpr uint32_t TC_SMEM_BYTES = 1024 + TC_STAGES * TC_STAGE_BYTES + 4 * 4 * (64 + 64 + 8 + 8) + 256;
constexpr uint32_t kSpinLimit = 1u << 28;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > kSpinLimit) __trap();   // turn a protocol bug into an error, not a hang
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16-byte row store: plain, or multimem.st to a multicast address (one store, every rank's copy)
__device__ __forceinline__ void store_row16(float* p, bool multicast, float a, float b, float c, float d) {
  if (multicast)
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
  else
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                       // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;  // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}

// ---- W^T hi/lo preparation (tiny): Wt[part][n][k], k padded to Fp ----------------------------------
__global__ void wt_split_kernel(const float* __restrict__ W, int64_t F, int64_t NC, int64_t Fp,
                                float* __restrict__ Wt_hi, float* __restrict__ Wt_lo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= NC * Fp) return;
  const int64_t n = idx / Fp, k = idx % Fp;
  const float w = (k < F) ? W[k * NC + n] : 0.f;
  const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
  Wt_hi[idx] = hi;
  Wt_lo[idx] = w - hi;
}

// ---- the projection kernel --------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
project_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                  const __grid_constant__ CUtensorMap tmBlo, int64_t n_rows, int nkb, int G,
                  const float* __restrict__ a1, const float* __restrict__ b1, float* __restrict__ T,
                  float* __restrict__ R,
                  float* __restrict__ Tmc, int64_t t_rows, int64_t t_row0, int64_t r_rows) {
  constexpr int D = 64, K = 8, H = 8, TS = 64, RS = 88;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  // stage s: [A_hi 16K][A_lo 16K][B_hi 32K][B_lo 32K]
  const uint32_t tail = base + TC_STAGES * TC_STAGE_BYTES;
  float* par = reinterpret_cast<float*>(gen_base + TC_STAGES * TC_STAGE_BYTES);   // [G][a1 64|b1 8]
  const uint32_t bars = tail + 4 * 4 * (64 + 8);
  const uint32_t full0 = bars, xform0 = bars + 16, empty0 = bars + 32, accum = bars + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + (bars + 64 - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM;
  const int NC = G * D;

  for (int i = threadIdx.x; i < G * 72; i += TC_THREADS) {
    const int g = i / 72, o = i % 72;
    par[i] = (o < 64) ? a1[g * 64 + o] : b1[g * 8 + (o - 64)];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(xform0 + 8 * s, 128);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(2 * TC_MAXN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t tx = TC_A_BYTES + (uint32_t)NC * TC_BK * 4 * (MODE == 3 ? 1 : 2);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t st = base + s * TC_STAGE_BYTES;
        mbar_expect_tx(full0 + 8 * s, tx);
        tma_load_2d(st, &tmA, kb * TC_BK, m0, full0 + 8 * s);
        tma_load_2d(st + 2 * TC_A_BYTES, &tmBhi, kb * TC_BK, 0, full0 + 8 * s);
        if (MODE != 3) tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &tmBlo, kb * TC_BK, 0, full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait((MODE == 1 ? xform0 : full0) + 8 * s, ph);
        tc_fence_after();
        const uint32_t st = base + s * TC_STAGE_BYTES;
        const uint64_t a_hi = make_kmajor_sw128_desc(st);
        const uint64_t a_lo = make_kmajor_sw128_desc(st + TC_A_BYTES);
        const uint64_t b_hi = make_kmajor_sw128_desc(st + 2 * TC_A_BYTES);
        const uint64_t b_lo = make_kmajor_sw128_desc(st + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);   // 32 B per K=8 step inside the swizzle span
          // Two accumulators: the tensor core adds into its fp32 accumulator with TRUNCATION (~0.5 ulp of the running
          // sum per instruction, biased), so the large hi*hi terms get a chain of their own (F/8 instructions instead
          // of 3F/8) and the 2^-11-times-smaller correction terms lo*hi + hi*lo accumulate separately; the epilogue
          // adds the two in fp32 (measured at the 2M-node scale: parity of the step 1.2e-5 -> see profiles/).
          const uint32_t corr = tmem_base + TC_MAXN;
          if (MODE == 1) umma_tf32(corr, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
          if (MODE != 3) umma_tf32(corr, a_hi + adv, b_lo + adv, idesc, (MODE == 1) || (kb | k) != 0);
          umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
        }
        umma_commit(empty0 + 8 * s);   // frees the stage when these MMAs have read it
      }
      umma_commit(accum);
    }
  } else {
    // ===== warps 2-5: X hi/lo split (mode 1), then epilogue =====
    const int t = threadIdx.x - 64;   // 0..127
    if (MODE == 1) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        uint4* hi = reinterpret_cast<uint4*>(gen_base + s * TC_STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(gen_base + s * TC_STAGE_BYTES + TC_A_BYTES);
#pragma unroll
        for (int i = 0; i < (int)(TC_A_BYTES / 16 / 128); ++i) {   // 8 x 16 B per thread, layout-agnostic
          const int idx = t + 128 * i;
          uint4 x = hi[idx];
          uint4 h, l;
          h.x = x.x & 0xFFFFE000u; l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(h.x));
          h.y = x.y & 0xFFFFE000u; l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(h.y));
          h.z = x.z & 0xFFFFE000u; l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(h.z));
          h.w = x.w & 0xFFFFE000u; l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(h.w));
          hi[idx] = h;
          lo[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (MMA)
        mbar_arrive(xform0 + 8 * s);
      }
    }
    // ---- epilogue: one accumulator row per thread ----
    mbar_wait(accum, 0);
    tc_fence_after();
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int64_t row = (int64_t)m0 + q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int g = 0; g < G; ++g) {
      uint32_t v0[32], v1[32];
      tmem_ld32(lane_addr + g * D, v0);
      tmem_ld32(lane_addr + g * D + 32, v1);
      tmem_ld_wait();
      if (MODE != 3) {   // + the correction accumulator (lo*hi + hi*lo)
        uint32_t c[32];
        tmem_ld32(lane_addr + TC_MAXN + g * D, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v0[i] = __float_as_uint(__uint_as_float(v0[i]) + __uint_as_float(c[i]));
        tmem_ld32(lane_addr + TC_MAXN + g * D + 32, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v1[i] = __float_as_uint(__uint_as_float(v1[i]) + __uint_as_float(c[i]));
      }
      if (row < n_rows) {
        const float* pg = par + g * 72;
        float f1[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float s1 = pg[64 + k];
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const int c = k * H + h;
            s1 = fmaf(__uint_as_float(c < 32 ? v0[c] : v1[c - 32]), pg[c], s1);
          }
          f1[k] = s1;
        }
        // node-table row: local store, or ONE multicast store per 16 B that NVSwitch replicates into the
        // same offset of every rank's table (GEMM epilogue fused with the all-gather)
        // (t_rows / r_rows: rows per meta-path of the destination tables; they exceed n_rows when the
        //  destination is this rank's slice of a full-size table shared with the other ranks)
        float* tbase = (Tmc != nullptr) ? Tmc + ((int64_t)g * t_rows + t_row0 + row) * TS
                                        : T + ((int64_t)g * t_rows + t_row0 + row) * TS;
        const bool mc = Tmc != nullptr;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          store_row16(tbase + 4 * c, mc, __uint_as_float(v0[4 * c]), __uint_as_float(v0[4 * c + 1]),
                      __uint_as_float(v0[4 * c + 2]), __uint_as_float(v0[4 * c + 3]));
#pragma unroll
        for (int c = 0; c < 8; ++c)
          store_row16(tbase + 32 + 4 * c, mc, __uint_as_float(v1[4 * c]), __uint_as_float(v1[4 * c + 1]),
                      __uint_as_float(v1[4 * c + 2]), __uint_as_float(v1[4 * c + 3]));
        float4* rp = reinterpret_cast<float4*>(R + ((int64_t)g * r_rows + row) * RS + D);
        rp[0] = make_float4(f1[0], f1[1], f1[2], f1[3]);
        rp[1] = make_float4(f1[4], f1[5], f1[6], f1[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TC_MAXN) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map: inner dimension `inner` elements (contiguous), `outer` rows with `ld` floats
// between rows; box = 32 x box_rows, 128-byte swizzle, out-of-bounds elements read as zero.
static int make_map(CUtensorMap* m, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail_arg("han_project_fwd", "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error), "han_project_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -2;
  }
  return 0;
}

}  // namespace han

using namespace han;

extern "C" {

// workspace for the transposed, hi/lo-split weights: 2 * NC * roundup(F,4) floats
size_t han_project_tc_workspace_bytes(int64_t F, int G, int K, int H) {
  const int64_t Fp = (F + 3) / 4 * 4;
  return (size_t)2 * G * K * H * Fp * sizeof(float);
}

int han_project_fwd_tc(const float* X, int64_t n, int64_t F, int64_t ldx, const float* W, int G, int K, int H,
                       const float* a1, const float* b1, float* T, float* R, float* T_mc, int64_t t_rows, int64_t t_row0, int64_t r_rows, int mode, void* ws,
                       size_t ws_bytes, han_stream_t stream) {
  HAN_REQUIRE(X && W && a1 && b1 && (T || T_mc) && R && ws, "null pointer");
  if (t_rows == 0) t_rows = n;   // destination tables are exactly [G][n][.]
  if (r_rows == 0) r_rows = n;
  HAN_REQUIRE(t_rows >= t_row0 + n && t_row0 >= 0 && r_rows >= n, "t_row0 + n <= t_rows and n <= r_rows required");
  HAN_REQUIRE(((uintptr_t)T_mc % 16 == 0) && ((uintptr_t)T % 16 == 0) && ((uintptr_t)R % 16 == 0), "16-byte alignment");
  HAN_REQUIRE(K == 8 && H == 8, "the tensor-core projection is built for K = H = 8");
  HAN_REQUIRE(G >= 1 && G <= 4, "1 <= G <= 4 meta-paths per launch (256 accumulator columns)");
  HAN_REQUIRE(mode >= 1 && mode <= 3, "mode 1 (3xTF32), 2 (2xTF32, tf32-exact X) or 3 (TF32)");
  HAN_REQUIRE(n > 0 && F > 0 && ldx >= F && n < ((int64_t)1 << 31), "sizes");
  HAN_REQUIRE(((uintptr_t)X % 16 == 0) && (ldx % 4 == 0), "X must be 16-byte aligned with ldx % 4 == 0 (TMA)");
  HAN_REQUIRE(ws_bytes >= han_project_tc_workspace_bytes(F, G, K, H), "workspace too small");
  cudaStream_t st = as_stream(stream);
  const int64_t NC = (int64_t)G * 64, Fp = (F + 3) / 4 * 4;
  float* Wt_hi = reinterpret_cast<float*>(ws);
  float* Wt_lo = Wt_hi + NC * Fp;
  wt_split_kernel<<<(unsigned)ceil_div64(NC * Fp, 256), 256, 0, st>>>(W, F, NC, Fp, Wt_hi, Wt_lo);

  CUtensorMap tmA, tmBhi, tmBlo;
  int rc = make_map(&tmA, X, (uint64_t)F, (uint64_t)n, (uint64_t)ldx, TC_BM);
  if (rc) return rc;
  rc = make_map(&tmBhi, Wt_hi, (uint64_t)Fp, (uint64_t)NC, (uint64_t)Fp, (uint32_t)NC);
  if (rc) return rc;
  rc = make_map(&tmBlo, Wt_lo, (uint64_t)Fp, (uint64_t)NC, (uint64_t)Fp, (uint32_t)NC);
  if (rc) return rc;

  const int nkb = (int)ceil_div64(F, TC_BK);
  const unsigned grid = (unsigned)ceil_div64(n, TC_BM);
  HAN_SMEM_ATTR_ONCE(project_tc_kernel<1>, TC_SMEM_BYTES);
  HAN_SMEM_ATTR_ONCE(project_tc_kernel<2>, TC_SMEM_BYTES);
  HAN_SMEM_ATTR_ONCE(project_tc_kernel<3>, TC_SMEM_BYTES);
  if (mode == 1)
    project_tc_kernel<1><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tmA, tmBhi, tmBlo, n, nkb, G, a1, b1, T, R,
                                                                  T_mc, t_rows, t_row0, r_rows);
  else if (mode == 2)
    project_tc_kernel<2><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tmA, tmBhi, tmBlo, n, nkb, G, a1, b1, T, R,
                                                                  T_mc, t_rows, t_row0, r_rows);
  else
    project_tc_kernel<3><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tmA, tmBhi, tmBlo, n, nkb, G, a1, b1, T, R,
                                                                  T_mc, t_rows, t_row0, r_rows);
  return check_launch(__func__);
}

}  // extern "C"
