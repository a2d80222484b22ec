// K-E on tensor cores: dW [F][G*64] = X^T dS, the backward of the projection (utils/layers.py:20 under
// TF autodiff), as a split-K tcgen05 GEMM whose reduction dimension is the node index.
//
// Both operands are "transposed" here (the reduction index n is the slow index of X [n][F] and of
// dS [g][n][64]), so TMA cannot deliver them K-major.  Instead four producer warps read 32-row slabs
// with fully coalesced loads, split every value into tf32 hi + lo in registers, and store the slab
// TRANSPOSED straight into the canonical K-major SWIZZLE_128B layout the UMMA descriptors expect
// (16-byte chunk c of row r lives at chunk position c ^ (r % 8) of its 128-byte row; conflict-free
// STS.128).  One elected thread issues tcgen05.mma.kind::tf32 (M=128 features x N=G*64 x K=8) with
// 3xTF32 accumulation into TMEM.
//
// Accumulation accuracy: the tensor core adds into its fp32 accumulator with TRUNCATION, so a chain of L node
// rows drifts by ~7e-9 * L relative to the sum (measured on B200, tools/dw_accuracy.py: 2e-4 over the 27k rows
// one CTA owns on the 2M-node graph, against 3e-6 for the FFMA kernel -- outside the 1e-5 contract).  The chain
// is therefore cut every BT_CHAIN_KB k-blocks (256 rows): two TMEM accumulators alternate, and while the MMAs
// fill one, four drain warps read the other with tcgen05.ld and add it (fp32, round-to-nearest) into this CTA's
// split-K partial in global memory (128 KB per CTA, L2-resident).  The deterministic second-stage reduce is
// shared with the FFMA path (project.cu).
#include "han_common.cuh"

namespace han {

constexpr int BT_BM = 128;      // features per CTA (MMA M)
constexpr int BT_BK = 32;       // rows of X / dS per stage (MMA K, one 128-byte swizzle span)
constexpr int BT_STAGES = 2;
constexpr int BT_MAXN = 256;
constexpr int BT_PRODUCERS = 256;                 // warps 1-8: every load of a stage is in flight before the first store
constexpr int BT_DRAINERS = 128;                  // warps 9-12: one per TMEM lane quarter (warp id % 4)
constexpr int BT_THREADS = 32 + BT_PRODUCERS + BT_DRAINERS;     // warp 0: MMA issuer + TMEM
constexpr int BT_CHAIN_KB = 8;                    // k-blocks (of 32 node rows) accumulated in TMEM before a drain
constexpr uint32_t BT_A_BYTES = BT_BM * BT_BK * 4;
constexpr uint32_t BT_B_BYTES = BT_MAXN * BT_BK * 4;
constexpr uint32_t BT_STAGE_BYTES = 2 * BT_A_BYTES + 2 * BT_B_BYTES;
constexpr uint32_t BT_SMEM_BYTES = 1024 + BT_STAGES * BT_STAGE_BYTES + 256;
constexpr uint32_t kBtSpinLimit = 1u << 28;

__device__ __forceinline__ uint32_t bt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bt_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bt_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bt_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > kBtSpinLimit) __trap();
  }
}
__device__ __forceinline__ void bt_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void bt_umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bt_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ uint64_t bt_desc(uint32_t smem_addr) {   // K-major, SWIZZLE_128B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// byte offset of 16-byte chunk kc (4 consecutive K values) of row m in a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_chunk(int m, int kc) {
  return (uint32_t)((m >> 3) * 1024 + (m & 7) * 128 + ((kc ^ (m & 7)) << 4));
}
__device__ __forceinline__ void split_store(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float x0, float x1,
                                            float x2, float x3, bool want_lo) {
  uint4 h, l;
  h.x = __float_as_uint(x0) & 0xFFFFE000u; h.y = __float_as_uint(x1) & 0xFFFFE000u;
  h.z = __float_as_uint(x2) & 0xFFFFE000u; h.w = __float_as_uint(x3) & 0xFFFFE000u;
  *reinterpret_cast<uint4*>(hi_tile + off) = h;
  if (want_lo) {
    l.x = __float_as_uint(x0 - __uint_as_float(h.x)); l.y = __float_as_uint(x1 - __uint_as_float(h.y));
    l.z = __float_as_uint(x2 - __uint_as_float(h.z)); l.w = __float_as_uint(x3 - __uint_as_float(h.w));
    *reinterpret_cast<uint4*>(lo_tile + off) = l;
  }
}

// MODE 1: 3xTF32 (X and dS split); MODE 2: X exactly tf32 (0/1 features), only dS split; MODE 3: plain TF32
template <int MODE>
__global__ void __launch_bounds__(BT_THREADS, 1)
project_bwd_tc_kernel(const float* __restrict__ X, int64_t n, int64_t F, int64_t ldx, const float* __restrict__ dS,
                      int G, int64_t rows_per_split, float* __restrict__ part) {
  constexpr int D = 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (bt_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - bt_smem_u32(smem_raw));
  const uint32_t bars = base + BT_STAGES * BT_STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 16, accfull0 = bars + 32, accfree0 = bars + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + BT_STAGES * BT_STAGE_BYTES + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NC = G * D;
  const int64_t f0 = (int64_t)blockIdx.x * BT_BM;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(n, r_begin + rows_per_split);
  const int nkb = (int)((max((int64_t)0, r_end - r_begin) + BT_BK - 1) / BT_BK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < BT_STAGES; ++s) {
      bt_mbar_init(full0 + 8 * s, BT_PRODUCERS);
      bt_mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      bt_mbar_init(accfull0 + 8 * a, 1);
      bt_mbar_init(accfree0 + 8 * a, BT_DRAINERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bt_smem_u32(tmem_slot)),
                 "n"(2 * BT_MAXN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== MMA issuer =====
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(BT_BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % BT_STAGES;
        const uint32_t ph = (kb / BT_STAGES) & 1;
        const int chain = kb / BT_CHAIN_KB, kc0 = kb % BT_CHAIN_KB;     // chain c accumulates into accumulator c & 1
        const uint32_t acc = tmem_base + (uint32_t)((chain & 1) * BT_MAXN);
        if (kc0 == 0 && chain >= 2) {                                   // the drain of chain c-2 released this accumulator
          bt_mbar_wait(accfree0 + 8 * (chain & 1), (uint32_t)(((chain >> 1) - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        bt_mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + s * BT_STAGE_BYTES;
        const uint64_t a_hi = bt_desc(st), a_lo = bt_desc(st + BT_A_BYTES);
        const uint64_t b_hi = bt_desc(st + 2 * BT_A_BYTES), b_lo = bt_desc(st + 2 * BT_A_BYTES + BT_B_BYTES);
#pragma unroll
        for (int k = 0; k < BT_BK / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          if (MODE == 1) bt_umma_tf32(acc, a_lo + adv, b_hi + adv, idesc, (kc0 | k) != 0);
          if (MODE != 3) bt_umma_tf32(acc, a_hi + adv, b_lo + adv, idesc, (MODE == 1) || (kc0 | k) != 0);
          bt_umma_tf32(acc, a_hi + adv, b_hi + adv, idesc, (MODE != 3) || (kc0 | k) != 0);
        }
        bt_umma_commit(empty0 + 8 * s);
        if (kc0 == BT_CHAIN_KB - 1 || kb == nkb - 1) bt_umma_commit(accfull0 + 8 * (chain & 1));
      }
    }
  } else if (warp > 8) {
    // ===== drain warps 9-12: thread = accumulator row = feature; partial += accumulator of every finished chain =====
    const int q = warp & 3;
    const int64_t f = f0 + q * 32 + lane;
    float* dst = part + ((int64_t)blockIdx.y * F + f) * NC;
    const int nchains = (nkb + BT_CHAIN_KB - 1) / BT_CHAIN_KB;
    // The partial was zeroed by the launcher; every chain is ADDED with red.global (fire-and-forget: no load round
    // trip, so the drain never paces the MMAs).  Each address belongs to one thread of one CTA and same-address
    // reductions of one thread apply in program order, so the sums are deterministic.
    for (int chain = 0; chain < nchains; ++chain) {
      const int a = chain & 1;
      bt_mbar_wait(accfull0 + 8 * a, (uint32_t)((chain >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BT_MAXN);
      for (int c0 = 0; c0 < NC; c0 += 32) {
        uint32_t v[32];
        bt_tmem_ld32(lane_addr + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (f < F) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + c0 + 4 * c), "f"(__uint_as_float(v[4 * c])),
                         "f"(__uint_as_float(v[4 * c + 1])), "f"(__uint_as_float(v[4 * c + 2])),
                         "f"(__uint_as_float(v[4 * c + 3]))
                         : "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      bt_mbar_arrive(accfree0 + 8 * a);
    }
  } else {
    // ===== producers (warps 1-8) =====
    // Per 32-row stage a thread owns half a feature column of X (16 values: feature fa, rows 16*half..) and one
    // column of dS (32 values); all 48 loads are issued before the first transposing store, so a CTA keeps a
    // whole 48 KB stage in flight instead of 4-8 loads per thread.
    const int t = threadIdx.x - 32;   // 0..255
    const int fa = t & 127, half = t >> 7;
    const int64_t f_a = f0 + fa;
    const bool fok = f_a < F;
    const bool cok = t < NC;
    const float* bsrc = dS + ((int64_t)(t >> 6) * n) * D + (t & 63);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % BT_STAGES;
      const uint32_t ph = (kb / BT_STAGES) & 1;
      const int64_t n0 = r_begin + (int64_t)kb * BT_BK;
      float va[16], vb[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int64_t r = n0 + 16 * half + j;
        va[j] = (fok && r < r_end) ? __ldg(X + r * ldx + f_a) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t r = n0 + j;
        vb[j] = (cok && r < r_end) ? __ldg(bsrc + r * D) : 0.f;
      }
      bt_mbar_wait(empty0 + 8 * s, ph ^ 1);
      uint8_t* a_hi = gen + s * BT_STAGE_BYTES;
      uint8_t* a_lo = a_hi + BT_A_BYTES;
      uint8_t* b_hi = a_hi + 2 * BT_A_BYTES;
      uint8_t* b_lo = b_hi + BT_B_BYTES;
      // A = X^T tile: row m = feature, K = 32 node rows (this thread: K chunks 4*half .. 4*half+3)
#pragma unroll
      for (int kc = 0; kc < 4; ++kc)
        split_store(a_hi, a_lo, sw128_chunk(fa, 4 * half + kc), va[4 * kc], va[4 * kc + 1], va[4 * kc + 2],
                    va[4 * kc + 3], MODE == 1);
      // B = dS^T tile: row = output column c (meta-path c/64, feature c%64), K = the same 32 node rows
      if (cok) {
#pragma unroll
        for (int kc = 0; kc < 8; ++kc)
          split_store(b_hi, b_lo, sw128_chunk(t, kc), vb[4 * kc], vb[4 * kc + 1], vb[4 * kc + 2], vb[4 * kc + 3],
                      MODE != 3);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core reads
      bt_mbar_arrive(full0 + 8 * s);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BT_MAXN) : "memory");
  }
}

__global__ void bt_reduce_kernel(const float* __restrict__ part, int splits, int64_t elems, float* __restrict__ outv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= elems) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(int64_t)k * elems + i];
  outv[i] = s;
}

static int bt_splits(int64_t n, int64_t F) {
  const int64_t ftiles = ceil_div64(F, BT_BM);
  int64_t s = kNumSMs / ftiles;
  if (s < 1) s = 1;
  const int64_t maxs = ceil_div64(n, 4 * BT_BK);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return (int)s;
}

}  // namespace han

using namespace han;

extern "C" {

size_t han_project_bwd_tc_workspace_bytes(int64_t n, int64_t F, int G) {
  return (size_t)bt_splits(n, F) * (size_t)F * G * 64 * sizeof(float);
}

int han_project_bwd_tc(const float* X, int64_t n, int64_t F, int64_t ldx, const float* dS, int G, float* dW,
                       void* ws, size_t ws_bytes, int mode, han_stream_t stream) {
  HAN_REQUIRE(X && dS && dW && ws, "null pointer");
  HAN_REQUIRE(n > 0 && F > 0 && ldx >= F, "sizes");
  HAN_REQUIRE(G >= 1 && G <= 4, "1 <= G <= 4 meta-paths per launch (256 accumulator columns)");
  HAN_REQUIRE(mode >= 1 && mode <= 3, "mode 1 (3xTF32), 2 (2xTF32, tf32-exact X) or 3 (TF32)");
  const int splits = bt_splits(n, F);
  HAN_REQUIRE(ws_bytes >= (size_t)splits * F * G * 64 * sizeof(float), "workspace too small");
  HAN_REQUIRE((uintptr_t)ws % 16 == 0, "workspace must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int64_t rows_per_split = ceil_div64(ceil_div64(n, splits), BT_BK) * BT_BK;
  HAN_SMEM_ATTR_ONCE(project_bwd_tc_kernel<1>, BT_SMEM_BYTES);
  HAN_SMEM_ATTR_ONCE(project_bwd_tc_kernel<2>, BT_SMEM_BYTES);
  HAN_SMEM_ATTR_ONCE(project_bwd_tc_kernel<3>, BT_SMEM_BYTES);
  dim3 grid((unsigned)ceil_div64(F, BT_BM), (unsigned)splits);
  float* part = reinterpret_cast<float*>(ws);
  cudaMemsetAsync(part, 0, (size_t)splits * F * G * 64 * sizeof(float), st);    // the drains accumulate into it
  if (mode == 1)
    project_bwd_tc_kernel<1><<<grid, BT_THREADS, BT_SMEM_BYTES, st>>>(X, n, F, ldx, dS, G, rows_per_split, part);
  else if (mode == 2)
    project_bwd_tc_kernel<2><<<grid, BT_THREADS, BT_SMEM_BYTES, st>>>(X, n, F, ldx, dS, G, rows_per_split, part);
  else
    project_bwd_tc_kernel<3><<<grid, BT_THREADS, BT_SMEM_BYTES, st>>>(X, n, F, ldx, dS, G, rows_per_split, part);
  const int64_t elems = F * (int64_t)G * 64;
  bt_reduce_kernel<<<(unsigned)ceil_div64(elems, 256), 256, 0, st>>>(part, splits, elems, dW);
  return check_launch(__func__);
}

}  // extern "C"
