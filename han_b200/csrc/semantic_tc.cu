// K-C on 5th-generation tensor cores (EXPERIMENTAL, opt-in: HAN_SEM_TC=1; the shipped path is the
// mma.sync kernel in semantic.cu).  Status: parity-green on B200 (tests/test_gpu_semantic.py), 2.60 ms vs
// 2.78 ms on the 2M-node config -- the four epilogue warps (tanh + scattered v stores) are the bound now.
// Semantic attention forward, utils/layers.py:152-159, for the shipped shape D = 64, A = 128:
//
//   v = tanh(Z w + b)   s = v . u   beta = softmax_P(s) per node   out[n] = sum_p beta[n,p] Z[n,p]
//
// Why: the warp-level MMA path issues one m16n8k8 per ~2.2 cycles per SM, so the three TF32 passes of
// Z w cost ~1.5 ms of tensor-pipe time on the 2M-node config (profiles/r1_ncu_summary.md section 2c);
// tcgen05 does the same 128 x 128 x 64 tile in 24 instructions (~0.8 us).
//
// One PERSISTENT CTA per SM walks tiles of 128 (node, meta-path) rows (whole nodes per tile):
//   warp 0      TMA producer: w^T hi/lo once (64 KB, SWIZZLE_128B, stays resident), then one Z tile per
//               step into a 2-deep ring (2 K-blocks of 32 floats each)
//   warp 1      TMEM allocator (256 columns = two 128-column accumulators) + single-thread
//               tcgen05.mma.kind::tf32 issuer, 3xTF32 (lo*hi + hi*lo + hi*hi)
//   warps 2-5   split the landed Z tile into hi / lo in place (generic proxy -> fence.proxy.async), and, one
//               tile behind, the epilogue: tcgen05.ld one accumulator row per thread, tanh, dot with u,
//               v written for the backward, per-node softmax over P through shared memory, and the
//               weighted sum from the Z tile still sitting in the ring (Z = hi + lo exactly).
// The epilogue of tile i runs under the MMAs of tile i+1 (second accumulator) and the TMA of tile i+2.
#include <cuda.h>
#include <stdlib.h>

#include "han_common.cuh"

namespace han {

constexpr int ST_BM = 128;                     // rows per tile
constexpr int ST_D = 64, ST_A = 128;
constexpr int ST_BK = 32;                      // fp32 K elements per 128-byte swizzle span
constexpr int ST_MAX_EG = 4;                   // epilogue groups: 4 warps each (one per TMEM lane quarter), splitting the columns
constexpr uint32_t ST_KB_BYTES = ST_BM * ST_BK * 4;          // one K-block of a 128-row operand: 16 KB
constexpr uint32_t ST_W_BYTES = 2 * 2 * ST_KB_BYTES;         // w^T: hi | lo, 2 K-blocks each: 64 KB
constexpr uint32_t ST_STAGE_BYTES = 2 * 2 * ST_KB_BYTES;     // Z tile: hi (2 K-blocks) | lo (2 K-blocks): 64 KB
constexpr int ST_PAR_FLOATS = 2 * ST_A + ST_MAX_EG * ST_BM + ST_BM;          // b | u | partial scores [EG][128] | beta [128]
constexpr uint32_t ST_SMEM_BYTES = 1024 + ST_W_BYTES + 2 * ST_STAGE_BYTES + 4 * ST_PAR_FLOATS + 256;
constexpr uint32_t kStSpinLimit = 1u << 28;

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void st_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > kStSpinLimit) __trap();   // a protocol bug becomes an error, not a hang
  }
}
__device__ __forceinline__ void st_tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void st_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void st_umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void st_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <int N>
__device__ __forceinline__ void st_tmem_ldN(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 32) st_tmem_ld32(taddr, r);
  else st_tmem_ld16(taddr, r);
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t st_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// byte offset of float column c (0..31) of row r inside one K-block tile in the SWIZZLE_128B layout
__device__ __forceinline__ uint32_t st_sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 2) ^ (r & 7)) << 4)) + (c & 3) * 4);
}
template <int NT>
__device__ __forceinline__ void st_bar_epi() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// w [D][A] -> w^T hi / lo [A][D] (K-major operand B)
__global__ void st_wt_split_kernel(const float* __restrict__ w, float* __restrict__ wt_hi, float* __restrict__ wt_lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ST_A * ST_D) return;
  const int a = idx / ST_D, d = idx % ST_D;
  const float x = w[d * ST_A + a];
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  wt_hi[idx] = hi;
  wt_lo[idx] = x - hi;
}

// EG = 1 is the configuration validated on hardware (four epilogue warps, each thread a whole row of 128
// columns); EG = 2 / 4 give every row to 2 / 4 threads of different warp groups (64 / 32 columns each) so that
// 2 / 4 epilogue warps per scheduler hide each other's tanh and store latencies.
// tanh to ~1e-7 absolute (the contract is max-norm relative 1e-5 on O(1) values) in five instructions:
// FMUL, MUFU.EX2, FADD, MUFU.RCP, FFMA.  e = inf -> 1, e = 0 -> -1.
__device__ __forceinline__ float st_tanh(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));      // exp(2x) = 2^(2x log2 e)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

template <int EG>
__global__ void __launch_bounds__(64 + 128 * EG, 1)
semantic_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmWhi,
                       const __grid_constant__ CUtensorMap tmWlo, int64_t n, int P, const float* __restrict__ b,
                       const float* __restrict__ u, int mode, float* __restrict__ out, float* __restrict__ beta,
                       float* __restrict__ vsave, float* __restrict__ scores) {
  constexpr int D = ST_D, A = ST_A;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (st_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - st_smem_u32(smem_raw));
  // [w^T hi kb0 | hi kb1 | lo kb0 | lo kb1] [stage 0: Z hi kb0 | hi kb1 | lo kb0 | lo kb1] [stage 1 ...] [b|u|ss|bts] [bars]
  const uint32_t w_base = base;
  const uint32_t st_base = base + ST_W_BYTES;
  float* par = reinterpret_cast<float*>(gen + ST_W_BYTES + 2 * ST_STAGE_BYTES);   // b[128] | u[128]
  float* ss = par + 2 * A;                                                        // [EG][128] partial scores of the tile
  float* bts = ss + ST_MAX_EG * ST_BM;                                            // [128] beta of the tile
  const uint32_t bars = base + ST_W_BYTES + 2 * ST_STAGE_BYTES + 4 * ST_PAR_FLOATS;
  constexpr int NEPI = 128 * EG;                                                  // splitter / epilogue threads
  const uint32_t w_full = bars, full0 = bars + 8, xform0 = bars + 24, mma0 = bars + 40, free0 = bars + 56;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bars + 80 - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nodes_per_tile = ST_BM / P;
  const int rows_per_tile = nodes_per_tile * P;
  const int64_t total_rows = n * P;
  const int64_t n_tiles = ceil_div64(n, nodes_per_tile);
  const int64_t my_tiles = (n_tiles > (int64_t)blockIdx.x) ? ceil_div64(n_tiles - blockIdx.x, gridDim.x) : 0;

  for (int i = threadIdx.x; i < A; i += 64 + NEPI) {
    par[i] = b[i];
    par[A + i] = u[i];
  }
  if (threadIdx.x == 0) {
    st_mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      st_mbar_init(full0 + 8 * s, 1);
      st_mbar_init(xform0 + 8 * s, NEPI);
      st_mbar_init(mma0 + 8 * s, 1);
      st_mbar_init(free0 + 8 * s, NEPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(st_smem_u32(tmem_slot)),
                 "n"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && my_tiles > 0) {
      st_mbar_expect_tx(w_full, ST_W_BYTES);
      for (int kb = 0; kb < 2; ++kb) {
        st_tma_load_2d(w_base + kb * ST_KB_BYTES, &tmWhi, kb * ST_BK, 0, w_full);
        st_tma_load_2d(w_base + (2 + kb) * ST_KB_BYTES, &tmWlo, kb * ST_BK, 0, w_full);
      }
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int s = (int)(i & 1);
        const uint32_t ph = (uint32_t)((i >> 1) & 1);
        st_mbar_wait(free0 + 8 * s, ph ^ 1);       // the epilogue two tiles back released this slot
        const int64_t tile = blockIdx.x + i * gridDim.x;
        const int row0 = (int)(tile * rows_per_tile);
        const uint32_t dst = st_base + s * ST_STAGE_BYTES;
        st_mbar_expect_tx(full0 + 8 * s, 2 * ST_KB_BYTES);
        st_tma_load_2d(dst, &tmZ, 0, row0, full0 + 8 * s);
        st_tma_load_2d(dst + ST_KB_BYTES, &tmZ, ST_BK, row0, full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && my_tiles > 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(A >> 3) << 17) | ((uint32_t)(ST_BM >> 4) << 24);
      st_mbar_wait(w_full, 0);
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int s = (int)(i & 1);
        const uint32_t ph = (uint32_t)((i >> 1) & 1);
        st_mbar_wait(xform0 + 8 * s, ph);           // hi / lo of this tile are in place (and, transitively,
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");   //  accumulator s was drained)
        const uint32_t zs = st_base + s * ST_STAGE_BYTES;
        const uint32_t acc = tmem_base + (uint32_t)(s * A);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t a_hi = st_desc(zs + kb * ST_KB_BYTES), a_lo = st_desc(zs + (2 + kb) * ST_KB_BYTES);
          const uint64_t b_hi = st_desc(w_base + kb * ST_KB_BYTES), b_lo = st_desc(w_base + (2 + kb) * ST_KB_BYTES);
#pragma unroll
          for (int k = 0; k < ST_BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            st_umma_tf32(acc, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
            st_umma_tf32(acc, a_hi + adv, b_lo + adv, idesc, 1u);
            st_umma_tf32(acc, a_hi + adv, b_hi + adv, idesc, 1u);
          }
        }
        st_umma_commit(mma0 + 8 * s);
      }
    }
  } else {
    // ===== warps 2..: split Z hi/lo (tile i+1), epilogue (tile i) =====
    const int t = threadIdx.x - 64;   // 0 .. NEPI-1
    const int grp = t >> 7;           // column group of this thread's warp group
    const int q = warp & 3;           // TMEM lane quarter this warp may read (warp id % 4)
    auto split_tile = [&](int64_t i) {
      const int s = (int)(i & 1);
      const uint32_t ph = (uint32_t)((i >> 1) & 1);
      st_mbar_wait(full0 + 8 * s, ph);
      uint4* hi = reinterpret_cast<uint4*>(gen + ST_W_BYTES + s * ST_STAGE_BYTES);
      uint4* lo = reinterpret_cast<uint4*>(gen + ST_W_BYTES + s * ST_STAGE_BYTES + 2 * ST_KB_BYTES);
#pragma unroll
      for (int j = 0; j < (int)(2 * ST_KB_BYTES / 16 / NEPI); ++j) {   // 16 / EG x 16 B per thread, layout-agnostic
        const int idx = t + NEPI * j;
        const uint4 x = hi[idx];
        uint4 h, l;
        h.x = x.x & 0xFFFFE000u; l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(h.x));
        h.y = x.y & 0xFFFFE000u; l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(h.y));
        h.z = x.z & 0xFFFFE000u; l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(h.z));
        h.w = x.w & 0xFFFFE000u; l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(h.w));
        hi[idx] = h;
        lo[idx] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      st_mbar_arrive(xform0 + 8 * s);
    };
    if (my_tiles > 0) split_tile(0);
    for (int64_t i = 0; i < my_tiles; ++i) {
      if (i + 1 < my_tiles) split_tile(i + 1);
      const int s = (int)(i & 1);
      const uint32_t ph = (uint32_t)((i >> 1) & 1);
      const int64_t tile = blockIdx.x + i * gridDim.x;
      const int64_t node0 = tile * nodes_per_tile;
      const int64_t row0 = node0 * P;
      const int rows_here = (int)min((int64_t)rows_per_tile, total_rows - row0);
      st_mbar_wait(mma0 + 8 * s, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- v = tanh(acc + b), s = v . u ; thread = row ----
      const int r = q * 32 + lane;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * A);
      float part = 0.f;
      float* vrow = (vsave != nullptr && r < rows_here) ? vsave + (row0 + r) * A : nullptr;
#pragma unroll 1
      for (int c0 = grp * (A / EG); c0 < (grp + 1) * (A / EG); c0 += 32) {
        uint32_t acc[32];
        st_tmem_ld32(lane_addr + c0, acc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 v;
          v.x = st_tanh(__uint_as_float(acc[c]) + par[c0 + c]);
          v.y = st_tanh(__uint_as_float(acc[c + 1]) + par[c0 + c + 1]);
          v.z = st_tanh(__uint_as_float(acc[c + 2]) + par[c0 + c + 2]);
          v.w = st_tanh(__uint_as_float(acc[c + 3]) + par[c0 + c + 3]);
          part = fmaf(v.x, par[A + c0 + c], part);
          part = fmaf(v.y, par[A + c0 + c + 1], part);
          part = fmaf(v.z, par[A + c0 + c + 2], part);
          part = fmaf(v.w, par[A + c0 + c + 3], part);
          if (vrow != nullptr) *reinterpret_cast<float4*>(vrow + c0 + c) = v;
        }
      }
      ss[grp * ST_BM + r] = part;
      st_bar_epi<NEPI>();
      auto score_of = [&](int row) {      // column groups summed in a fixed order
        float sc = ss[row];
#pragma unroll
        for (int g = 1; g < EG; ++g) sc += ss[g * ST_BM + row];
        return sc;
      };
      if (grp == 0 && scores != nullptr && r < rows_here) scores[row0 + r] = score_of(r);
      if (mode == HAN_SEM_REFERENCE) {
        // per-node softmax over the P meta-paths (utils/layers.py:156): thread = row (group 0 only)
        if (grp == 0) {
          float bt = 0.f;
          if (r < rows_here) {
            const int nl = r / P;
            float mx = -INFINITY;
            for (int p = 0; p < P; ++p) mx = fmaxf(mx, score_of(nl * P + p));
            float sum = 0.f;
            for (int p = 0; p < P; ++p) sum += expf(score_of(nl * P + p) - mx);
            bt = expf(score_of(r) - mx) / sum;
            beta[row0 + r] = bt;
          }
          bts[r] = bt;
        }
        st_bar_epi<NEPI>();
        // out[n] = sum_p beta[n,p] Z[n,p] (:159) from the ring: Z = hi + lo exactly
        const uint8_t* zhi = gen + ST_W_BYTES + s * ST_STAGE_BYTES;
        const uint8_t* zlo = zhi + 2 * ST_KB_BYTES;
        const int nodes_here = rows_here / P;
        for (int item = t; item < nodes_here * (D / 4); item += NEPI) {
          const int nl = item / (D / 4), c4 = item % (D / 4);
          const int kb = c4 >> 3, cc = (c4 & 7) * 4;     // K-block, float column inside it
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int p = 0; p < P; ++p) {
            const int rr = nl * P + p;
            const uint32_t off = kb * ST_KB_BYTES + st_sw128(rr, cc);
            const float4 h = *reinterpret_cast<const float4*>(zhi + off);
            const float4 l = *reinterpret_cast<const float4*>(zlo + off);
            const float bp = bts[rr];
            o.x = fmaf(bp, h.x + l.x, o.x);
            o.y = fmaf(bp, h.y + l.y, o.y);
            o.z = fmaf(bp, h.z + l.z, o.z);
            o.w = fmaf(bp, h.w + l.w, o.w);
          }
          *reinterpret_cast<float4*>(out + (node0 + nl) * D + 4 * c4) = o;
        }
      }
      // this thread is done with accumulator s (tcgen05.ld completed above) and with ring slot s
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      st_mbar_arrive(free0 + 8 * s);
      st_bar_epi<NEPI>();      // ss / bts are reused by the next tile
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
  }
}


// =====================================================================================================================
// K-F on tcgen05: backward of the semantic layer (TF autodiff of utils/layers.py:152-159), D = 64, A = 128.
//
//   g[r] = <dout[n], Z[r]>      ds = beta (g - sum_q beta_q g_q)   (per node; paper mode: ds = dsbar[p])
//   v = tanh(Z w + b)           RECOMPUTED here (GEMM1) instead of being stored by the forward and read back:
//                               4.1 GB of writes + 4.1 GB of reads per step on the 2M-node config buy 24 MMAs per tile
//   dv = ds u (1 - v^2)         du += ds v     db += dv
//   dZ = beta dout + dv w^T     (GEMM2)        dw += Z^T dv   (GEMM3, accumulated over the CTA's tiles)
//
// FP32-grade products as BF16x3: every fp32 operand is cut into three bf16 pieces x = x1 + x2 + x3 (8 + 8 + 8 mantissa
// bits, exact), and a b = a1b1 + a1b2 + a2b1 + a1b3 + a2b2 + a3b1 + O(2^-24) accumulates in fp32 in TMEM (six
// kind::f16 MMAs of K = 16 per 16 reduction elements: the same instruction count as 3xTF32 with K = 8).  Why bf16 and
// not tf32 here: each operand is needed in TWO orientations -- Z as the K-major A of GEMM1 (K = d) and the MN-major A
// of GEMM3 (M = d, K = r); w^T as the K-major B of GEMM1 and the MN-major B of GEMM2; dv as the K-major A of GEMM2
// and the MN-major B of GEMM3 -- and for 16-bit types ONE 128-byte-swizzled shared-memory tile serves both (a
// [rows][64] tile read K-major has rows = M/N, read MN-major has rows = K), whereas MN-major tf32 operands need a
// different swizzle (128B with 32-byte atoms), i.e. a second copy of everything, which does not fit.
// Shared memory: w^T 3 x 16 KB, Z 3 x 16 KB, one dv panel (64 columns of a) 3 x 16 KB, raw-Z landing 32 KB.
//
// One persistent CTA per SM walks tiles of 128 (node, meta-path) rows:
//   warp 0    TMA: w^T pieces once; the raw fp32 Z tile of tile i+1 into the landing buffer while tile i is processed
//   warp 1    TMEM allocator + single-thread MMA issuer
//   warps 2-5 (thread = row): cut Z into bf16 pieces, g, ds; per panel: v, dv -> shared memory, du / db column sums
//             (warp transpose-reduce, registers across tiles); dZ epilogue; dw drain
// GEMM3 is M = 64 (d) x N = 64 (a panel) x K = 128 (r); its accumulator (TMEM half-subpartition layout: row m in lane
// (m % 16) + 32 (m / 16)) is drained every second tile with red.global.add into this CTA's partial, because the tensor
// core's truncating accumulate drifts with the chain length (see project_bwd_tc.cu).
constexpr uint32_t SB_T16 = ST_BM * 64 * 2;                         // [128 rows][64 bf16] = 16 KB: one piece of a tile
constexpr uint32_t SB_W_OFF = 0;                                    // w^T pieces 1..3                        48 KB
constexpr uint32_t SB_Z_OFF = SB_W_OFF + 3 * SB_T16;                // Z pieces                               48 KB
constexpr uint32_t SB_DV_OFF = SB_Z_OFF + 3 * SB_T16;               // dv panel pieces                        48 KB
constexpr uint32_t SB_LAND_OFF = SB_DV_OFF + 3 * SB_T16;            // raw fp32 Z of the next tile (2 x 16 KB) 32 KB
constexpr uint32_t SB_PAR_OFF = SB_LAND_OFF + 2 * ST_KB_BYTES;      // b[128] | u[128] | gs[128] | bts[128] | red[4][2][128]
constexpr uint32_t SB_PAR_BYTES = 4 * (4 * 128 + 4 * 2 * 128);
constexpr uint32_t SB_BAR_OFF = SB_PAR_OFF + SB_PAR_BYTES;
constexpr uint32_t SB_SMEM_BYTES = 1024 + SB_BAR_OFF + 256;
constexpr int SB_CHAIN_TILES = 2;                                   // dw accumulation chain: 2 tiles = 256 rows

// MN-major, SWIZZLE_128B descriptor for 16-bit operands: 8 K-rows of 128 B (64 elements along M/N) form one
// swizzle atom, SBO = 1024 B between atoms along K; a single 64-element block along M/N here (LBO unused)
__device__ __forceinline__ uint64_t st_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void st_umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the six partial products of a BF16x3 multiply, smallest terms first; a[i], b[i] = descriptors of piece i
__device__ __forceinline__ void st_mma_bf16x3(uint32_t tmem_d, const uint64_t (&a)[3], const uint64_t (&b)[3], uint32_t idesc,
                                              uint32_t accumulate) {
  st_umma_bf16(tmem_d, a[0], b[2], idesc, accumulate);
  st_umma_bf16(tmem_d, a[2], b[0], idesc, 1u);
  st_umma_bf16(tmem_d, a[1], b[1], idesc, 1u);
  st_umma_bf16(tmem_d, a[0], b[1], idesc, 1u);
  st_umma_bf16(tmem_d, a[1], b[0], idesc, 1u);
  st_umma_bf16(tmem_d, a[0], b[0], idesc, 1u);
}
// x = p1 + p2 + p3 exactly, each piece a bf16 (kept in the top half of a 32-bit word)
__device__ __forceinline__ void st_cut3(float x, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  p1 = __float_as_uint(x) & 0xFFFF0000u;
  const float r1 = x - __uint_as_float(p1);
  p2 = __float_as_uint(r1) & 0xFFFF0000u;
  const float r2 = r1 - __uint_as_float(p2);
  p3 = __float_as_uint(r2) & 0xFFFF0000u;
}
__device__ __forceinline__ uint32_t st_pack(uint32_t lo_elem, uint32_t hi_elem) {   // two bf16 (top halves) -> one word
  return __byte_perm(lo_elem, hi_elem, 0x7632);
}
// eight consecutive fp32 values of row r -> one 16-byte chunk (index c8) of each of the three bf16 piece tiles
// Eight floats -> their three bf16 pieces, one 16-byte chunk per piece tile.  Two floats at a time: the packed piece is
// the two top halves (one PRMT, truncation), the remainder x - piece is exact and taken with one packed FADD2 against
// the sign-flipped truncations ((x & 0xFFFF0000) ^ 0x80000000: one LOP3 each).
__device__ __forceinline__ void st_store8(uint8_t* t1, uint8_t* t2, uint8_t* t3, int r, int c8, const float (&x)[8]) {
  uint32_t k1[4], k2[4], k3[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x0 = __float_as_uint(x[2 * i]), x1 = __float_as_uint(x[2 * i + 1]);
    k1[i] = st_pack(x0, x1);
    const float2 r1 = __fadd2_rn(make_float2(x[2 * i], x[2 * i + 1]),
                                 make_float2(__uint_as_float((x0 & 0xFFFF0000u) ^ 0x80000000u),
                                             __uint_as_float((x1 & 0xFFFF0000u) ^ 0x80000000u)));
    const uint32_t y0 = __float_as_uint(r1.x), y1 = __float_as_uint(r1.y);
    k2[i] = st_pack(y0, y1);
    const float2 r2 = __fadd2_rn(r1, make_float2(__uint_as_float((y0 & 0xFFFF0000u) ^ 0x80000000u),
                                                 __uint_as_float((y1 & 0xFFFF0000u) ^ 0x80000000u)));
    k3[i] = st_pack(__float_as_uint(r2.x), __float_as_uint(r2.y));
  }
  const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4));
  *reinterpret_cast<uint4*>(t1 + off) = make_uint4(k1[0], k1[1], k1[2], k1[3]);
  *reinterpret_cast<uint4*>(t2 + off) = make_uint4(k2[0], k2[1], k2[2], k2[3]);
  *reinterpret_cast<uint4*>(t3 + off) = make_uint4(k3[0], k3[1], k3[2], k3[3]);
}
// Sum over the 32 lanes of each of the N columns held as x[0..N): N = 32 leaves column l in lane l, N = 16 leaves
// column l >> 1 in lanes l (both lanes of a pair hold it).  Reduce-scatter butterfly: N - 1 (+1) shuffles.
template <int N>
__device__ __forceinline__ float st_col_sums(float (&x)[N], int lane) {
  int n = N;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (n > 1) {
      const int h = n >> 1;
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        if (i < h) {
          const float send = up ? x[i] : x[i + h];
          const float keep = up ? x[i + h] : x[i];
          x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      n = h;
    } else {
      x[0] += __shfl_xor_sync(0xffffffffu, x[0], off);
    }
  }
  return x[0];
}

// w [D][A] fp32 -> w^T [A][D] as three bf16 piece matrices
__global__ void st_wt_cut3_kernel(const float* __restrict__ w, uint16_t* __restrict__ w1, uint16_t* __restrict__ w2,
                                  uint16_t* __restrict__ w3) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ST_A * ST_D) return;
  const int a = idx / ST_D, d = idx % ST_D;
  uint32_t p1, p2, p3;
  st_cut3(w[d * ST_A + a], p1, p2, p3);
  w1[idx] = (uint16_t)(p1 >> 16);
  w2[idx] = (uint16_t)(p2 >> 16);
  w3[idx] = (uint16_t)(p3 >> 16);
}

// NH = epilogue threads per tile row (2 or 4): with one warp per scheduler the epilogue is latency-bound (ncu:
// issue slots 24 % busy, 6.3 cycles between issues); NH warps per TMEM lane quarter share each row's columns.
template <int NH>
__global__ void __launch_bounds__(64 + 128 * NH, 1)
semantic_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW1,
                       const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3, int64_t n, int P,
                       const float* __restrict__ dout, const float* __restrict__ beta, const float* __restrict__ b,
                       const float* __restrict__ u, int mode, const float* __restrict__ dsbar, float* __restrict__ dZ,
                       float* const* __restrict__ dz_tab, int64_t dz_stride, float* __restrict__ part) {
  constexpr int D = ST_D, A = ST_A;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (st_smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - st_smem_u32(smem_raw));
  float* par = reinterpret_cast<float*>(gen + SB_PAR_OFF);
  float* bs = par;              // b[128]
  float* us = par + 128;        // u[128]
  float* gs = par + 256;        // g of the tile's rows
  float* bts = par + 384;       // beta of the tile's rows
  float* red = par + 512;       // [4 quarters][du 128 | db 128]   (also: g partials [NH][128] inside a tile)
  constexpr int NE = 128 * NH;  // epilogue threads
  constexpr int CW = 64 / NH;   // panel columns per epilogue thread
  const uint32_t bars = base + SB_BAR_OFF;
  const uint32_t w_full = bars, land_full = bars + 8, land_free = bars + 16, z_ready = bars + 24, acc1_full = bars + 32,
                 acc1_free = bars + 40, dvp_full = bars + 48, dvp_free = bars + 56, dz_full = bars + 64, dz_free = bars + 72,
                 g3_done = bars + 80, dw_full = bars + 88, dw_free = bars + 96;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + SB_BAR_OFF + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nodes_per_tile = ST_BM / P;
  const int rows_per_tile = nodes_per_tile * P;
  const int64_t total_rows = n * P;
  const int64_t n_tiles = ceil_div64(n, nodes_per_tile);
  const int64_t my_tiles = (n_tiles > (int64_t)blockIdx.x) ? ceil_div64(n_tiles - blockIdx.x, gridDim.x) : 0;

  for (int i = threadIdx.x; i < A; i += 64 + NE) {
    bs[i] = b[i];
    us[i] = u[i];
  }
  if (threadIdx.x == 0) {
    st_mbar_init(w_full, 1);
    st_mbar_init(land_full, 1);
    st_mbar_init(land_free, NE);
    st_mbar_init(z_ready, NE);
    st_mbar_init(acc1_full, 1);
    st_mbar_init(acc1_free, NE);
    st_mbar_init(dvp_full, NE);
    st_mbar_init(dvp_free, 1);
    st_mbar_init(dz_full, 1);
    st_mbar_init(dz_free, NE);
    st_mbar_init(g3_done, 1);
    st_mbar_init(dw_full, 1);
    st_mbar_init(dw_free, NE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(st_smem_u32(tmem_slot)),
                 "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_acc1 = tmem_base, t_dz = tmem_base + 128, t_dw = tmem_base + 192;   // columns

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && my_tiles > 0) {
      st_mbar_expect_tx(w_full, 3 * SB_T16);
      st_tma_load_2d(base + SB_W_OFF, &tmW1, 0, 0, w_full);
      st_tma_load_2d(base + SB_W_OFF + SB_T16, &tmW2, 0, 0, w_full);
      st_tma_load_2d(base + SB_W_OFF + 2 * SB_T16, &tmW3, 0, 0, w_full);
      for (int64_t it = 0; it < my_tiles; ++it) {
        if (it > 0) st_mbar_wait(land_free, (uint32_t)((it - 1) & 1));
        const int64_t tile = blockIdx.x + it * gridDim.x;
        const int row0 = (int)(tile * rows_per_tile);
        st_mbar_expect_tx(land_full, 2 * ST_KB_BYTES);
        st_tma_load_2d(base + SB_LAND_OFF, &tmZ, 0, row0, land_full);
        st_tma_load_2d(base + SB_LAND_OFF + ST_KB_BYTES, &tmZ, ST_BK, row0, land_full);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t kBf16 = (1u << 4) | (1u << 7) | (1u << 10);      // D = f32, A = B = bf16
      const uint32_t id1 = kBf16 | ((uint32_t)(A >> 3) << 17) | ((uint32_t)(ST_BM >> 4) << 24);                       // 128 x 128, K | K
      const uint32_t id2 = kBf16 | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(ST_BM >> 4) << 24);           // 128 x 64,  K | MN
      const uint32_t id3 = kBf16 | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24); // 64 x 64,  MN | MN
      const uint32_t wT = base + SB_W_OFF, zt = base + SB_Z_OFF, dvt = base + SB_DV_OFF;
      st_mbar_wait(w_full, 0);
      for (int64_t it = 0; it < my_tiles; ++it) {
        const bool last = it + 1 == my_tiles;
        // ---- GEMM1: acc1 = Z w  (recompute of the pre-activation); K = d = 64 = 4 steps of 16 ----
        st_mbar_wait(z_ready, (uint32_t)(it & 1));
        if (it > 0) st_mbar_wait(acc1_free, (uint32_t)((it - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          const uint64_t a[3] = {st_desc(zt) + adv, st_desc(zt + SB_T16) + adv, st_desc(zt + 2 * SB_T16) + adv};
          const uint64_t bb[3] = {st_desc(wT) + adv, st_desc(wT + SB_T16) + adv, st_desc(wT + 2 * SB_T16) + adv};
          st_mma_bf16x3(t_acc1, a, bb, id1, k != 0);
        }
        st_umma_commit(acc1_full);
        // ---- per dv panel (64 columns of a): GEMM2 (dZ += dv w^T) and GEMM3 (dw += Z^T dv) ----
        for (int j = 0; j < 2; ++j) {
          const int64_t pc = it * 2 + j;
          st_mbar_wait(dvp_full, (uint32_t)(pc & 1));
          if (j == 0 && it > 0) st_mbar_wait(dz_free, (uint32_t)((it - 1) & 1));
          if (j == 0 && (it % SB_CHAIN_TILES) == 0 && it >= SB_CHAIN_TILES)
            st_mbar_wait(dw_free, (uint32_t)(((it / SB_CHAIN_TILES) - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // GEMM2: A = dv panel (K-major, K = 64 a), B = w^T rows a in [64j, 64j+64) read MN-major (N = d = 64)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            const uint32_t wrow = (uint32_t)((64 * j + 16 * k) * 128);
            const uint64_t a[3] = {st_desc(dvt) + adv, st_desc(dvt + SB_T16) + adv, st_desc(dvt + 2 * SB_T16) + adv};
            const uint64_t bb[3] = {st_desc_mn(wT + wrow), st_desc_mn(wT + SB_T16 + wrow), st_desc_mn(wT + 2 * SB_T16 + wrow)};
            st_mma_bf16x3(t_dz, a, bb, id2, (j | k) != 0);
          }
          // GEMM3: A = Z^T (MN-major: M = d = 64, 16 rows of r per step), B = dv panel (MN-major, N = 64 a)
          const uint32_t first = ((it % SB_CHAIN_TILES) == 0) ? 0u : 1u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t koff = (uint32_t)(k * 16 * 128);
            const uint64_t a[3] = {st_desc_mn(zt + koff), st_desc_mn(zt + SB_T16 + koff), st_desc_mn(zt + 2 * SB_T16 + koff)};
            const uint64_t bb[3] = {st_desc_mn(dvt + koff), st_desc_mn(dvt + SB_T16 + koff), st_desc_mn(dvt + 2 * SB_T16 + koff)};
            st_mma_bf16x3(t_dw + (uint32_t)(64 * j), a, bb, id3, (k != 0) ? 1u : first);
          }
          st_umma_commit(dvp_free);
        }
        st_umma_commit(dz_full);
        st_umma_commit(g3_done);
        if ((it % SB_CHAIN_TILES) == SB_CHAIN_TILES - 1 || last) st_umma_commit(dw_full);
      }
    }
  } else {
    // ===== epilogue warps: NH threads per row of the tile (sub = which share of the row's columns) =====
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int sub = (warp - 2) >> 2;            // 0 .. NH-1
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float du_acc[2] = {0.f, 0.f}, db_acc[2] = {0.f, 0.f};      // per panel: column 64 j + CW sub + (lane or lane / 2)
    float* my_part = part + (size_t)blockIdx.x * ((size_t)D * A + 2 * A);
    uint8_t* z1 = gen + SB_Z_OFF;
    uint8_t* dv1 = gen + SB_DV_OFF;
    float* gpart = red;                         // [NH][128] partial g of this tile (red is only needed at the very end)
    for (int64_t it = 0; it < my_tiles; ++it) {
      const bool last = it + 1 == my_tiles;
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int64_t node0 = tile * nodes_per_tile;
      const int64_t row0 = node0 * P;
      const int rows_here = (int)min((int64_t)rows_per_tile, total_rows - row0);
      const bool live = r < rows_here;
      const int64_t node = node0 + r / P;
      const float* drow = dout + node * D;
      // ---- cut the landed raw tile into bf16 pieces (8 / NH chunks of 8 floats per thread), g = <dout, Z> on the way ----
      st_mbar_wait(land_full, (uint32_t)(it & 1));
      if (it > 0) st_mbar_wait(g3_done, (uint32_t)((it - 1) & 1));        // GEMM3 of the previous tile has read the Z pieces
      float g = 0.f;
#pragma unroll
      for (int cc = 0; cc < 8 / NH; ++cc) {     // 8 floats = d in [8 c8, 8 c8 + 8): K-block c8 / 4, 16-byte chunks 2 (c8 % 4), +1
        const int c8 = sub * (8 / NH) + cc;
        float x[8];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t off = (uint32_t)(c8 >> 2) * ST_KB_BYTES + st_sw128(r, 4 * (2 * (c8 & 3) + hh));
          const float4 v4 = *reinterpret_cast<const float4*>(gen + SB_LAND_OFF + off);
          x[4 * hh] = v4.x; x[4 * hh + 1] = v4.y; x[4 * hh + 2] = v4.z; x[4 * hh + 3] = v4.w;
        }
        st_store8(z1, z1 + SB_T16, z1 + 2 * SB_T16, r, c8, x);
        if (live) {
          const float4 d0 = ldg4(drow + 8 * c8), d1 = ldg4(drow + 8 * c8 + 4);
          g = fmaf(x[0], d0.x, g); g = fmaf(x[1], d0.y, g); g = fmaf(x[2], d0.z, g); g = fmaf(x[3], d0.w, g);
          g = fmaf(x[4], d1.x, g); g = fmaf(x[5], d1.y, g); g = fmaf(x[6], d1.z, g); g = fmaf(x[7], d1.w, g);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      st_mbar_arrive(z_ready);
      st_mbar_arrive(land_free);
      const float bt = live ? beta[row0 + r] : 0.f;
      gpart[sub * 128 + r] = g;
      if (sub == 0) bts[r] = bt;
      st_bar_epi<NE>();
      if (sub == 0) {
        float gt = gpart[r];
#pragma unroll
        for (int k = 1; k < NH; ++k) gt += gpart[k * 128 + r];
        gs[r] = gt;
      }
      st_bar_epi<NE>();
      float ds = 0.f;
      if (live) {
        if (mode == HAN_SEM_REFERENCE) {
          const int nl = r / P;
          float dot = 0.f;
          for (int pp = 0; pp < P; ++pp) dot = fmaf(bts[nl * P + pp], gs[nl * P + pp], dot);
          ds = bt * (gs[r] - dot);
        } else {
          ds = dsbar[r % P];
        }
      }
      st_bar_epi<NE>();          // gs / bts / gpart are rewritten by the next tile
      // ---- panels: v = tanh(acc1 + b), dv = ds u (1 - v^2) -> shared memory; du, db column sums ----
      st_mbar_wait(acc1_full, (uint32_t)(it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        const int64_t pc = it * 2 + j;
        const int c0 = 64 * j + CW * sub;
        uint32_t acc[CW];
        st_tmem_ldN<CW>(lane_base + (uint32_t)c0, acc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (j == 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          st_mbar_arrive(acc1_free);
        }
        float y[CW], dv[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const float v = st_tanh(__uint_as_float(acc[c]) + bs[c0 + c]);
          y[c] = ds * v;
          dv[c] = ds * us[c0 + c] * (1.f - v * v);
        }
        if (pc > 0) st_mbar_wait(dvp_free, (uint32_t)((pc - 1) & 1));     // the previous panel's MMAs have read it
#pragma unroll
        for (int c8 = 0; c8 < CW / 8; ++c8) {
          const float x[8] = {dv[8 * c8], dv[8 * c8 + 1], dv[8 * c8 + 2], dv[8 * c8 + 3],
                              dv[8 * c8 + 4], dv[8 * c8 + 5], dv[8 * c8 + 6], dv[8 * c8 + 7]};
          st_store8(dv1, dv1 + SB_T16, dv1 + 2 * SB_T16, r, sub * (CW / 8) + c8, x);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        st_mbar_arrive(dvp_full);
        du_acc[j] += st_col_sums<CW>(y, lane);
        db_acc[j] += st_col_sums<CW>(dv, lane);
      }
      // ---- dZ = beta dout + dv w^T : this thread's 64 / NH columns, staged through shared memory (the dv panel buffer is
      //      idle once dz_full has fired) so that the global stores are whole 256-byte rows: two rows per warp
      //      instruction instead of 32 scattered 16-byte pieces -- tile-sharded, these stores cross NVLink, where a
      //      16-byte write costs a packet of its own (semantic backward at 8 GPUs: 1.44 ms for 1/8 of the rows) ----
      st_mbar_wait(dz_full, (uint32_t)(it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      constexpr int SROW = D + 4;                 // staging row stride in floats: conflict-free 16-byte stores per quarter warp
      float* stg = reinterpret_cast<float*>(gen + SB_DV_OFF);
      {
        uint32_t acc[CW];
        st_tmem_ldN<CW>(lane_base + 128u + (uint32_t)(CW * sub), acc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (live) {
#pragma unroll
          for (int c = 0; c < CW / 4; ++c) {
            const float4 dd = ldg4(drow + CW * sub + 4 * c);
            *reinterpret_cast<float4*>(stg + r * SROW + CW * sub + 4 * c) =
                make_float4(fmaf(bt, dd.x, __uint_as_float(acc[4 * c])), fmaf(bt, dd.y, __uint_as_float(acc[4 * c + 1])),
                            fmaf(bt, dd.z, __uint_as_float(acc[4 * c + 2])), fmaf(bt, dd.w, __uint_as_float(acc[4 * c + 3])));
          }
        }
      }
      st_bar_epi<NE>();
      {
        const int te = threadIdx.x - 64;          // 0 .. NE-1
#pragma unroll 1
        for (int ch = te; ch < rows_here * (D / 4); ch += NE) {
          const int rw = ch / (D / 4), c4 = ch % (D / 4);
          float* dst = (dz_tab != nullptr) ? dz_tab[rw % P] + (node0 + rw / P) * dz_stride : dZ + (row0 + rw) * D;
          *reinterpret_cast<float4*>(dst + 4 * c4) = *reinterpret_cast<const float4*>(stg + rw * SROW + 4 * c4);
        }
      }
      st_bar_epi<NE>();                           // the staging area is the next panel's dv buffer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      st_mbar_arrive(dz_free);
      // ---- every second tile: drain the dw accumulator (rows d = 16 q + lane for lane < 16; 128 / NH columns per thread) ----
      if ((it % SB_CHAIN_TILES) == SB_CHAIN_TILES - 1 || last) {
        st_mbar_wait(dw_full, (uint32_t)((it / SB_CHAIN_TILES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int jj = 0; jj < 4 / NH; ++jj) {
          const int j = sub * (4 / NH) + jj;
          uint32_t acc[32];
          st_tmem_ld32(lane_base + 192u + (uint32_t)(32 * j), acc);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (lane < 16) {
            float* dst = my_part + (size_t)(16 * q + lane) * A + 32 * j;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * c), "f"(__uint_as_float(acc[4 * c])),
                           "f"(__uint_as_float(acc[4 * c + 1])), "f"(__uint_as_float(acc[4 * c + 2])),
                           "f"(__uint_as_float(acc[4 * c + 3]))
                           : "memory");
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        st_mbar_arrive(dw_free);
      }
    }
    // ---- du / db: the four quarters' column sums in a fixed order ----
    st_bar_epi<NE>();
    if (CW == 32 || (lane & 1) == 0) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = 64 * j + CW * sub + (CW == 32 ? lane : (lane >> 1));
        red[q * 256 + col] = du_acc[j];
        red[q * 256 + 128 + col] = db_acc[j];
      }
    }
    st_bar_epi<NE>();
    if (threadIdx.x - 64 < 128) {
      const int t = threadIdx.x - 64;     // 0..127 = column a
      const float du = red[0 * 256 + t] + red[1 * 256 + t] + red[2 * 256 + t] + red[3 * 256 + t];
      const float db = red[0 * 256 + 128 + t] + red[1 * 256 + 128 + t] + red[2 * 256 + 128 + t] + red[3 * 256 + 128 + t];
      my_part[(size_t)D * A + t] = db;
      my_part[(size_t)D * A + A + t] = du;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// per-CTA partials [dw (D*A) | db (A) | du (A)] -> dw, db, du
__global__ void st_bwd_reduce_kernel(const float* __restrict__ part, int nblocks, float* __restrict__ dw,
                                     float* __restrict__ db, float* __restrict__ du) {
  const int64_t cols = (int64_t)ST_D * ST_A + 2 * ST_A;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int k = 0; k < nblocks; ++k) s += part[(int64_t)k * cols + c];
  if (c < (int64_t)ST_D * ST_A) dw[c] = s;
  else if (c < (int64_t)ST_D * ST_A + ST_A) db[c - (int64_t)ST_D * ST_A] = s;
  else du[c - (int64_t)ST_D * ST_A - ST_A] = s;
}

typedef CUresult (*StEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int st_make_map(CUtensorMap* m, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_rows) {
  static StEncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StEncodeTiledFn>(p);
  }
  if (!fn) return fail_arg("han_semantic_fwd_tc", "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {ST_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error), "han_semantic_fwd_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -2;
  }
  return 0;
}

// [rows][64 bf16] row-major, box = 64 x box_rows, 128-byte swizzle
static int st_make_map_bf16(CUtensorMap* m, const uint16_t* ptr, uint64_t rows, uint32_t box_rows) {
  static StEncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StEncodeTiledFn>(p);
  }
  if (!fn) return fail_arg("han_semantic_bwd_tc", "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {64 * sizeof(uint16_t)};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error), "han_semantic_bwd_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -2;
  }
  return 0;
}

}  // namespace han

using namespace han;

extern "C" {

size_t han_semantic_tc_workspace_bytes(void) { return (size_t)2 * ST_A * ST_D * sizeof(float); }

int han_semantic_fwd_tc(const float* Z, int64_t n, int P, int D, int A, const float* w, const float* b,
                        const float* u, int mode, float* out, float* beta, float* vsave, float* scores, void* ws,
                        size_t ws_bytes, int epilogue_groups, han_stream_t stream) {
  HAN_REQUIRE(epilogue_groups == 1 || epilogue_groups == 2 || epilogue_groups == 4, "epilogue_groups in {1, 2, 4}");
  HAN_REQUIRE(Z && w && b && u && ws, "null pointer");
  HAN_REQUIRE(D == ST_D && A == ST_A, "the tensor-core semantic forward is built for D = 64, A = 128");
  HAN_REQUIRE(n > 0 && P > 0 && P <= 64 && n * P < ((int64_t)1 << 31), "n > 0, 1 <= P <= 64, n*P < 2^31");
  HAN_REQUIRE(mode == HAN_SEM_REFERENCE || mode == HAN_SEM_PAPER, "mode");
  HAN_REQUIRE(mode == HAN_SEM_PAPER || (out && beta), "reference mode needs out and beta");
  HAN_REQUIRE(mode == HAN_SEM_REFERENCE || scores, "paper mode needs scores");
  HAN_REQUIRE(ws_bytes >= han_semantic_tc_workspace_bytes(), "workspace too small");
  HAN_REQUIRE(((uintptr_t)Z % 16 == 0) && ((uintptr_t)ws % 16 == 0) && ((uintptr_t)vsave % 16 == 0) &&
              ((uintptr_t)out % 16 == 0), "16-byte alignment");
  cudaStream_t st = as_stream(stream);
  float* wt_hi = reinterpret_cast<float*>(ws);
  float* wt_lo = wt_hi + ST_A * ST_D;
  st_wt_split_kernel<<<(ST_A * ST_D + 255) / 256, 256, 0, st>>>(w, wt_hi, wt_lo);
  CUtensorMap tmZ, tmWhi, tmWlo;
  int rc = st_make_map(&tmZ, Z, ST_D, (uint64_t)(n * P), ST_D, ST_BM);
  if (rc) return rc;
  rc = st_make_map(&tmWhi, wt_hi, ST_D, ST_A, ST_D, ST_A);
  if (rc) return rc;
  rc = st_make_map(&tmWlo, wt_lo, ST_D, ST_A, ST_D, ST_A);
  if (rc) return rc;
  HAN_SMEM_ATTR_ONCE(semantic_fwd_tc_kernel<1>, ST_SMEM_BYTES);
  HAN_SMEM_ATTR_ONCE(semantic_fwd_tc_kernel<2>, ST_SMEM_BYTES);
  HAN_SMEM_ATTR_ONCE(semantic_fwd_tc_kernel<4>, ST_SMEM_BYTES);
  const int64_t n_tiles = ceil_div64(n, ST_BM / P);
  const unsigned grid = (unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
  if (epilogue_groups == 1)
    semantic_fwd_tc_kernel<1><<<grid, 64 + 128, ST_SMEM_BYTES, st>>>(tmZ, tmWhi, tmWlo, n, P, b, u, mode, out, beta,
                                                                   vsave, scores);
  else if (epilogue_groups == 2)
    semantic_fwd_tc_kernel<2><<<grid, 64 + 256, ST_SMEM_BYTES, st>>>(tmZ, tmWhi, tmWlo, n, P, b, u, mode, out, beta,
                                                                   vsave, scores);
  else
    semantic_fwd_tc_kernel<4><<<grid, 64 + 512, ST_SMEM_BYTES, st>>>(tmZ, tmWhi, tmWlo, n, P, b, u, mode, out, beta,
                                                                   vsave, scores);
  return check_launch(__func__);
}

size_t han_semantic_bwd_tc_workspace_bytes(void) {
  return (size_t)3 * ST_A * ST_D * sizeof(uint16_t) + (size_t)kNumSMs * ((size_t)ST_D * ST_A + 2 * ST_A) * sizeof(float);
}

int han_semantic_bwd_tc(const float* dout, const float* Z, const float* beta, int64_t n, int P, int D, int A,
                        const float* w, const float* b, const float* u, int mode, const float* dsbar, float* dZ,
                        float* dw, float* db, float* du, void* ws, size_t ws_bytes, float* const* dz_tab,
                        int64_t dz_stride, han_stream_t stream) {
  HAN_REQUIRE(dout && Z && beta && w && b && u && (dZ || dz_tab) && dw && db && du && ws, "null pointer");
  HAN_REQUIRE(D == ST_D && A == ST_A, "the tensor-core semantic backward is built for D = 64, A = 128");
  HAN_REQUIRE(n > 0 && P > 0 && P <= 64 && n * P < ((int64_t)1 << 31), "n > 0, 1 <= P <= 64, n*P < 2^31");
  HAN_REQUIRE(mode == HAN_SEM_REFERENCE || dsbar, "paper mode needs dsbar");
  HAN_REQUIRE(!dz_tab || (dz_stride >= D && dz_stride % 4 == 0), "dz_stride");
  HAN_REQUIRE(ws_bytes >= han_semantic_bwd_tc_workspace_bytes(), "workspace too small");
  HAN_REQUIRE(((uintptr_t)Z % 16 == 0) && ((uintptr_t)ws % 16 == 0) && ((uintptr_t)dout % 16 == 0) &&
              ((uintptr_t)dZ % 16 == 0), "16-byte alignment");
  cudaStream_t st = as_stream(stream);
  uint16_t* w1 = reinterpret_cast<uint16_t*>(ws);
  uint16_t* w2 = w1 + ST_A * ST_D;
  uint16_t* w3 = w2 + ST_A * ST_D;
  float* part = reinterpret_cast<float*>(w3 + ST_A * ST_D);
  st_wt_cut3_kernel<<<(ST_A * ST_D + 255) / 256, 256, 0, st>>>(w, w1, w2, w3);
  CUtensorMap tmZ, tmW1, tmW2, tmW3;
  int rc = st_make_map(&tmZ, Z, ST_D, (uint64_t)(n * P), ST_D, ST_BM);
  if (rc) return rc;
  rc = st_make_map_bf16(&tmW1, w1, ST_A, ST_A);
  if (rc) return rc;
  rc = st_make_map_bf16(&tmW2, w2, ST_A, ST_A);
  if (rc) return rc;
  rc = st_make_map_bf16(&tmW3, w3, ST_A, ST_A);
  if (rc) return rc;
  HAN_SMEM_ATTR_ONCE(semantic_bwd_tc_kernel<2>, SB_SMEM_BYTES);
  const int64_t n_tiles = ceil_div64(n, ST_BM / P);
  const unsigned grid = (unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
  cudaMemsetAsync(part, 0, (size_t)grid * ((size_t)ST_D * ST_A + 2 * ST_A) * sizeof(float), st);   // the dw drains accumulate
  // NH = 2 epilogue threads per row (4 was measured slower on the 2M config, 5.3 vs 4.8 ms, and removed)
  semantic_bwd_tc_kernel<2><<<grid, 64 + 256, SB_SMEM_BYTES, st>>>(tmZ, tmW1, tmW2, tmW3, n, P, dout, beta, b, u, mode, dsbar,
                                                                 dZ, dz_tab, dz_stride, part);
  const int64_t cols = (int64_t)ST_D * ST_A + 2 * ST_A;
  st_bwd_reduce_kernel<<<(unsigned)ceil_div64(cols, 128), 128, 0, st>>>(part, (int)grid, dw, db, du);
  return check_launch(__func__);
}

}  // extern "C"
