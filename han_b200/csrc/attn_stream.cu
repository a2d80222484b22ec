// Chunked edge-stream versions of the two gather kernels (K-B forward, K-D by-source backward).
//
// Why: a warp-per-row kernel (round 1's first version, removed) holds every gathered row in registers,
// so registers cap both occupancy and the number of rows in flight (ncu on B200: 35% / 21% warps
// active, 67% / 35% DRAM utilisation).  Here a warp owns a contiguous CHUNK of the CSR edge stream
// (whole rows, ~chunk_edges edges, boundaries precomputed per graph by han_csr_chunk_rows) and
// pulls the gathered rows through a per-warp shared-memory ring with cp.async (LDGSTS, 16 B per
// lane, L2-only): STAGES-1 batches of B records are always in flight per warp, independent of
// registers (ring geometry: template parameters, chosen by a sweep on the B200, see gather_cfg below), and work is
// balanced by edges instead of by rows, so degree skew does not matter.
// Row boundaries inside the stream are handled by the consumer (segmented softmax states / sums).
#include "han_common.cuh"
#include "han_rng.cuh"

#include <stdlib.h>

namespace han {

constexpr int kMaxChunkEdges = 2048;   // edges per work item (whole rows; boundaries from chunk_rows)
constexpr int kMinChunkEdges = 128;
constexpr int kBatch = 16;          // records per cp.async group (default; template parameter B of the kernels)
constexpr int kStreamWarps = 4;     // warps per CTA

// attention-coefficient dropout arguments (thr == 0: disabled)
struct DropCoef {
  const uint32_t* seed_ptr;   // device word holding the step seed (a captured CUDA graph can bump it per replay)
  uint32_t thr;               // keep probability * 2^24
  float inv_keep;
  uint32_t metapath;          // stream id
  int64_t row0;               // global id of local row 0 (destination rows forward, source rows backward)
  uint32_t in_thr;            // feature dropout of the gathered / own S rows (utils/layers.py:31-32); 0: disabled
  float in_inv_keep;
};

// f2_j = S_j a2 + b2 (utils/layers.py:24) is recomputed from the table row wherever it is needed
struct Scorer {
  const float* a2;   // [K][H]
  const float* b2;   // [K]
};

// Heavy rows (power-law graphs: a row with 10^5..10^6 edges would keep ONE warp busy for milliseconds) are cut
// into segments of at most `split` edges.  The kernels then walk a VIRTUAL-row CSR (same column array, extra
// offsets): a virtual row that is a whole real row is finished as usual; a segment of a cut row leaves its
// partial state (forward: running max, normaliser, un-normalised aggregate; backward: plain sums) in
// part[slot][K][H+2], and a small merge kernel (one warp per cut row) combines the segments.
struct SplitRows {
  const int2* vmap;   // per virtual row: x = real row, y = partial slot, or -1 when the row is whole
  float* part;        // [n_slots][K][H + 2]
};

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// chunk_rows[c] = first row whose start offset indptr[r] >= c * chunk_edges, c in [0, n_chunks);
// chunk_rows[n_chunks] = n_rows.  n_chunks = nnz / chunk_edges + 1.
__global__ void chunk_rows_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, int64_t n_chunks,
                                  int64_t chunk_edges, int32_t* __restrict__ chunk_rows) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_chunks) return;
  if (c == n_chunks) {
    chunk_rows[c] = (int32_t)n_rows;
    return;
  }
  const int64_t target = indptr[0] + c * chunk_edges;   // indptr may point INTO a larger CSR (row sub-range)
  int64_t lo = 0, hi = n_rows;  // lower_bound over indptr[0 .. n_rows)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (indptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  chunk_rows[c] = (int32_t)lo;
}

// -------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------
// Per warp and batch of B gathered table rows (S_j, D floats each -- f2 is not stored, see project.cu):
//   edge      lane = (slot, head): loads the head's H features of S_j once (2 LDS.128 for H = 8), takes
//             f2_j = S_j a2 + b2 from them (packed FFMA2), the logit, and accumulates exp(e - m) S_j into V and V'
//             with ONE packed FFMA2 per two accumulator elements.  m is a LAZY reference exponent: the state is
//             rescaled only when an edge exceeds it by kLazyExp (first edge of a row, then almost never), so the
//             common path has no rescale multiplies; terms stay below e^kLazyExp, far from overflow, and the
//             reference never exceeds the running maximum, so nothing underflows that the exact form would keep.
//             With feature dropout (utils/layers.py:31-32) the loaded features are masked after f2 was taken: f2
//             sees the un-dropped S (:24) and the aggregate the dropped one (:33), from one table.
//   row end   the SLOTS partial states of each head are brought to the largest reference and reduce-scattered over
//             the slot lanes, so that every lane finishes H / SLOTS outputs (normalise, bias, residual, ELU, stores).
// Edge positions are 32-bit offsets from the chunk's first edge (a chunk is ~2048 edges plus at most one row).
// ncu, 2M-node config: the per-edge online softmax of round 1 issued 211 warp instructions per 4 edges and was
// issue-bound (74 % issue slots busy at 4.9 TB/s); see profiles/ for this version.
constexpr float kLazyExp = 14.f;
// a[0..N) on every slot lane of a head -> the sums over the slot lanes, lane `slot` keeping elements
// [slot * N / SLOTS, (slot + 1) * N / SLOTS) in a[0 .. N / SLOTS): recursive halving over lane bits OFF, OFF/2, .. KK
template <int N, int OFF, int KK>
__device__ __forceinline__ void slot_reduce_scatter(float* a, int lane) {
  if constexpr (OFF >= KK) {
    const bool upper = (lane & OFF) != 0;
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      const float send = upper ? a[j] : a[N / 2 + j];
      const float keep = upper ? a[N / 2 + j] : a[j];
      a[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    slot_reduce_scatter<N / 2, OFF / 2, KK>(a, lane);
  }
}
template <int N>
__device__ __forceinline__ void st_vec(float* p, const float* v) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
      *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else if constexpr (N == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) p[q] = v[q];
  }
}
template <int N>
__device__ __forceinline__ void ld_vec(const float* p, float* v) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 t = ldg4(p + 4 * q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  } else if constexpr (N == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int q = 0; q < N; ++q) v[q] = __ldg(p + q);
  }
}

// PLAIN: the instantiation for 0/1 adjacencies without dropout (no edge weights, no masks): the flags fold away
template <int K, int H, int STAGES, int B, int MINB, bool SPLIT, bool PLAIN>
__global__ void __launch_bounds__(kStreamWarps * 32, MINB)
attn_fwd_chunked_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                        const int32_t* __restrict__ chunk_rows, int64_t n_chunks,
                        const float* __restrict__ T, Scorer f2w, float* __restrict__ R, const float* __restrict__ bias,
                        int act, float* __restrict__ out, int64_t out_stride, float* __restrict__ vsave,
                        const float* __restrict__ colmean, const float* __restrict__ ew,
                        const float* __restrict__ resid, int64_t resid_stride, float* const* __restrict__ out2_tab,
                        int64_t out2_block_rows, int64_t out2_stride, float* __restrict__ vsave2,
                        float* __restrict__ csave, DropCoef dc, SplitRows sp) {
  constexpr int D = K * H;
  constexpr int TS = D;
  constexpr int RS = ((D + 3 * K + 3) / 4) * 4;
  constexpr int SLOTS = 32 / K;
  constexpr int HV = H / 4;
  constexpr int REC_CHUNKS = TS / 4;               // 16-byte pieces of a table row: a power of two <= 32
  constexpr int RPI = 32 / REC_CHUNKS;             // rows fetched by one cp.async of the whole warp
  constexpr int PASSES = (B + RPI - 1) / RPI;
  constexpr int H2 = H / 2;
  constexpr bool RSF = SLOTS <= H;                 // row end: reduce-scatter over the slot lanes (else slot 0 finishes)
  constexpr int HF = RSF ? H / SLOTS : H;          // outputs a lane finishes
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int head = lane % K, slot = lane / K;
  float* ring = smem + (size_t)w * STAGES * B * TS;
  int* col_s = reinterpret_cast<int*>(smem + (size_t)kStreamWarps * STAGES * B * TS) + w * STAGES * B;
  // sp_attn_head (utils/layers.py:95-96): stored adjacency value w_ij scales the logit; staged like the columns.
  // The region exists only when ew != nullptr (the launcher sizes the dynamic shared memory accordingly).
  float* w_s = reinterpret_cast<float*>(col_s - w * STAGES * B + kStreamWarps * STAGES * B) + w * STAGES * B;
  // attention-coefficient dropout (utils/layers.py:29-30): mask bit from (seed, dst, src, head);
  // feature dropout (:31-32): mask bit from (seed, src, column)
  const uint32_t c_thr = PLAIN ? 0u : dc.thr, in_thr = PLAIN ? 0u : dc.in_thr;
  const uint32_t cseed = c_thr ? stream_seed(*dc.seed_ptr, 3u, dc.metapath, (uint32_t)head) : 0u;
  const uint32_t sseed = in_thr ? stream_seed(*dc.seed_ptr, 2u, dc.metapath, 0u) : 0u;

  const int64_t chunk = (int64_t)blockIdx.x * kStreamWarps + w;
  if (chunk >= n_chunks) return;
  const int r_lo = chunk_rows[chunk], r_hi = chunk_rows[chunk + 1];
  if (r_lo >= r_hi) return;
  const int64_t e_lo = indptr[r_lo];
  const int ne = (int)(indptr[r_hi] - e_lo);
  const int nb = (ne + B - 1) / B;
  const int32_t* __restrict__ idx = indices + e_lo;
  const float* __restrict__ ewp = (!PLAIN && ew) ? ew + e_lo : nullptr;

  int q = 0;  // next batch to issue
  int col_pref = (lane < B && lane < ne) ? ldg_stream_i32(idx + lane) : 0;
  float w_pref = (ewp && lane < B && lane < ne) ? __ldg(ewp + lane) : 1.f;
  const int rsub = lane / REC_CHUNKS, off4 = (lane % REC_CHUNKS) * 4;   // this lane's row within a pass, float offset
  const float* tsrc = T + off4;
  auto issue = [&]() {
    if (q < nb) {
      const int bs = q * B;
      const int cnt = min(B, ne - bs);
      float* dst = ring + (q % STAGES) * (B * TS) + rsub * TS + off4;
      if (B % RPI == 0 && cnt == B) {
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
          const int col = __shfl_sync(0xffffffffu, col_pref, i * RPI + rsub);
          cp_async16(dst + i * RPI * TS, tsrc + (int64_t)col * TS);
        }
      } else {
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
          const int rec = i * RPI + rsub;
          const int col = __shfl_sync(0xffffffffu, col_pref, rec & 31);
          if (rec < cnt) cp_async16(dst + i * RPI * TS, tsrc + (int64_t)col * TS);
        }
      }
      if (!PLAIN && lane < B) col_s[(q % STAGES) * B + lane] = col_pref;      // source ids: needed by the masks only
      if (ewp && lane < B) w_s[(q % STAGES) * B + lane] = w_pref;
      const int nbs = bs + B;
      col_pref = (lane < B && nbs + lane < ne) ? ldg_stream_i32(idx + nbs + lane) : 0;
      if (ewp) w_pref = (lane < B && nbs + lane < ne) ? __ldg(ewp + nbs + lane) : 1.f;
    }
    cp_async_commit();
    ++q;
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue();

  int row = r_lo;     // row of the (virtual) CSR being streamed; rr = the real row it belongs to
  int rr = row, pslot = -1;
  if constexpr (SPLIT) {
    const int2 v = __ldg(sp.vmap + row);
    rr = v.x;
    pslot = v.y;
  }
  int row_start = 0;
  int row_end = (int)(indptr[row + 1] - e_lo);
  const float b2v = __ldg(f2w.b2 + head);
  float2 a2p[H2];                                   // this head's a2, for f2_j = S_j a2 + b2
#pragma unroll
  for (int i = 0; i < H2; ++i) a2p[i] = __ldg(reinterpret_cast<const float2*>(f2w.a2 + head * H) + i);
  float f1v = R[(int64_t)rr * RS + D + head] + b2v;   // f1_i + b2: the logit is f1v + S_j a2
  float m = -INFINITY, l = 0.f;
  float2 acc_pk[H2];
#pragma unroll
  for (int i = 0; i < H2; ++i) acc_pk[i] = make_float2(0.f, 0.f);
  // Second aggregate, kept for the backward: with k_ij = leaky'(l_ij) (* w_ij for sp_attn_head),
  //   V'_i = sum_j alpha~_ij k_ij S_j   and   c_i = sum_j alpha_ij k_ij
  // make df1_i = sum_j dl_ij = <dV_i, V'_i> - delta_i c_i a ROW-LOCAL quantity: the backward needs neither a
  // per-edge dl array nor a by-destination pass (nor, sharded, a reduce-scatter).
  const bool train = vsave2 != nullptr;
  float2 acc2p[H2];
  float cacc = 0.f;
#pragma unroll
  for (int i = 0; i < H2; ++i) acc2p[i] = make_float2(0.f, 0.f);

  auto finalize_row = [&]() {
    // bring the SLOTS partial states of each head to the largest reference exponent, then plain sums
    float mrow = m;
#pragma unroll
    for (int off = K; off < 32; off <<= 1) mrow = fmaxf(mrow, __shfl_xor_sync(0xffffffffu, mrow, off));
    const float sc = (m == mrow) ? 1.f : fast_exp(m - mrow);    // m = -inf (no edge on this lane): 0, or 1 if the row is empty
    l *= sc;
    cacc *= sc;
    float acc_[H], acc2[H];
    float* acc = acc_;     // (the packed accumulators are reset by next_row)
#pragma unroll
    for (int i = 0; i < H2; ++i) {
      acc[2 * i] = acc_pk[i].x * sc;
      acc[2 * i + 1] = acc_pk[i].y * sc;
      acc2[2 * i] = acc2p[i].x * sc;
      acc2[2 * i + 1] = acc2p[i].y * sc;
    }
#pragma unroll
    for (int off = K; off < 32; off <<= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, off);
      if (train) cacc += __shfl_xor_sync(0xffffffffu, cacc, off);
    }
    if constexpr (RSF) {
      slot_reduce_scatter<H, 16, K>(acc, lane);
      if (train) slot_reduce_scatter<H, 16, K>(acc2, lane);
    } else {
#pragma unroll
      for (int off = K; off < 32; off <<= 1) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], off);
          if (train) acc2[h] += __shfl_xor_sync(0xffffffffu, acc2[h], off);
        }
      }
    }
    // this lane now holds columns [c0, c0 + HF) of the row (of its head: h0 .. h0 + HF)
    const int h0 = RSF ? slot * HF : 0;
    const bool storer = RSF || slot == 0;
    const int c0 = head * H + h0;
    if constexpr (SPLIT) {
      if (pslot >= 0) {   // a segment of a cut row: leave (max, normaliser, un-normalised aggregates) for the merge
        float* pp = sp.part + ((int64_t)pslot * K + head) * (2 * H + 3);
        if (slot == 0) {
          pp[0] = mrow;
          pp[1] = l;
          pp[2] = cacc;
        }
        if (storer) {
#pragma unroll
          for (int j = 0; j < HF; ++j) {
            pp[3 + h0 + j] = acc[j];
            pp[3 + H + h0 + j] = acc2[j];
          }
        }
        return;
      }
    }
    float lse = 0.f;
    float fa[HF];
    if (row_end > row_start) {
      const float rinv = 1.f / l;
      lse = mrow + __logf(l);
#pragma unroll
      for (int j = 0; j < HF; ++j) fa[j] = acc[j] * rinv;
      if (train) {
        float f2v[HF];
#pragma unroll
        for (int j = 0; j < HF; ++j) f2v[j] = acc2[j] * rinv;
        if (storer) st_vec<HF>(vsave2 + (int64_t)rr * D + c0, f2v);
        if (slot == 0) csave[(int64_t)rr * K + head] = cacc * rinv;
      }
    } else {
      // row without any edge: dense path = uniform 1/N over ALL nodes; its scores have no effect
      if (colmean) ld_vec<HF>(colmean + c0, fa);
      else {
#pragma unroll
        for (int j = 0; j < HF; ++j) fa[j] = 0.f;
      }
      if (train) {
        float z[HF];
#pragma unroll
        for (int j = 0; j < HF; ++j) z[j] = 0.f;
        if (storer) st_vec<HF>(vsave2 + (int64_t)rr * D + c0, z);
        if (slot == 0) csave[(int64_t)rr * K + head] = 0.f;
      }
    }
    if (slot == 0) R[(int64_t)rr * RS + D + K + head] = lse;
    if (storer) {
      st_vec<HF>(vsave + (int64_t)rr * D + c0, fa);
      float bv[HF];
      ld_vec<HF>(bias + c0, bv);
#pragma unroll
      for (int j = 0; j < HF; ++j) fa[j] += bv[j];
      if (resid) {   // residual branch, utils/layers.py:38-40: ret + conv1d(seq, H, 1), then the activation
        ld_vec<HF>(resid + (int64_t)rr * resid_stride + c0, bv);
#pragma unroll
        for (int j = 0; j < HF; ++j) fa[j] += bv[j];
      }
      if (act == HAN_ACT_ELU) {
#pragma unroll
        for (int j = 0; j < HF; ++j) fa[j] = fa[j] > 0.f ? fa[j] : expm1f(fa[j]);
      }
      st_vec<HF>(out + (int64_t)rr * out_stride + c0, fa);
      // second destination of the same row: a peer GPU's semantic-layer input over NVLink (tile sharding) -- the
      // all-to-all of Z is this store, fire-and-forget, overlapped with the rest of the gather.  Rows
      // [m * out2_block_rows, (m+1) * out2_block_rows) belong to the rank whose buffer is out2_tab[m].
      if (out2_tab) {
        const int64_t mb = rr / out2_block_rows;
        st_vec<HF>(out2_tab[mb] + ((int64_t)rr - mb * out2_block_rows) * out2_stride + c0, fa);
      }
    }
  };
  auto next_row = [&]() {
    ++row;
    row_start = row_end;
    if (row < r_hi) {
      row_end = (int)(indptr[row + 1] - e_lo);
      rr = row;
      if constexpr (SPLIT) {
        const int2 v = __ldg(sp.vmap + row);
        rr = v.x;
        pslot = v.y;
      }
      f1v = R[(int64_t)rr * RS + D + head] + b2v;
    }
    m = -INFINITY;
    l = 0.f;
    cacc = 0.f;
#pragma unroll
    for (int i = 0; i < H2; ++i) {
      acc_pk[i] = make_float2(0.f, 0.f);
      acc2p[i] = make_float2(0.f, 0.f);
    }
  };

  int pos = 0;
  for (int b = 0; b < nb; ++b) {
    issue();
    cp_async_wait<STAGES - 1>();
    __syncwarp();
    const float* buf = ring + (b % STAGES) * (B * TS) + head * H;
    const int* cbuf = col_s + (b % STAGES) * B;
    const float* wbuf = w_s + (b % STAGES) * B;
    const int bs = b * B;
    const int be = min(ne, bs + B);
    while (pos < be) {
      const int seg_end = min(row_end, be);
      const int lim = seg_end - bs;
      for (int r0 = pos - bs; r0 < lim; r0 += SLOTS) {      // warp-uniform trip count
        const int rec = r0 + slot;
        if (rec < lim) {
          const float* rp = buf + rec * TS;
          float2 v[H2];
#pragma unroll
          for (int qv = 0; qv < HV; ++qv) {
            const float4 t4 = *reinterpret_cast<const float4*>(rp + 4 * qv);
            v[2 * qv] = make_float2(t4.x, t4.y);
            v[2 * qv + 1] = make_float2(t4.z, t4.w);
          }
          float2 d2 = __fmul2_rn(v[0], a2p[0]);
#pragma unroll
          for (int i = 1; i < H2; ++i) d2 = __ffma2_rn(v[i], a2p[i], d2);
          float lg = f1v + (d2.x + d2.y);            // f1_i + f2_j
          float kfac = 1.f;                          // d l_ij / d (f1_i + f2_j): leaky'(l_ij) (* w_ij)
          if (ewp) {
            kfac = wbuf[rec];
            lg *= kfac;
          }
          if (lg <= 0.f) kfac *= kLeakySlope;
          const float e = leaky(lg);
          if (e > m + kLazyExp) {                    // first edge of the row (m = -inf), then almost never
            const float sc = fast_exp(m - e);          // -inf -> 0
            const float2 sc2 = make_float2(sc, sc);
            l *= sc;
            cacc *= sc;
#pragma unroll
            for (int i = 0; i < H2; ++i) {
              acc_pk[i] = __fmul2_rn(acc_pk[i], sc2);
              acc2p[i] = __fmul2_rn(acc2p[i], sc2);
            }
            m = e;
          }
          const float p = fast_exp(e - m);
          l += p;                    // the softmax normaliser always sees every neighbour
          float pk = p;              // ... the aggregate only the kept ones, scaled 1/keep (no re-normalisation)
          if (c_thr) pk = keep24(cseed, (uint32_t)(rr + dc.row0), (uint32_t)cbuf[rec], c_thr) ? p * dc.inv_keep : 0.f;
          if (in_thr) {              // feature dropout of S_j (after f2 was taken from the un-dropped row)
            const uint32_t node = (uint32_t)cbuf[rec];
#pragma unroll
            for (int i = 0; i < H2; ++i) {
              const uint32_t d = (uint32_t)(head * H + 2 * i);
              v[i].x = keep24(sseed, node, d, in_thr) ? v[i].x * dc.in_inv_keep : 0.f;
              v[i].y = keep24(sseed, node, d + 1u, in_thr) ? v[i].y * dc.in_inv_keep : 0.f;
            }
          }
          const float2 pkk = make_float2(pk, pk);
#pragma unroll
          for (int i = 0; i < H2; ++i) acc_pk[i] = __ffma2_rn(pkk, v[i], acc_pk[i]);
          if (train) {
            const float pk2 = pk * kfac;
            const float2 pkk2 = make_float2(pk2, pk2);
#pragma unroll
            for (int i = 0; i < H2; ++i) acc2p[i] = __ffma2_rn(pkk2, v[i], acc2p[i]);
            cacc = fmaf(p, kfac, cacc);
          }
        }
      }
      pos = seg_end;
      while (row < r_hi && pos == row_end) {   // row complete (also sweeps up empty rows)
        finalize_row();
        next_row();
      }
    }
    __syncwarp();   // all lanes done with this stage before it is refilled
  }
  while (row < r_hi) {   // trailing rows without edges
    finalize_row();
    next_row();
  }
  cp_async_wait<0>();
}

// -------------------------------------------------------------------------------------------------
// backward, by source (transposed structure)
// -------------------------------------------------------------------------------------------------
template <int K, int H, int STAGES, int B, int MINB, bool SPLIT>
__global__ void __launch_bounds__(kStreamWarps * 32, MINB)
attn_bwd_src_chunked_kernel(const int64_t* __restrict__ t_indptr, const int32_t* __restrict__ t_indices,
                            const int32_t* __restrict__ chunk_rows,
                            int64_t n_chunks, const float* __restrict__ Tsrc, Scorer f2w,
                            const float* __restrict__ R, float* __restrict__ dS_agg, float* __restrict__ df2,
                            const float* __restrict__ ew_t, DropCoef dc, SplitRows sp) {
  constexpr int D = K * H;
  constexpr int TS = D;
  constexpr int RS = ((D + 3 * K + 3) / 4) * 4;
  constexpr int SLOTS = 32 / K;
  constexpr int HV = H / 4;
  constexpr int REC_CHUNKS = RS / 4;
  constexpr int TOT = B * REC_CHUNKS;
  constexpr int PER_LANE = (TOT + 31) / 32;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int head = lane % K, slot = lane / K;
  float* ring = smem + (size_t)w * STAGES * B * RS;
  int* row_s = reinterpret_cast<int*>(smem + (size_t)kStreamWarps * STAGES * B * RS) + w * STAGES * B;
  // edge weights in transposed-edge order (sp_attn_head); region present only when ew_t != nullptr
  float* w_s = reinterpret_cast<float*>(row_s - w * STAGES * B + kStreamWarps * STAGES * B) + w * STAGES * B;
  const uint32_t cseed = dc.thr ? stream_seed(*dc.seed_ptr, 3u, dc.metapath, (uint32_t)head) : 0u;
  const uint32_t sseed = dc.in_thr ? stream_seed(*dc.seed_ptr, 2u, dc.metapath, 0u) : 0u;

  const int64_t chunk = (int64_t)blockIdx.x * kStreamWarps + w;
  if (chunk >= n_chunks) return;
  const int r_lo = chunk_rows[chunk], r_hi = chunk_rows[chunk + 1];
  if (r_lo >= r_hi) return;
  const int64_t e_lo = t_indptr[r_lo], e_hi = t_indptr[r_hi];
  const int nb = (int)((e_hi - e_lo + B - 1) / B);

  int q = 0;
  int row_pref = 0;
  float w_pref = 1.f;
  if (lane < B && e_lo + lane < e_hi) {
    row_pref = ldg_stream_i32(t_indices + e_lo + lane);
    if (ew_t) w_pref = __ldg(ew_t + e_lo + lane);
  }
  auto issue = [&]() {
    if (q < nb) {
      const int64_t bs = e_lo + (int64_t)q * B;
      const int cnt = (int)min((int64_t)B, e_hi - bs);
      const int st = q % STAGES;
      float* dst = ring + (size_t)st * B * RS;
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i) {
        const int c = lane + 32 * i;
        const int rec = c / REC_CHUNKS, off = c - rec * REC_CHUNKS;
        const int i_row = __shfl_sync(0xffffffffu, row_pref, rec & 31);
        if (c < TOT && rec < cnt) cp_async16(dst + rec * RS + off * 4, R + (int64_t)i_row * RS + off * 4);
      }
      if (lane < B) {
        row_s[st * B + lane] = row_pref;
        if (ew_t) w_s[st * B + lane] = w_pref;
      }
      const int64_t nbs = bs + B;
      if (lane < B && nbs + lane < e_hi) {
        row_pref = ldg_stream_i32(t_indices + nbs + lane);
        if (ew_t) w_pref = __ldg(ew_t + nbs + lane);
      }
    }
    cp_async_commit();
    ++q;
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue();

  int row = r_lo;
  int64_t row_end = t_indptr[row + 1];
  const float b2v = __ldg(f2w.b2 + head);
  float sj[H], f2;
  auto load_src = [&](int r) {
#pragma unroll
    for (int qv = 0; qv < HV; ++qv) {
      const float4 s4 = ldg4(Tsrc + (int64_t)r * TS + head * H + 4 * qv);
      sj[4 * qv] = s4.x; sj[4 * qv + 1] = s4.y; sj[4 * qv + 2] = s4.z; sj[4 * qv + 3] = s4.w;
    }
    // f2_j = S_j a2 + b2 from the UN-dropped row (utils/layers.py:24); the aggregate saw the dropped one (:31-33).
    // f2 here = b2 + S_j a2 is added to f1 in the forward's order: (f1 + b2) + S_j a2
    float a2v[H];
#pragma unroll
    for (int qv = 0; qv < HV; ++qv) {
      const float4 x = ldg4(f2w.a2 + head * H + 4 * qv);
      a2v[4 * qv] = x.x; a2v[4 * qv + 1] = x.y; a2v[4 * qv + 2] = x.z; a2v[4 * qv + 3] = x.w;
    }
    f2 = score_dot<H>(sj, a2v);
    if (dc.in_thr) {
      const uint32_t node = (uint32_t)(r + dc.row0);
#pragma unroll
      for (int h = 0; h < H; ++h)
        sj[h] = keep24(sseed, node, (uint32_t)(head * H + h), dc.in_thr) ? sj[h] * dc.in_inv_keep : 0.f;
    }
  };
  int rr = row, pslot = -1;
  if constexpr (SPLIT) {
    const int2 v = __ldg(sp.vmap + row);
    rr = v.x;
    pslot = v.y;
  }
  load_src(rr);
  float acc[H];
#pragma unroll
  for (int h = 0; h < H; ++h) acc[h] = 0.f;
  float df2acc = 0.f;

  auto finalize_row = [&]() {
#pragma unroll
    for (int off = K; off < 32; off <<= 1) {
      df2acc += __shfl_xor_sync(0xffffffffu, df2acc, off);
#pragma unroll
      for (int h = 0; h < H; ++h) acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], off);
    }
    if constexpr (SPLIT) {
      if (pslot >= 0) {   // a segment of a cut row: partial sums for the merge
        if (slot == 0) {
          float* pp = sp.part + ((int64_t)pslot * K + head) * (H + 2);
          pp[0] = df2acc;
#pragma unroll
          for (int h = 0; h < H; ++h) pp[2 + h] = acc[h];
        }
        return;
      }
    }
    if (slot == 0) {
      df2[(int64_t)rr * K + head] = df2acc;
#pragma unroll
      for (int qv = 0; qv < HV; ++qv)
        *reinterpret_cast<float4*>(dS_agg + (int64_t)rr * D + head * H + 4 * qv) =
            make_float4(acc[4 * qv], acc[4 * qv + 1], acc[4 * qv + 2], acc[4 * qv + 3]);
    }
  };
  auto next_row = [&]() {
    ++row;
    if (row < r_hi) {
      row_end = t_indptr[row + 1];
      rr = row;
      if constexpr (SPLIT) {
        const int2 v = __ldg(sp.vmap + row);
        rr = v.x;
        pslot = v.y;
      }
      load_src(rr);
    }
    df2acc = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] = 0.f;
  };

  int64_t pos = e_lo;
  for (int b = 0; b < nb; ++b) {
    issue();
    cp_async_wait<STAGES - 1>();
    __syncwarp();
    const int st = b % STAGES;
    const float* buf = ring + (size_t)st * B * RS;
    const int* rbuf = row_s + st * B;
    const float* wbuf = w_s + st * B;
    const int64_t bs = e_lo + (int64_t)b * B;
    const int64_t be = min(e_hi, bs + B);
    while (pos < be) {
      const int64_t seg_end = min(row_end, be);
      for (int64_t g = pos; g < seg_end; g += SLOTS) {
        const int64_t ei = g + slot;
        if (ei < seg_end) {
          const int rec = (int)(ei - bs);
          const float* rp = buf + rec * RS;
          const float wt = ew_t ? wbuf[rec] : 1.f;     // l_ij = w_ij (f1_i + f2_j): d l / d f1 = d l / d f2 = w_ij
          const float lg = ((rp[D + head] + b2v) + f2) * wt;
          const float a = __expf(leaky(lg) - rp[D + K + head]);
          // coefficient dropout: alpha~ = alpha * m / keep feeds the aggregate; d alpha = d alpha~ * m / keep
          float mk = 1.f;
          if (dc.thr)
            mk = keep24(cseed, (uint32_t)rbuf[rec], (uint32_t)(rr + dc.row0), dc.thr) ? dc.inv_keep : 0.f;
          const float am = a * mk;
          float da = 0.f;
#pragma unroll
          for (int qv = 0; qv < HV; ++qv) {
            const float4 g4 = *reinterpret_cast<const float4*>(rp + head * H + 4 * qv);
            da = fmaf(g4.x, sj[4 * qv], da);
            da = fmaf(g4.y, sj[4 * qv + 1], da);
            da = fmaf(g4.z, sj[4 * qv + 2], da);
            da = fmaf(g4.w, sj[4 * qv + 3], da);
            acc[4 * qv] = fmaf(am, g4.x, acc[4 * qv]);
            acc[4 * qv + 1] = fmaf(am, g4.y, acc[4 * qv + 1]);
            acc[4 * qv + 2] = fmaf(am, g4.z, acc[4 * qv + 2]);
            acc[4 * qv + 3] = fmaf(am, g4.w, acc[4 * qv + 3]);
          }
          // dl_ij feeds df2_j only: df1_i = sum_j dl_ij is row-local since the forward kept V' and c (prep kernel)
          df2acc += a * (da * mk - rp[D + 2 * K + head]) * (lg > 0.f ? 1.f : kLeakySlope) * wt;
        }
      }
      pos = seg_end;
      while (row < r_hi && pos == row_end) {
        finalize_row();
        next_row();
      }
    }
    __syncwarp();
  }
  while (row < r_hi) {
    finalize_row();
    next_row();
  }
  cp_async_wait<0>();
}

// one warp per cut row: combine its segments' partial states, then the forward's row epilogue
template <int K, int H>
__global__ void __launch_bounds__(128)
attn_fwd_merge_kernel(const int32_t* __restrict__ heavy_rows, const int32_t* __restrict__ heavy_ptr, int n_heavy,
                      const float* __restrict__ part, float* __restrict__ R, const float* __restrict__ bias, int act,
                      float* __restrict__ out, int64_t out_stride, float* __restrict__ vsave,
                      const float* __restrict__ resid, int64_t resid_stride, float* const* __restrict__ out2_tab,
                      int64_t out2_block_rows, int64_t out2_stride, float* __restrict__ vsave2,
                      float* __restrict__ csave) {
  constexpr int D = K * H;
  constexpr int RS = ((D + 3 * K + 3) / 4) * 4;
  constexpr int SLOTS = 32 / K;
  const int hw = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (hw >= n_heavy) return;
  const int lane = threadIdx.x & 31, head = lane % K, slot = lane / K;
  const int rr = heavy_rows[hw];
  const int s_lo = heavy_ptr[hw], s_hi = heavy_ptr[hw + 1];
  float m = -INFINITY, l = 0.f, cc = 0.f, acc[H], acc2[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    acc[h] = 0.f;
    acc2[h] = 0.f;
  }
  auto combine = [&](float mo, float lo, float co, const float* ao, const float* a2o) {
    const float mn = fmaxf(m, mo);
    const float s0 = (m == -INFINITY) ? 0.f : __expf(m - mn);
    const float s1 = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
    l = l * s0 + lo * s1;
    cc = cc * s0 + co * s1;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      acc[h] = acc[h] * s0 + ao[h] * s1;
      acc2[h] = acc2[h] * s0 + a2o[h] * s1;
    }
    m = mn;
  };
  for (int sgm = s_lo + slot; sgm < s_hi; sgm += SLOTS) {
    const float* pp = part + ((int64_t)sgm * K + head) * (2 * H + 3);
    float ao[H], a2o[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      ao[h] = pp[3 + h];
      a2o[h] = pp[3 + H + h];
    }
    combine(pp[0], pp[1], pp[2], ao, a2o);
  }
#pragma unroll
  for (int off = K; off < 32; off <<= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, m, off);
    const float lo = __shfl_xor_sync(0xffffffffu, l, off);
    const float co = __shfl_xor_sync(0xffffffffu, cc, off);
    float ao[H], a2o[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      ao[h] = __shfl_xor_sync(0xffffffffu, acc[h], off);
      a2o[h] = __shfl_xor_sync(0xffffffffu, acc2[h], off);
    }
    combine(mo, lo, co, ao, a2o);
  }
  if (slot == 0) {
    const float rinv = 1.f / l;
    R[(int64_t)rr * RS + D + K + head] = m + __logf(l);
    if (vsave2 != nullptr) {
      csave[(int64_t)rr * K + head] = cc * rinv;
#pragma unroll
      for (int h = 0; h < H; ++h) vsave2[(int64_t)rr * D + head * H + h] = acc2[h] * rinv;
    }
    float* vp = vsave + (int64_t)rr * D + head * H;
    float* op = out + (int64_t)rr * out_stride + head * H;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float a = acc[h] * rinv;
      vp[h] = a;
      float z = a + bias[head * H + h];
      if (resid) z += resid[(int64_t)rr * resid_stride + head * H + h];
      const float o = (act == HAN_ACT_ELU && z <= 0.f) ? expm1f(z) : z;
      op[h] = o;
      if (out2_tab) {
        const int64_t m = rr / out2_block_rows;
        out2_tab[m][((int64_t)rr - m * out2_block_rows) * out2_stride + head * H + h] = o;
      }
    }
  }
}

template <int K, int H>
__global__ void __launch_bounds__(128)
attn_bwd_src_merge_kernel(const int32_t* __restrict__ heavy_rows, const int32_t* __restrict__ heavy_ptr, int n_heavy,
                          const float* __restrict__ part, float* __restrict__ dS_agg, float* __restrict__ df2) {
  constexpr int D = K * H;
  constexpr int SLOTS = 32 / K;
  const int hw = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (hw >= n_heavy) return;
  const int lane = threadIdx.x & 31, head = lane % K, slot = lane / K;
  const int rr = heavy_rows[hw];
  const int s_lo = heavy_ptr[hw], s_hi = heavy_ptr[hw + 1];
  float d2 = 0.f, acc[H];
#pragma unroll
  for (int h = 0; h < H; ++h) acc[h] = 0.f;
  for (int sgm = s_lo + slot; sgm < s_hi; sgm += SLOTS) {
    const float* pp = part + ((int64_t)sgm * K + head) * (H + 2);
    d2 += pp[0];
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] += pp[2 + h];
  }
#pragma unroll
  for (int off = K; off < 32; off <<= 1) {
    d2 += __shfl_xor_sync(0xffffffffu, d2, off);
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], off);
  }
  if (slot == 0) {
    df2[(int64_t)rr * K + head] = d2;
#pragma unroll
    for (int h = 0; h < H; ++h) dS_agg[(int64_t)rr * D + head * H + h] = acc[h];
  }
}

static DropCoef make_drop(const uint32_t* seed_ptr, float keep, float in_keep, int metapath, int64_t row0) {
  DropCoef dc;
  dc.in_thr = (in_keep < 1.f) ? (uint32_t)(in_keep * 16777216.f + 0.5f) : 0u;
  dc.in_inv_keep = (in_keep < 1.f) ? 1.f / ((float)dc.in_thr / 16777216.f) : 1.f;
  dc.seed_ptr = seed_ptr;
  dc.thr = (keep < 1.f) ? (uint32_t)(keep * 16777216.f + 0.5f) : 0u;
  dc.inv_keep = (keep < 1.f) ? 1.f / ((float)dc.thr / 16777216.f) : 1.f;   // unbiased for the quantised keep
  dc.metapath = (uint32_t)metapath;
  dc.row0 = row0;
  return dc;
}

// ring geometry: STAGES batches of B records per warp (STAGES - 1 always in flight), MINB CTAs of 4 warps per SM
template <int K, int H, int STAGES, int B>
struct RingCfg {
  static constexpr int D = K * H;
  static constexpr int TS = D;
  static constexpr int RS = ((D + 3 * K + 3) / 4) * 4;
  static constexpr size_t fwd_smem = (size_t)kStreamWarps * STAGES * B * (TS * 4 + 4);   // ring, columns
  static constexpr size_t bwd_smem = (size_t)kStreamWarps * STAGES * B * (RS * 4 + 4);
  static constexpr size_t w_smem = (size_t)kStreamWarps * STAGES * B * 4;    // + staged edge weights
};

struct HeavyRows {   // the cut rows of a virtual-row view (all null / 0: no splitting)
  const int32_t* rows;
  const int32_t* ptr;
  int n;
};

// Ring geometry of the (K,H) = (8,8) kernels, from the sweep in profiles/r2_gather_sweep.txt (2M-node bench, ms per step
// of 4 launches): more resident warps with ONE batch in flight each beat deeper rings -- forward 2 stages x 16 records at
// 6 CTAs/SM 25.1 (3x16 at 4 CTAs/SM: 29.2), backward 2x16 at 5 CTAs/SM 24.7 (3x16 at 3: 29.3).
// HAN_GATHER_CFG="f,b" picks an alternative for tuning runs (tools/gather_sweep.sh): 0 = the default above,
// 1 = 3x16 (4 / 3 CTAs per SM), 2 = forward 2x24 at 4, backward 2x8 at 6, 3 = forward 2x16 at 5, backward 2x16 at 4.
static void gather_cfg(int* f, int* b) {
  static int cf = -1, cb = -1;
  if (cf < 0) {
    cf = 0;
    cb = 0;
    const char* e = getenv("HAN_GATHER_CFG");
    if (e && e[0] >= '0' && e[0] <= '3') cf = e[0] - '0';
    if (e && e[0] && e[1] == ',' && e[2] >= '0' && e[2] <= '3') cb = e[2] - '0';
  }
  *f = cf;
  *b = cb;
}

template <int K, int H, int STAGES, int B, int MINB, bool SPLIT, bool PLAIN>
static int launch_fwd_cfg2(const int64_t* indptr, const int32_t* indices, const int32_t* chunk_rows,
                          int64_t n_chunks, const float* T, Scorer f2w, float* R, const float* bias, int act,
                          float* out, int64_t out_stride, float* vsave, const float* colmean,
                          const float* ew, const float* resid, int64_t resid_stride, float* const* out2_tab,
                          int64_t out2_block_rows, int64_t out2_stride, float* vsave2, float* csave, DropCoef dc,
                          SplitRows sp, HeavyRows hv, cudaStream_t st) {
  using C = RingCfg<K, H, STAGES, B>;
  HAN_SMEM_ATTR_ONCE((attn_fwd_chunked_kernel<K, H, STAGES, B, MINB, SPLIT, PLAIN>), C::fwd_smem + C::w_smem);
  unsigned grid = (unsigned)ceil_div64(n_chunks, kStreamWarps);
  const size_t smem = C::fwd_smem + (ew ? C::w_smem : 0);
  attn_fwd_chunked_kernel<K, H, STAGES, B, MINB, SPLIT, PLAIN><<<grid, kStreamWarps * 32, smem, st>>>(
      indptr, indices, chunk_rows, n_chunks, T, f2w, R, bias, act, out, out_stride, vsave, colmean, ew, resid, resid_stride,
      out2_tab, out2_block_rows, out2_stride, vsave2, csave, dc, sp);
  if (SPLIT && hv.n > 0)
    attn_fwd_merge_kernel<K, H><<<(unsigned)ceil_div64(hv.n, 4), 128, 0, st>>>(hv.rows, hv.ptr, hv.n, sp.part, R, bias, act,
                                                                            out, out_stride, vsave, resid, resid_stride, out2_tab,
                                                                            out2_block_rows, out2_stride, vsave2, csave);
  return check_launch("han_attn_fwd_chunked");
}

template <int K, int H, int STAGES, int B, int MINB, bool SPLIT>
static int launch_fwd_cfg(const int64_t* indptr, const int32_t* indices, const int32_t* chunk_rows,
                          int64_t n_chunks, const float* T, Scorer f2w, float* R, const float* bias, int act,
                          float* out, int64_t out_stride, float* vsave, const float* colmean,
                          const float* ew, const float* resid, int64_t resid_stride, float* const* out2_tab,
                          int64_t out2_block_rows, int64_t out2_stride, float* vsave2, float* csave, DropCoef dc,
                          SplitRows sp, HeavyRows hv, cudaStream_t st) {
  if constexpr (K == 8 && H == 8) {
    if (!ew && !dc.thr && !dc.in_thr)
      return launch_fwd_cfg2<K, H, STAGES, B, MINB, SPLIT, true>(indptr, indices, chunk_rows, n_chunks, T, f2w, R, bias, act, out,
                                                                 out_stride, vsave, colmean, ew, resid, resid_stride, out2_tab,
                                                                 out2_block_rows, out2_stride, vsave2, csave, dc, sp, hv, st);
  }
  return launch_fwd_cfg2<K, H, STAGES, B, MINB, SPLIT, false>(indptr, indices, chunk_rows, n_chunks, T, f2w, R, bias, act, out,
                                                              out_stride, vsave, colmean, ew, resid, resid_stride, out2_tab,
                                                              out2_block_rows, out2_stride, vsave2, csave, dc, sp, hv, st);
}

template <int K, int H, bool SPLIT>
static int launch_fwd_chunked(const int64_t* indptr, const int32_t* indices, const int32_t* chunk_rows,
                              int64_t n_chunks, const float* T, Scorer f2w, float* R, const float* bias, int act,
                              float* out, int64_t out_stride, float* vsave, const float* colmean,
                              const float* ew, const float* resid, int64_t resid_stride, float* const* out2_tab,
                              int64_t out2_block_rows, int64_t out2_stride, float* vsave2, float* csave, DropCoef dc,
                              SplitRows sp, HeavyRows hv, cudaStream_t st) {
#define HAN_FWD_ARGS indptr, indices, chunk_rows, n_chunks, T, f2w, R, bias, act, out, out_stride, vsave, colmean, ew, resid, \
                     resid_stride, out2_tab, out2_block_rows, out2_stride, vsave2, csave, dc, sp, hv, st
  if constexpr (K == 8 && H == 8) {
    int f, b;
    gather_cfg(&f, &b);
    if (f == 1) return launch_fwd_cfg<K, H, 3, 16, 4, SPLIT>(HAN_FWD_ARGS);
    if (f == 2) return launch_fwd_cfg<K, H, 2, 24, 4, SPLIT>(HAN_FWD_ARGS);
    if (f == 3) return launch_fwd_cfg<K, H, 2, 16, 5, SPLIT>(HAN_FWD_ARGS);
    return launch_fwd_cfg<K, H, 2, 16, 6, SPLIT>(HAN_FWD_ARGS);
  }
  return launch_fwd_cfg<K, H, 3, kBatch, 4, SPLIT>(HAN_FWD_ARGS);
#undef HAN_FWD_ARGS
}

template <int K, int H, int STAGES, int B, int MINB, bool SPLIT>
static int launch_bwd_cfg(const int64_t* t_indptr, const int32_t* t_indices,
                          const int32_t* chunk_rows, int64_t n_chunks, const float* Tsrc, Scorer f2w,
                          const float* R, float* dS_agg, float* df2,
                          const float* ew_t, DropCoef dc, SplitRows sp, HeavyRows hv, cudaStream_t st) {
  using C = RingCfg<K, H, STAGES, B>;
  HAN_SMEM_ATTR_ONCE((attn_bwd_src_chunked_kernel<K, H, STAGES, B, MINB, SPLIT>), C::bwd_smem + C::w_smem);
  unsigned grid = (unsigned)ceil_div64(n_chunks, kStreamWarps);
  const size_t smem = C::bwd_smem + (ew_t ? C::w_smem : 0);
  attn_bwd_src_chunked_kernel<K, H, STAGES, B, MINB, SPLIT><<<grid, kStreamWarps * 32, smem, st>>>(
      t_indptr, t_indices, chunk_rows, n_chunks, Tsrc, f2w, R, dS_agg, df2, ew_t, dc, sp);
  if (SPLIT && hv.n > 0)
    attn_bwd_src_merge_kernel<K, H><<<(unsigned)ceil_div64(hv.n, 4), 128, 0, st>>>(hv.rows, hv.ptr, hv.n, sp.part, dS_agg, df2);
  return check_launch("han_attn_bwd_src_chunked");
}

template <int K, int H, bool SPLIT>
static int launch_bwd_src_chunked(const int64_t* t_indptr, const int32_t* t_indices,
                                  const int32_t* chunk_rows, int64_t n_chunks, const float* Tsrc, Scorer f2w,
                                  const float* R, float* dS_agg, float* df2,
                                  const float* ew_t, DropCoef dc, SplitRows sp, HeavyRows hv, cudaStream_t st) {
#define HAN_BWD_ARGS t_indptr, t_indices, chunk_rows, n_chunks, Tsrc, f2w, R, dS_agg, df2, ew_t, dc, sp, hv, st
  if constexpr (K == 8 && H == 8) {
    int f, b;
    gather_cfg(&f, &b);
    if (b == 1) return launch_bwd_cfg<K, H, 3, 16, 3, SPLIT>(HAN_BWD_ARGS);
    if (b == 2) return launch_bwd_cfg<K, H, 2, 8, 6, SPLIT>(HAN_BWD_ARGS);
    if (b == 3) return launch_bwd_cfg<K, H, 2, 16, 4, SPLIT>(HAN_BWD_ARGS);
    return launch_bwd_cfg<K, H, 2, 16, 5, SPLIT>(HAN_BWD_ARGS);
  }
  return launch_bwd_cfg<K, H, 3, kBatch, 3, SPLIT>(HAN_BWD_ARGS);
#undef HAN_BWD_ARGS
}

}  // namespace han

using namespace han;

#define HAN_FOR_SHAPES(X) X(8, 8) X(4, 8) X(2, 8) X(1, 8) X(8, 4) X(4, 4) X(1, 4) X(8, 16) X(4, 16) X(1, 16) X(16, 4) X(16, 8)

extern "C" {

// Work-item size: ~2048 edges on large graphs (amortises the pipeline fill), smaller on small graphs so
// that there are at least ~32 work items per SM.
int64_t han_csr_chunk_edges(int64_t nnz) {
  int64_t c = nnz / ((int64_t)kNumSMs * 32);
  if (c > kMaxChunkEdges) c = kMaxChunkEdges;
  if (c < kMinChunkEdges) c = kMinChunkEdges;
  return c;
}

int64_t han_csr_num_chunks(int64_t nnz) { return nnz / han_csr_chunk_edges(nnz) + 1; }

int han_csr_chunk_rows(const int64_t* indptr, int64_t n_rows, int64_t nnz, int32_t* chunk_rows,
                       han_stream_t stream) {
  HAN_REQUIRE(indptr && chunk_rows, "null pointer");
  HAN_REQUIRE(n_rows > 0 && nnz >= 0 && n_rows < ((int64_t)1 << 31), "sizes");
  const int64_t ce = han_csr_chunk_edges(nnz);
  const int64_t n_chunks = nnz / ce + 1;
  chunk_rows_kernel<<<(unsigned)ceil_div64(n_chunks + 1, 256), 256, 0, as_stream(stream)>>>(indptr, n_rows, n_chunks, ce,
                                                                                           chunk_rows);
  return check_launch(__func__);
}

int han_attn_fwd_chunked(const int64_t* indptr, const int32_t* indices, const int32_t* chunk_rows,
                         int64_t n_chunks, int64_t n_dst, const float* T, const float* a2, const float* b2,
                         float* R, const float* bias,
                         int K, int H, int act, float* out, int64_t out_stride, float* vsave,
                         const float* colmean, const float* edge_w, const float* resid, int64_t resid_stride,
                         float* const* out2_tab, int64_t out2_block_rows, int64_t out2_stride, float* vsave2,
                         float* csave, const uint32_t* seed_ptr, float coef_keep, float in_keep, int metapath,
                         int64_t row0,
                         han_stream_t stream) {
  HAN_REQUIRE((vsave2 == nullptr) == (csave == nullptr) && (uintptr_t)vsave2 % 16 == 0, "vsave2 / csave: both or neither");
  HAN_REQUIRE(!out2_tab || (out2_stride >= (int64_t)K * H && out2_stride % 4 == 0 && out2_block_rows > 0), "out2");
  HAN_REQUIRE(!resid || (resid_stride >= (int64_t)K * H && resid_stride % 4 == 0 && (uintptr_t)resid % 16 == 0), "resid");
  HAN_REQUIRE(indptr && chunk_rows && T && a2 && b2 && R && bias && out && vsave, "null pointer");
  HAN_REQUIRE(coef_keep > 0.f && coef_keep <= 1.f && in_keep > 0.f && in_keep <= 1.f &&
              ((coef_keep == 1.f && in_keep == 1.f) || seed_ptr), "coef_keep, in_keep in (0,1]; dropout needs seed_ptr");
  const DropCoef dc = make_drop(seed_ptr, coef_keep, in_keep, metapath, row0);
  const Scorer f2w{a2, b2};
  HAN_REQUIRE(n_dst > 0 && n_chunks > 0, "sizes");
  HAN_REQUIRE(act == HAN_ACT_ELU || act == HAN_ACT_IDENTITY, "activation");
  HAN_REQUIRE(out_stride >= (int64_t)K * H && out_stride % 4 == 0, "out_stride");
  HAN_REQUIRE(((uintptr_t)T % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)vsave % 16 == 0) &&
              ((uintptr_t)bias % 16 == 0), "16-byte alignment");
#define X(k, h)         \
  if (K == k && H == h) \
    return launch_fwd_chunked<k, h, false>(indptr, indices, chunk_rows, n_chunks, T, f2w, R, bias, act, out, out_stride, vsave, colmean, edge_w, resid, resid_stride, out2_tab, out2_block_rows, out2_stride, vsave2, csave, dc, SplitRows{nullptr, nullptr}, HeavyRows{nullptr, nullptr, 0}, as_stream(stream));
  HAN_FOR_SHAPES(X)
#undef X
  return fail_arg(__func__, "unsupported (K,H); see han_attn_shape_supported");
}

int han_attn_bwd_src_chunked(const int64_t* t_indptr, const int32_t* t_indices,
                             const int32_t* chunk_rows, int64_t n_chunks, int64_t n_src,
                             const float* Tsrc, const float* a2, const float* b2, const float* R, int K, int H,
                             float* dS_agg, float* df2,
                             const float* edge_w_t, const uint32_t* seed_ptr,
                             float coef_keep, float in_keep, int metapath, int64_t row0, han_stream_t stream) {
  HAN_REQUIRE(t_indptr && chunk_rows && Tsrc && a2 && b2 && R && dS_agg && df2, "null pointer");
  HAN_REQUIRE(coef_keep > 0.f && coef_keep <= 1.f && in_keep > 0.f && in_keep <= 1.f &&
              ((coef_keep == 1.f && in_keep == 1.f) || seed_ptr), "coef_keep, in_keep in (0,1]; dropout needs seed_ptr");
  const DropCoef dc = make_drop(seed_ptr, coef_keep, in_keep, metapath, row0);
  const Scorer f2w{a2, b2};
  HAN_REQUIRE(n_src > 0 && n_chunks > 0, "sizes");
  HAN_REQUIRE(((uintptr_t)R % 16 == 0) && ((uintptr_t)Tsrc % 16 == 0) && ((uintptr_t)dS_agg % 16 == 0), "16-byte alignment");
#define X(k, h)         \
  if (K == k && H == h) \
    return launch_bwd_src_chunked<k, h, false>(t_indptr, t_indices, chunk_rows, n_chunks, Tsrc, f2w, R, dS_agg, df2, edge_w_t, dc, SplitRows{nullptr, nullptr}, HeavyRows{nullptr, nullptr, 0}, as_stream(stream));
  HAN_FOR_SHAPES(X)
#undef X
  return fail_arg(__func__, "unsupported (K,H); see han_attn_shape_supported");
}

int han_attn_fwd_chunked_split(const int64_t* indptr_v, const int32_t* indices, const int32_t* chunk_rows,
                               int64_t n_chunks, int64_t n_dst, const float* T, const float* a2, const float* b2,
                         float* R, const float* bias,
                               int K, int H, int act, float* out, int64_t out_stride, float* vsave,
                               const float* colmean, const float* edge_w, const float* resid, int64_t resid_stride,
                               float* const* out2_tab, int64_t out2_block_rows, int64_t out2_stride, float* vsave2,
                               float* csave, const uint32_t* seed_ptr, float coef_keep, float in_keep, int metapath,
                         int64_t row0,
                               const int32_t* vmap,
                               float* part, const int32_t* heavy_rows, const int32_t* heavy_ptr, int n_heavy,
                               han_stream_t stream) {
  HAN_REQUIRE(!resid || (resid_stride >= (int64_t)K * H && resid_stride % 4 == 0 && (uintptr_t)resid % 16 == 0), "resid");
  HAN_REQUIRE(indptr_v && chunk_rows && T && a2 && b2 && R && bias && out && vsave, "null pointer");
  HAN_REQUIRE(vmap && part && n_heavy >= 0 && (n_heavy == 0 || (heavy_rows && heavy_ptr)), "split view: vmap, part, heavy rows");
  HAN_REQUIRE(!out2_tab || (out2_stride >= (int64_t)K * H && out2_stride % 4 == 0 && out2_block_rows > 0), "out2");
  HAN_REQUIRE(coef_keep > 0.f && coef_keep <= 1.f && in_keep > 0.f && in_keep <= 1.f &&
              ((coef_keep == 1.f && in_keep == 1.f) || seed_ptr), "coef_keep, in_keep in (0,1]; dropout needs seed_ptr");
  const DropCoef dc = make_drop(seed_ptr, coef_keep, in_keep, metapath, row0);
  const Scorer f2w{a2, b2};
  HAN_REQUIRE(n_dst > 0 && n_chunks > 0, "sizes");
  HAN_REQUIRE(act == HAN_ACT_ELU || act == HAN_ACT_IDENTITY, "activation");
  HAN_REQUIRE(out_stride >= (int64_t)K * H && out_stride % 4 == 0, "out_stride");
  HAN_REQUIRE(((uintptr_t)T % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)vsave % 16 == 0) &&
              ((uintptr_t)bias % 16 == 0) && ((uintptr_t)vmap % 8 == 0), "alignment");
  const SplitRows sp{reinterpret_cast<const int2*>(vmap), part};
  const HeavyRows hv{heavy_rows, heavy_ptr, n_heavy};
#define X(k, h)         \
  if (K == k && H == h) \
    return launch_fwd_chunked<k, h, true>(indptr_v, indices, chunk_rows, n_chunks, T, f2w, R, bias, act, out, out_stride, vsave, colmean, edge_w, resid, resid_stride, out2_tab, out2_block_rows, out2_stride, vsave2, csave, dc, sp, hv, as_stream(stream));
  HAN_FOR_SHAPES(X)
#undef X
  return fail_arg(__func__, "unsupported (K,H); see han_attn_shape_supported");
}

int han_attn_bwd_src_chunked_split(const int64_t* t_indptr_v, const int32_t* t_indices,
                                   const int32_t* chunk_rows, int64_t n_chunks, int64_t n_src,
                                   const float* Tsrc, const float* a2, const float* b2, const float* R, int K, int H,
                             float* dS_agg, float* df2,
                                   const float* edge_w_t, const uint32_t* seed_ptr,
                                   float coef_keep, float in_keep, int metapath, int64_t row0, const int32_t* vmap, float* part,
                                   const int32_t* heavy_rows, const int32_t* heavy_ptr, int n_heavy,
                                   han_stream_t stream) {
  HAN_REQUIRE(t_indptr_v && chunk_rows && Tsrc && a2 && b2 && R && dS_agg && df2, "null pointer");
  HAN_REQUIRE(vmap && part && n_heavy >= 0 && (n_heavy == 0 || (heavy_rows && heavy_ptr)), "split view: vmap, part, heavy rows");
  HAN_REQUIRE(coef_keep > 0.f && coef_keep <= 1.f && in_keep > 0.f && in_keep <= 1.f &&
              ((coef_keep == 1.f && in_keep == 1.f) || seed_ptr), "coef_keep, in_keep in (0,1]; dropout needs seed_ptr");
  const DropCoef dc = make_drop(seed_ptr, coef_keep, in_keep, metapath, row0);
  const Scorer f2w{a2, b2};
  HAN_REQUIRE(n_src > 0 && n_chunks > 0, "sizes");
  HAN_REQUIRE(((uintptr_t)R % 16 == 0) && ((uintptr_t)Tsrc % 16 == 0) && ((uintptr_t)dS_agg % 16 == 0) &&
              ((uintptr_t)vmap % 8 == 0), "alignment");
  const SplitRows sp{reinterpret_cast<const int2*>(vmap), part};
  const HeavyRows hv{heavy_rows, heavy_ptr, n_heavy};
#define X(k, h)         \
  if (K == k && H == h) \
    return launch_bwd_src_chunked<k, h, true>(t_indptr_v, t_indices, chunk_rows, n_chunks, Tsrc, f2w, R, dS_agg, df2, edge_w_t, dc, sp, hv, as_stream(stream));
  HAN_FOR_SHAPES(X)
#undef X
  return fail_arg(__func__, "unsupported (K,H); see han_attn_shape_supported");
}

}  // extern "C"
