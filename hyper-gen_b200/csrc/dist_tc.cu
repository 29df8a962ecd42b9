// dist_tc.cu — tensor-pipe dist path: every ref x query dot product as an exact integer
// contraction on tcgen05 (kind::i8, s32 accumulators in TMEM) fed by TMA, with the ANI /
// threshold epilogue fused behind it.
//
// Replaces the pair loop of dist::compute_hv_ani + compute_pairwise_ani (reference
// src/dist.rs:139-161,231-294).  Sketch HV elements need 9-11 bits at the BASELINE configs
// (SURVEY.md F5), which does not fit the s8 MMA input type, so every i16 value is split into
// two s8 limbs
//        x = 128 * h + l,   l = ((x + 64) mod 128) - 64 in [-64, 63],   h = (x - l) / 128
// (exact for |x| <= 8127, i.e. hv_quant_bits <= 13; wider inputs take the SIMT path), and
//        r . q = 2^14 (Hr.Hq) + 2^7 (Hr.Lq + Lr.Hq) + (Lr.Lq)
// is accumulated as three s32 accumulators (the two cross products share one), each < 2^31
// for D <= 32768.  The recombination wraps in i32 exactly like the reference's i32 sum.
//
// One CTA per 128 x 128 output tile, 10 warps: warp 0 issues TMA loads of the four 128 x 128 B
// operand tiles (128B-swizzled, K-major) through a 3-stage mbarrier ring, one thread of warp 1
// issues the tcgen05.mma stream and commits stage releases, warps 2-9 drain the three TMEM
// accumulators with tcgen05.ld, apply a division-free bound and run the shared exact epilogue
// (dist_common.cuh) only where a pair can reach the threshold.
#include <cuda.h>

#include <cstdlib>

#include "dist_common.cuh"

namespace {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 128;  // BK in bytes == int8 elements
constexpr int TC_STAGES = 3;
constexpr int TC_TILE_BYTES = 128 * TC_BK;            // one operand tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;     // A_hi, A_lo, B_hi, B_lo
constexpr int TC_EPI_WARPS = 8;                      // two per TMEM lane quarter, half the columns each
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 512 /*col bounds*/;
constexpr uint32_t TC_TMEM_COLS = 512;

// instruction descriptor, kind::i8: D = s32, A = B = signed 8-bit, both K-major, M = 128, N = 128
constexpr uint32_t TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(dst),
      "l"(map), "r"(bar), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// smem matrix descriptor: K-major, 128B swizzle, 8-row atoms 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// CL = 1: stand-alone CTAs.  CL = 2: 2 x 2 thread-block clusters — the two CTAs of a cluster row
// share their A (ref) tile and the two of a cluster column their B (query) tile; each CTA loads
// half of each shared tile and TMA-multicasts it to its peer, halving the L2 -> SM operand
// traffic that bounds this kernel.
template <int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
dist_tc_kernel(const __grid_constant__ CUtensorMap tm_ref, const __grid_constant__ CUtensorMap tm_qry,
               uint32_t ref_plane_rows, uint32_t ref_row_base, uint32_t qry_plane_rows, uint32_t hv_d,
               hg::DistEpilogue ep) {
  const uint32_t row0 = blockIdx.y * TC_BM, col0 = blockIdx.x * TC_BN;
  // symmetric: a tile (cluster of tiles) whose largest global j is not above its smallest global
  // i is empty.  The test is uniform over a cluster, so whole clusters leave together.
  {
    const uint32_t crow0 = (blockIdx.y / CL) * CL * TC_BM, ccol_end = (blockIdx.x / CL + 1) * CL * TC_BN;
    if (ep.symmetric && (uint64_t)ep.j0 + ccol_end - 1 <= (uint64_t)ep.i0 + crow0) return;
  }
  const bool own_tile_empty = ep.symmetric && (uint64_t)ep.j0 + col0 + TC_BN - 1 <= (uint64_t)ep.i0 + row0;
  // position inside the cluster: cx along N (blockIdx.x), cy along M (blockIdx.y); rank = cx + CL * cy
  const uint32_t cx = CL == 1 ? 0u : (blockIdx.x % CL), cy = CL == 1 ? 0u : (blockIdx.y % CL);
  const uint32_t crank = cx + CL * cy;
  const uint16_t mask_a = CL == 1 ? 1 : (uint16_t)(((1u << CL) - 1u) << (CL * cy));               // same cluster row
  const uint16_t mask_b = CL == 1 ? 1 : (uint16_t)((1u << cx) | (1u << (cx + CL)));               // same cluster column
  const uint16_t mask_e = (uint16_t)(mask_a | mask_b);  // everyone that writes into my stages
  (void)crank;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle wants 1024 B alignment
  uint8_t *aligned = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + TC_STAGES * TC_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * TC_STAGES);
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(aligned + TC_STAGES * TC_STAGE_BYTES + 8 * (2 * TC_STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    // a stage is released when my own MMAs and those of every CTA I multicast into have read it
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL == 1 ? 1 : 2 * CL - 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every peer's barriers are initialised before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t num_kb = hv_d / TC_BK;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (uint32_t kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), TC_STAGE_BYTES);
        const uint32_t st = base + s * TC_STAGE_BYTES;
        const int k0 = (int)(kb * TC_BK);
        if (CL == 1) {
          tma_load_2d(st + 0 * TC_TILE_BYTES, &tm_ref, full_bar(s), k0, (int)(ref_row_base + row0));  // ref hi limbs
          tma_load_2d(st + 1 * TC_TILE_BYTES, &tm_ref, full_bar(s), k0, (int)(ref_plane_rows + ref_row_base + row0));  // ref lo
          tma_load_2d(st + 2 * TC_TILE_BYTES, &tm_qry, full_bar(s), k0, (int)col0);                    // qry hi limbs
          tma_load_2d(st + 3 * TC_TILE_BYTES, &tm_qry, full_bar(s), k0, (int)(qry_plane_rows + col0));  // qry lo limbs
        } else {
          // my 1/CL slice of the row-shared A tile and of the column-shared B tile, multicast
          constexpr int HR = 128 / CL;                     // rows per slice
          constexpr int HB = HR * TC_BK;                   // bytes per slice
          const int ar = (int)(ref_row_base + row0 + cx * HR), br = (int)(col0 + cy * HR);
          tma_load_2d_mc(st + 0 * TC_TILE_BYTES + cx * HB, &tm_ref, full_bar(s), k0, ar, mask_a);
          tma_load_2d_mc(st + 1 * TC_TILE_BYTES + cx * HB, &tm_ref, full_bar(s), k0, (int)ref_plane_rows + ar, mask_a);
          tma_load_2d_mc(st + 2 * TC_TILE_BYTES + cy * HB, &tm_qry, full_bar(s), k0, br, mask_b);
          tma_load_2d_mc(st + 3 * TC_TILE_BYTES + cy * HB, &tm_qry, full_bar(s), k0, (int)qry_plane_rows + br, mask_b);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (a single thread drives the tensor core) =====
    if (lane == 0) {
      for (uint32_t kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + s * TC_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < TC_BK / 32; ++ks) {  // one MMA covers K = 32 int8
          const uint64_t a_hi = umma_desc(st + 0 * TC_TILE_BYTES + 32 * ks), a_lo = umma_desc(st + 1 * TC_TILE_BYTES + 32 * ks);
          const uint64_t b_hi = umma_desc(st + 2 * TC_TILE_BYTES + 32 * ks), b_lo = umma_desc(st + 3 * TC_TILE_BYTES + 32 * ks);
          const uint32_t acc = (kb | (uint32_t)ks) != 0u;
          umma_i8(tmem_base + 0 * TC_BN, a_hi, b_hi, acc);  // Hr.Hq
          umma_i8(tmem_base + 1 * TC_BN, a_hi, b_lo, acc);  // Hr.Lq
          umma_i8(tmem_base + 1 * TC_BN, a_lo, b_hi, 1u);   //   + Lr.Hq
          umma_i8(tmem_base + 2 * TC_BN, a_lo, b_lo, acc);  // Lr.Lq
        }
        if (CL == 1) umma_commit(empty_bar(s));  // the stage is free once these MMAs have read it
        else umma_commit_mc(empty_bar(s), mask_e);
      }
      umma_commit(accum_bar);       // all accumulators final
    }
  } else {
    // ===== epilogue: TMEM -> registers -> exact i32 dot -> bound test -> (rare) ANI + append =====
    // Cheap per-element test first: ani >= ani_th implies dot >= cfrac * (norm_r + norm_q)
    // (dist_common.cuh), split into a per-row and a per-column integer so that an element
    // costs two shift-adds, one add and one compare.  Only 16-column chunks in which some
    // lane passes run the exact f32 ANI sequence and the compacted append.
    const int ew = warp - 2;                 // 0..7
    const uint32_t q = warp & 3;             // TMEM lane quarter this warp may read
    const int half = (ew >> 2);              // which 64 columns this warp drains
    int32_t *s_tq = reinterpret_cast<int32_t *>(aligned + TC_STAGES * TC_STAGE_BYTES + 256);
    const bool use_bound = ep.cfrac > 0.0f;
    {  // per-column bounds, one column per epilogue thread of the first four warps
      const int t = threadIdx.x - 64;
      if (t < TC_BN) {
        const uint32_t lj = col0 + t;
        int32_t v = INT32_MIN;  // out-of-range / degenerate columns always go to the exact path
        if (use_bound && lj < ep.n_qry) {
          const int32_t nq = ep.qry_norm[lj];
          if (nq > 0) v = __float2int_rd(ep.cfrac * __int2float_rz(nq)) - 1;
        }
        s_tq[t] = v;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(32 * TC_EPI_WARPS) : "memory");  // epilogue warps only
    const uint32_t li = row0 + q * 32 + lane;
    const bool row_live = li < ep.n_ref;
    int32_t tr = INT32_MIN / 2;
    if (use_bound && row_live) {
      const int32_t nr = ep.ref_norm[li];
      if (nr > 0) tr = __float2int_rd(ep.cfrac * __int2float_rz(nr)) - 1;
    }
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((q * 32u) << 16);
#pragma unroll 1
    for (int c = half * (TC_BN / 2); c < (half + 1) * (TC_BN / 2) && !own_tile_empty; c += 16) {
      uint32_t hh[16], cr[16], ll[16];
      tmem_ld16(taddr + 0 * TC_BN + c, hh);
      tmem_ld16(taddr + 1 * TC_BN + c, cr);
      tmem_ld16(taddr + 2 * TC_BN + c, ll);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      int32_t dot[16];
      uint32_t cand = 0;  // bit j: column c + j of my row can reach the threshold
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        dot[j] = (int32_t)((((hh[j] << 7) + cr[j]) << 7) + ll[j]);  // wrapping i32, as dist.rs:147-151
        const int32_t tq = s_tq[c + j];
        // tq = INT32_MIN or tr = INT32_MIN / 2 (bound off / degenerate norm) always qualify
        const bool cj = dot[j] >= (int32_t)((uint32_t)tr + (uint32_t)tq) || tq == INT32_MIN || tr == INT32_MIN / 2;
        cand |= (uint32_t)cj << j;
      }
      if (!row_live) cand = 0;
      // exact f32 ANI + compacted append, one candidate per lane per round (candidates are rare:
      // the number of rounds is the largest per-lane count, not 16)
      while (__any_sync(0xffffffffu, cand != 0)) {
        const bool have = cand != 0;
        const int j = have ? __ffs(cand) - 1 : 0;
        cand &= cand - 1;
        int32_t d = dot[0];
#pragma unroll
        for (int jj = 1; jj < 16; ++jj) d = (jj == j) ? dot[jj] : d;
        const uint32_t lj = col0 + c + j;
        hg::dist_emit(ep, have && lj < ep.n_qry, li, lj, d);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // nobody leaves while a peer may still multicast into it or signal its barriers
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// i16 -> two s8 limb planes: out[0][r][d] = h, out[1][r][d] = l  (plane stride = rows * hv_d)
__global__ void split_limbs_kernel(const int16_t *__restrict__ hv, uint64_t n_elems, int8_t *__restrict__ planes) {
  const uint64_t n8 = n_elems / 8;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4 *>(hv)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t hi[2] = {0, 0}, lo[2] = {0, 0};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int x = (int)(int16_t)(w[e >> 1] >> (16 * (e & 1)));
      const int l = ((x + 64) & 127) - 64;
      const int h = (x - l) >> 7;
      hi[e >> 2] |= (uint32_t)(h & 0xFF) << (8 * (e & 3));
      lo[e >> 2] |= (uint32_t)(l & 0xFF) << (8 * (e & 3));
    }
    reinterpret_cast<uint2 *>(planes)[i] = make_uint2(hi[0], hi[1]);
    reinterpret_cast<uint2 *>(planes + n_elems)[i] = make_uint2(lo[0], lo[1]);
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_plane_map(CUtensorMap *map, const int8_t *planes, uint64_t rows2, uint32_t hv_d, uint32_t box_rows) {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    HG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) { hg_set_error("cuTensorMapEncodeTiled not available"); return HG_E_CUDA; }
    fn = (encode_tiled_fn)p;
  }
  const cuuint64_t dims[2] = {hv_d, rows2};          // innermost first: K bytes, then rows of both planes
  const cuuint64_t strides[1] = {hv_d};              // bytes between rows
  const cuuint32_t box[2] = {TC_BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)planes, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { hg_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return HG_E_CUDA; }
  return HG_OK;
}

}  // namespace

int hg_launch_dist_tc(hg_ctx *ctx, const int16_t *d_ref, const int32_t *d_ref_norm, uint32_t n_ref, uint32_t i0,
                      const int16_t *d_qry, const int32_t *d_qry_norm, uint32_t n_qry, uint32_t j0, uint32_t hv_d,
                      uint32_t ksize, float ani_th, int symmetric, hg_hit *d_hits, uint64_t cap,
                      unsigned long long *d_n_hits) {
  if (n_ref == 0 || n_qry == 0) return HG_OK;
  if (hv_d % TC_BK != 0 || hv_d > 32768) {
    hg_set_error("tensor path needs hv_d %% 128 == 0 and hv_d <= 32768 (s32 accumulator bound), got %u", hv_d);
    return HG_E_UNSUPPORTED;
  }
  if (((uintptr_t)d_ref | (uintptr_t)d_qry) & 15) {
    hg_set_error("tensor path needs 16-byte aligned HV matrices");
    return HG_E_UNSUPPORTED;
  }
  int rc;
  // ---- limb planes (scratch 8 / 9) ----
  const uint64_t ref_elems = (uint64_t)n_ref * hv_d, qry_elems = (uint64_t)n_qry * hv_d;
  // the query block may alias the ref block (all-vs-all) or contain it (row shard of the same matrix)
  const bool qry_covers_ref = d_ref >= d_qry && d_ref + ref_elems <= d_qry + qry_elems &&
                              ((d_ref - d_qry) % hv_d) == 0;
  void *p_q, *p_r = nullptr;
  if ((rc = hg_scratch(ctx, HG_S_QRY_LIMBS, 2 * qry_elems + 1024, &p_q))) return rc;
  if (!qry_covers_ref && (rc = hg_scratch(ctx, HG_S_REF_LIMBS, 2 * ref_elems + 1024, &p_r))) return rc;
  const uint32_t gx = (n_qry + TC_BN - 1) / TC_BN, gy_total = (n_ref + TC_BM - 1) / TC_BM;
  // Stand-alone CTAs by default.  The 2 x 2 cluster + TMA multicast variant (HG_DIST_CLUSTER=2) halves
  // the L2 reads but measured 10-15 % SLOWER on B200: the kernel is bound by bytes arriving per SM
  // (~37 B/clk/SM), which multicast does not change.  Kept for the record and for larger-L2-pressure shapes.
  int cl = 1;
  if (const char *e = getenv("HG_DIST_CLUSTER")) cl = (atoi(e) == 2 && gx >= 2 && gy_total >= 2) ? 2 : 1;
  const uint32_t box_rows = 128 / cl;

  CUtensorMap tm_ref, tm_qry;
  if ((rc = make_plane_map(&tm_qry, (const int8_t *)p_q, 2ull * n_qry, hv_d, box_rows))) return rc;
  uint32_t ref_plane_rows = n_ref, ref_row_off = 0;
  if (qry_covers_ref) {  // the ref rows are a window of the query planes
    tm_ref = tm_qry;
    ref_plane_rows = n_qry;
    ref_row_off = (uint32_t)((d_ref - d_qry) / hv_d);
  } else if ((rc = make_plane_map(&tm_ref, (const int8_t *)p_r, 2ull * n_ref, hv_d, box_rows))) {
    return rc;
  }

  if (!ctx->tc_attr_set) {
    HG_CUDA(cudaFuncSetAttribute(dist_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    HG_CUDA(cudaFuncSetAttribute(dist_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    ctx->tc_attr_set = 1;
  }
  // ---- GPU work starts here (stage timer 4..5 brackets exactly this) ----
  auto split = [&](const int16_t *src, uint64_t elems, void *dst) {
    uint64_t blocks = (elems / 8 + 255) / 256;
    if (blocks > (uint64_t)ctx->sm_count * 16) blocks = (uint64_t)ctx->sm_count * 16;
    split_limbs_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(src, elems, (int8_t *)dst);
    ctx->launches++;
  };

  HG_PROF(ctx, 4);
  split(d_qry, qry_elems, p_q);
  if (!qry_covers_ref) split(d_ref, ref_elems, p_r);
  HG_CUDA(cudaGetLastError());
  const uint32_t y_step = 65534;  // even, <= the gridDim.y limit
  for (uint32_t y0 = 0; y0 < gy_total; y0 += y_step) {
    const uint32_t gy = gy_total - y0 < y_step ? gy_total - y0 : y_step;
    hg::DistEpilogue ep;
    ep.ref_norm = d_ref_norm + (size_t)y0 * TC_BM;
    ep.qry_norm = d_qry_norm;
    ep.n_ref = n_ref - y0 * TC_BM;
    ep.n_qry = n_qry;
    ep.i0 = i0 + y0 * TC_BM;
    ep.j0 = j0;
    ep.ksize_f = (float)ksize;
    ep.ani_th = ani_th;
    ep.jmin = hg::dist_jmin(ani_th, ksize);
    ep.cfrac = ep.jmin > 0.0f ? (float)((double)ep.jmin / (1.0 + (double)ep.jmin)) : 0.0f;
    ep.symmetric = symmetric;
    ep.hits = d_hits;
    ep.cap = cap;
    ep.n_hits = d_n_hits;
    const uint32_t row_base = ref_row_off + y0 * TC_BM;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((gx + cl - 1) / cl * cl, (gy + cl - 1) / cl * cl, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = TC_SMEM_BYTES;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = cl;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cl == 2)
      HG_CUDA(cudaLaunchKernelEx(&cfg, dist_tc_kernel<2>, tm_ref, tm_qry, ref_plane_rows, row_base, n_qry, hv_d, ep));
    else
      HG_CUDA(cudaLaunchKernelEx(&cfg, dist_tc_kernel<1>, tm_ref, tm_qry, ref_plane_rows, row_base, n_qry, hv_d, ep));
    ctx->launches++;
  }
  HG_PROF(ctx, 5);
  HG_CUDA(cudaGetLastError());
  return HG_OK;
}
