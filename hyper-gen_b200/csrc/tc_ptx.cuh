// tc_ptx.cuh — PTX wrappers shared by the tcgen05 dist kernels (dist_tc.cu: two-limb planes,
// dist_narrow.cu: single s8 plane): mbarriers, TMA loads, tcgen05.mma / commit / ld for kind::i8,
// K-major 128B-swizzled shared-memory descriptors, and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>

#include "hg_common.cuh"

namespace hgtc {

constexpr int TC_BK = 128;  // K block in bytes == int8 elements (one 128B-swizzle row)

// (shared-memory addresses as 32-bit values throughout)
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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes, uint32_t on) {
  asm volatile(
      "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %2, 0;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar),
      "r"(bytes), "r"(on)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, uint32_t on) {
  asm volatile(
      "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %5, 0;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(on)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// smem matrix descriptor: K-major, 128B swizzle, 8-row atoms 1024 B apart (SBO), version 1
// Low word: start address >> 4 (14 bits) | LBO field = 1; high word: SBO = 1024 >> 4, version 1 (bit 46),
// layout 2 = SWIZZLE_128B (bits 61-63).  Shared addresses are < 2^18, so the low word of (addr + off) is
// desc_lo(addr) + (off >> 4): the issue loop adds compile-time constants instead of re-encoding.
constexpr uint32_t TC_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
// The single-thread instructions (tcgen05.mma / commit, TMA) are issued from WARP-UNIFORM code with the
// issuing lane selected by a predicate (`on` is 1 in exactly one lane, from elect_one()).  Inside an
// `if (lane == 0)` branch ptxas cannot keep the descriptors in uniform registers and wraps every MMA in
// an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~10 extra dependent instructions), which left the
// issuing thread, not the tensor pipe, as the bottleneck (~126 clk per 64-clk MMA).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t on;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(on));
  return on;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t on) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 e, %5, 0;\n\t"
      "mov.b64 da, {%1, %6};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "@e tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(on), "r"(TC_DESC_HI)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar, uint32_t on) {
  asm volatile(
      "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %1, 0;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar),
      "r"(on)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// tcgen05.ld writes its destination registers asynchronously until tcgen05.wait::ld.  This empty volatile asm
// is ordered after the wait (volatile asms keep their order) and "redefines" the registers, so no use of the
// loaded values can be scheduled ahead of the wait.
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&v)[16]) {
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                    "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
// TMA load whose completion bytes go to an mbarrier of either CTA of the pair (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1,
                                                 uint32_t on) {
  asm volatile(
      "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %5, 0;\n\t"
      "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(on)
      : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, uint32_t on) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 e, %5, 0;\n\t"
      "mov.b64 da, {%1, %6};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "@e tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(on), "r"(TC_DESC_HI)
      : "memory");
}
// arrive on the barrier at this offset in both CTAs once the pair's MMAs issued so far are done
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint32_t on) {
  asm volatile(
      "{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %2, 0;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar),
      "h"((uint16_t)3), "r"(on)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // acquire at cluster scope
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int make_plane_map(CUtensorMap *map, const int8_t *planes, uint64_t rows2, uint32_t hv_d, uint32_t box_rows) {
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

}  // namespace hgtc
