// dist_common.cuh — the fused epilogue shared by both dist kernels: exact i32 dot ->
// Jaccard -> ANI (reference src/dist.rs:153-160) -> `ani >= ani_th` filter
// (src/utils.rs:274-285) -> compacted append of (i, j, dot, ani).
#pragma once
#include <math.h>

#include "hg_common.cuh"

namespace hg {

struct DistEpilogue {
  const int32_t *__restrict__ ref_norm;
  const int32_t *__restrict__ qry_norm;
  uint32_t n_ref, n_qry;
  uint32_t i0, j0;  // global index offsets of this shard
  float ksize_f;
  float ani_th;
  float jmin;  // conservative Jaccard pre-filter (0 = off): pairs below it cannot reach ani_th
  float cfrac; // jmin / (1 + jmin): dot >= cfrac * (norm_r + norm_q) is the same bound, division-free
  int symmetric;
  hg_hit *__restrict__ hits;
  unsigned long long cap;
  unsigned long long *__restrict__ n_hits;
};

// Called by every lane of a warp (converged) with its own candidate; `live` marks lanes
// whose (li, lj) is inside the matrices.  Survivors are appended with one atomic per warp.
__device__ __forceinline__ void dist_emit(const DistEpilogue &e, bool live, uint32_t li, uint32_t lj, int32_t dot) {
  bool keep = false;
  float ani = 0.0f;
  uint32_t gi = 0, gj = 0;
  if (live) {
    gi = e.i0 + li;
    gj = e.j0 + lj;
    if (!e.symmetric || gj > gi) {  // dist.rs:253-265: only j > i when ref == query
      const int32_t nr = e.ref_norm[li], nq = e.qry_norm[lj];
      bool cand = true;
      if (e.jmin > 0.0f) {  // cheap monotone bound first; the exact f32 sequence decides
        const int32_t den = (int32_t)((uint32_t)nr + (uint32_t)nq - (uint32_t)dot);
        cand = den <= 0 || __int2float_rn(dot) >= e.jmin * __int2float_rn(den);
      }
      if (cand) {
        ani = ani_from_dot(dot, nr, nq, e.ksize_f);
        keep = ani >= e.ani_th;
      }
    }
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, keep);
  if (bal == 0) return;
  const uint32_t lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == (uint32_t)(__ffs(bal) - 1)) base = atomicAdd(e.n_hits, (unsigned long long)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
  if (keep) {
    const unsigned long long idx = base + __popc(bal & ((1u << lane) - 1u));
    if (idx < e.cap) {
      hg_hit h;
      h.i = gi; h.j = gj; h.dot = dot; h.ani = ani;
      e.hits[idx] = h;
    }
  }
}

// Host: the Jaccard value below which ANI (dist.rs:154) cannot reach ani_th, lowered by a safety
// margin (0.01 ANI points and 0.1 % relative) that dwarfs every f32 rounding in the exact path.
inline float dist_jmin(float ani_th, uint32_t ksize) {
  const double a = (double)ani_th / 100.0 - 1e-4;
  if (!(a > 0.0)) return 0.0f;            // everything (ani >= 0 always) must be evaluated exactly
  if (a >= 1.0) return 1e30f;             // ani is clamped to 100: nothing can pass, J bound irrelevant
  const double E = exp((double)ksize * (a - 1.0));  // 2J/(1+J) = E
  const double J = E / (2.0 - E);
  return (float)(J * (1.0 - 1e-3));
}

}  // namespace hg
