// Output stage of `dist` on the device: put the hit records in the order the reference prints them.
//
// Reference: utils::dump_ani_file (src/utils.rs:260-285) sorts ALL pair indices ascending by ANI with
// a stable sort and then reverses the vector, so the TSV runs ANI-descending and equal ANIs appear in
// DESCENDING pair-enumeration index.  Both enumerations (upper triangle row by row, src/dist.rs:243-265,
// and the full R x Q grid) are lexicographic in (i, j), so the order is: ani desc, then i desc, then j desc.
// Only pairs with ani >= ani_th survive the dist kernel, so this sorts thousands to millions of 16-byte
// records instead of the reference's 5e7-2e8 index entries.
//
// Method: least-significant-digit radix sort over the 96-bit key (ani bits | i | j), 8 bits per pass,
// digit = 255 - byte for the descending order.  ANI is a non-negative f32, so its bit pattern orders
// like the value.  A first kernel ORs (key ^ key[0]) over all records; byte positions where every
// record agrees are skipped (for 10 000 sketches and ani_th = 85 that leaves 7 of 12 passes).
// Each pass: per-block digit counts -> exclusive scan (digit-major, then block) -> stable scatter.
// In the scatter every warp owns a contiguous run (128 or 512 records) of the block's chunk and walks it in order,
// ranking equal digits inside a 32-record round with match_any, so no block barrier sits in the loop.
//
// Also here: the `{:.3}` field of the TSV as an integer (thousandths), see ani_milli_kernel.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "hg_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_PASSES = 12;
// ROUNDS = 32-record rounds per warp: a block owns 256 * ROUNDS records (4096; 1024 for very small inputs)

// byte `pass` of the key, least significant first: j (0-3), i (4-7), ani bits (8-11); complemented
__device__ __forceinline__ uint32_t rs_digit(const uint4 &h, int pass) {
  const uint32_t w = pass < 4 ? h.y : (pass < 8 ? h.x : h.w);  // hg_hit = {i, j, dot, ani}
  return 255u - ((w >> ((pass & 3) * 8)) & 255u);
}

// which key bytes differ anywhere: diff[0..2] |= key ^ key[0]  (words: j, i, ani)
__global__ void rs_diff_kernel(const uint4 *__restrict__ hits, uint64_t n, uint32_t *__restrict__ diff) {
  const uint4 h0 = hits[0];
  uint32_t dj = 0, di = 0, da = 0;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 h = hits[t];
    dj |= h.y ^ h0.y;
    di |= h.x ^ h0.x;
    da |= h.w ^ h0.w;
  }
  dj = __reduce_or_sync(0xffffffffu, dj);
  di = __reduce_or_sync(0xffffffffu, di);
  da = __reduce_or_sync(0xffffffffu, da);
  if ((threadIdx.x & 31) == 0) {
    if (dj) atomicOr(diff + 0, dj);
    if (di) atomicOr(diff + 1, di);
    if (da) atomicOr(diff + 2, da);
  }
}

// counts[d * n_blocks + b] = records of block b's chunk with digit d
template <int ROUNDS>
__global__ void __launch_bounds__(RS_THREADS)
rs_count_kernel(const uint4 *__restrict__ hits, uint64_t n, int pass, uint32_t n_blocks, uint32_t *__restrict__ counts) {
  constexpr int RS_CHUNK = RS_THREADS * ROUNDS;
  __shared__ uint32_t s_cnt[256];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * RS_CHUNK;
#pragma unroll 4
  for (int r = 0; r < RS_CHUNK / RS_THREADS; ++r) {
    const uint64_t t = base + (uint64_t)r * RS_THREADS + threadIdx.x;
    if (t < n) atomicAdd(&s_cnt[rs_digit(hits[t], pass)], 1u);
  }
  __syncthreads();
  counts[(uint64_t)threadIdx.x * n_blocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// in-place exclusive scan of `len` counters by one block (len = 256 * n_blocks; small next to the records):
// every thread owns a contiguous run, so there is one block-wide scan (two barriers) whatever the length
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *__restrict__ counts, uint64_t len) {
  __shared__ uint32_t s_warp[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t items = (len + 1023) / 1024;
  const uint64_t b = (uint64_t)threadIdx.x * items;
  const uint64_t e = b + items < len ? b + items : len;
  uint32_t sum = 0;
#pragma unroll 4
  for (uint64_t i = b; i < e; ++i) sum += counts[i];
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = s_warp[lane];
    uint32_t ws = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= (uint32_t)o) ws += y;
    }
    s_warp[lane] = ws - w;  // exclusive over warps
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + x - sum;
  for (uint64_t i = b; i < e; ++i) {
    const uint32_t v = counts[i];
    counts[i] = run;
    run += v;
  }
}

// stable scatter of block b's chunk to the scanned offsets
template <int RS_ROUNDS>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint64_t n, int pass, uint32_t n_blocks,
                  const uint32_t *__restrict__ offsets) {
  constexpr int RS_WARP_ITEMS = 32 * RS_ROUNDS, RS_CHUNK = RS_WARPS * RS_WARP_ITEMS;
  __shared__ uint32_t s_base[RS_WARPS][256];  // next output slot per (warp, digit)
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = 0; w < RS_WARPS; ++w) s_base[w][threadIdx.x] = 0;
  __syncthreads();
  const uint64_t wbase = (uint64_t)blockIdx.x * RS_CHUNK + (uint64_t)warp * RS_WARP_ITEMS;
  uint4 rec[RS_ROUNDS];
  // the warp's digit counts
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const uint64_t t = wbase + (uint64_t)r * 32 + lane;
    if (t < n) {
      rec[r] = in[t];
      atomicAdd(&s_base[warp][rs_digit(rec[r], pass)], 1u);
    }
  }
  __syncthreads();
  {  // digit d = threadIdx.x: running start per warp, in warp order
    uint32_t run = offsets[(uint64_t)threadIdx.x * n_blocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = s_base[w][threadIdx.x];
      s_base[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const uint64_t t = wbase + (uint64_t)r * 32 + lane;
    const bool live = t < n;
    const uint32_t d = live ? rs_digit(rec[r], pass) : 256u + lane;  // dead lanes match nobody
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t before = __popc(peers & ((1u << lane) - 1u));
    uint32_t slot = 0;
    if (live) slot = s_base[warp][d] + before;
    __syncwarp();
    if (live && before == 0) s_base[warp][d] += (uint32_t)__popc(peers);
    __syncwarp();
    if (live) out[slot] = rec[r];
  }
}

// ---- the whole sort as ONE cooperative kernel (small and medium inputs) -------------------------
// The multi-kernel path above spends its time in launch gaps: at 185 k records every one of its 22 kernels is a
// 10-15 us affair on 46 blocks.  Here every block keeps its positional chunk of 256 * ROUNDS records in registers,
// and the passes are separated by grid-wide barriers instead of kernel boundaries: which key bytes are live
// (rs_diff), per pass the block histograms -> barrier -> every block derives its own scatter offsets from the
// count matrix (no single-block scan) -> stable scatter -> barrier -> reload the chunk.  The last phase copies
// back if the records ended in the ping-pong buffer and writes the `{:.3}` thousandths.  Data written by other
// blocks is read with ld.global.cg: L1 may still hold the lines of an earlier pass.
template <int ROUNDS>
__global__ void __launch_bounds__(RS_THREADS)
rs_fused_kernel(uint4 *__restrict__ a, uint4 *__restrict__ b, uint64_t n, uint32_t *__restrict__ counts,
                uint32_t *__restrict__ diff, uint32_t *__restrict__ milli) {
  constexpr int WARP_ITEMS = 32 * ROUNDS, CHUNK = RS_WARPS * WARP_ITEMS;
  __shared__ uint32_t s_base[RS_WARPS][256];  // per (warp, digit): count, then next output slot
  __shared__ uint32_t s_scan[RS_WARPS];
  cg::grid_group grid = cg::this_grid();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_blocks = gridDim.x;
  const uint64_t wbase = (uint64_t)blockIdx.x * CHUNK + (uint64_t)warp * WARP_ITEMS;
  uint4 rec[ROUNDS];
  // which key bytes differ anywhere
  {
    const uint4 h0 = __ldcg(a);
    uint32_t dj = 0, di = 0, da = 0;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint64_t t = wbase + (uint64_t)r * 32 + lane;
      if (t < n) {
        rec[r] = __ldcg(a + t);
        dj |= rec[r].y ^ h0.y;
        di |= rec[r].x ^ h0.x;
        da |= rec[r].w ^ h0.w;
      }
    }
    dj = __reduce_or_sync(0xffffffffu, dj);
    di = __reduce_or_sync(0xffffffffu, di);
    da = __reduce_or_sync(0xffffffffu, da);
    if (lane == 0) {
      if (dj) atomicOr(diff + 0, dj);
      if (di) atomicOr(diff + 1, di);
      if (da) atomicOr(diff + 2, da);
    }
  }
  grid.sync();
  const uint32_t dmask[3] = {__ldcg(diff + 0), __ldcg(diff + 1), __ldcg(diff + 2)};
  uint4 *src = a, *dst = b;
  for (int pass = 0; pass < RS_PASSES; ++pass) {
    if (((dmask[pass >> 2] >> ((pass & 3) * 8)) & 255u) == 0) continue;  // every record has the same byte here (grid-uniform)
    // per-warp digit histograms of my chunk
    for (int w = 0; w < RS_WARPS; ++w) s_base[w][threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint64_t t = wbase + (uint64_t)r * 32 + lane;
      if (t < n) atomicAdd(&s_base[warp][rs_digit(rec[r], pass)], 1u);
    }
    __syncthreads();
    {
      uint32_t c = 0;
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) c += s_base[w][threadIdx.x];
      counts[(uint64_t)blockIdx.x * 256 + threadIdx.x] = c;  // block-major: the reads below are coalesced
    }
    grid.sync();
    // my scatter offsets: everything with a smaller digit + my digit in the blocks before me
    uint32_t tot = 0, pre = 0;
    {
      const uint32_t *col = counts + threadIdx.x;
#pragma unroll 8
      for (uint32_t bb = 0; bb < n_blocks; ++bb) {
        const uint32_t c = __ldcg(col + (uint64_t)bb * 256);
        pre += bb < blockIdx.x ? c : 0u;
        tot += c;
      }
    }
    uint32_t x = tot;  // exclusive scan of tot over the 256 digits (thread = digit)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) before += (uint32_t)w < warp ? s_scan[w] : 0u;
    {
      uint32_t run = before + x - tot + pre;  // first slot of (my block, digit threadIdx.x); then warp by warp
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = s_base[w][threadIdx.x];
        s_base[w][threadIdx.x] = run;
        run += c;
      }
    }
    __syncthreads();
    // stable scatter: every warp walks its run in order, equal digits of a round ranked with match_any
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint64_t t = wbase + (uint64_t)r * 32 + lane;
      const bool live = t < n;
      const uint32_t d = live ? rs_digit(rec[r], pass) : 256u + lane;  // dead lanes match nobody
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const uint32_t ahead = __popc(peers & ((1u << lane) - 1u));
      uint32_t slot = 0;
      if (live) slot = s_base[warp][d] + ahead;
      __syncwarp();
      if (live && ahead == 0) s_base[warp][d] += (uint32_t)__popc(peers);
      __syncwarp();
      if (live) dst[slot] = rec[r];
    }
    grid.sync();
    { uint4 *tmp = src; src = dst; dst = tmp; }
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {  // my positional chunk of the new order
      const uint64_t t = wbase + (uint64_t)r * 32 + lane;
      if (t < n) rec[r] = __ldcg(src + t);
    }
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint64_t t = wbase + (uint64_t)r * 32 + lane;
    if (t < n) {
      if (src != a) a[t] = rec[r];
      if (milli) milli[t] = (uint32_t)__double2ll_rn((double)__uint_as_float(rec[r].w) * 1000.0);  // see ani_milli_kernel
    }
  }
}

// thousandths of the ANI, rounded as `{:.3}` rounds: ani * 1000 is exact in binary64 (24-bit x 10-bit
// significands), so rint() - round half to even - is the correctly rounded decimal
__global__ void ani_milli_kernel(const uint4 *__restrict__ hits, uint64_t n, uint32_t *__restrict__ milli) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) milli[t] = (uint32_t)__double2ll_rn((double)__uint_as_float(hits[t].w) * 1000.0);
}

}  // namespace

// Sorts d_hits[0..n) in place (ani desc, i desc, j desc).  One small D2H + sync to learn the live digits.
int hg_launch_sort_hits(hg_ctx *ctx, hg_hit *d_hits, uint64_t n, uint32_t *d_milli) {
  static_assert(sizeof(hg_hit) == 16, "hg_hit must be 16 bytes");
  if (n == 0) return HG_OK;
  if (n > 0xffffffffull) { hg_set_error("hg_sort_hits: more than 2^32 - 1 records"); return HG_E_UNSUPPORTED; }
  int rc;
  const int rounds = n <= (1u << 15) ? 4 : 16;  // measured: at 185 k records the longer single-block scan of 1 KiB blocks costs more than it wins
  const uint32_t chunk = (uint32_t)RS_THREADS * rounds;
  const uint32_t n_blocks = (uint32_t)((n + chunk - 1) / chunk);
  void *d_tmp, *d_cnt;
  if (n > 1) {
    // small and medium inputs: the whole sort (and the thousandths) as one cooperative kernel
    if (!getenv("HG_SORT_MULTI")) {
      // few blocks matter more than short blocks (measured at 185 k records: 46 blocks of 4096 sort in 0.17 ms,
      // 181 blocks of 1024 in 0.48 ms, the multi-kernel path in 0.26 ms): 16 rounds per warp unless the input is tiny
      int only = n <= (1u << 14) ? 4 : 16;
      if (const char *e = getenv("HG_SORT_ROUNDS")) only = atoi(e);
      for (int fr : {4, 16}) {
        if (fr != only) continue;
        const uint32_t fchunk = (uint32_t)RS_THREADS * fr;
        const uint64_t fblocks = (n + fchunk - 1) / fchunk;
        int per_sm = 0;
        if (fr == 4) HG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rs_fused_kernel<4>, RS_THREADS, 0));
        else HG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rs_fused_kernel<16>, RS_THREADS, 0));
        if (fblocks > (uint64_t)per_sm * ctx->sm_count) continue;  // all blocks must be resident for the grid barriers
        if ((rc = hg_scratch(ctx, HG_S_SORT_TMP, n * sizeof(hg_hit), &d_tmp))) return rc;
        if ((rc = hg_scratch(ctx, HG_S_SORT_CNT, (size_t)256 * fblocks * 4 + 256, &d_cnt))) return rc;
        uint32_t *d_diff = (uint32_t *)d_cnt + (size_t)256 * fblocks;
        HG_CUDA(cudaMemsetAsync(d_diff, 0, 16, ctx->stream));
        uint4 *pa = (uint4 *)d_hits, *pb = (uint4 *)d_tmp;
        uint64_t nn = n;
        uint32_t *pc = (uint32_t *)d_cnt, *pm = d_milli;
        void *args[] = {&pa, &pb, &nn, &pc, &d_diff, &pm};
        if (fr == 4) HG_CUDA(cudaLaunchCooperativeKernel((void *)rs_fused_kernel<4>, dim3((unsigned)fblocks), dim3(RS_THREADS), args, 0, ctx->stream));
        else HG_CUDA(cudaLaunchCooperativeKernel((void *)rs_fused_kernel<16>, dim3((unsigned)fblocks), dim3(RS_THREADS), args, 0, ctx->stream));
        ctx->launches++;
        HG_CUDA(cudaGetLastError());
        return HG_OK;
      }
    }
    if ((rc = hg_scratch(ctx, HG_S_SORT_TMP, n * sizeof(hg_hit), &d_tmp))) return rc;
    if ((rc = hg_scratch(ctx, HG_S_SORT_CNT, (size_t)256 * n_blocks * 4 + 256, &d_cnt))) return rc;
    uint32_t *d_diff = (uint32_t *)d_cnt;  // first 3 words, read back before the counters are used
    HG_CUDA(cudaMemsetAsync(d_diff, 0, 16, ctx->stream));
    const uint32_t dgrid = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
    rs_diff_kernel<<<dgrid, 256, 0, ctx->stream>>>((const uint4 *)d_hits, n, d_diff);
    ctx->launches++;
    uint32_t diff[3] = {0, 0, 0};
    HG_CUDA(cudaMemcpyAsync(diff, d_diff, sizeof(diff), cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    uint4 *src = (uint4 *)d_hits, *dst = (uint4 *)d_tmp;
    for (int pass = 0; pass < RS_PASSES; ++pass) {
      if (((diff[pass >> 2] >> ((pass & 3) * 8)) & 255u) == 0) continue;  // every record has the same byte here
      if (rounds == 4) rs_count_kernel<4><<<n_blocks, RS_THREADS, 0, ctx->stream>>>(src, n, pass, n_blocks, (uint32_t *)d_cnt);
      else rs_count_kernel<16><<<n_blocks, RS_THREADS, 0, ctx->stream>>>(src, n, pass, n_blocks, (uint32_t *)d_cnt);
      rs_scan_kernel<<<1, 1024, 0, ctx->stream>>>((uint32_t *)d_cnt, (uint64_t)256 * n_blocks);
      if (rounds == 4)
        rs_scatter_kernel<4><<<n_blocks, RS_THREADS, 0, ctx->stream>>>(src, dst, n, pass, n_blocks, (const uint32_t *)d_cnt);
      else
        rs_scatter_kernel<16><<<n_blocks, RS_THREADS, 0, ctx->stream>>>(src, dst, n, pass, n_blocks, (const uint32_t *)d_cnt);
      ctx->launches += 3;
      std::swap(src, dst);
    }
    if (src != (uint4 *)d_hits)
      HG_CUDA(cudaMemcpyAsync(d_hits, src, n * sizeof(hg_hit), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (d_milli) {
    ani_milli_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint4 *)d_hits, n, d_milli);
    ctx->launches++;
  }
  HG_CUDA(cudaGetLastError());
  return HG_OK;
}
