// fasta.cu — FASTA file bytes -> merged sequence, on the GPU.
//
// Replaces fastx_reader::read_merge_seq (reference src/fastx_reader.rs:6-29), the reader of the
// reference's GPU path: every line that starts with '>' contributes one 'N' (so k-mers never span
// records), every other line contributes its bytes minus one trailing '\n' and, before it, one
// trailing '\r'.  A '\r' elsewhere is kept (it then breaks k-mers, exactly as in the reference).
//
// This is a stream compaction with one bit of carried state ("inside a header line").  A run of
// bytes is summarised by (bytes emitted if it starts outside a header, bytes emitted if it starts
// inside one, contains a line start, state handed on); such summaries compose associatively, so one
// scan of one packed word per thread resolves counts and state together.  Three kernels:
//   fasta_scan_kernel    one CTA per 4 KB block of one file: the block's summary (16 bytes per thread
//                        classified with SWAR byte compares, the header state filled bit-parallel);
//   fasta_chain_kernel   one warp per file: chains the block summaries (carry-in state and output
//                        offset of every block, merged length of the file);
//   fasta_emit_kernel    same decomposition: exclusive scan -> each thread's output offset and state,
//                        bytes compacted in shared memory, 16-byte coalesced stores.
// The merged sequence of file f is written at the file's own offset in an output buffer of the
// same size as the input (it can only be shorter), so no cross-file prefix is needed.
#include "hg_common.cuh"

namespace {

constexpr int FA_THREADS = 256;
constexpr int FA_BPT = 16;                       // bytes per thread
constexpr int FA_BLOCK = FA_THREADS * FA_BPT;    // 4096 bytes per CTA

// Per 4 KB block: how many bytes it emits if the carried state at its first byte is "outside a header"
// (c0) or "inside a header line" (c1), whether it contains a line start, and - if so - whether the last
// one opens a header line (the state it hands on).
struct BlockSum {
  uint32_t c0, c1, has_ls, last_hdr;
};

// The same four values for any run of bytes, packed in one word so that a scan moves one register:
// bits 0-13 c0, 14-27 c1 (<= 4096 each), 28 has_ls, 29 last_hdr.  Runs compose left to right.
__device__ __forceinline__ uint32_t seg_make(uint32_t c0, uint32_t c1, uint32_t has, uint32_t hdr) {
  return c0 | (c1 << 14) | (has << 28) | (hdr << 29);
}
__device__ __forceinline__ uint32_t seg_c0(uint32_t x) { return x & 0x3FFFu; }
__device__ __forceinline__ uint32_t seg_c1(uint32_t x) { return (x >> 14) & 0x3FFFu; }
__device__ __forceinline__ uint32_t seg_has(uint32_t x) { return (x >> 28) & 1u; }
__device__ __forceinline__ uint32_t seg_hdr(uint32_t x) { return (x >> 29) & 1u; }
__device__ __forceinline__ uint32_t seg_compose(uint32_t a, uint32_t b) {  // a, then b
  const uint32_t after0 = seg_has(a) ? seg_hdr(a) : 0u, after1 = seg_has(a) ? seg_hdr(a) : 1u;  // state after a
  const uint32_t c0 = seg_c0(a) + (after0 ? seg_c1(b) : seg_c0(b));
  const uint32_t c1 = seg_c1(a) + (after1 ? seg_c1(b) : seg_c0(b));
  return seg_make(c0, c1, seg_has(a) | seg_has(b), seg_has(b) ? seg_hdr(b) : seg_hdr(a));
}

// what one thread learns about its 16 bytes (bit i = byte i)
struct Lane {
  uint32_t w[4];       // the bytes
  uint32_t emit_in;    // emitted if NOT inside a header (includes the 'N' of a header start)
  uint32_t hdr_start;  // a '>' at a line start (always emits 'N')
  uint32_t ls;         // starts a line
};

// 4 flags (bit i: byte i of w equals the byte replicated in c4)
__device__ __forceinline__ uint32_t eq4(uint32_t w, uint32_t c4) {
  const uint32_t v = w ^ c4;
  const uint32_t nz = (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;  // exact: bit 7 set where the byte differs
  return ((((nz ^ 0x80808080u) >> 7) * 0x00204081u) >> 21) & 0xFu;
}
__device__ __forceinline__ uint32_t eq16(const uint32_t (&w)[4], uint32_t c4) {
  return eq4(w[0], c4) | (eq4(w[1], c4) << 4) | (eq4(w[2], c4) << 8) | (eq4(w[3], c4) << 12);
}
__device__ __forceinline__ bool any_eq16(const uint32_t (&w)[4], uint32_t c4) {  // cheap presence test
  uint32_t z = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t v = w[j] ^ c4;
    z |= (v - 0x01010101u) & ~v;
  }
  return (z & 0x80808080u) != 0;
}

// The thread's 16 bytes come in with one 16-byte load when the file is 16-byte aligned in the
// buffer (the host entry stages files that way), else with byte loads; the byte before and the byte
// after come from the neighbouring lanes (one extra byte load at the warp edges).
__device__ __forceinline__ Lane classify(const uint8_t *__restrict__ f, uint64_t len, uint64_t p0) {
  Lane L;
  L.w[0] = L.w[1] = L.w[2] = L.w[3] = 0x0A0A0A0Au;  // bytes past the end read as '\n'
  if (p0 + FA_BPT <= len && (((uintptr_t)(f + p0)) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4 *>(f + p0);
    L.w[0] = v.x; L.w[1] = v.y; L.w[2] = v.z; L.w[3] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i)
      if (p0 + i < len) L.w[i >> 2] = (L.w[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | ((uint32_t)f[p0 + i] << (8 * (i & 3)));
  }
  const int lane = threadIdx.x & 31;
  // previous byte ('\n' before the first byte of the file) and next byte ('\n' after the last)
  uint32_t prev = __shfl_up_sync(0xffffffffu, L.w[3] >> 24, 1);
  uint32_t next = __shfl_down_sync(0xffffffffu, L.w[0] & 0xFFu, 1);
  if (lane == 0) prev = (p0 == 0 || p0 > len) ? (uint32_t)'\n' : (uint32_t)f[p0 - 1];
  if (lane == 31) next = (p0 + FA_BPT < len) ? (uint32_t)f[p0 + FA_BPT] : (uint32_t)'\n';
  const uint32_t valid = p0 + FA_BPT <= len ? 0xFFFFu : (p0 < len ? ((1u << (uint32_t)(len - p0)) - 1u) : 0u);
  const uint32_t nl = eq16(L.w, 0x0A0A0A0Au);
  L.ls = ((nl << 1) | (prev == '\n' ? 1u : 0u)) & valid;
  L.hdr_start = (L.ls && any_eq16(L.w, 0x3E3E3E3Eu)) ? (L.ls & eq16(L.w, 0x3E3E3E3Eu)) : 0u;
  uint32_t keep = valid & ~nl;
  if (any_eq16(L.w, 0x0D0D0D0Du)) {  // a '\r' right before a '\n' goes with it (fastx_reader.rs:13-17)
    const uint32_t nl_next = (nl >> 1) | (next == '\n' ? 0x8000u : 0u);
    keep &= ~(eq16(L.w, 0x0D0D0D0Du) & nl_next);
  }
  L.emit_in = keep | L.hdr_start;
  return L;
}

// For the thread's 16 bytes, given whether it starts inside a header: the emit mask and the state
// after its last byte.  A header line runs from its '>' to the next line start; the "inside a header"
// bit of every byte is the hdr_start bit of the last line start at or before it (a segmented fill,
// four doubling steps), or the incoming state before the first line start.
__device__ __forceinline__ uint32_t emit_mask(const Lane &L, bool in_hdr, bool &out_hdr) {
  uint32_t h = L.hdr_start, p = ~L.ls & 0xFFFFu;
  h |= p & (h << 1); p &= p << 1;
  h |= p & (h << 2); p &= p << 2;
  h |= p & (h << 4); p &= p << 4;
  h |= p & (h << 8);
  if (in_hdr) h |= L.ls ? ((L.ls & (0u - L.ls)) - 1u) : 0xFFFFu;  // the bytes before the first line start
  out_hdr = (h >> 15) & 1u;
  return (L.hdr_start | (~h & L.emit_in)) & 0xFFFFu;
}

// the thread's bytes as a run: counts under both incoming states, line start seen, state handed on
__device__ __forceinline__ uint32_t lane_seg(const Lane &L, uint32_t &m0) {
  bool oh;
  m0 = emit_mask(L, false, oh);
  const uint32_t pre = L.ls ? ((L.ls & (0u - L.ls)) - 1u) : 0xFFFFu;
  return seg_make(__popc(m0), __popc(m0 & ~pre), L.ls != 0, oh ? 1u : 0u);
}

// inclusive scan of the runs over the CTA; returns this thread's EXCLUSIVE prefix, `total` = the whole block
__device__ __forceinline__ uint32_t block_seg_scan(uint32_t mine, uint32_t &total) {
  __shared__ uint32_t s_w[FA_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t pv = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x = seg_compose(pv, x);
  }
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  uint32_t ex = __shfl_up_sync(0xffffffffu, x, 1);
  if (lane == 0) ex = 0;  // the identity run
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < FA_THREADS / 32; ++w) {
    if (w == warp) before = all;
    all = seg_compose(all, s_w[w]);
  }
  total = all;
  return seg_compose(before, ex);
}

__global__ void __launch_bounds__(FA_THREADS)
fasta_scan_kernel(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ file_off, const uint64_t *__restrict__ blk_off,
                  BlockSum *__restrict__ sums) {
  const uint32_t f = blockIdx.y;
  const uint64_t len = file_off[f + 1] - file_off[f];
  const uint64_t b0 = (uint64_t)blockIdx.x * FA_BLOCK;
  if (b0 >= len) return;
  const Lane L = classify(raw + file_off[f], len, b0 + (uint64_t)threadIdx.x * FA_BPT);
  uint32_t m0, total;
  (void)block_seg_scan(lane_seg(L, m0), total);
  if (threadIdx.x == 0) {
    BlockSum s;
    s.c0 = seg_c0(total);
    s.c1 = seg_c1(total);
    s.has_ls = seg_has(total);
    s.last_hdr = seg_hdr(total);
    sums[blk_off[f] + blockIdx.x] = s;
  }
}

// one warp per file: sequential chain over the file's block summaries, 32 at a time
__global__ void fasta_chain_kernel(const uint64_t *__restrict__ file_off, const uint64_t *__restrict__ blk_off,
                                   uint32_t n_files, const BlockSum *__restrict__ sums, uint8_t *__restrict__ carry,
                                   uint64_t *__restrict__ out_off, uint64_t *__restrict__ merged_len) {
  const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (f >= n_files) return;
  const uint64_t nb = blk_off[f + 1] - blk_off[f];
  uint32_t state = 0;  // a file starts at a line start, so the initial state is never inherited
  uint64_t off = 0;
  for (uint64_t base = 0; base < nb; base += 32) {
    const uint64_t b = base + lane;
    BlockSum s = {0, 0, 0, 0};
    if (b < nb) s = sums[blk_off[f] + b];
    // walk the 32 summaries in order (state is a 1-bit recurrence; cheap enough serially via shuffles)
    for (int l = 0; l < 32 && base + l < nb; ++l) {
      const uint32_t c0 = __shfl_sync(0xffffffffu, s.c0, l), c1 = __shfl_sync(0xffffffffu, s.c1, l);
      const uint32_t hl = __shfl_sync(0xffffffffu, s.has_ls, l), lh = __shfl_sync(0xffffffffu, s.last_hdr, l);
      if (lane == l) {
        carry[blk_off[f] + b] = (uint8_t)state;
        out_off[blk_off[f] + b] = off;
      }
      off += state ? c1 : c0;
      if (hl) state = lh;
    }
  }
  if (lane == 0) merged_len[f] = off;
}

// Same decomposition as the scan.  The block's output bytes are compacted in shared memory, at the
// 16-byte phase of their destination, and leave with 16-byte stores (byte stores only at the two
// ragged ends, which neighbouring blocks own the rest of).
__global__ void __launch_bounds__(FA_THREADS)
fasta_emit_kernel(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ file_off, const uint64_t *__restrict__ blk_off,
                  const uint8_t *__restrict__ carry, const uint64_t *__restrict__ out_off, uint8_t *__restrict__ merged) {
  __shared__ __align__(16) uint8_t s_out[FA_BLOCK + 32];
  const uint32_t f = blockIdx.y;
  const uint64_t len = file_off[f + 1] - file_off[f];
  const uint64_t b0 = (uint64_t)blockIdx.x * FA_BLOCK;
  if (b0 >= len) return;
  const uint64_t p0 = b0 + (uint64_t)threadIdx.x * FA_BPT;
  const Lane L = classify(raw + file_off[f], len, p0);
  const bool carry_in = carry[blk_off[f] + blockIdx.x] != 0;
  uint32_t m0, total;
  const uint32_t ex = block_seg_scan(lane_seg(L, m0), total);
  bool oh;
  const uint32_t m = emit_mask(L, seg_has(ex) ? (seg_hdr(ex) != 0) : carry_in, oh);
  uint8_t *dst = merged + file_off[f] + out_off[blk_off[f] + blockIdx.x];
  const uint32_t al = (uint32_t)((uintptr_t)dst & 15);  // staged at the destination's phase
  uint32_t k = al + (carry_in ? seg_c1(ex) : seg_c0(ex));
  if (m == 0xFFFFu && L.hdr_start == 0 && (k & 3u) == 0) {  // every byte kept, word-aligned target
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint32_t *>(s_out + k + 4 * j) = L.w[j];
  } else {
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i)
      if ((m >> i) & 1u) s_out[k++] = ((L.hdr_start >> i) & 1u) ? (uint8_t)'N' : (uint8_t)(L.w[i >> 2] >> (8 * (i & 3)));
  }
  __syncthreads();
  const uint32_t n_out = carry_in ? seg_c1(total) : seg_c0(total);
  const uint32_t end = al + n_out;  // staged bytes are s_out[al, end)
  uint8_t *dst0 = dst - al;         // 16-byte aligned
  for (uint32_t c = threadIdx.x; 16 * c < end; c += FA_THREADS) {
    const uint32_t lo = 16 * c, hi = lo + 16;
    if (lo >= al && hi <= end) {
      *reinterpret_cast<uint4 *>(dst0 + lo) = *reinterpret_cast<const uint4 *>(s_out + lo);
    } else {
      for (uint32_t q = (lo > al ? lo : al); q < (hi < end ? hi : end); ++q) dst0[q] = s_out[q];
    }
  }
}

}  // namespace

uint32_t hg_fasta_block_bytes() { return FA_BLOCK; }

// d_file_off / d_blk_off: device copies of the per-file byte offsets and block prefix (n_files + 1 each)
int hg_launch_fasta_merge(hg_ctx *ctx, const uint8_t *d_raw, const uint64_t *d_file_off, const uint64_t *d_blk_off,
                          uint32_t n_files, uint32_t max_blocks, void *d_sums, uint8_t *d_carry, uint64_t *d_out_off,
                          uint8_t *d_merged, uint64_t *d_merged_len) {
  if (n_files == 0) return HG_OK;
  for (uint32_t f0 = 0; f0 < n_files; f0 += 65535) {
    const uint32_t nf = n_files - f0 < 65535 ? n_files - f0 : 65535;
    if (max_blocks)
      fasta_scan_kernel<<<dim3(max_blocks, nf), FA_THREADS, 0, ctx->stream>>>(d_raw, d_file_off + f0, d_blk_off + f0,
                                                                             (BlockSum *)d_sums);
    ctx->launches++;
  }
  fasta_chain_kernel<<<(n_files * 32 + 127) / 128, 128, 0, ctx->stream>>>(d_file_off, d_blk_off, n_files,
                                                                          (const BlockSum *)d_sums, d_carry, d_out_off,
                                                                          d_merged_len);
  ctx->launches++;
  for (uint32_t f0 = 0; f0 < n_files; f0 += 65535) {
    const uint32_t nf = n_files - f0 < 65535 ? n_files - f0 : 65535;
    if (max_blocks)
      fasta_emit_kernel<<<dim3(max_blocks, nf), FA_THREADS, 0, ctx->stream>>>(d_raw, d_file_off + f0, d_blk_off + f0,
                                                                             d_carry, d_out_off, d_merged);
    ctx->launches++;
  }
  HG_CUDA(cudaGetLastError());
  return HG_OK;
}
