// fasta.cu — FASTA file bytes -> merged sequence, on the GPU.
//
// Replaces fastx_reader::read_merge_seq (reference src/fastx_reader.rs:6-29), the reader of the
// reference's GPU path: every line that starts with '>' contributes one 'N' (so k-mers never span
// records), every other line contributes its bytes minus one trailing '\n' and, before it, one
// trailing '\r'.  A '\r' elsewhere is kept (it then breaks k-mers, exactly as in the reference).
//
// This is a stream compaction with one bit of carried state ("inside a header line").  Three
// small kernels, all HBM-bound and < 2 % of the hashing time:
//   fasta_scan_kernel    one CTA per 4 KB block of one file: how many bytes the block emits
//                        (split into the part before its first line start, which depends on the
//                        carried state, and the rest), and the state it hands on;
//   fasta_chain_kernel   one warp per file: chains the block summaries (carry-in state and output
//                        offset of every block, merged length of the file);
//   fasta_emit_kernel    same decomposition as the scan: block-wide exclusive scan of the emit
//                        flags and a coalesced-ish scatter.
// The merged sequence of file f is written at the file's own offset in an output buffer of the
// same size as the input (it can only be shorter), so no cross-file prefix is needed.
#include "hg_common.cuh"

namespace {

constexpr int FA_THREADS = 256;
constexpr int FA_BPT = 16;                       // bytes per thread
constexpr int FA_BLOCK = FA_THREADS * FA_BPT;    // 4096 bytes per CTA

struct BlockSum {      // per 4 KB block
  uint32_t cnt_pre;    // bytes emitted before the first line start, if the block starts outside a header
  uint32_t cnt_post;   // bytes emitted from the first line start on (independent of the carried state)
  uint32_t has_ls;     // the block contains a line start
  uint32_t last_hdr;   // ... and the last one opens a header line
};

// what one thread learns about its 16 bytes
struct Lane {
  uint32_t emit_in;    // bit i: byte i is emitted if it is NOT inside a header (includes 'N' of a header start)
  uint32_t hdr_start;  // bit i: byte i is a '>' at a line start (always emits 'N')
  uint32_t ls;         // bit i: byte i starts a line
};

// The thread's 16 bytes come in with one 16-byte load when the file is 16-byte aligned in the
// buffer (the host entry stages files that way), else with byte loads; the byte before and the byte
// after come from the neighbouring lanes (one extra byte load at the warp edges).
__device__ __forceinline__ Lane classify(const uint8_t *__restrict__ f, uint64_t len, uint64_t p0) {
  uint32_t w[4] = {0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au};  // bytes past the end read as '\n'
  if (p0 + FA_BPT <= len && (((uintptr_t)(f + p0)) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4 *>(f + p0);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < FA_BPT; ++i)
      if (p0 + i < len) w[i >> 2] = (w[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | ((uint32_t)f[p0 + i] << (8 * (i & 3)));
  }
  const int lane = threadIdx.x & 31;
  // previous byte ('\n' before the first byte of the file) and next byte ('\n' after the last)
  uint32_t prev = __shfl_up_sync(0xffffffffu, w[3] >> 24, 1);
  uint32_t next = __shfl_down_sync(0xffffffffu, w[0] & 0xFFu, 1);
  if (lane == 0) prev = (p0 == 0 || p0 > len) ? (uint32_t)'\n' : (uint32_t)f[p0 - 1];
  if (lane == 31) next = (p0 + FA_BPT < len) ? (uint32_t)f[p0 + FA_BPT] : (uint32_t)'\n';
  Lane L = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < FA_BPT; ++i) {
    const uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
    const uint32_t pc = i == 0 ? prev : ((w[(i - 1) >> 2] >> (8 * ((i - 1) & 3))) & 0xFFu);
    const uint32_t nc = i == FA_BPT - 1 ? next : ((w[(i + 1) >> 2] >> (8 * ((i + 1) & 3))) & 0xFFu);
    const bool inside = p0 + i < len;
    const bool ls = inside && pc == '\n';
    const bool hs = ls && c == '>';
    const bool keep = inside && c != '\n' && !(c == '\r' && nc == '\n');
    L.ls |= (uint32_t)ls << i;
    L.hdr_start |= (uint32_t)hs << i;
    L.emit_in |= (uint32_t)(keep || hs) << i;
  }
  return L;
}

// For the thread's 16 bytes, given whether it starts inside a header: the emit mask and the state
// after its last byte.  A header line runs from its '>' to the next line start.
__device__ __forceinline__ uint32_t emit_mask(const Lane &L, bool in_hdr, bool &out_hdr) {
  uint32_t m = 0;
  bool h = in_hdr;
#pragma unroll
  for (int i = 0; i < FA_BPT; ++i) {
    if ((L.ls >> i) & 1u) h = (L.hdr_start >> i) & 1u;
    const bool e = ((L.hdr_start >> i) & 1u) || (!h && ((L.emit_in >> i) & 1u));
    m |= (uint32_t)e << i;
  }
  out_hdr = h;
  return m;
}

// block-wide exclusive scan of (has_ls, last_hdr) "last writer wins" and of a count
__device__ __forceinline__ void block_scan(uint32_t my_has, uint32_t my_hdr, uint32_t my_cnt, uint32_t &ex_has,
                                           uint32_t &ex_hdr, uint32_t &ex_cnt, uint32_t &tot_has, uint32_t &tot_hdr,
                                           uint32_t &tot_cnt) {
  __shared__ uint32_t s_has[FA_THREADS / 32], s_hdr[FA_THREADS / 32], s_cnt[FA_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t has = my_has, hdr = my_hdr, cnt = my_cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t ph = __shfl_up_sync(0xffffffffu, has, o), pd = __shfl_up_sync(0xffffffffu, hdr, o),
                   pc = __shfl_up_sync(0xffffffffu, cnt, o);
    if (lane >= o) {
      if (!has) { has = ph; hdr = pd; }
      cnt += pc;
    }
  }
  if (lane == 31) { s_has[warp] = has; s_hdr[warp] = hdr; s_cnt[warp] = cnt; }
  __syncthreads();
  uint32_t bh = 0, bd = 0, bc = 0;  // everything in the warps before mine
  for (int w = 0; w < warp; ++w) {
    if (s_has[w]) { bh = 1; bd = s_hdr[w]; }
    bc += s_cnt[w];
  }
  // exclusive = inclusive of the previous lane, combined with the previous warps
  uint32_t eh = __shfl_up_sync(0xffffffffu, has, 1), ed = __shfl_up_sync(0xffffffffu, hdr, 1),
           ec = __shfl_up_sync(0xffffffffu, cnt, 1);
  if (lane == 0) { eh = 0; ed = 0; ec = 0; }
  ex_has = eh | bh;
  ex_hdr = eh ? ed : bd;
  ex_cnt = ec + bc;
  tot_has = 0; tot_hdr = 0; tot_cnt = 0;
  for (int w = 0; w < FA_THREADS / 32; ++w) {
    if (s_has[w]) { tot_has = 1; tot_hdr = s_hdr[w]; }
    tot_cnt += s_cnt[w];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FA_THREADS)
fasta_scan_kernel(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ file_off, const uint64_t *__restrict__ blk_off,
                  BlockSum *__restrict__ sums) {
  const uint32_t f = blockIdx.y;
  const uint64_t len = file_off[f + 1] - file_off[f];
  const uint64_t b0 = (uint64_t)blockIdx.x * FA_BLOCK;
  if (b0 >= len) return;
  const uint8_t *fp = raw + file_off[f];
  const Lane L = classify(fp, len, b0 + (uint64_t)threadIdx.x * FA_BPT);
  // two hypotheses for the bytes before the block's first line start: outside / inside a header.
  // After the first line start the carried state no longer matters.
  bool oh;
  const uint32_t m_out = emit_mask(L, false, oh);
  const uint32_t my_has = L.ls != 0, my_hdr = oh ? 1u : 0u;
  // bytes of this thread before its own first line start (they inherit the carried state)
  const uint32_t first_ls = L.ls ? (uint32_t)(__ffs(L.ls) - 1) : FA_BPT;
  const uint32_t pre_mask = first_ls >= 32 ? 0xffffffffu : ((1u << first_ls) - 1u);
  uint32_t ex_has, ex_hdr, ex_cnt, th, td, tc;
  // count 1: emitted bytes assuming "outside header" wherever the state is inherited from before the block
  (void)m_out;
  block_scan(my_has, my_hdr, 0u, ex_has, ex_hdr, ex_cnt, th, td, tc);
  // a thread's bytes are "pre" (before the block's first line start) when no earlier thread has a line start
  // and they precede its own first line start; with the true in-block state for later threads:
  bool oh2;
  const uint32_t m_true = emit_mask(L, ex_has ? (ex_hdr != 0) : false, oh2);
  const uint32_t pre_bytes = ex_has ? 0u : (uint32_t)__popc(m_true & pre_mask);
  const uint32_t all_bytes = (uint32_t)__popc(m_true);
  uint32_t a, b, c2, t1h, t1d, pre_tot, d1, d2, d3, all_tot;
  block_scan(0, 0, pre_bytes, a, b, c2, t1h, t1d, pre_tot);
  block_scan(0, 0, all_bytes, d1, d2, d3, t1h, t1d, all_tot);
  if (threadIdx.x == 0) {
    BlockSum s;
    s.cnt_pre = pre_tot;
    s.cnt_post = all_tot - pre_tot;
    s.has_ls = th;
    s.last_hdr = td;
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
      const uint32_t cp = __shfl_sync(0xffffffffu, s.cnt_pre, l), cq = __shfl_sync(0xffffffffu, s.cnt_post, l);
      const uint32_t hl = __shfl_sync(0xffffffffu, s.has_ls, l), lh = __shfl_sync(0xffffffffu, s.last_hdr, l);
      if (lane == l) {
        carry[blk_off[f] + b] = (uint8_t)state;
        out_off[blk_off[f] + b] = off;
      }
      off += (state ? 0u : cp) + cq;
      if (hl) state = lh;
    }
  }
  if (lane == 0) merged_len[f] = off;
}

__global__ void __launch_bounds__(FA_THREADS)
fasta_emit_kernel(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ file_off, const uint64_t *__restrict__ blk_off,
                  const uint8_t *__restrict__ carry, const uint64_t *__restrict__ out_off, uint8_t *__restrict__ merged) {
  const uint32_t f = blockIdx.y;
  const uint64_t len = file_off[f + 1] - file_off[f];
  const uint64_t b0 = (uint64_t)blockIdx.x * FA_BLOCK;
  if (b0 >= len) return;
  const uint8_t *fp = raw + file_off[f];
  const uint64_t p0 = b0 + (uint64_t)threadIdx.x * FA_BPT;
  const Lane L = classify(fp, len, p0);
  const bool carry_in = carry[blk_off[f] + blockIdx.x] != 0;
  bool oh;
  (void)emit_mask(L, false, oh);
  uint32_t ex_has, ex_hdr, ex_cnt, th, td, tc;
  block_scan(L.ls != 0, oh ? 1u : 0u, 0, ex_has, ex_hdr, ex_cnt, th, td, tc);
  bool oh2;
  const uint32_t m = emit_mask(L, ex_has ? (ex_hdr != 0) : carry_in, oh2);
  uint32_t a, b, pos, t1, t2, t3;
  block_scan(0, 0, (uint32_t)__popc(m), a, b, pos, t1, t2, t3);
  uint8_t *dst = merged + file_off[f] + out_off[blk_off[f] + blockIdx.x] + pos;
  uint32_t k = 0;
#pragma unroll
  for (int i = 0; i < FA_BPT; ++i)
    if ((m >> i) & 1u) dst[k++] = ((L.hdr_start >> i) & 1u) ? (uint8_t)'N' : fp[p0 + i];
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
