// kmer_hash.cu — stage 1a: canonical k-mer extraction, t1ha2 hashing, FracMinHash filter and
// set insertion for a batch of genomes, in one pass over the sequence bytes.
//
// Replaces cuda_kmer_t1ha2 (reference src/cuda_kernel.cu:250-321) and its host-side result
// scan (src/sketch_cuda.rs:156-163); the result equals the CPU path's HashSet
// (src/sketch.rs:71-98).  Design (not the reference's one-thread-per-512-k-mers walk):
//   * one CTA per tile of 8192 k-mer start positions of ONE genome; the tile's bytes
//     (+ k-1 halo) are read once with aligned, coalesced 16-byte loads, converted SWAR-style
//     to 2-bit codes + a validity bit per base, and staged in 3 KB of shared memory;
//   * each thread then owns 32 consecutive start positions: rolling forward / reverse-
//     complement 2-bit k-mers in registers, canonical = min() on the packed value (same order
//     as the reference's byte compare, cuda_kernel.cu:84-89,306-311), expansion of the chosen
//     k-mer back to its upper-case ASCII bytes with PRMT table lookups (the hash is defined
//     over the ASCII k-mer, sketch.rs:90), t1ha2_atonce in registers, threshold compare;
//   * the rare survivors (1/scaled) are inserted into the genome's open-addressing table in
//     HBM with atomicCAS, which both removes duplicates (the reference keeps a set) and has
//     no per-thread capacity cliff (the reference drops hits beyond 8 per thread,
//     cuda_kernel.cu:316-317).
#include "hg_common.cuh"

namespace {

constexpr int KH_THREADS = 256;
constexpr int KH_PPT = 32;                       // k-mer start positions per thread
constexpr int KH_TILE = KH_THREADS * KH_PPT;     // start positions per CTA
constexpr int KH_CHUNKS = (KH_TILE + 64) / 16;   // 16-base chunks staged per tile (halo + align)

__device__ __forceinline__ uint4 ld_stream16(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// 4 ASCII bytes -> 8 bits of 2-bit codes (A0 C1 G2 T3, base 0 in the low bits) and 4
// validity bits (only A,C,G,T in either case are valid — cuda_kernel.cu:277-296).
__device__ __forceinline__ void encode4(uint32_t w, uint32_t &codes8, uint32_t &valid4) {
  const uint32_t u = w & 0xDFDFDFDFu;                               // fold case
  const uint32_t code = ((u >> 1) ^ (u >> 2)) & 0x03030303u;        // per byte: A0 C1 G2 T3
  codes8 = (code * 0x01041040u) >> 24;                              // gather the four fields
  // re-expand the codes to upper-case ASCII and compare: anything else is not a base
  const uint32_t sel = (code | (code >> 12)) & 0xFFFFu;             // nibbles: b0 b2 b1 b3
  const uint32_t expect = __byte_perm(0x54474341u, 0u, sel);        // 'A','C','G','T'
  const uint32_t have = __byte_perm(u, 0u, 0x3120u);                // same byte order
  const uint32_t d = expect ^ have;
  const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
  const uint32_t ok = (nz ^ 0x80808080u) >> 7;                      // flags at bits 0,8,16,24
  valid4 = ((ok * 0x01040208u) >> 24) & 0xFu;                       // back to base order
}

__device__ __forceinline__ uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
  return __funnelshift_r(lo, hi, s);
}

// Insert h into the genome's table (linear probing).  Duplicates collapse: F3 set semantics.
__device__ __noinline__ void table_insert(uint64_t *__restrict__ table, uint32_t mask, uint64_t h,
                                          uint32_t *count, uint32_t *status) {
  uint32_t slot = (uint32_t)h & mask;
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    const unsigned long long prev =
        atomicCAS((unsigned long long *)(table + slot), (unsigned long long)HG_EMPTY_SLOT,
                  (unsigned long long)h);
    if (prev == HG_EMPTY_SLOT) {
      atomicAdd(count, 1u);
      return;
    }
    if (prev == h) return;
    slot = (slot + 1) & mask;
  }
  atomicOr(status, 1u);  // table full: reported by hg_sketch_status, never silently dropped
}

template <int K, bool CANON>
__global__ void __launch_bounds__(KH_THREADS)
kmer_hash_kernel(const uint8_t *__restrict__ seq, const hg_genome_desc *__restrict__ desc,
                 uint32_t n_genomes, uint64_t threshold, uint64_t seed,
                 uint64_t *__restrict__ tables, uint32_t *__restrict__ counts,
                 uint32_t *__restrict__ status) {
  constexpr int NW = (K + 7) / 8;
  __shared__ uint32_t s_codes[KH_CHUNKS + 4];
  __shared__ uint32_t s_valid[KH_CHUNKS / 2 + 4];  // 16 bits per chunk
  __shared__ uint32_t s_genome;

  // ---- which genome does this tile belong to (binary search over first_tile) ----
  if (threadIdx.x == 0) {
    uint32_t lo = 0, hi = n_genomes - 1;
    while (lo < hi) {
      const uint32_t mid = (lo + hi + 1) >> 1;
      if (desc[mid].first_tile <= blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_genome = lo;
  }
  __syncthreads();
  const uint32_t g = s_genome;
  const hg_genome_desc gd = desc[g];
  const uint64_t tile_base = (uint64_t)(blockIdx.x - gd.first_tile) * KH_TILE;  // first start position
  uint64_t need_end = tile_base + KH_TILE + (K - 1);                            // bases [tile_base, need_end)
  if (need_end > gd.seq_len) need_end = gd.seq_len;

  // ---- phase A: bytes -> 2-bit codes + validity, staged in shared memory ----
  const uintptr_t p_lo = (uintptr_t)(seq + gd.seq_begin + tile_base);
  const uintptr_t p_hi = (uintptr_t)(seq + gd.seq_begin + need_end);
  const uintptr_t p_al = p_lo & ~(uintptr_t)15;
  const uint32_t sh = (uint32_t)(p_lo - p_al);  // 0..15 bases of alignment slack
  uint16_t *s_valid16 = reinterpret_cast<uint16_t *>(s_valid);
  for (int c = threadIdx.x; c < KH_CHUNKS + 4; c += KH_THREADS) {
    const uintptr_t pc = p_al + (uintptr_t)16 * c;
    uint32_t codes = 0, valid = 0;
    if (c < KH_CHUNKS && pc < p_hi && pc + 16 > p_lo) {
      const uint4 v = ld_stream16(reinterpret_cast<const void *>(pc));
      uint32_t c8, v4;
      encode4(v.x, c8, v4); codes = c8;        valid = v4;
      encode4(v.y, c8, v4); codes |= c8 << 8;  valid |= v4 << 4;
      encode4(v.z, c8, v4); codes |= c8 << 16; valid |= v4 << 8;
      encode4(v.w, c8, v4); codes |= c8 << 24; valid |= v4 << 12;
      // bytes of this chunk outside [p_lo, p_hi) belong to someone else (or nobody)
      const int first = pc < p_lo ? (int)(p_lo - pc) : 0;
      const int last = pc + 16 > p_hi ? (int)(p_hi - pc) : 16;
      valid &= ((1u << last) - 1u) & ~((1u << first) - 1u);
    }
    s_codes[c] = codes;
    s_valid16[c] = (uint16_t)valid;
  }
  __syncthreads();

  // ---- phase B: 32 start positions per thread ----
  const int t = threadIdx.x;
  // 64 bases of codes starting at this thread's first base
  uint32_t cw[4];
  {
    const uint32_t a0 = s_codes[2 * t], a1 = s_codes[2 * t + 1], a2 = s_codes[2 * t + 2],
                   a3 = s_codes[2 * t + 3], a4 = s_codes[2 * t + 4];
    const uint32_t s2 = 2 * sh;
    cw[0] = funnel_r(a0, a1, s2); cw[1] = funnel_r(a1, a2, s2);
    cw[2] = funnel_r(a2, a3, s2); cw[3] = funnel_r(a3, a4, s2);
  }
  uint64_t v;
  {
    const uint32_t b0 = s_valid[t], b1 = s_valid[t + 1], b2 = s_valid[t + 2];
    v = (uint64_t)funnel_r(b0, b1, sh) | ((uint64_t)funnel_r(b1, b2, sh) << 32);
  }
  // kv bit j: all K bases of the k-mer starting at thread-local position j are valid
  uint64_t kv = ~0ull;
  {
    uint64_t r = v;
    int done = 0;
#pragma unroll
    for (int len = 1; len <= 32; len <<= 1) {
      if (K & len) { kv &= (r >> done); done += len; }
      r &= r >> len;
    }
  }
  const uint32_t kv32 = (uint32_t)kv;
  if (kv32 == 0) return;

  constexpr uint64_t KMASK = (K == 32) ? ~0ull : ((1ull << (2 * K)) - 1ull);
  const uint64_t codes_lo = (uint64_t)cw[0] | ((uint64_t)cw[1] << 32);  // bases 0..31, LSB first

  // rolling state after the first K-1 bases: fwd is MSB-first (oldest base on top) so that an
  // unsigned compare is the lexicographic compare; rc holds the reverse complement likewise.
  uint64_t fwd = 0, rc = 0;
  if constexpr (K > 1) {
    constexpr uint64_t PMASK = (K == 1) ? 0ull : ((1ull << (2 * (K - 1))) - 1ull);
    const uint64_t x = codes_lo & PMASK;                       // bases 0..K-2, base i at bits 2i
    rc = ((~x) & PMASK) << 2;                                  // complement, base i at 2(i+1)
    uint64_t y = __brevll(x) >> (64 - 2 * (K - 1));            // order reversed, bit pairs swapped
    fwd = ((y & 0x5555555555555555ull) << 1) | ((y >> 1) & 0x5555555555555555ull);
  }

  uint64_t *table = tables + gd.table_begin;
  uint32_t *count = counts + g;

  // the 32 incoming bases (window positions K-1 .. K+30), LSB first
  uint32_t up[2];
  {
    constexpr int S = 2 * (K - 1);  // bit offset of base K-1 in the 128-bit code window
    const uint32_t q0 = cw[0], q1 = cw[1], q2 = cw[2], q3 = cw[3];
    if constexpr (S < 32) { up[0] = funnel_r(q0, q1, S);      up[1] = funnel_r(q1, q2, S); (void)q3; }
    else                  { up[0] = funnel_r(q1, q2, S - 32); up[1] = funnel_r(q2, q3, S - 32); (void)q0; }
  }

#pragma unroll 1
  for (int o = 0; o < KH_PPT / 8; ++o) {
    const uint32_t in16 = ((o & 2) ? up[1] : up[0]) >> (16 * (o & 1));
    const uint32_t kv8 = kv32 >> (8 * o);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const uint32_t c = (in16 >> (2 * jj)) & 3u;
      fwd = ((fwd << 2) | c) & KMASK;
      rc = (rc >> 2) | ((uint64_t)(3u - c) << (2 * (K - 1)));
      const uint64_t canon = CANON ? (fwd < rc ? fwd : rc) : fwd;

      // ASCII expansion: reverse so base 0 sits in the low bits (bit pairs come out swapped,
      // which the lookup table absorbs: 0->'A' 1->'G' 2->'C' 3->'T'), spread each 2-bit
      // field to a nibble, and let PRMT pick the bytes.
      const uint64_t y = __brevll(canon) >> (64 - 2 * K);
      uint64_t w[NW];
#pragma unroll
      for (int m = 0; m < NW; ++m) {
        constexpr uint32_t LUT = 0x54434741u;
        const int nb = (K - 8 * m) < 8 ? (K - 8 * m) : 8;  // bases in this word
        uint32_t x = (uint32_t)(y >> (16 * m)) & 0xFFFFu;
        x = (x | (x << 8)) & 0x00FF00FFu;
        x = (x | (x << 4)) & 0x0F0F0F0Fu;
        x = (x | (x << 2)) & 0x33333333u;
        if (nb < 8) x |= 0x44444444u << (4 * nb);         // bytes past the k-mer read as zero
        const uint32_t lo = __byte_perm(LUT, 0u, x);
        const uint32_t hi = __byte_perm(LUT, 0u, x >> 16);
        w[m] = (uint64_t)lo | ((uint64_t)hi << 32);
      }
      const uint64_t h = hg::t1ha2_kmer<K>(w, seed);
      if (h < threshold && ((kv8 >> jj) & 1u)) table_insert(table, gd.table_mask, h, count, status);
    }
  }
}

template <int K>
int launch_k(hg_ctx *ctx, bool canon, uint32_t n_tiles, const uint8_t *d_seq, const hg_genome_desc *d_desc,
             uint32_t n_genomes, uint64_t threshold, uint64_t seed, uint64_t *d_tables, uint32_t *d_counts) {
  if (canon)
    kmer_hash_kernel<K, true><<<n_tiles, KH_THREADS, 0, ctx->stream>>>(d_seq, d_desc, n_genomes, threshold,
                                                                       seed, d_tables, d_counts, ctx->d_status);
  else
    kmer_hash_kernel<K, false><<<n_tiles, KH_THREADS, 0, ctx->stream>>>(d_seq, d_desc, n_genomes, threshold,
                                                                        seed, d_tables, d_counts, ctx->d_status);
  ctx->launches++;
  HG_CUDA(cudaGetLastError());
  return HG_OK;
}

}  // namespace

uint32_t hg_kmer_tile_positions() { return KH_TILE; }

int hg_launch_kmer_hash(hg_ctx *ctx, const uint8_t *d_seq, const hg_genome_desc *d_desc, uint32_t n_genomes,
                        uint32_t n_tiles, const hg_sketch_params *p, uint64_t *d_tables, uint32_t *d_counts) {
  if (n_tiles == 0) return HG_OK;
  const uint64_t threshold = UINT64_MAX / p->scaled;  // sketch.rs:73
  const bool canon = p->canonical != 0;
  switch (p->ksize) {
#define HG_K(KK) \
  case KK: return launch_k<KK>(ctx, canon, n_tiles, d_seq, d_desc, n_genomes, threshold, p->seed, d_tables, d_counts);
    HG_K(1) HG_K(2) HG_K(3) HG_K(4) HG_K(5) HG_K(6) HG_K(7) HG_K(8)
    HG_K(9) HG_K(10) HG_K(11) HG_K(12) HG_K(13) HG_K(14) HG_K(15) HG_K(16)
    HG_K(17) HG_K(18) HG_K(19) HG_K(20) HG_K(21) HG_K(22) HG_K(23) HG_K(24)
    HG_K(25) HG_K(26) HG_K(27) HG_K(28) HG_K(29) HG_K(30) HG_K(31) HG_K(32)
#undef HG_K
    default:
      hg_set_error("ksize %u unsupported (1..32)", (unsigned)p->ksize);
      return HG_E_INVALID;
  }
}
