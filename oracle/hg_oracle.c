/*
 * hg_oracle.c — CPU ORACLE for the HyperGen sketch → dist hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` legs may load it.  The product path
 * (hyper-gen_b200/, include/hypergen_b200.h) never links, imports or calls anything here.
 *
 * It is an operation-for-operation restatement, in plain C, of the reference's CPU
 * algorithm (wh-xu/Hyper-Gen, mounted at /root/reference when this was written).  Every
 * function cites the reference file:line it follows.  Third-party crates the reference
 * calls are not vendored there (no Cargo.lock, caret versions in Cargo.toml:26-57); their
 * published algorithms are restated and named where used:
 *     t1ha      0.1.x  t1ha2_atonce      (in-tree definition: src/cuda_kernel.cu:71-246)
 *     wyhash    0.5.x  WyRng             (v1 wyrng: state += P0; wymum(state ^ P1, state))
 *     bitpacking 0.9.x BitPacker8x       (8-lane vertical layout, 256-value blocks)
 *     bincode   1.3.x  default options   (LE, fixed-width ints, u64 length prefixes)
 *     needletail 0.5.x normalize / canonical_kmers (ACGT-only windows, min(fwd, rc))
 *     glibc     logf   (Rust f32::ln lowers to libm logf on x86_64-linux-gnu)
 *
 * PARITY PINNING (see tests/test_oracle_kat.py):
 *   - t1ha2: upstream t1ha self-check vectors + the reference's own device function
 *     compiled as host code from /root/reference/src/cuda_kernel.cu (oracle/_ref).
 *   - k-mer extraction: the reference's cuda_kmer_t1ha2 kernel body run as host code
 *     (oracle/_ref) on the same inputs.
 *   - wyrng: wyhash-rs README vectors.  HV layout: an AVX2-intrinsics restatement of
 *     hd.rs:15-92 executed on this host against the closed-form permutation.
 *   - BitPacker8x bytes, bincode framing: restated from the crates' published layout;
 *     the reference holds only round-trip tests for them (src/lib.rs:229-299) — the
 *     on-disk byte layout is therefore "parity unpinned" against a real hyper-gen binary.
 *
 * Build: see oracle/Makefile (gcc -O2 -mavx2 -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if defined(__AVX2__)
#include <immintrin.h>
#endif
#ifdef _OPENMP
#include <omp.h>
#endif

#define HGO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * t1ha2_atonce for length <= 32  — src/cuda_kernel.cu:71-77 (primes), :91-153 (rot64,
 * mul_64x64_128, mixup64, mux64, final64), :155-194 (tail64_le_unaligned), :196-246
 * (the length switch).  Call site on the CPU path: src/sketch.rs:90.
 * ---------------------------------------------------------------------------------- */
static const uint64_t T1_P0 = UINT64_C(0xEC99BF0D8372CAAB);
static const uint64_t T1_P1 = UINT64_C(0x82434FE90EDCEF39);
static const uint64_t T1_P2 = UINT64_C(0xD4F06DB99D67BE4B);
static const uint64_t T1_P3 = UINT64_C(0xBD9CACC22C6E9571);
static const uint64_t T1_P4 = UINT64_C(0x9C06FAF4D023E3AB);
static const uint64_t T1_P5 = UINT64_C(0xC060724A8424F345);
static const uint64_t T1_P6 = UINT64_C(0xCB5AF53AE3AAAC31);

static inline uint64_t rot64(uint64_t v, unsigned s) { return (v >> s) | (v << (64 - s)); }

static inline uint64_t mul128(uint64_t a, uint64_t b, uint64_t *hi) {
  unsigned __int128 p = (unsigned __int128)a * b;
  *hi = (uint64_t)(p >> 64);
  return (uint64_t)p;
}

/* cuda_kernel.cu:136-141 */
static inline void mixup64(uint64_t *a, uint64_t *b, uint64_t v, uint64_t prime) {
  uint64_t h;
  *a ^= mul128(*b + v, prime, &h);
  *b += h;
}

/* cuda_kernel.cu:143-153 */
static inline uint64_t final64(uint64_t a, uint64_t b) {
  uint64_t x = (a + rot64(b, 41)) * T1_P0;
  uint64_t y = (rot64(a, 23) + b) * T1_P6;
  uint64_t h, l = mul128(x ^ y, T1_P5, &h);
  return l ^ h;
}

/* cuda_kernel.cu:155-194: little-endian read of (tail & 7) bytes, 8 when that is 0 */
static inline uint64_t tail64_le(const uint8_t *p, size_t tail) {
  unsigned n = (unsigned)(tail & 7);
  if (n == 0) n = 8;
  uint64_t r = 0;
  for (unsigned i = 0; i < n; i++) r |= (uint64_t)p[i] << (8 * i);
  return r;
}

HGO_API uint64_t hgo_t1ha2_atonce(const uint8_t *data, uint64_t length, uint64_t seed) {
  uint64_t a = seed, b = length; /* cuda_kernel.cu:200-202 */
  const uint8_t *v = data;
  if (length > 32) return 0; /* out of scope: the in-tree definition has no >32 loop */
  if (length > 24) { mixup64(&a, &b, tail64_le(v, 8), T1_P4); v += 8; } /* :207-209 */
  if (length > 16) { mixup64(&b, &a, tail64_le(v, 8), T1_P3); v += 8; } /* :211-220 */
  if (length > 8)  { mixup64(&a, &b, tail64_le(v, 8), T1_P2); v += 8; } /* :222-231 */
  if (length > 0)  { mixup64(&b, &a, tail64_le(v, length), T1_P1); }     /* :233-242 */
  return final64(a, b);                                                 /* :243-244 */
}

/* ------------------------------------------------------------------------------------
 * FASTA → one merged sequence: src/fastx_reader.rs:6-29.  Header lines ('>' first)
 * contribute a single 'N'; other lines are appended minus one trailing '\n' then one
 * trailing '\r'.  Returns the number of bytes written (call with out=NULL to size).
 * ---------------------------------------------------------------------------------- */
HGO_API uint64_t hgo_read_merge_seq(const uint8_t *file, uint64_t n, uint8_t *out) {
  uint64_t w = 0, i = 0;
  while (i < n) {
    uint64_t e = i;
    while (e < n && file[e] != '\n') e++;
    uint64_t line_end = (e < n) ? e + 1 : e; /* read_line keeps the '\n' */
    if (file[i] == '>') {
      if (out) out[w] = 'N';
      w++;
    } else {
      uint64_t le = line_end;
      if (le > i && file[le - 1] == '\n') le--;
      if (le > i && file[le - 1] == '\r') le--;
      if (out) memcpy(out + w, file + i, le - i);
      w += le - i;
    }
    i = line_end;
  }
  return w;
}

/* ------------------------------------------------------------------------------------
 * k-mer hashing + FracMinHash filter: src/sketch.rs:71-98 (CPU path) — for every window
 * of k ACGT bases (upper-cased; needletail normalize + canonical_kmers skip windows with
 * any other byte), hash the lexicographically smaller of the k-mer and its reverse
 * complement (always canonical on the CPU path, sketch.rs:89; the GPU path honours the
 * flag, cuda_kernel.cu:306-314) with t1ha2_atonce(seed) and keep h < u64::MAX / scaled.
 * The result is a set (HashSet, sketch.rs:79,93): returned sorted and de-duplicated.
 * ---------------------------------------------------------------------------------- */
static int cmp_u64(const void *x, const void *y) {
  uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
  return (a > b) - (a < b);
}

static inline int base_up(uint8_t c) { /* returns upper-case ACGT or 0 */
  switch (c) {
    case 'A': case 'a': return 'A';
    case 'C': case 'c': return 'C';
    case 'G': case 'g': return 'G';
    case 'T': case 't': return 'T';
    default: return 0;
  }
}
static inline uint8_t comp_up(uint8_t c) {
  return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
}

/* returns the number of distinct sampled hashes; writes min(n, cap) of them, sorted */
HGO_API uint64_t hgo_kmer_hash_set(const uint8_t *seq, uint64_t n, uint32_t k, uint64_t scaled,
                                   uint64_t seed, int canonical, uint64_t *out, uint64_t cap) {
  if (k == 0 || k > 32 || n < k) return 0;
  const uint64_t threshold = UINT64_MAX / scaled; /* sketch.rs:73 */
  uint64_t cnt = 0, alloc = 1024;
  uint64_t *buf = (uint64_t *)malloc(alloc * sizeof(uint64_t));
  /* sketch.rs:84,87: one normalised copy and one reverse-complemented copy per sequence */
  uint8_t *norm = (uint8_t *)malloc(n), *rcs = (uint8_t *)malloc(n);
  for (uint64_t i = 0; i < n; i++) {
    int c = base_up(seq[i]);
    norm[i] = c ? (uint8_t)c : 'N';
    rcs[n - 1 - i] = c ? comp_up((uint8_t)c) : 'N';
  }
  uint64_t run = 0;
  for (uint64_t i = 0; i < n; i++) {
    run = norm[i] != 'N' ? run + 1 : 0;
    if (run < k) continue;
    const uint8_t *fwd = norm + (i + 1 - k);
    const uint8_t *use = fwd;
    if (canonical) {
      const uint8_t *rc = rcs + (n - 1 - i); /* revcomp of this window */
      if (memcmp(rc, fwd, k) < 0) use = rc;
    }
    uint64_t h = hgo_t1ha2_atonce(use, k, seed);
    if (h < threshold) {
      if (cnt == alloc) { alloc *= 2; buf = (uint64_t *)realloc(buf, alloc * sizeof(uint64_t)); }
      buf[cnt++] = h;
    }
  }
  free(norm);
  free(rcs);
  qsort(buf, cnt, sizeof(uint64_t), cmp_u64);
  uint64_t u = 0;
  for (uint64_t i = 0; i < cnt; i++)
    if (i == 0 || buf[i] != buf[i - 1]) { if (u < cap && out) out[u] = buf[i]; u++; }
  free(buf);
  return u;
}

/* ------------------------------------------------------------------------------------
 * WyRng (wyhash crate 0.5, v1 "wyrng"), call sites src/hd.rs:24,44,51,100,103.
 * seed_from_u64(x) sets the state to x (crate override of rand_core's default; pinned by
 * the crate's README vector seed 3 → 0x3e99a772750dcbe).
 * ---------------------------------------------------------------------------------- */
#define WY_P0 UINT64_C(0xa0761d6478bd642f)
#define WY_P1 UINT64_C(0xe7037ed1a0b428db)
HGO_API uint64_t hgo_wyrng_next(uint64_t *state) {
  *state += WY_P0;
  uint64_t hi, lo = mul128(*state ^ WY_P1, *state, &hi);
  return lo ^ hi;
}

/* ------------------------------------------------------------------------------------
 * HD encode, scalar form of the AVX2 routine the reference runs on every AVX2 host
 * (src/sketch.rs:40-41 → src/hd.rs:15-92).  hv[64*i + p] = -n + 2 * sum_h bit_{pi(p)} of
 * the i-th WyRng word of h, pi(p) = (p % 4) * 16 + p / 4, in wrapping i16 (hd.rs:29).
 * layout 0 = AVX2 (the reference layout), 1 = scalar fallback hd.rs:94-112 (pi = id).
 * ---------------------------------------------------------------------------------- */
HGO_API void hgo_encode_hd(const uint64_t *hashes, uint64_t n, uint32_t hv_d, int layout,
                           int16_t *hv) {
  uint32_t chunks = hv_d / 64; /* hd.rs:34 — trailing dims (hv_d % 64) stay at -n */
  int32_t *acc = (int32_t *)calloc(hv_d, sizeof(int32_t));
  for (uint64_t s = 0; s < n; s++) {
    uint64_t st = hashes[s];
    for (uint32_t i = 0; i < chunks; i++) {
      uint64_t r = hgo_wyrng_next(&st);
      for (uint32_t p = 0; p < 64; p++) {
        uint32_t bit = layout == 0 ? (p % 4) * 16 + p / 4 : p;
        acc[i * 64 + p] += (int32_t)((r >> bit) & 1);
      }
    }
  }
  for (uint32_t d = 0; d < hv_d; d++)
    hv[d] = (int16_t)(uint16_t)((uint32_t)(2 * acc[d]) - (uint32_t)n); /* wrapping i16 */
  free(acc);
}

#if defined(__AVX2__)
/* The same routine with the reference's own intrinsic sequence (hd.rs:17-22 mask, :41-58
 * batches of four seeds with zero padding, :61-69 shuffle, :71-87 per-bit hadd tree).
 * Exists only to pin the permutation pi against real AVX2 execution on this host. */
HGO_API void hgo_encode_hd_avx2(const uint64_t *hashes, uint64_t n, uint32_t hv_d, int16_t *hv) {
  const __m256i one = _mm256_set1_epi16(1), zero = _mm256_setzero_si256();
  const __m256i mask = _mm256_set_epi8(15, 14, 7, 6, 13, 12, 5, 4, 11, 10, 3, 2, 9, 8, 1, 0, 15, 14,
                                       7, 6, 13, 12, 5, 4, 11, 10, 3, 2, 9, 8, 1, 0);
  for (uint32_t d = 0; d < hv_d; d++) hv[d] = (int16_t)(-(int16_t)(uint16_t)n);
  uint64_t n4 = (n + 3) / 4 * 4, tail = n % 4;
  uint32_t chunks = hv_d / 64;
  for (uint64_t b = 0; b < n4 / 4; b++) {
    uint64_t st[4], rnd[4];
    for (int j = 0; j < 4; j++) st[j] = (b * 4 + j < n) ? hashes[b * 4 + j] : 0;
    for (uint32_t i = 0; i < chunks; i++) {
      for (int j = 0; j < 4; j++) rnd[j] = hgo_wyrng_next(&st[j]);
      if (b == n4 / 4 - 1 && tail > 0)
        for (uint64_t j = tail; j < 4; j++) rnd[j] = 0;
      __m256i v = _mm256_shuffle_epi8(
          _mm256_set_epi64x((long long)rnd[0], (long long)rnd[1], (long long)rnd[2], (long long)rnd[3]),
          mask);
      for (int kbit = 0; kbit < 16; kbit++) {
        __m256i t = _mm256_and_si256(_mm256_srl_epi16(v, _mm_set1_epi64x(kbit)), one);
        __m256i h = _mm256_hadd_epi16(t, zero);
        h = _mm256_permute4x64_epi64(h, 0xD8);
        h = _mm256_shuffle_epi8(h, mask);
        h = _mm256_hadd_epi16(h, zero);
        h = _mm256_slli_epi16(h, 1);
        int16_t lanes[16];
        _mm256_storeu_si256((__m256i *)lanes, h);
        for (int w = 0; w < 4; w++)
          hv[i * 64 + kbit * 4 + w] = (int16_t)(hv[i * 64 + kbit * 4 + w] + lanes[w]);
      }
    }
  }
}
#endif

/* src/dist.rs:132-137 — wrapping i32 sum of squares */
HGO_API int32_t hgo_hv_l2_norm_sq(const int16_t *hv, uint32_t hv_d) {
  uint32_t s = 0;
  for (uint32_t d = 0; d < hv_d; d++) s += (uint32_t)((int32_t)hv[d] * (int32_t)hv[d]);
  return (int32_t)s;
}

/* ------------------------------------------------------------------------------------
 * Quantise + bit-pack: src/hd.rs:116-157.  Returns the bit width b (6..16); writes
 * b * hv_d / 8 bytes.  BitPacker8x (bitpacking 0.9): per block of 256 values, lane
 * l = idx % 8, row r = idx / 8; the lane's rows are concatenated LSB-first into a
 * 32*b-bit stream whose j-th 32-bit word is stored little-endian at byte 32*j + 4*l.
 * ---------------------------------------------------------------------------------- */
HGO_API uint32_t hgo_quant_bits(const int16_t *hv, uint32_t hv_d) {
  int16_t mn = hv[0], mx = hv[0];
  for (uint32_t d = 1; d < hv_d; d++) { if (hv[d] < mn) mn = hv[d]; if (hv[d] > mx) mx = hv[d]; }
  int b = 6; /* hd.rs:123-136 */
  for (;;) {
    int qmin = -(1 << (b - 1)), qmax = (1 << (b - 1)) - 1;
    if (qmin <= mn && qmax >= mx) break;
    if (b == 16) break;
    b++;
  }
  return (uint32_t)b;
}

static void bitpack8x_block(const uint32_t *in256, uint8_t *out, uint32_t b) {
  memset(out, 0, 32 * b);
  for (uint32_t l = 0; l < 8; l++)
    for (uint32_t r = 0; r < 32; r++) {
      uint64_t v = in256[8 * r + l];
      uint32_t bit = b * r, j = bit / 32, sh = bit % 32;
      uint64_t spread = v << sh; /* like the crate: no masking of over-wide values */
      uint32_t lo = (uint32_t)spread, hi = (uint32_t)(spread >> 32);
      uint32_t w;
      memcpy(&w, out + 32 * j + 4 * l, 4); w |= lo; memcpy(out + 32 * j + 4 * l, &w, 4);
      if (sh + b > 32 && j + 1 < b) { /* the crate carries only a straddling value */
        memcpy(&w, out + 32 * (j + 1) + 4 * l, 4); w |= hi; memcpy(out + 32 * (j + 1) + 4 * l, &w, 4);
      }
    }
}

HGO_API uint32_t hgo_compress_hd_sketch(const int16_t *hv, uint32_t hv_d, uint8_t *packed) {
  uint32_t b = hgo_quant_bits(hv, hv_d);
  int16_t offset = (int16_t)(1 << (b - 1)); /* hd.rs:140 (wraps to i16::MIN at b = 16) */
  uint32_t u[256];
  for (uint32_t blk = 0; blk < hv_d / 256; blk++) { /* hd.rs:147 */
    for (uint32_t i = 0; i < 256; i++)
      u[i] = (uint32_t)(int32_t)(int16_t)(hv[blk * 256 + i] + offset); /* (i + offset) as u32 */
    bitpack8x_block(u, packed + (size_t)32 * b * blk, b);
  }
  return b;
}

/* src/hd.rs:184-212 */
HGO_API void hgo_decompress_hd_sketch(const uint8_t *packed, uint32_t hv_d, uint32_t b, int16_t *hv) {
  int16_t offset = (int16_t)(1 << (b - 1));
  uint64_t mask = (b >= 32) ? 0xffffffffu : ((1u << b) - 1);
  for (uint32_t blk = 0; blk < hv_d / 256; blk++) {
    const uint8_t *in = packed + (size_t)32 * b * blk;
    for (uint32_t l = 0; l < 8; l++)
      for (uint32_t r = 0; r < 32; r++) {
        uint32_t bit = b * r, j = bit / 32, sh = bit % 32;
        uint32_t w0, w1 = 0;
        memcpy(&w0, in + 32 * j + 4 * l, 4);
        if (j + 1 < b) memcpy(&w1, in + 32 * (j + 1) + 4 * l, 4);
        uint64_t both = ((uint64_t)w1 << 32) | w0;
        uint32_t u = (uint32_t)((both >> sh) & mask);
        hv[blk * 256 + 8 * r + l] = (int16_t)((int16_t)u - offset);
      }
  }
}

/* ------------------------------------------------------------------------------------
 * ANI of one pair: src/dist.rs:139-161.  i32 dot (wrapping), f32 everything else;
 * `.ln()` is libm logf on the reference's platform.
 * ---------------------------------------------------------------------------------- */
HGO_API int32_t hgo_dot_i16(const int16_t *r, const int16_t *q, uint32_t hv_d) {
  uint32_t s = 0;
  for (uint32_t d = 0; d < hv_d; d++) s += (uint32_t)((int32_t)r[d] * (int32_t)q[d]);
  return (int32_t)s;
}

HGO_API float hgo_ani_from_dot(int32_t dot, int32_t norm2_r, int32_t norm2_q, uint32_t ksize) {
  int32_t den = (int32_t)((uint32_t)norm2_r + (uint32_t)norm2_q - (uint32_t)dot);
  volatile float jaccard = (float)dot / (float)den;
  volatile float t = 1.0f / jaccard;
  t = t + 1.0f;
  t = 2.0f / t;
  volatile float l = logf(t);
  volatile float ani = l / (float)ksize;
  ani = 1.0f + ani;
  if (ani != ani) return 0.0f;
  float a = ani;
  a = a < 1.0f ? a : 1.0f; /* f32::min */
  a = a > 0.0f ? a : 0.0f; /* f32::max */
  return a * 100.0f;
}

HGO_API float hgo_pairwise_ani(const int16_t *r, int32_t norm2_r, const int16_t *q, int32_t norm2_q,
                               uint32_t hv_d, uint32_t ksize) {
  return hgo_ani_from_dot(hgo_dot_i16(r, q, hv_d), norm2_r, norm2_q, ksize);
}

/* ------------------------------------------------------------------------------------
 * A restatement of glibc 2.27+ logf (sysdeps/ieee754/flt-32/e_logf.c, the FMA ifunc
 * variant every FMA-capable x86-64 selects).  The device epilogue mirrors THIS function;
 * tests check it bit-for-bit against the host libm's logf so that GPU ANI == reference
 * ANI in every f32 bit.  Constants read out of this image's libm.so.6 (.rodata).
 * ---------------------------------------------------------------------------------- */
static const double LOGF_T[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2}, {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

HGO_API float hgo_logf_glibc(float x) {
  uint32_t ix;
  memcpy(&ix, &x, 4);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return -INFINITY;
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
    float s = x * 0x1p23f; /* subnormal: normalise */
    memcpy(&ix, &s, 4);
    ix -= 23u << 23;
  }
  uint32_t tmp = ix - 0x3f330000u;
  uint32_t i = (tmp >> 19) & 15;
  int32_t k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  float zf;
  memcpy(&zf, &iz, 4);
  double z = (double)zf;
  double r = fma(z, LOGF_T[i][0], -1.0);
  double y0 = fma((double)k, 0x1.62e42fefa39efp-1, LOGF_T[i][1]);
  double r2 = r * r;
  double y = fma(r, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2);
  y = fma(r2, -0x1.00ea348b88334p-2, y);
  y = fma(r2, y, y0 + r);
  return (float)y;
}

/* exhaustive / strided self-check against the host libm; returns the mismatch count */
HGO_API uint64_t hgo_logf_selfcheck(uint32_t first, uint32_t last, uint32_t stride) {
  uint64_t bad = 0;
  if (stride == 0) stride = 1;
  for (uint64_t u = first; u <= last; u += stride) {
    uint32_t b = (uint32_t)u;
    float x, a, c;
    memcpy(&x, &b, 4);
    a = logf(x);
    c = hgo_logf_glibc(x);
    uint32_t ba, bc;
    memcpy(&ba, &a, 4);
    memcpy(&bc, &c, 4);
    if (ba != bc && !(a != a && c != c)) bad++;
  }
  return bad;
}

/* ------------------------------------------------------------------------------------
 * Whole-stage drivers (one genome → FileSketch fields; all-pairs ANI), OpenMP-parallel
 * in the reference's own granularity: per genome (rayon par_iter_mut, sketch.rs:35) and
 * per pair (dist.rs:268-272).  These double as the timed CPU baseline.
 * ---------------------------------------------------------------------------------- */
HGO_API int hgo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* sketch.rs:35-52 for n_genomes sequences laid out back to back (seg_off[n+1]).
 * hv_out may be NULL; packed capacity is 2*hv_d bytes per genome. */
HGO_API void hgo_sketch_batch(const uint8_t *seq, const uint64_t *seg_off, uint32_t n_genomes,
                              uint32_t k, uint64_t scaled, uint64_t seed, int canonical,
                              uint32_t hv_d, int16_t *hv_out, uint8_t *packed, uint8_t *quant_bits,
                              int32_t *norm2, uint32_t *n_hashes) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t g = 0; g < (int64_t)n_genomes; g++) {
    const uint8_t *s = seq + seg_off[g];
    uint64_t len = seg_off[g + 1] - seg_off[g];
    uint64_t cap = len / scaled * 2 + 4096;
    uint64_t *set = (uint64_t *)malloc(cap * sizeof(uint64_t));
    uint64_t n = hgo_kmer_hash_set(s, len, k, scaled, seed, canonical, set, cap);
    if (n > cap) { /* cannot happen at cap = 2x expectation + 4096, but never truncate */
      set = (uint64_t *)realloc(set, n * sizeof(uint64_t));
      n = hgo_kmer_hash_set(s, len, k, scaled, seed, canonical, set, n);
    }
    int16_t *hv = (int16_t *)malloc(hv_d * sizeof(int16_t));
    hgo_encode_hd(set, n, hv_d, 0, hv);
    if (norm2) norm2[g] = hgo_hv_l2_norm_sq(hv, hv_d);
    if (packed) {
      uint32_t b = hgo_compress_hd_sketch(hv, hv_d, packed + (size_t)g * 2 * hv_d);
      if (quant_bits) quant_bits[g] = (uint8_t)b;
    } else if (quant_bits) {
      quant_bits[g] = (uint8_t)hgo_quant_bits(hv, hv_d);
    }
    if (hv_out) memcpy(hv_out + (size_t)g * hv_d, hv, hv_d * sizeof(int16_t));
    if (n_hashes) n_hashes[g] = (uint32_t)n;
    free(hv);
    free(set);
  }
}

/* k-mer stage alone, for stage-level timing: returns total sampled (non-unique) hits */
HGO_API uint64_t hgo_kmer_count_batch(const uint8_t *seq, const uint64_t *seg_off, uint32_t n_genomes,
                                      uint32_t k, uint64_t scaled, uint64_t seed, int canonical) {
  uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int64_t g = 0; g < (int64_t)n_genomes; g++)
    total += hgo_kmer_hash_set(seq + seg_off[g], seg_off[g + 1] - seg_off[g], k, scaled, seed,
                               canonical, NULL, 0);
  return total;
}

/* dist.rs:231-294: ANI of every enumerated pair.  symmetric → (i, j > i) in row-major
 * order (dist.rs:253-265), else the full R x Q grid.  ani_out has n_pairs entries
 * (R*(Q-1)/2 when symmetric, dist.rs:243-247).  dot_out may be NULL. */
HGO_API uint64_t hgo_dist_all(const int16_t *ref, const int32_t *norm_r, uint32_t n_ref,
                              const int16_t *qry, const int32_t *norm_q, uint32_t n_qry,
                              uint32_t hv_d, uint32_t ksize, int symmetric, float *ani_out,
                              int32_t *dot_out) {
  uint64_t n_pairs = symmetric ? (uint64_t)n_ref * (n_qry - 1) / 2 : (uint64_t)n_ref * n_qry;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < (int64_t)n_ref; i++) {
    uint64_t base;
    uint32_t j0;
    if (symmetric) {
      /* pairs before row i: sum_{t<i} (n_qry - 1 - t) */
      base = (uint64_t)i * (n_qry - 1) - (uint64_t)i * (i - 1) / 2;
      j0 = (uint32_t)i + 1;
    } else {
      base = (uint64_t)i * n_qry;
      j0 = 0;
    }
    for (uint32_t j = j0; j < n_qry; j++) {
      int32_t dot = hgo_dot_i16(ref + (size_t)i * hv_d, qry + (size_t)j * hv_d, hv_d);
      uint64_t p = base + (j - j0);
      if (p < n_pairs) {
        ani_out[p] = hgo_ani_from_dot(dot, norm_r[i], norm_q[j], ksize);
        if (dot_out) dot_out[p] = dot;
      }
    }
  }
  return n_pairs;
}

/* ------------------------------------------------------------------------------------
 * Output order + threshold filter: src/utils.rs:260-286.  Stable ascending sort of the
 * pair indices by ANI, reversed, emitted while ani >= ani_th.  order_out receives the
 * pair indices in emission order; returns how many are emitted.
 * ---------------------------------------------------------------------------------- */
typedef struct { float ani; uint64_t idx; } hgo_key;
static int cmp_key(const void *x, const void *y) {
  const hgo_key *a = (const hgo_key *)x, *b = (const hgo_key *)y;
  if (a->ani < b->ani) return -1;
  if (a->ani > b->ani) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx); /* stable */
}
HGO_API uint64_t hgo_ani_output_order(const float *ani, uint64_t n_pairs, float ani_th,
                                      uint64_t *order_out) {
  hgo_key *keys = (hgo_key *)malloc((n_pairs ? n_pairs : 1) * sizeof(hgo_key));
  for (uint64_t i = 0; i < n_pairs; i++) { keys[i].ani = ani[i]; keys[i].idx = i; }
  qsort(keys, n_pairs, sizeof(hgo_key), cmp_key);
  uint64_t emitted = 0;
  for (uint64_t t = 0; t < n_pairs; t++) {
    const hgo_key *kk = &keys[n_pairs - 1 - t]; /* indices.reverse() */
    if (kk->ani >= ani_th) order_out[emitted++] = kk->idx; else break;
  }
  free(keys);
  return emitted;
}
