// hg_common.cuh — shared declarations for the sm_100a HyperGen hot path.
// Arithmetic building blocks restate, for the device, the functions the reference defines in
// src/cuda_kernel.cu:71-246 (t1ha2), the wyhash crate's wyrng (call sites src/hd.rs:44,51)
// and glibc's logf (behind f32::ln in src/dist.rs:154).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hypergen_b200.h"

#define HG_EMPTY_SLOT 0xFFFFFFFFFFFFFFFFull

// ---------------------------------------------------------------------------------------
// t1ha2_atonce (length <= 32) on pre-assembled little-endian words.
//   reference: cuda_kernel.cu:196-246 (switch), :136-141 mixup64, :149-153 final64.
// ---------------------------------------------------------------------------------------
namespace hg {

__host__ __device__ constexpr uint64_t t1_p0() { return 0xEC99BF0D8372CAABull; }
__host__ __device__ constexpr uint64_t t1_p1() { return 0x82434FE90EDCEF39ull; }
__host__ __device__ constexpr uint64_t t1_p2() { return 0xD4F06DB99D67BE4Bull; }
__host__ __device__ constexpr uint64_t t1_p3() { return 0xBD9CACC22C6E9571ull; }
__host__ __device__ constexpr uint64_t t1_p4() { return 0x9C06FAF4D023E3ABull; }
__host__ __device__ constexpr uint64_t t1_p5() { return 0xC060724A8424F345ull; }
__host__ __device__ constexpr uint64_t t1_p6() { return 0xCB5AF53AE3AAAC31ull; }

// rotate right by a compile-time amount: two funnel shifts (written out, nvcc makes four instructions of the shift-or form)
template <unsigned S>
__device__ __forceinline__ uint64_t rot64(uint64_t v) {
  static_assert(S > 0 && S < 64 && S != 32, "rot64");
  const uint32_t lo = S < 32 ? (uint32_t)v : (uint32_t)(v >> 32), hi = S < 32 ? (uint32_t)(v >> 32) : (uint32_t)v;
  const uint32_t a = __funnelshift_r(lo, hi, S & 31), b = __funnelshift_r(hi, lo, S & 31);
  return (uint64_t)a | ((uint64_t)b << 32);
}

__device__ __forceinline__ void mixup64(uint64_t &a, uint64_t &b, uint64_t v, uint64_t prime) {
  const uint64_t t = b + v;
  a ^= t * prime;
  b += __umul64hi(t, prime);
}

__device__ __forceinline__ uint64_t final64(uint64_t a, uint64_t b) {
  const uint64_t x = (a + rot64<41>(b)) * t1_p0();
  const uint64_t y = (rot64<23>(a) + b) * t1_p6();
  const uint64_t v = x ^ y;
  return (v * t1_p5()) ^ __umul64hi(v, t1_p5());
}

// NW = ceil(K / 8) words; word m holds bytes 8m..8m+7 of the k-mer, little-endian, the last
// one zero-extended (tail64_le_unaligned, cuda_kernel.cu:155-194).
template <int K>
__device__ __forceinline__ uint64_t t1ha2_kmer(const uint64_t (&w)[(K + 7) / 8], uint64_t seed) {
  uint64_t a = seed, b = (uint64_t)K;
  int m = 0;
  if (K > 24) mixup64(a, b, w[m++], t1_p4());
  if (K > 16) mixup64(b, a, w[m++], t1_p3());
  if (K > 8) mixup64(a, b, w[m++], t1_p2());
  if (K > 0) mixup64(b, a, w[m++], t1_p1());
  return final64(a, b);
}

// ---------------------------------------------------------------------------------------
// WyRng word i (0-based) of the generator seeded with h: the state advances by a constant,
// so the i-th output is a closed form of (h, i) — no sequential dependency (hd.rs:44,51).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t wyrng_word(uint64_t h, uint32_t i) {
  const uint64_t s = h + (uint64_t)(i + 1) * 0xa0761d6478bd642full;
  const uint64_t t = s ^ 0xe7037ed1a0b428dbull;
  return (s * t) ^ __umul64hi(s, t);
}

// ---------------------------------------------------------------------------------------
// glibc logf (FMA ifunc variant), bit-for-bit: double-precision table + polynomial, one
// final rounding to float.  Restated in oracle/hg_oracle.c:hgo_logf_glibc and checked there
// against the host libm over every positive float.
// ---------------------------------------------------------------------------------------
static __device__ __constant__ double c_logf_tab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
    0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1.0000000000000p+0, 0x0.0p+0,
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3,
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};

__device__ __forceinline__ float logf_glibc(float x) {
  uint32_t ix = __float_as_uint(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return __uint_as_float(0xff800000u);  // -inf
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return __uint_as_float(0x7fc00000u);
    ix = __float_as_uint(__fmul_rn(x, 8388608.0f));  // subnormal: scale by 2^23
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const uint32_t i = (tmp >> 19) & 15;
  const int k = (int)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const double z = (double)__uint_as_float(iz);
  const double invc = c_logf_tab[2 * i], logc = c_logf_tab[2 * i + 1];
  const double r = __fma_rn(z, invc, -1.0);
  const double y0 = __fma_rn((double)k, 0x1.62e42fefa39efp-1, logc);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(r, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2);
  y = __fma_rn(r2, -0x1.00ea348b88334p-2, y);
  y = __fma_rn(r2, y, __dadd_rn(y0, r));
  return __double2float_rn(y);
}

// compute_pairwise_ani after the dot product: src/dist.rs:153-160, every operation a
// correctly rounded f32 op exactly as rustc emits it (no contraction, no fast-math).
__device__ __forceinline__ float ani_from_dot(int32_t dot, int32_t norm2_r, int32_t norm2_q, float ksize_f) {
  const int32_t den = (int32_t)((uint32_t)norm2_r + (uint32_t)norm2_q - (uint32_t)dot);
  const float jaccard = __fdiv_rn(__int2float_rn(dot), __int2float_rn(den));
  float t = __fdiv_rn(1.0f, jaccard);
  t = __fadd_rn(t, 1.0f);
  t = __fdiv_rn(2.0f, t);
  float ani = __fadd_rn(1.0f, __fdiv_rn(logf_glibc(t), ksize_f));
  if (ani != ani) return 0.0f;
  ani = ani < 1.0f ? ani : 1.0f;  // f32::min(1.0)
  ani = ani > 0.0f ? ani : 0.0f;  // f32::max(0.0)
  return __fmul_rn(ani, 100.0f);
}

}  // namespace hg

// ---------------------------------------------------------------------------------------
// host-side plumbing shared by the .cu files
// ---------------------------------------------------------------------------------------
struct hg_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;
  unsigned long long launches;
  // grow-only device scratch
  void *d_scratch[14];        // indexed by hg_scratch_slot
  size_t d_scratch_bytes[14];
  // pinned host scratch
  void *h_pinned[4];
  size_t h_pinned_bytes[4];
  // last dist decision
  int dist_path;
  char dist_reason[288];
  // status words of the last sketch batch (device + host copy)
  uint32_t *d_status;  // [0] table overflow, [1] quant range overflow
  uint32_t h_status[4];
  // optional stage timing (hg_set_profiling): events on ctx->stream around each stage
  int prof;
  cudaEvent_t ev[8];
  int ev_used;           // which stage boundaries were recorded in the last call
  // H2D pipeline of the host-pointer sketch entry
  int tc_attr_set;
  int tc_is_hint;   // the two-limb path was chosen by the library (packed rows of 11..13 bits), not forced by the caller: SIMT fallback allowed
  int n1_attr_set;
  const uint64_t *d_actual_len;  // optional per-genome true lengths for the next k-mer launch (raw-FASTA path)
  cudaStream_t copy_stream;
  cudaEvent_t ev_copied[2], ev_done[2];
  cudaEvent_t ev_chunk[10];  // dist streaming: fork + one per row chunk (created with the copy stream)
  // hg_dist_dev with the single-plane path forced: the pre-pass verdict is not waited for (hg_dist_status reads it)
  const uint32_t *pending_stats[2];
  uint32_t pending_nsets[2];
};

// stage boundaries: 0 start, 1 after staging/memsets, 2 after k-mer hash, 3 after encode,
// 4 dist start, 5 dist end
#define HG_PROF(ctx, idx)                                              \
  do {                                                                 \
    if ((ctx)->prof) {                                                 \
      cudaEventRecord((ctx)->ev[idx], (ctx)->stream);                  \
      (ctx)->ev_used |= 1 << (idx);                                    \
    }                                                                  \
  } while (0)

void hg_set_error(const char *fmt, ...);
int hg_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define HG_CUDA(call)                                                   \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return hg_cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

// Grow-only device scratch slots.  A slot has one tenant per entry point; entry points of one
// context never run concurrently (one stream), so slots may be shared between them.
enum hg_scratch_slot {
  HG_S_SEQ = 0,        // staged sequence bytes (two halves when the H2D pipeline is active)
  HG_S_DESC = 1,       // hg_genome_desc[] (+ sentinels)
  HG_S_TABLES = 2,     // per-genome hash tables / explicit hash sets
  HG_S_COUNTS = 3,     // distinct-hash counters; dist: the hit counter of the host entry
  HG_S_HV = 4,         // int16 HVs (sketch output staging; dist: both matrices of the host entry)
  HG_S_PACKED = 5,     // packed sketches; dist: hit records of the host entry
  HG_S_SMALL = 6,      // norms / bits / counts per sketch
  HG_S_MISC = 7,       // CTA -> genome map, |hv| max, probe sink
  HG_S_REF_LIMBS = 8,  // s8 limb planes of the ref matrix (dist_tc)
  HG_S_QRY_LIMBS = 9,  // s8 limb planes of the query matrix (dist_tc)
  HG_S_SORT_TMP = 10,  // ping-pong buffer of the hit sort
  HG_S_SORT_CNT = 11,  // digit counters of the hit sort
  HG_S_NARROW_META = 12,  // per-row constants + outlier lists of the single-plane dist path (dist_narrow)
  HG_S_TILES = 13,        // tile list of a multi-GPU dist launch (peer.cu)
  HG_S_COUNT = 14
};
// scratch slot `slot` grown to at least `bytes` (contents not preserved)
int hg_scratch(hg_ctx *ctx, int slot, size_t bytes, void **out);
int hg_pinned(hg_ctx *ctx, int slot, size_t bytes, void **out);

// ---- kernel launchers (all asynchronous on ctx->stream) ----

struct hg_genome_desc {   // per genome, device-visible
  uint64_t seq_begin;     // byte offset of the genome in the sequence buffer
  uint64_t seq_len;       // bases
  uint64_t table_begin;   // first slot of this genome's hash table
  uint32_t table_mask;    // slots - 1 (slots is a power of two) — used by the inserts
  uint32_t first_tile;    // index of this genome's first tile in the launch
  uint32_t table_slots;   // slots the encoder scans (== mask + 1, or a dense list length)
  uint32_t pad_;
};

int hg_launch_kmer_hash(hg_ctx *ctx, const uint8_t *d_seq, const hg_genome_desc *d_desc,
                        uint32_t n_genomes, uint32_t n_tiles, const hg_sketch_params *p,
                        uint64_t *d_tables, uint32_t *d_counts);
int hg_launch_sort_tables(hg_ctx *ctx, const hg_genome_desc *d_desc, uint32_t n_genomes,
                          uint64_t *d_tables, uint32_t max_slots);
int hg_launch_encode(hg_ctx *ctx, const hg_genome_desc *d_desc, uint32_t n_genomes,
                     const uint64_t *d_tables, const uint32_t *d_counts, uint32_t hv_d,
                     int16_t *d_hv, uint8_t *d_packed, uint8_t *d_quant_bits, int32_t *d_norm2,
                     uint32_t *d_n_hashes);
int hg_launch_unpack(hg_ctx *ctx, const uint8_t *d_packed, uint64_t row_stride,
                     const uint8_t *d_quant_bits, uint32_t n, uint32_t hv_d, int16_t *d_hv);
int hg_launch_dist_simt(hg_ctx *ctx, const int16_t *d_ref, const int32_t *d_ref_norm, uint32_t n_ref,
                        uint32_t i0, const int16_t *d_qry, const int32_t *d_qry_norm, uint32_t n_qry,
                        uint32_t j0, uint32_t hv_d, uint32_t ksize, float ani_th, int symmetric,
                        hg_hit *d_hits, uint64_t cap, unsigned long long *d_n_hits);
// A dist launch that is one member of several GPUs (csrc/peer.cu) walks a host-built list of ITS tiles instead of the
// arithmetic enumeration, ordered by when the rows they read arrive, and its TMA producer waits for the arrival flags
// of those rows: the operand exchange over NVLink overlaps the tile computation.
// What a multi-GPU member pushes to the other members WHILE its dist kernel runs: the kernel carries extra "pusher"
// warps that copy this member's operand rows (byte ranges of its window) to the same offsets of other windows with
// 16-byte stores over NVLink.  The work is a list of UNITS, done in order: unit u sends one SET of byte ranges (set
// 0..3 = a chunk of this member's rows, set 4 = what every member needs before it starts: pre-pass statistics, outlier
// entries) to the members in its destination mask and then raises a flag in those windows - the chunk's arrival flag,
// or this member's start flag - when all pusher warps of the grid are through with it.  All-vs-all on three or more
// GPUs sends a member's rows only to the floor(N/2) members that compute tiles with them, nearest ring neighbour first
// (csrc/peer.cu); otherwise every chunk goes to everybody.  One kernel computes tiles and moves operands.
// TPCs a multi-GPU dist launch leaves free when this member's pushes run as kernels of their own NEXT TO it (on a
// second, lower-priority stream): the dist kernel is resident and waiting for the other members' chunks while they run, so
// the pushes must never need one of its SMs.
#define HG_PEER_RESERVED_TPCS 8
#define HG_PUSH_CHUNKS 4
#define HG_PUSH_SETS (HG_PUSH_CHUNKS + 1)
#define HG_PUSH_START_SET HG_PUSH_CHUNKS
#define HG_PUSH_RANGES 8
#define HG_PUSH_UNITS (1 + HG_PUSH_CHUNKS * (HG_MAX_PEERS / 2))
struct hg_push_plan {
  uint8_t *win[HG_MAX_PEERS];  // every member's window as this member addresses it
  int rank, world;
  uint64_t off[HG_PUSH_SETS][HG_PUSH_RANGES];   // byte ranges (relative to the window base), 4-byte granular
  uint32_t bytes[HG_PUSH_SETS][HG_PUSH_RANGES];
  int n[HG_PUSH_SETS];
  int n_units;
  uint8_t unit_set[HG_PUSH_UNITS];   // which set unit u sends
  uint8_t unit_dest[HG_PUSH_UNITS];  // bit m: to member m
  int8_t unit_stamp[HG_PUSH_UNITS];  // timeline stamp this unit's completion writes (-1: none)
  uint32_t *done;       // HG_PUSH_UNITS counters in this member's window: pusher warps through with the unit
  uint64_t ready_off;   // offset of the arrival flags in every window; chunk c raises flag rank * HG_PUSH_CHUNKS + c
  uint64_t start_off;   // offset of the start flags in every window; the start set raises flag `rank`
  const uint32_t *seq;  // the call's sequence number lives in device memory (the launch sequence is replayed as a CUDA graph)
  const uint32_t *dyn_count;  // not NULL: range 0 of the start set holds (*dyn_count - dyn_base) 4-byte entries, not its full size
  uint32_t dyn_base;
  unsigned long long *dbg;  // timeline stamps (HG_PEER_TIMELINE=1, hg_peer_timeline), else NULL
};
// pusher warps per CTA of a multi-GPU dist launch: 2, 4 or 6 (HG_PEER_PUSH_WARPS overrides the kernels' default)
inline int hg_push_warps(int dflt) {
  int w = dflt;
  if (const char *e = getenv("HG_PEER_PUSH_WARPS")) w = atoi(e);
  return w <= 2 ? 2 : (w <= 4 ? 4 : 6);
}
struct hg_tile_feed {
  const uint2 *list;        // x = tile row | tile column << 16, y = mask of the arrival flags the tile needs; NULL: arithmetic walk
  uint32_t n_list;
  const uint32_t *ready;    // 32 arrival flags (in this member's window); a flag is set when (int)(flag - seq) >= 0
  const uint32_t *seq_ptr;  // where the call's sequence number lives (device memory: the launch sequence is replayed as a CUDA graph)
  uint32_t seq;             // filled in by the kernel from *seq_ptr
  const uint32_t *start;    // HG_MAX_PEERS start flags (in this member's window): member m's start set has landed
  uint32_t start_need;      // bit m: wait for member m's start flag before anything is read (its pre-pass statistics and
                            // outlier entries are here; the root has reset its gather counter)
  uint32_t *status;         // set to 1 if a wait times out (the host turns it into an error)
  unsigned long long timeout_ns;
  int reserve_tpcs;         // TPCs the launch leaves free (for the push kernel running next to it)
  unsigned long long *dbg;  // timeline stamps (HG_PEER_TIMELINE=1, hg_peer_timeline), else NULL
};

// two s8 limb planes of one matrix (dist_tc.cu), split piecewise or at once
struct hg_tc_mat {
  const int16_t *hv;
  int8_t *planes;  // [2][n_rows][hv_d]
  uint32_t n_rows, hv_d;
};
int hg_tc_shape_ok(uint32_t hv_d, const void *d_a, const void *d_b);
int hg_tc_setup(hg_ctx *ctx, const int16_t *d_hv, uint32_t n_rows, uint32_t hv_d, int plane_slot, hg_tc_mat *m);
int hg_tc_split_rows(hg_ctx *ctx, const hg_tc_mat *m, uint32_t row0, uint32_t rows);
int hg_tc_launch(hg_ctx *ctx, const hg_tc_mat *R, uint32_t r0, uint32_t n_ref, uint32_t i0, const int32_t *d_ref_norm,
                 const hg_tc_mat *Q, uint32_t q0, uint32_t n_qry, uint32_t j0, const int32_t *d_qry_norm, uint32_t ksize,
                 float ani_th, int symmetric, hg_hit *d_hits, uint64_t cap, unsigned long long *d_n_hits);
int hg_tc_attach(hg_ctx *ctx, const int16_t *d_hv, uint32_t n_rows, uint32_t hv_d, int8_t *planes, hg_tc_mat *m);
int hg_tc_launch_ex(hg_ctx *ctx, const hg_tc_mat *R, uint32_t r0, uint32_t n_ref, uint32_t i0, const int32_t *d_ref_norm,
                    const hg_tc_mat *Q, uint32_t q0, uint32_t n_qry, uint32_t j0, const int32_t *d_qry_norm, uint32_t ksize,
                    float ani_th, int symmetric, hg_hit *d_hits, uint64_t cap, unsigned long long *d_n_hits, uint32_t walk_mul,
                    uint32_t walk_add, const hg_tile_feed *feed = nullptr, const hg_push_plan *push = nullptr);
void hg_tc_tile_shape(uint32_t *rows, uint32_t *cols);
int hg_launch_dist_tc(hg_ctx *ctx, const int16_t *d_ref, const int32_t *d_ref_norm, uint32_t n_ref,
                      uint32_t i0, const int16_t *d_qry, const int32_t *d_qry_norm, uint32_t n_qry,
                      uint32_t j0, uint32_t hv_d, uint32_t ksize, float ani_th, int symmetric,
                      hg_hit *d_hits, uint64_t cap, unsigned long long *d_n_hits);
// single s8 plane (x = 2a + s) + sparse outlier corrections (dist_narrow.cu): one matrix in that form, prepared
// piecewise (rows as they arrive over PCIe) or at once
struct hg_narrow_mat {
  const int16_t *hv;  // the i16 matrix in HBM (n_rows x hv_d); rows need to be there only when they are prepared
  int8_t *plane;
  int32_t *s, *a2;
  uint32_t *e, *out_off, *out_cnt, *entries;
  // pre-pass statistics, 4 words per set: [0] max |x|, [1] max (|s| + 256), [2] entry cursor, [3] declined.
  // `stats` is the set THIS context's pre-pass writes; the dist kernel and the verdict reduce over the
  // n_sets sets at stats_all (one per member when the rows of the matrix were prepared on several GPUs).
  uint32_t *stats, *stats_all;
  uint32_t n_sets;
  uint32_t entry_base;  // first entry index this context's pre-pass may use ...
  uint32_t cap;         // ... and the limit (exclusive)
  uint32_t n_rows, hv_d;
};
int hg_narrow_shape_ok(uint32_t hv_d, const void *d_a, const void *d_b);
size_t hg_narrow_meta_bytes(uint32_t n_rows);
int hg_narrow_setup(hg_ctx *ctx, const int16_t *d_hv, uint32_t n_rows, uint32_t hv_d, int plane_slot, void *meta,
                    hg_narrow_mat *m);
// the same over caller-provided memory (a peer window): nothing is allocated; entries_per_set as hg_narrow_set_cap()
size_t hg_narrow_arrays_bytes(uint32_t n_rows);  // s | a2 | e | out_off | out_cnt, n_rows words each
uint32_t hg_narrow_set_cap(uint32_t n_rows);
int hg_narrow_attach(hg_ctx *ctx, const int16_t *d_hv, uint32_t n_rows, uint32_t hv_d, int8_t *plane, void *arrays,
                     uint32_t *entries, uint32_t *stats_all, uint32_t n_sets, uint32_t my_set, uint32_t entry_base,
                     uint32_t set_cap, hg_narrow_mat *m, bool init_stats = true);
int hg_narrow_prep_rows(hg_ctx *ctx, const hg_narrow_mat *m, uint32_t row0, uint32_t rows);
int hg_narrow_launch(hg_ctx *ctx, const hg_narrow_mat *R, uint32_t r0, uint32_t n_ref, uint32_t i0, const int32_t *d_ref_norm,
                     const hg_narrow_mat *Q, uint32_t q0, uint32_t n_qry, uint32_t j0, const int32_t *d_qry_norm,
                     uint32_t ksize, float ani_th, int symmetric, hg_hit *d_hits, uint64_t cap,
                     unsigned long long *d_n_hits);
// as hg_narrow_launch for one member of `walk_mul` GPUs sharing the tile enumeration (takes tiles walk_add, walk_add + walk_mul, ...)
int hg_narrow_launch_ex(hg_ctx *ctx, const hg_narrow_mat *R, uint32_t r0, uint32_t n_ref, uint32_t i0, const int32_t *d_ref_norm,
                        const hg_narrow_mat *Q, uint32_t q0, uint32_t n_qry, uint32_t j0, const int32_t *d_qry_norm,
                        uint32_t ksize, float ani_th, int symmetric, hg_hit *d_hits, uint64_t cap,
                        unsigned long long *d_n_hits, uint32_t walk_mul, uint32_t walk_add, const hg_tile_feed *feed = nullptr,
                        const hg_push_plan *push = nullptr);
void hg_narrow_tile_shape(uint32_t *rows, uint32_t *cols);
int hg_narrow_verdict(hg_ctx *ctx, const hg_narrow_mat *A, const hg_narrow_mat *B, int32_t *absmax_out, uint64_t *outliers_out);
// one-shot form of the above; HG_E_UNSUPPORTED when the rows are not narrow
int hg_launch_dist_narrow(hg_ctx *ctx, const int16_t *d_ref, const int32_t *d_ref_norm, uint32_t n_ref,
                          uint32_t i0, const int16_t *d_qry, const int32_t *d_qry_norm, uint32_t n_qry,
                          uint32_t j0, uint32_t hv_d, uint32_t ksize, float ani_th, int symmetric,
                          hg_hit *d_hits, uint64_t cap, unsigned long long *d_n_hits, int32_t *absmax_out,
                          uint64_t *outliers_out, int defer_verdict);
int hg_launch_int_peak(hg_ctx *ctx, int which, uint32_t iters, uint32_t *d_sink, uint32_t blocks);
// max |hv| over a device matrix (decides the dist path); result in *d_out (int32)
int hg_launch_sort_hits(hg_ctx *ctx, hg_hit *d_hits, uint64_t n, uint32_t *d_milli);
int hg_launch_absmax(hg_ctx *ctx, const int16_t *d_hv, uint64_t n_elems, int32_t *d_out);
