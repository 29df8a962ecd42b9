// api.cu — the extern "C" boundary declared in include/hypergen_b200.h: context, host-buffer
// entry points (H2D / D2H inside), device-pointer entry points (inputs resident in HBM).
// Replaces the cudarc plumbing of reference src/sketch_cuda.rs:52-60,119-166.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

#include "hg_common.cuh"

uint32_t hg_kmer_tile_positions();
uint32_t hg_kmer_tiles_per_cta();

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void hg_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int hg_cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  hg_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return HG_E_CUDA;
}

extern "C" const char *hg_last_error(void) { return g_err; }
extern "C" const char *hg_version(void) { return "hypergen_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
extern "C" int hg_init(int device, hg_ctx **out) {
  if (!out) { hg_set_error("hg_init: out is NULL"); return HG_E_INVALID; }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    hg_set_error("hg_init: no CUDA device (%s); this library has no CPU fallback",
                 e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return HG_E_CUDA;
  }
  if (device < 0 || device >= n) { hg_set_error("hg_init: device %d out of range (%d)", device, n); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(device));
  hg_ctx *c = new hg_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  struct Guard { hg_ctx *c; ~Guard() { if (c) hg_destroy(c); } } guard{c};  // nothing leaks when a step below fails
  cudaDeviceProp prop;
  HG_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  {  // highest priority: work other streams of the library run NEXT TO this one (a multi-GPU member's operand pushes,
     // peer.cu) must not get in front of the kernels launched here
    int lo = 0, hi = 0;
    HG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    HG_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
  }
  HG_CUDA(cudaMalloc(&c->d_status, 4 * sizeof(uint32_t)));
  HG_CUDA(cudaMemset(c->d_status, 0, 4 * sizeof(uint32_t)));
  for (int i = 0; i < 8; i++) HG_CUDA(cudaEventCreate(&c->ev[i]));
  guard.c = nullptr;
  *out = c;
  return HG_OK;
}

extern "C" void hg_destroy(hg_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 10; i++) if (c->ev_chunk[i]) cudaEventDestroy(c->ev_chunk[i]);
  for (int i = 0; i < HG_S_COUNT; i++) if (c->d_scratch[i]) cudaFree(c->d_scratch[i]);
  for (int i = 0; i < 4; i++) if (c->h_pinned[i]) cudaFreeHost(c->h_pinned[i]);
  if (c->d_status) cudaFree(c->d_status);
  for (int i = 0; i < 8; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->copy_stream) {
    cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(c->ev_copied[i]); cudaEventDestroy(c->ev_done[i]); }
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int hg_sync(hg_ctx *c) {
  if (!c) { hg_set_error("hg_sync: ctx is NULL"); return HG_E_INVALID; }
  HG_CUDA(cudaStreamSynchronize(c->stream));
  return HG_OK;
}

extern "C" uint64_t hg_stream_handle(hg_ctx *c) { return c ? (uint64_t)(uintptr_t)c->stream : 0; }
extern "C" uint64_t hg_launch_count(hg_ctx *c) { return c ? c->launches : 0; }

extern "C" int hg_set_profiling(hg_ctx *c, int enabled) {
  if (!c) { hg_set_error("ctx is NULL"); return HG_E_INVALID; }
  c->prof = enabled != 0;
  c->ev_used = 0;
  return HG_OK;
}

extern "C" int hg_stage_ms(hg_ctx *c, float out_ms[4]) {
  if (!c || !out_ms) { hg_set_error("hg_stage_ms: NULL argument"); return HG_E_INVALID; }
  HG_CUDA(cudaStreamSynchronize(c->stream));
  const int pairs[4][2] = {{0, 1}, {1, 2}, {2, 3}, {4, 5}};
  for (int s = 0; s < 4; s++) {
    out_ms[s] = -1.0f;
    const int a = pairs[s][0], b = pairs[s][1];
    if ((c->ev_used >> a & 1) && (c->ev_used >> b & 1)) HG_CUDA(cudaEventElapsedTime(&out_ms[s], c->ev[a], c->ev[b]));
  }
  return HG_OK;
}

extern "C" int hg_int_peak(hg_ctx *c, int which, double *lane_ops_per_s) {
  if (!c || !lane_ops_per_s) { hg_set_error("hg_int_peak: NULL argument"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  void *d_sink;
  int rc;
  if ((rc = hg_scratch(c, HG_S_MISC, 256, &d_sink))) return rc;
  const uint32_t blocks = (uint32_t)c->sm_count * 8, iters = 4096;
  if ((rc = hg_launch_int_peak(c, which, 64, (uint32_t *)d_sink, blocks))) return rc;  // warm-up
  cudaEvent_t e0, e1;
  HG_CUDA(cudaEventCreate(&e0));
  HG_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    HG_CUDA(cudaEventRecord(e0, c->stream));
    if ((rc = hg_launch_int_peak(c, which, iters, (uint32_t *)d_sink, blocks))) return rc;
    HG_CUDA(cudaEventRecord(e1, c->stream));
    HG_CUDA(cudaEventSynchronize(e1));
    float ms;
    HG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double per_thread = (double)iters * 64.0 * (which == 2 ? 2.0 : 1.0);
  *lane_ops_per_s = per_thread * 256.0 * blocks / (best * 1e-3);
  return HG_OK;
}

extern "C" int hg_host_alloc(uint64_t bytes, void **out) {
  if (!out) { hg_set_error("hg_host_alloc: out is NULL"); return HG_E_INVALID; }
  *out = nullptr;
  if (bytes == 0) return HG_OK;
  HG_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return HG_OK;
}

extern "C" int hg_host_free(void *p) {
  if (p) HG_CUDA(cudaFreeHost(p));
  return HG_OK;
}

extern "C" int hg_host_register(void *p, uint64_t bytes, void **dev_ptr) {
  if (!p || !dev_ptr) { hg_set_error("hg_host_register: NULL argument"); return HG_E_INVALID; }
  *dev_ptr = nullptr;
  HG_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  HG_CUDA(cudaHostGetDevicePointer(dev_ptr, p, 0));
  return HG_OK;
}

extern "C" int hg_host_unregister(void *p) {
  if (p) HG_CUDA(cudaHostUnregister(p));
  return HG_OK;
}

int hg_scratch(hg_ctx *c, int slot, size_t bytes, void **out) {
  if (bytes == 0) bytes = 256;
  if (c->d_scratch_bytes[slot] < bytes) {
    HG_CUDA(cudaStreamSynchronize(c->stream));
    if (c->d_scratch[slot]) HG_CUDA(cudaFree(c->d_scratch[slot]));
    c->d_scratch[slot] = nullptr;
    c->d_scratch_bytes[slot] = 0;
    size_t want = bytes + bytes / 8 + 4096;
    HG_CUDA(cudaMalloc(&c->d_scratch[slot], want));
    c->d_scratch_bytes[slot] = want;
  }
  *out = c->d_scratch[slot];
  return HG_OK;
}

int hg_pinned(hg_ctx *c, int slot, size_t bytes, void **out) {
  if (bytes == 0) bytes = 256;
  if (c->h_pinned_bytes[slot] < bytes) {
    HG_CUDA(cudaStreamSynchronize(c->stream));
    if (c->h_pinned[slot]) HG_CUDA(cudaFreeHost(c->h_pinned[slot]));
    c->h_pinned[slot] = nullptr;
    c->h_pinned_bytes[slot] = 0;
    size_t want = bytes + bytes / 8 + 4096;
    HG_CUDA(cudaMallocHost(&c->h_pinned[slot], want));
    c->h_pinned_bytes[slot] = want;
  }
  *out = c->h_pinned[slot];
  return HG_OK;
}

// ---------------------------------------------------------------------------------------
// sketch
// ---------------------------------------------------------------------------------------
static int check_params(const hg_sketch_params *p) {
  if (!p) { hg_set_error("params is NULL"); return HG_E_INVALID; }
  if (p->ksize == 0 || p->ksize > 32) { hg_set_error("ksize %u out of range 1..32", (unsigned)p->ksize); return HG_E_INVALID; }
  if (p->scaled == 0) { hg_set_error("scaled must be >= 1"); return HG_E_INVALID; }
  if (p->hv_d == 0 || p->hv_d % 256 != 0) {
    hg_set_error("hv_d %u must be a positive multiple of 256 (BitPacker8x block, src/hd.rs:147)", p->hv_d);
    return HG_E_INVALID;
  }
  return HG_OK;
}

struct SketchPlan {
  std::vector<hg_genome_desc> desc;
  uint64_t total_slots = 0;
  uint32_t n_tiles = 0;
  uint32_t max_slots = 0;
};

static int make_plan_bl(const uint64_t *begin, const uint64_t *lens, uint32_t n, const hg_sketch_params *p, SketchPlan &pl);

static int make_plan(const uint64_t *seg_off, uint32_t n, const hg_sketch_params *p, SketchPlan &pl) {
  std::vector<uint64_t> lens(n);
  for (uint32_t g = 0; g < n; g++) {
    if (seg_off[g + 1] < seg_off[g]) { hg_set_error("seg_off not monotone at %u", g); return HG_E_INVALID; }
    lens[g] = seg_off[g + 1] - seg_off[g];
  }
  return make_plan_bl(seg_off, lens.data(), n, p, pl);
}

// genome g = bytes [begin[g], begin[g] + lens[g]) of the sequence buffer
static int make_plan_bl(const uint64_t *begin, const uint64_t *lens, uint32_t n, const hg_sketch_params *p, SketchPlan &pl) {
  const uint32_t tile = hg_kmer_tile_positions();
  pl.desc.resize(n + 1);  // + sentinel
  pl.max_slots = 0;
  uint64_t slots_total = 0, tiles_total = 0;
  for (uint32_t g = 0; g < n; g++) {
    const uint64_t len = lens[g];
    const uint64_t n_kmers = len >= p->ksize ? len - p->ksize + 1 : 0;
    // open-addressing table at <= 50 % load for the expected FracMinHash sample size
    uint64_t want = 2 * (len / p->scaled + 1) + 32, slots = 64;
    while (slots < want) slots <<= 1;
    if (slots > (1ull << 31)) { hg_set_error("genome %u too large for one table", g); return HG_E_UNSUPPORTED; }
    hg_genome_desc &d = pl.desc[g];
    d.seq_begin = begin[g];
    d.seq_len = len;
    d.table_begin = slots_total;
    d.table_mask = (uint32_t)(slots - 1);
    d.first_tile = (uint32_t)tiles_total;
    d.table_slots = (uint32_t)slots;
    d.pad_ = 0;
    slots_total += slots;
    tiles_total += (n_kmers + tile - 1) / tile;
    if (slots > pl.max_slots) pl.max_slots = (uint32_t)slots;
    if (tiles_total > 0x7fffffffull) { hg_set_error("batch too large: split it"); return HG_E_UNSUPPORTED; }
  }
  memset(&pl.desc[n], 0, sizeof(hg_genome_desc));
  pl.desc[n].first_tile = (uint32_t)tiles_total;  // sentinel: ends the kernel's forward walk
  pl.total_slots = slots_total;
  pl.n_tiles = (uint32_t)tiles_total;
  return HG_OK;
}

// stage the plan, clear the tables and hash every genome; tables/counts stay in scratch 2/3
static int run_hash_stage(hg_ctx *c, const uint8_t *d_seq, const SketchPlan &pl, uint32_t n,
                          const hg_sketch_params *p, hg_genome_desc **d_desc_out, uint64_t **d_tables_out,
                          uint32_t **d_counts_out) {
  void *d_desc, *d_tables, *d_counts, *h_desc;
  int rc;
  if ((rc = hg_scratch(c, HG_S_DESC, sizeof(hg_genome_desc) * (n + 1), &d_desc))) return rc;
  if ((rc = hg_scratch(c, HG_S_TABLES, pl.total_slots * 8, &d_tables))) return rc;
  if ((rc = hg_scratch(c, HG_S_COUNTS, sizeof(uint32_t) * n, &d_counts))) return rc;
  if ((rc = hg_pinned(c, 0, sizeof(hg_genome_desc) * (n + 1), &h_desc))) return rc;
  memcpy(h_desc, pl.desc.data(), sizeof(hg_genome_desc) * (n + 1));
  c->ev_used = 0;
  HG_PROF(c, 0);
  HG_CUDA(cudaMemcpyAsync(d_desc, h_desc, sizeof(hg_genome_desc) * (n + 1), cudaMemcpyHostToDevice, c->stream));
  HG_CUDA(cudaMemsetAsync(d_tables, 0xFF, pl.total_slots * 8, c->stream));
  HG_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * n, c->stream));
  HG_CUDA(cudaMemsetAsync(c->d_status, 0, 4 * sizeof(uint32_t), c->stream));
  HG_PROF(c, 1);
  if ((rc = hg_launch_kmer_hash(c, d_seq, (const hg_genome_desc *)d_desc, n, pl.n_tiles, p, (uint64_t *)d_tables,
                                (uint32_t *)d_counts)))
    return rc;
  HG_PROF(c, 2);
  *d_desc_out = (hg_genome_desc *)d_desc;
  *d_tables_out = (uint64_t *)d_tables;
  *d_counts_out = (uint32_t *)d_counts;
  return HG_OK;
}

static int status_to_rc(const uint32_t *st) {
  if (st[0]) { hg_set_error("a genome's hash table overflowed (more distinct sampled hashes than 2x the FracMinHash expectation)"); return HG_E_CAPACITY; }
  if (st[1]) { hg_set_error("a sketch needs hv_quant_bits = 16: outside the reference's representable range (src/hd.rs:140)"); return HG_E_RANGE; }
  return HG_OK;
}

extern "C" int hg_sketch_batch_dev(hg_ctx *c, const uint8_t *d_seq, const uint64_t *seg_off, uint32_t n,
                                   const hg_sketch_params *p, int16_t *d_hv, uint8_t *d_packed,
                                   uint8_t *d_quant_bits, int32_t *d_norm2, uint32_t *d_n_hashes) {
  if (!c || !seg_off || (!d_seq && n && seg_off[n] > 0)) { hg_set_error("hg_sketch_batch_dev: NULL argument"); return HG_E_INVALID; }
  if (!d_quant_bits || !d_norm2 || !d_n_hashes) { hg_set_error("hg_sketch_batch_dev: quant_bits/norm2/n_hashes are required"); return HG_E_INVALID; }
  int rc = check_params(p);
  if (rc) return rc;
  if (n == 0) return HG_OK;
  HG_CUDA(cudaSetDevice(c->device));
  SketchPlan pl;
  if ((rc = make_plan(seg_off, n, p, pl))) return rc;
  hg_genome_desc *d_desc; uint64_t *d_tables; uint32_t *d_counts;
  if ((rc = run_hash_stage(c, d_seq, pl, n, p, &d_desc, &d_tables, &d_counts))) return rc;
  rc = hg_launch_encode(c, d_desc, n, d_tables, d_counts, p->hv_d, d_hv, d_packed, d_quant_bits, d_norm2, d_n_hashes);
  HG_PROF(c, 3);
  return rc;
}

// ---- stage hook: encode from explicit hash sets (hd.rs:15 takes a HashSet) ----
extern "C" int hg_encode_sets_dev(hg_ctx *c, const uint64_t *d_hashes, const uint64_t *hash_off, uint32_t n,
                                  uint32_t hv_d, int16_t *d_hv, uint8_t *d_packed, uint8_t *d_quant_bits,
                                  int32_t *d_norm2) {
  if (!c || !hash_off || !d_quant_bits || !d_norm2) { hg_set_error("hg_encode_sets_dev: NULL argument"); return HG_E_INVALID; }
  if (hv_d == 0 || hv_d % 256 != 0) { hg_set_error("hv_d %u must be a positive multiple of 256", hv_d); return HG_E_INVALID; }
  if (n == 0) return HG_OK;
  HG_CUDA(cudaSetDevice(c->device));
  std::vector<hg_genome_desc> desc(n);
  for (uint32_t g = 0; g < n; g++) {
    if (hash_off[g + 1] < hash_off[g] || hash_off[g + 1] - hash_off[g] > 0x7fffffffull) { hg_set_error("hash_off invalid at %u", g); return HG_E_INVALID; }
    memset(&desc[g], 0, sizeof(hg_genome_desc));
    desc[g].table_begin = hash_off[g];
    desc[g].table_slots = (uint32_t)(hash_off[g + 1] - hash_off[g]);
  }
  void *d_desc, *h_desc;
  int rc;
  if ((rc = hg_scratch(c, HG_S_DESC, sizeof(hg_genome_desc) * n, &d_desc))) return rc;
  if ((rc = hg_pinned(c, 0, sizeof(hg_genome_desc) * n, &h_desc))) return rc;
  memcpy(h_desc, desc.data(), sizeof(hg_genome_desc) * n);
  c->ev_used = 0;
  HG_CUDA(cudaMemcpyAsync(d_desc, h_desc, sizeof(hg_genome_desc) * n, cudaMemcpyHostToDevice, c->stream));
  HG_CUDA(cudaMemsetAsync(c->d_status, 0, 4 * sizeof(uint32_t), c->stream));
  HG_PROF(c, 2);
  rc = hg_launch_encode(c, (const hg_genome_desc *)d_desc, n, d_hashes, nullptr, hv_d, d_hv, d_packed, d_quant_bits,
                        d_norm2, nullptr);
  HG_PROF(c, 3);
  return rc;
}

extern "C" int hg_encode_sets(hg_ctx *c, const uint64_t *hashes, const uint64_t *hash_off, uint32_t n, uint32_t hv_d,
                              int16_t *hv, uint8_t *packed, uint8_t *quant_bits, int32_t *norm2) {
  if (!c || !hash_off) { hg_set_error("hg_encode_sets: NULL argument"); return HG_E_INVALID; }
  if (n == 0) return HG_OK;
  if (hv_d == 0 || hv_d % 256 != 0) { hg_set_error("hv_d %u must be a positive multiple of 256", hv_d); return HG_E_INVALID; }
  const uint64_t total = hash_off[n] - hash_off[0];
  if (total && !hashes) { hg_set_error("hg_encode_sets: hashes is NULL"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  int rc;
  void *d_hashes, *d_hv = nullptr, *d_packed = nullptr, *d_small;
  if ((rc = hg_scratch(c, HG_S_TABLES, total * 8 + 64, &d_hashes))) return rc;
  if (hv && (rc = hg_scratch(c, HG_S_HV, (size_t)n * hv_d * 2, &d_hv))) return rc;
  if (packed && (rc = hg_scratch(c, HG_S_PACKED, (size_t)n * hv_d * 2, &d_packed))) return rc;
  if ((rc = hg_scratch(c, HG_S_SMALL, (size_t)n * 12, &d_small))) return rc;
  uint8_t *d_bits = (uint8_t *)d_small + (size_t)n * 8;
  int32_t *d_norm = (int32_t *)d_small;
  HG_CUDA(cudaMemcpyAsync(d_hashes, hashes + hash_off[0], total * 8, cudaMemcpyHostToDevice, c->stream));
  std::vector<uint64_t> rel(n + 1);
  for (uint32_t g = 0; g <= n; g++) rel[g] = hash_off[g] - hash_off[0];
  if ((rc = hg_encode_sets_dev(c, (const uint64_t *)d_hashes, rel.data(), n, hv_d, (int16_t *)d_hv, (uint8_t *)d_packed,
                               d_bits, d_norm)))
    return rc;
  if (hv) HG_CUDA(cudaMemcpyAsync(hv, d_hv, (size_t)n * hv_d * 2, cudaMemcpyDeviceToHost, c->stream));
  if (packed) HG_CUDA(cudaMemcpyAsync(packed, d_packed, (size_t)n * hv_d * 2, cudaMemcpyDeviceToHost, c->stream));
  if (quant_bits) HG_CUDA(cudaMemcpyAsync(quant_bits, d_bits, n, cudaMemcpyDeviceToHost, c->stream));
  if (norm2) HG_CUDA(cudaMemcpyAsync(norm2, d_norm, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  return hg_sketch_status(c);
}

extern "C" int hg_sketch_status(hg_ctx *c) {
  if (!c) { hg_set_error("ctx is NULL"); return HG_E_INVALID; }
  HG_CUDA(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(c->h_status), cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  return status_to_rc(c->h_status);
}

// Host-pointer entry: the batch is cut into ~64 MB chunks on genome boundaries and pipelined —
// chunk c+1 crosses PCIe on a copy stream (double-buffered device staging) while chunk c is hashed
// and encoded on the compute stream; the sketches come back in one D2H at the end.  (Measured on
// B200: 32-64 MB chunks sustain the full 55 GB/s PCIe rate; 256 MB chunks drop it to 43 GB/s.)
static uint64_t hg_chunk_bytes() {
  const char *e = getenv("HG_CHUNK_MB");
  const uint64_t mb = e ? strtoull(e, nullptr, 10) : 64;
  return (mb ? mb : 64) << 20;
}

extern "C" int hg_sketch_batch(hg_ctx *c, const uint8_t *seq, const uint64_t *seg_off, uint32_t n,
                               const hg_sketch_params *p, int16_t *hv, uint8_t *packed, uint8_t *quant_bits,
                               int32_t *norm2, uint32_t *n_hashes) {
  if (!c || !seg_off) { hg_set_error("hg_sketch_batch: NULL argument"); return HG_E_INVALID; }
  int rc = check_params(p);
  if (rc) return rc;
  if (n == 0) return HG_OK;
  if (!seq && seg_off[n] > seg_off[0]) { hg_set_error("hg_sketch_batch: seq is NULL"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  const uint32_t D = p->hv_d;

  const uint64_t chunk_bytes = hg_chunk_bytes();
  struct Chunk { uint32_t g0, g1; uint64_t lo, hi, shift; SketchPlan pl; };
  std::vector<Chunk> chunks;
  for (uint32_t g = 0; g < n;) {
    Chunk ch;
    ch.g0 = g;
    ch.lo = seg_off[g];
    do { ++g; } while (g < n && seg_off[g + 1] - ch.lo <= chunk_bytes);
    ch.g1 = g;
    ch.hi = seg_off[g];
    ch.shift = (uint64_t)((uintptr_t)(seq + ch.lo) & 15);  // keep the caller's alignment modulo 16
    std::vector<uint64_t> rel(ch.g1 - ch.g0 + 1);
    for (uint32_t t = 0; t <= ch.g1 - ch.g0; t++) {
      if (seg_off[ch.g0 + t] < ch.lo) { hg_set_error("seg_off not monotone"); return HG_E_INVALID; }
      rel[t] = seg_off[ch.g0 + t] - ch.lo + ch.shift;
    }
    if ((rc = make_plan(rel.data(), ch.g1 - ch.g0, p, ch.pl))) return rc;
    chunks.push_back(std::move(ch));
  }
  uint64_t max_bytes = 0, max_slots = 0, max_tiles = 0;
  for (const Chunk &ch : chunks) {
    max_bytes = std::max(max_bytes, ch.hi - ch.lo);
    max_slots = std::max<uint64_t>(max_slots, ch.pl.total_slots);
    max_tiles = std::max<uint64_t>(max_tiles, ch.pl.n_tiles);
  }
  const uint64_t slot_bytes = (max_bytes + 64 + 255) & ~255ull;
  const size_t n_desc = (size_t)n + chunks.size();  // one sentinel per chunk

  // size every scratch buffer before anything is enqueued (growing one synchronises)
  void *d_seq, *d_desc, *d_tables, *d_counts, *d_hv = nullptr, *d_packed, *d_small, *d_map, *h_desc;
  if ((rc = hg_scratch(c, HG_S_SEQ, 2 * slot_bytes, &d_seq))) return rc;
  if ((rc = hg_scratch(c, HG_S_DESC, sizeof(hg_genome_desc) * n_desc, &d_desc))) return rc;
  if ((rc = hg_scratch(c, HG_S_TABLES, max_slots * 8, &d_tables))) return rc;
  if ((rc = hg_scratch(c, HG_S_COUNTS, sizeof(uint32_t) * n, &d_counts))) return rc;
  if (hv && (rc = hg_scratch(c, HG_S_HV, (size_t)n * D * 2, &d_hv))) return rc;
  if ((rc = hg_scratch(c, HG_S_PACKED, (size_t)n * D * 2, &d_packed))) return rc;
  if ((rc = hg_scratch(c, HG_S_SMALL, (size_t)n * 12, &d_small))) return rc;
  if ((rc = hg_scratch(c, HG_S_MISC, (max_tiles / hg_kmer_tiles_per_cta() + 2) * 4 + 256, &d_map))) return rc;
  if ((rc = hg_pinned(c, 0, sizeof(hg_genome_desc) * n_desc, &h_desc))) return rc;
  uint8_t *d_bits = (uint8_t *)d_small + (size_t)n * 8;
  int32_t *d_norm = (int32_t *)d_small;
  uint32_t *d_nh = (uint32_t *)d_small + n;
  if (!c->copy_stream) {
    HG_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      HG_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
      HG_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
  }
  c->ev_used = 0;
  HG_CUDA(cudaMemsetAsync(c->d_status, 0, 4 * sizeof(uint32_t), c->stream));
  HG_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * n, c->stream));
  // the copy stream may not overwrite a staging slot before everything already queued is done with it
  HG_CUDA(cudaEventRecord(c->ev_done[0], c->stream));
  HG_CUDA(cudaEventRecord(c->ev_done[1], c->stream));

  const bool dbg = getenv("HG_DEBUG_PIPE") != nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr, t2 = nullptr;
  if (dbg) {
    cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventCreate(&t2);
    cudaEventRecord(t0, c->stream);
  }
  size_t desc_pos = 0;
  for (size_t ci = 0; ci < chunks.size(); ci++) {
    const Chunk &ch = chunks[ci];
    const int slot = (int)(ci & 1);
    const uint32_t m = ch.g1 - ch.g0;
    uint8_t *d_slot = (uint8_t *)d_seq + slot * slot_bytes;
    // ---- copy stream: H2D of this chunk once the slot's previous tenant has been consumed ----
    HG_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_done[slot], 0));
    if (ch.hi > ch.lo)
      HG_CUDA(cudaMemcpyAsync(d_slot + ch.shift, seq + ch.lo, ch.hi - ch.lo, cudaMemcpyHostToDevice, c->copy_stream));
    HG_CUDA(cudaEventRecord(c->ev_copied[slot], c->copy_stream));
    // ---- compute stream: descriptors, table clear, hash, encode ----
    hg_genome_desc *hd = (hg_genome_desc *)h_desc + desc_pos, *dd = (hg_genome_desc *)d_desc + desc_pos;
    memcpy(hd, ch.pl.desc.data(), sizeof(hg_genome_desc) * (m + 1));
    desc_pos += m + 1;
    HG_CUDA(cudaMemcpyAsync(dd, hd, sizeof(hg_genome_desc) * (m + 1), cudaMemcpyHostToDevice, c->stream));
    HG_CUDA(cudaMemsetAsync(d_tables, 0xFF, ch.pl.total_slots * 8, c->stream));
    HG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[slot], 0));
    if ((rc = hg_launch_kmer_hash(c, d_slot, dd, m, ch.pl.n_tiles, p, (uint64_t *)d_tables, (uint32_t *)d_counts + ch.g0)))
      return rc;
    HG_CUDA(cudaEventRecord(c->ev_done[slot], c->stream));
    if ((rc = hg_launch_encode(c, dd, m, (const uint64_t *)d_tables, (const uint32_t *)d_counts + ch.g0, D,
                               d_hv ? (int16_t *)d_hv + (size_t)ch.g0 * D : nullptr,
                               (uint8_t *)d_packed + (size_t)ch.g0 * 2 * D, d_bits + ch.g0, d_norm + ch.g0, d_nh + ch.g0)))
      return rc;
  }
  if (dbg) { cudaEventRecord(t1, c->copy_stream); cudaEventRecord(t2, c->stream); }
  if (hv) HG_CUDA(cudaMemcpyAsync(hv, d_hv, (size_t)n * D * 2, cudaMemcpyDeviceToHost, c->stream));
  if (packed) HG_CUDA(cudaMemcpyAsync(packed, d_packed, (size_t)n * D * 2, cudaMemcpyDeviceToHost, c->stream));
  if (quant_bits) HG_CUDA(cudaMemcpyAsync(quant_bits, d_bits, n, cudaMemcpyDeviceToHost, c->stream));
  if (norm2) HG_CUDA(cudaMemcpyAsync(norm2, d_norm, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  if (n_hashes) HG_CUDA(cudaMemcpyAsync(n_hashes, d_nh, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  rc = hg_sketch_status(c);
  if (dbg) {
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, t0, t1); cudaEventElapsedTime(&b, t0, t2);
    fprintf(stderr, "[hg pipe] chunks=%zu copies done at %.2f ms, compute done at %.2f ms\n", chunks.size(), a, b);
    cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(t2);
  }
  return rc;
}

extern "C" int hg_kmer_hash(hg_ctx *c, const uint8_t *seq, const uint64_t *seg_off, uint32_t n,
                            const hg_sketch_params *p, uint64_t *hashes, uint64_t cap, uint64_t *hash_off) {
  if (!c || !seg_off || !hash_off) { hg_set_error("hg_kmer_hash: NULL argument"); return HG_E_INVALID; }
  int rc = check_params(p);
  if (rc) return rc;
  hash_off[0] = 0;
  if (n == 0) return HG_OK;
  if (!seq && seg_off[n] > 0) { hg_set_error("hg_kmer_hash: seq is NULL"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  const uint64_t lo = seg_off[0], hi = seg_off[n];
  void *d_seq;
  if ((rc = hg_scratch(c, HG_S_SEQ, hi - lo + 64, &d_seq))) return rc;
  const uint64_t shift = (uint64_t)((uintptr_t)(seq + lo) & 15);
  HG_CUDA(cudaMemcpyAsync((uint8_t *)d_seq + shift, seq + lo, hi - lo, cudaMemcpyHostToDevice, c->stream));
  std::vector<uint64_t> rel(n + 1);
  for (uint32_t g = 0; g <= n; g++) rel[g] = seg_off[g] - lo + shift;
  SketchPlan pl;
  if ((rc = make_plan(rel.data(), n, p, pl))) return rc;
  hg_genome_desc *d_desc; uint64_t *d_tables; uint32_t *d_counts;
  if ((rc = run_hash_stage(c, (const uint8_t *)d_seq, pl, n, p, &d_desc, &d_tables, &d_counts))) return rc;
  if ((rc = hg_launch_sort_tables(c, d_desc, n, d_tables, pl.max_slots))) return rc;
  std::vector<uint32_t> counts(n);
  HG_CUDA(cudaMemcpyAsync(counts.data(), d_counts, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  if ((rc = hg_sketch_status(c))) return rc;  // synchronises
  uint64_t total = 0;
  for (uint32_t g = 0; g < n; g++) { total += counts[g]; hash_off[g + 1] = total; }
  if (total > cap || (!hashes && total)) { hg_set_error("hg_kmer_hash: need room for %llu hashes", (unsigned long long)total); return HG_E_CAPACITY; }
  for (uint32_t g = 0; g < n; g++)
    if (counts[g])
      HG_CUDA(cudaMemcpyAsync(hashes + hash_off[g], d_tables + pl.desc[g].table_begin, (size_t)counts[g] * 8,
                              cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  return HG_OK;
}

// ---------------------------------------------------------------------------------------
// FASTA file bytes in (fastx_reader::read_merge_seq on the GPU)
// ---------------------------------------------------------------------------------------
uint32_t hg_fasta_block_bytes();
int hg_launch_fasta_merge(hg_ctx *ctx, const uint8_t *d_raw, const uint64_t *d_file_off, const uint64_t *d_blk_off,
                          uint32_t n_files, uint32_t max_blocks, void *d_sums, uint8_t *d_carry, uint64_t *d_out_off,
                          uint8_t *d_merged, uint64_t *d_merged_len);

struct FastaStage {
  std::vector<uint64_t> dev_off;   // device offset of every file (+ end)
  std::vector<uint64_t> merged_len;
  uint8_t *d_merged = nullptr;
};

// H2D of the raw files (one copy, files back to back), merge on the device, lengths back.
static int fasta_stage(hg_ctx *c, const uint8_t *raw, const uint64_t *file_off, uint32_t n, FastaStage &st) {
  const uint32_t B = hg_fasta_block_bytes();
  st.dev_off.assign(n + 1, 0);
  std::vector<uint64_t> blk_off(n + 1, 0);
  uint64_t max_blocks = 0;
  for (uint32_t f = 0; f < n; f++) {
    if (file_off[f + 1] < file_off[f]) { hg_set_error("file_off not monotone at %u", f); return HG_E_INVALID; }
    const uint64_t len = file_off[f + 1] - file_off[f];
    st.dev_off[f + 1] = st.dev_off[f] + len;  // files back to back, as in the host buffer: one copy (the kernels take any byte phase)
    const uint64_t nb = (len + B - 1) / B;
    blk_off[f + 1] = blk_off[f] + nb;
    max_blocks = std::max(max_blocks, nb);
  }
  if (max_blocks > 0x7fffffffull) { hg_set_error("file too large"); return HG_E_UNSUPPORTED; }
  const uint64_t total = st.dev_off[n], nblk = blk_off[n];
  void *d_raw, *d_merged, *d_meta, *d_sums, *h_meta;
  int rc;
  // device layout of the small arrays: file_off (n+1) | blk_off (n+1) | merged_len (n) | out_off (nblk) | carry (nblk bytes)
  const size_t meta_words = (size_t)(3 * n + 2) + nblk;
  if ((rc = hg_scratch(c, HG_S_PACKED, total + 64, &d_raw))) return rc;      // raw bytes (free until the sketches are packed)
  if ((rc = hg_scratch(c, HG_S_SEQ, total + 64, &d_merged))) return rc;      // merged sequences
  if ((rc = hg_scratch(c, HG_S_MISC, meta_words * 8 + nblk + 256, &d_meta))) return rc;
  if ((rc = hg_scratch(c, HG_S_TABLES, nblk * 16 + 256, &d_sums))) return rc;
  if ((rc = hg_pinned(c, 1, (size_t)(2 * n + 2) * 8 + (size_t)n * 8, &h_meta))) return rc;
  uint64_t *h_off = (uint64_t *)h_meta, *h_blk = h_off + (n + 1), *h_len = h_blk + (n + 1);
  memcpy(h_off, st.dev_off.data(), (n + 1) * 8);
  memcpy(h_blk, blk_off.data(), (n + 1) * 8);
  uint64_t *d_off = (uint64_t *)d_meta, *d_blk = d_off + (n + 1), *d_len = d_blk + (n + 1), *d_out = d_len + n;
  uint8_t *d_carry = (uint8_t *)(d_out + nblk);
  HG_CUDA(cudaMemcpyAsync(d_off, h_off, (size_t)(2 * n + 2) * 8, cudaMemcpyHostToDevice, c->stream));
  if (total) HG_CUDA(cudaMemcpyAsync(d_raw, raw + file_off[0], total, cudaMemcpyHostToDevice, c->stream));
  if ((rc = hg_launch_fasta_merge(c, (const uint8_t *)d_raw, d_off, d_blk, n, (uint32_t)max_blocks, d_sums, d_carry, d_out,
                                  (uint8_t *)d_merged, d_len)))
    return rc;
  HG_CUDA(cudaMemcpyAsync(h_len, d_len, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  st.merged_len.assign(h_len, h_len + n);
  st.d_merged = (uint8_t *)d_merged;
  return HG_OK;
}

extern "C" int hg_fasta_merge(hg_ctx *c, const uint8_t *raw, const uint64_t *file_off, uint32_t n, uint8_t *merged,
                              uint64_t cap, uint64_t *merged_off) {
  if (!c || !file_off || !merged_off) { hg_set_error("hg_fasta_merge: NULL argument"); return HG_E_INVALID; }
  merged_off[0] = 0;
  if (n == 0) return HG_OK;
  if (!raw && file_off[n] > file_off[0]) { hg_set_error("hg_fasta_merge: raw is NULL"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  FastaStage st;
  int rc = fasta_stage(c, raw, file_off, n, st);
  if (rc) return rc;
  for (uint32_t f = 0; f < n; f++) merged_off[f + 1] = merged_off[f] + st.merged_len[f];
  if (merged_off[n] > cap || (!merged && merged_off[n])) { hg_set_error("hg_fasta_merge: need room for %llu bytes", (unsigned long long)merged_off[n]); return HG_E_CAPACITY; }
  for (uint32_t f = 0; f < n; f++)
    if (st.merged_len[f])
      HG_CUDA(cudaMemcpyAsync(merged + merged_off[f], st.d_merged + st.dev_off[f], st.merged_len[f], cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  return HG_OK;
}

// Raw FASTA files in, sketches out.  Same ~64 MB chunk pipeline as hg_sketch_batch (copy stream +
// double-buffered staging), with the merge kernels in front of the hash kernel.  Nothing comes back
// to the host between the stages: tiles and tables are planned from the raw file sizes (an upper
// bound of the merged lengths) and the hash kernel reads the true lengths the merge left in HBM.
extern "C" int hg_sketch_fasta_batch(hg_ctx *c, const uint8_t *raw, const uint64_t *file_off, uint32_t n,
                                     const hg_sketch_params *p, int16_t *hv, uint8_t *packed, uint8_t *quant_bits,
                                     int32_t *norm2, uint32_t *n_hashes) {
  if (!c || !file_off) { hg_set_error("hg_sketch_fasta_batch: NULL argument"); return HG_E_INVALID; }
  int rc = check_params(p);
  if (rc) return rc;
  if (n == 0) return HG_OK;
  if (!raw && file_off[n] > file_off[0]) { hg_set_error("hg_sketch_fasta_batch: raw is NULL"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  const uint32_t D = p->hv_d, B = hg_fasta_block_bytes();
  const uint64_t chunk_bytes = hg_chunk_bytes();

  struct Chunk {
    uint32_t g0, g1;
    std::vector<uint64_t> dev_off, blk_off;  // per file of the chunk (+ end): offset in the slot, block prefix
    uint64_t max_blocks = 0;
    SketchPlan pl;
  };
  std::vector<Chunk> chunks;
  for (uint32_t g = 0; g < n;) {
    Chunk ch;
    ch.g0 = g;
    const uint64_t lo = file_off[g];
    do {
      if (file_off[g + 1] < file_off[g]) { hg_set_error("file_off not monotone at %u", g); return HG_E_INVALID; }
      ++g;
    } while (g < n && file_off[g + 1] - lo <= chunk_bytes);
    ch.g1 = g;
    const uint32_t m = ch.g1 - ch.g0;
    ch.dev_off.assign(m + 1, 0);
    ch.blk_off.assign(m + 1, 0);
    std::vector<uint64_t> lens(m);
    for (uint32_t t = 0; t < m; t++) {
      lens[t] = file_off[ch.g0 + t + 1] - file_off[ch.g0 + t];
      ch.dev_off[t + 1] = ch.dev_off[t] + lens[t];  // back to back, as in the host buffer
      const uint64_t nb = (lens[t] + B - 1) / B;
      ch.blk_off[t + 1] = ch.blk_off[t] + nb;
      ch.max_blocks = std::max(ch.max_blocks, nb);
    }
    if (ch.max_blocks > 0x7fffffffull) { hg_set_error("file too large"); return HG_E_UNSUPPORTED; }
    if ((rc = make_plan_bl(ch.dev_off.data(), lens.data(), m, p, ch.pl))) return rc;  // upper bounds
    chunks.push_back(std::move(ch));
  }
  uint64_t max_bytes = 0, max_slots = 0, max_tiles = 0, max_blk = 0;
  uint32_t max_m = 0;
  for (const Chunk &ch : chunks) {
    max_bytes = std::max(max_bytes, ch.dev_off.back());
    max_slots = std::max<uint64_t>(max_slots, ch.pl.total_slots);
    max_tiles = std::max<uint64_t>(max_tiles, ch.pl.n_tiles);
    max_blk = std::max(max_blk, ch.blk_off.back());
    max_m = std::max(max_m, ch.g1 - ch.g0);
  }
  const uint64_t slot_bytes = (max_bytes + 64 + 255) & ~255ull;
  const size_t n_desc = (size_t)n + chunks.size();
  // per-chunk merge metadata in one buffer: file_off | blk_off | out_off | sums | carry
  const size_t meta_bytes = ((size_t)(2 * (max_m + 1)) * 8 + max_blk * 8 + max_blk * 16 + max_blk + 256 + 15) & ~(size_t)15;
  const size_t h_meta_per_chunk = (size_t)(2 * (max_m + 1)) * 8;

  void *d_rawbuf, *d_merged, *d_meta, *d_desc, *d_tables, *d_counts, *d_hv = nullptr, *d_packed, *d_small, *d_map, *h_desc,
      *h_meta;
  if ((rc = hg_scratch(c, HG_S_REF_LIMBS, 2 * slot_bytes, &d_rawbuf))) return rc;  // raw files, double buffered
  if ((rc = hg_scratch(c, HG_S_SEQ, slot_bytes, &d_merged))) return rc;
  if ((rc = hg_scratch(c, HG_S_QRY_LIMBS, meta_bytes + (size_t)n * 8, &d_meta))) return rc;
  if ((rc = hg_scratch(c, HG_S_DESC, sizeof(hg_genome_desc) * n_desc, &d_desc))) return rc;
  if ((rc = hg_scratch(c, HG_S_TABLES, max_slots * 8, &d_tables))) return rc;
  if ((rc = hg_scratch(c, HG_S_COUNTS, sizeof(uint32_t) * n, &d_counts))) return rc;
  if (hv && (rc = hg_scratch(c, HG_S_HV, (size_t)n * D * 2, &d_hv))) return rc;
  if ((rc = hg_scratch(c, HG_S_PACKED, (size_t)n * D * 2, &d_packed))) return rc;
  if ((rc = hg_scratch(c, HG_S_SMALL, (size_t)n * 12, &d_small))) return rc;
  if ((rc = hg_scratch(c, HG_S_MISC, (max_tiles / hg_kmer_tiles_per_cta() + 2) * 4 + 256, &d_map))) return rc;
  if ((rc = hg_pinned(c, 0, sizeof(hg_genome_desc) * n_desc, &h_desc))) return rc;
  if ((rc = hg_pinned(c, 1, h_meta_per_chunk * chunks.size(), &h_meta))) return rc;
  uint8_t *d_bits = (uint8_t *)d_small + (size_t)n * 8;
  int32_t *d_norm = (int32_t *)d_small;
  uint32_t *d_nh = (uint32_t *)d_small + n;
  uint64_t *d_len = (uint64_t *)((uint8_t *)d_meta + meta_bytes);  // true merged lengths, all files
  uint64_t *dm_off = (uint64_t *)d_meta, *dm_blk = dm_off + (max_m + 1), *dm_out = dm_blk + (max_m + 1);
  uint8_t *dm_sums = (uint8_t *)(dm_out + max_blk);
  uint8_t *dm_carry = dm_sums + max_blk * 16;
  if (!c->copy_stream) {
    HG_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      HG_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
      HG_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
  }
  c->ev_used = 0;
  HG_CUDA(cudaMemsetAsync(c->d_status, 0, 4 * sizeof(uint32_t), c->stream));
  HG_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * n, c->stream));
  HG_CUDA(cudaEventRecord(c->ev_done[0], c->stream));
  HG_CUDA(cudaEventRecord(c->ev_done[1], c->stream));

  size_t desc_pos = 0;
  for (size_t ci = 0; ci < chunks.size(); ci++) {
    const Chunk &ch = chunks[ci];
    const int slot = (int)(ci & 1);
    const uint32_t m = ch.g1 - ch.g0;
    uint8_t *d_raw = (uint8_t *)d_rawbuf + slot * slot_bytes;
    // ---- copy stream: the chunk's files in one copy (per-file copies cost ~6 us each of PCIe idle time) ----
    HG_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_done[slot], 0));
    if (ch.dev_off[m]) HG_CUDA(cudaMemcpyAsync(d_raw, raw + file_off[ch.g0], ch.dev_off[m], cudaMemcpyHostToDevice, c->copy_stream));
    HG_CUDA(cudaEventRecord(c->ev_copied[slot], c->copy_stream));
    // ---- compute stream ----
    uint64_t *hm = (uint64_t *)((uint8_t *)h_meta + ci * h_meta_per_chunk);
    memcpy(hm, ch.dev_off.data(), (m + 1) * 8);
    memcpy(hm + (max_m + 1), ch.blk_off.data(), (m + 1) * 8);
    HG_CUDA(cudaMemcpyAsync(dm_off, hm, h_meta_per_chunk, cudaMemcpyHostToDevice, c->stream));
    hg_genome_desc *hd = (hg_genome_desc *)h_desc + desc_pos, *dd = (hg_genome_desc *)d_desc + desc_pos;
    memcpy(hd, ch.pl.desc.data(), sizeof(hg_genome_desc) * (m + 1));
    desc_pos += m + 1;
    HG_CUDA(cudaMemcpyAsync(dd, hd, sizeof(hg_genome_desc) * (m + 1), cudaMemcpyHostToDevice, c->stream));
    HG_CUDA(cudaMemsetAsync(d_tables, 0xFF, ch.pl.total_slots * 8, c->stream));
    HG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[slot], 0));
    if ((rc = hg_launch_fasta_merge(c, d_raw, dm_off, dm_blk, m, (uint32_t)ch.max_blocks, dm_sums, dm_carry, dm_out,
                                    (uint8_t *)d_merged, d_len + ch.g0)))
      return rc;
    HG_CUDA(cudaEventRecord(c->ev_done[slot], c->stream));  // the raw slot may be overwritten now
    c->d_actual_len = d_len + ch.g0;
    rc = hg_launch_kmer_hash(c, (const uint8_t *)d_merged, dd, m, ch.pl.n_tiles, p, (uint64_t *)d_tables,
                             (uint32_t *)d_counts + ch.g0);
    c->d_actual_len = nullptr;
    if (rc) return rc;
    if ((rc = hg_launch_encode(c, dd, m, (const uint64_t *)d_tables, (const uint32_t *)d_counts + ch.g0, D,
                               d_hv ? (int16_t *)d_hv + (size_t)ch.g0 * D : nullptr,
                               (uint8_t *)d_packed + (size_t)ch.g0 * 2 * D, d_bits + ch.g0, d_norm + ch.g0, d_nh + ch.g0)))
      return rc;
  }
  if (hv) HG_CUDA(cudaMemcpyAsync(hv, d_hv, (size_t)n * D * 2, cudaMemcpyDeviceToHost, c->stream));
  if (packed) HG_CUDA(cudaMemcpyAsync(packed, d_packed, (size_t)n * D * 2, cudaMemcpyDeviceToHost, c->stream));
  if (quant_bits) HG_CUDA(cudaMemcpyAsync(quant_bits, d_bits, n, cudaMemcpyDeviceToHost, c->stream));
  if (norm2) HG_CUDA(cudaMemcpyAsync(norm2, d_norm, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  if (n_hashes) HG_CUDA(cudaMemcpyAsync(n_hashes, d_nh, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  return hg_sketch_status(c);
}

// ---------------------------------------------------------------------------------------
// sketch format
// ---------------------------------------------------------------------------------------
extern "C" int hg_unpack_dev(hg_ctx *c, const uint8_t *d_packed, uint64_t row_stride, const uint8_t *d_bits,
                             uint32_t n, uint32_t hv_d, int16_t *d_hv) {
  if (!c || !d_packed || !d_bits || !d_hv) { hg_set_error("hg_unpack_dev: NULL argument"); return HG_E_INVALID; }
  if (hv_d == 0 || hv_d % 256 != 0 || row_stride % 4 != 0) { hg_set_error("hg_unpack_dev: hv_d %% 256 or row_stride %% 4"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  return hg_launch_unpack(c, d_packed, row_stride, d_bits, n, hv_d, d_hv);
}

extern "C" int hg_unpack(hg_ctx *c, const uint8_t *packed, uint64_t row_stride, const uint8_t *bits, uint32_t n,
                         uint32_t hv_d, int16_t *hv) {
  if (!c || !packed || !bits || !hv) { hg_set_error("hg_unpack: NULL argument"); return HG_E_INVALID; }
  if (n == 0) return HG_OK;
  for (uint32_t g = 0; g < n; g++)
    if (bits[g] < 1 || bits[g] > 16 || (uint64_t)bits[g] * hv_d / 8 > row_stride) {
      hg_set_error("hg_unpack: sketch %u has hv_quant_bits %u (row_stride %llu)", g, (unsigned)bits[g], (unsigned long long)row_stride);
      return HG_E_INVALID;
    }
  HG_CUDA(cudaSetDevice(c->device));
  int rc;
  void *d_p, *d_b, *d_h;
  if ((rc = hg_scratch(c, HG_S_PACKED, (size_t)n * row_stride, &d_p))) return rc;
  if ((rc = hg_scratch(c, HG_S_SMALL, n, &d_b))) return rc;
  if ((rc = hg_scratch(c, HG_S_HV, (size_t)n * hv_d * 2, &d_h))) return rc;
  HG_CUDA(cudaMemcpyAsync(d_p, packed, (size_t)n * row_stride, cudaMemcpyHostToDevice, c->stream));
  HG_CUDA(cudaMemcpyAsync(d_b, bits, n, cudaMemcpyHostToDevice, c->stream));
  if ((rc = hg_unpack_dev(c, (const uint8_t *)d_p, row_stride, (const uint8_t *)d_b, n, hv_d, (int16_t *)d_h))) return rc;
  HG_CUDA(cudaMemcpyAsync(hv, d_h, (size_t)n * hv_d * 2, cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  return HG_OK;
}

// ---------------------------------------------------------------------------------------
// dist
// ---------------------------------------------------------------------------------------
static const int32_t HG_TC_MAX_ABS = 8127;  // |x| <= 8127 splits into two s8 limbs (x = 128 h + l)

// allow_narrow = false: the single-plane path has already been tried (and declined) on these matrices, absmax is
// max |hv| from its pre-pass
static int dist_dev_impl(hg_ctx *c, const int16_t *d_ref, const int32_t *d_ref_norm, uint32_t n_ref, uint32_t i0,
                         const int16_t *d_qry, const int32_t *d_qry_norm, uint32_t n_qry, uint32_t j0,
                         uint32_t hv_d, uint32_t ksize, float ani_th, int symmetric, int path, hg_hit *d_hits,
                         uint64_t cap, unsigned long long *d_n_hits, bool allow_narrow, int32_t absmax, bool defer_forced = false) {
  if (!c || !d_n_hits) { hg_set_error("hg_dist_dev: NULL argument"); return HG_E_INVALID; }
  c->pending_stats[0] = c->pending_stats[1] = nullptr;
  if ((n_ref && (!d_ref || !d_ref_norm)) || (n_qry && (!d_qry || !d_qry_norm)) || (cap && !d_hits)) {
    hg_set_error("hg_dist_dev: NULL argument"); return HG_E_INVALID;
  }
  if (hv_d == 0 || hv_d % 256 != 0) { hg_set_error("hg_dist_dev: hv_d %u must be a multiple of 256", hv_d); return HG_E_INVALID; }
  if (path < 0 || path > 3) { hg_set_error("hg_dist_dev: path %d", path); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  HG_CUDA(cudaMemsetAsync(d_n_hits, 0, sizeof(unsigned long long), c->stream));
  if (n_ref == 0 || n_qry == 0) return HG_OK;
  c->ev_used &= ~(3 << 4);
  int rc;

  // ---- choose: 3 = single-plane tensor kernel, 2 = two-limb tensor kernel, 1 = SIMT ----
  int use = path;  // absmax: max |hv| once some scan has produced it (-1: not yet)
  if (path == 0 && (uint64_t)n_ref * n_qry < 128ull * 128ull) {
    use = 1;
    snprintf(c->dist_reason, sizeof(c->dist_reason), "SIMT: %u x %u pairs do not fill one 128x128 tensor tile", n_ref, n_qry);
  } else if (!allow_narrow && (path == 0 || path == 3)) {
    if (path == 3) { hg_set_error("rows are not narrow (single-plane path declined)"); return HG_E_UNSUPPORTED; }
    snprintf(c->dist_reason, sizeof(c->dist_reason), "rows are not narrow");
    hg_set_error("rows are not narrow");
    use = 0;
  } else if (path == 0 || path == 3) {
    // the narrow path's own pre-pass (one read of both matrices) tells whether the rows fit one s8 plane
    uint64_t outliers = 0;
    rc = hg_launch_dist_narrow(c, d_ref, d_ref_norm, n_ref, i0, d_qry, d_qry_norm, n_qry, j0, hv_d, ksize, ani_th, symmetric,
                               d_hits, cap, d_n_hits, &absmax, &outliers, path == 3 && defer_forced);
    if (rc == HG_OK) {
      c->dist_path = 3;
      if (path == 3)
        snprintf(c->dist_reason, sizeof(c->dist_reason), "tensor-narrow: forced by caller%s", defer_forced ? " (verdict deferred to hg_dist_status)" : "");
      else
        snprintf(c->dist_reason, sizeof(c->dist_reason),
                 "tensor-narrow: rows fit one s8 plane as x = 2a + s (max |hv| = %d, %llu outlier elements corrected per candidate); "
                 "tcgen05 kind::i8, one MMA per K step", absmax, (unsigned long long)outliers);
      return HG_OK;
    }
    if (rc != HG_E_UNSUPPORTED || path == 3) return rc;
    use = 0;  // declined: two-limb tensor kernel if every |hv| fits 13 bits, else SIMT
  }
  if (use == 0) {
    char why[96];
    snprintf(why, sizeof(why), "%s", hg_last_error());
    if (absmax < 0) {  // the narrow path declined before its scan: one read of both matrices and one 4-byte D2H each
      void *d_max;
      if ((rc = hg_scratch(c, HG_S_MISC, 256, &d_max))) return rc;
      int32_t m1 = 0, m2 = 0;
      if ((rc = hg_launch_absmax(c, d_ref, (uint64_t)n_ref * hv_d, (int32_t *)d_max))) return rc;
      HG_CUDA(cudaMemcpyAsync(&m1, d_max, 4, cudaMemcpyDeviceToHost, c->stream));
      HG_CUDA(cudaStreamSynchronize(c->stream));
      if (d_qry != d_ref) {
        if ((rc = hg_launch_absmax(c, d_qry, (uint64_t)n_qry * hv_d, (int32_t *)d_max))) return rc;
        HG_CUDA(cudaMemcpyAsync(&m2, d_max, 4, cudaMemcpyDeviceToHost, c->stream));
        HG_CUDA(cudaStreamSynchronize(c->stream));
      }
      absmax = m1 > m2 ? m1 : m2;
    }
    if (absmax > HG_TC_MAX_ABS) {
      use = 1;
      snprintf(c->dist_reason, sizeof(c->dist_reason),
               "SIMT: max |hv| = %d exceeds the 13-bit budget (%d) of the int8 limb split", absmax, HG_TC_MAX_ABS);
    } else {
      use = 2;
      snprintf(c->dist_reason, sizeof(c->dist_reason),
               "tensor: max |hv| = %d fits two s8 limbs, tcgen05 kind::i8 (one s8 plane declined: %.80s)", absmax, why);
    }
  } else if (path != 0) {
    snprintf(c->dist_reason, sizeof(c->dist_reason), "%s: forced by caller", use == 1 ? "SIMT" : "tensor");
  }
  c->dist_path = use;
  int rc2 = HG_OK;
  if (use == 2) {
    rc2 = hg_launch_dist_tc(c, d_ref, d_ref_norm, n_ref, i0, d_qry, d_qry_norm, n_qry, j0, hv_d, ksize, ani_th,
                            symmetric, d_hits, cap, d_n_hits);
    if (rc2 == HG_E_UNSUPPORTED && (path == 0 || c->tc_is_hint)) {  // shape outside the tensor kernel's tiling: exact SIMT path
      use = c->dist_path = 1;
      snprintf(c->dist_reason, sizeof(c->dist_reason), "SIMT: tensor kernel declined this shape (%s)", hg_last_error());
    }
  }
  if (use != 2) {
    HG_PROF(c, 4);
    rc2 = hg_launch_dist_simt(c, d_ref, d_ref_norm, n_ref, i0, d_qry, d_qry_norm, n_qry, j0, hv_d, ksize, ani_th,
                              symmetric, d_hits, cap, d_n_hits);
    HG_PROF(c, 5);
  }
  return rc2;
}

extern "C" int hg_dist_dev(hg_ctx *c, const int16_t *d_ref, const int32_t *d_ref_norm, uint32_t n_ref, uint32_t i0,
                           const int16_t *d_qry, const int32_t *d_qry_norm, uint32_t n_qry, uint32_t j0,
                           uint32_t hv_d, uint32_t ksize, float ani_th, int symmetric, int path, hg_hit *d_hits,
                           uint64_t cap, unsigned long long *d_n_hits) {
  return dist_dev_impl(c, d_ref, d_ref_norm, n_ref, i0, d_qry, d_qry_norm, n_qry, j0, hv_d, ksize, ani_th, symmetric, path, d_hits,
                       cap, d_n_hits, true, -1, true);
}

// After hg_dist_dev(path = 3): the pre-pass verdict that call did not wait for.
extern "C" int hg_dist_status(hg_ctx *c) {
  if (!c) { hg_set_error("ctx is NULL"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  uint32_t st[2][4 * HG_MAX_PEERS] = {{0}, {0}};
  for (int k = 0; k < 2; ++k)
    if (c->pending_stats[k] && c->pending_nsets[k] <= HG_MAX_PEERS)
      HG_CUDA(cudaMemcpyAsync(st[k], c->pending_stats[k], 16 * c->pending_nsets[k], cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 2; ++k)
    for (uint32_t t = 0; c->pending_stats[k] && t < c->pending_nsets[k]; ++t)
      if (st[k][4 * t + 3]) {
        hg_set_error("the forced single-plane dist did nothing: its rows are not narrow (outlier budget exceeded)");
        return HG_E_UNSUPPORTED;
      }
  return HG_OK;
}

extern "C" int hg_dist_last_path(hg_ctx *c) { return c ? c->dist_path : 0; }
extern "C" const char *hg_dist_last_reason(hg_ctx *c) { return c ? c->dist_reason : ""; }

// shared tail of the host-pointer dist entries: hit count back, output-stage sort if asked, hits back
static int dist_tail(hg_ctx *c, void *d_hits, void *d_cnt, hg_hit *hits, uint64_t cap, uint64_t *n_hits, bool sorted,
                     uint32_t *ani_milli) {
  int rc;
  unsigned long long cnt = 0;
  HG_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  *n_hits = cnt;
  if (cnt > cap) {
    hg_set_error("hg_dist: %llu pairs pass the threshold, capacity is %llu", cnt, (unsigned long long)cap);
    return HG_E_CAPACITY;
  }
  if (cnt) {
    void *d_milli = nullptr;
    if (sorted) {  // output stage on the device (src/utils.rs:262-269)
      if (ani_milli && (rc = hg_scratch(c, HG_S_MISC, cnt * 4 + 256, &d_milli))) return rc;
      if ((rc = hg_launch_sort_hits(c, (hg_hit *)d_hits, cnt, (uint32_t *)d_milli))) return rc;
    }
    HG_CUDA(cudaMemcpyAsync(hits, d_hits, cnt * sizeof(hg_hit), cudaMemcpyDeviceToHost, c->stream));
    if (d_milli) HG_CUDA(cudaMemcpyAsync(ani_milli, d_milli, cnt * 4, cudaMemcpyDeviceToHost, c->stream));
    HG_CUDA(cudaStreamSynchronize(c->stream));  // unsorted: append order is unspecified
  }
  return HG_OK;
}

static int dist_out_buffers(hg_ctx *c, uint64_t cap, void **d_hits, void **d_cnt) {
  int rc;
  if ((rc = hg_scratch(c, HG_S_PACKED, cap * sizeof(hg_hit) + 256, d_hits))) return rc;
  return hg_scratch(c, HG_S_COUNTS, 256, d_cnt);
}

// matrices already on the device: dist kernel (auto / forced path) + tail
static int dist_finish(hg_ctx *c, const int16_t *d_ref, const int32_t *d_rn, uint32_t n_ref, const int16_t *d_qry,
                       const int32_t *d_qn, uint32_t n_qry, uint32_t hv_d, uint32_t ksize, float ani_th, int symmetric, int path,
                       hg_hit *hits, uint64_t cap, uint64_t *n_hits, bool sorted, uint32_t *ani_milli, bool allow_narrow = true,
                       int32_t absmax = -1) {
  int rc;
  void *d_hits, *d_cnt;
  if ((rc = dist_out_buffers(c, cap, &d_hits, &d_cnt))) return rc;
  rc = dist_dev_impl(c, d_ref, d_rn, n_ref, 0, d_qry, d_qn, n_qry, 0, hv_d, ksize, ani_th, symmetric, path, (hg_hit *)d_hits, cap,
                     (unsigned long long *)d_cnt, allow_narrow, absmax);
  if (rc) return rc;
  return dist_tail(c, d_hits, d_cnt, hits, cap, n_hits, sorted, ani_milli);
}

// ---- host matrices streamed in row chunks: H2D of chunk c + 1 under unpack / pre-pass / dist kernel of chunk c ----
struct HostRows {
  const uint8_t *base;  // first row on the host
  uint64_t stride;      // bytes between host rows
  uint64_t width;       // bytes of a row that travel (int16 rows: 2 hv_d; packed rows: max hv_quant_bits * hv_d / 8)
  const uint8_t *bits;  // packed rows: hv_quant_bits per row (host); int16 rows: NULL
  const int32_t *norm;  // host
  uint32_t n;
  uint8_t *d_stage;     // packed rows: where they land (width apart); int16 rows: the matrix itself
  uint8_t *d_bits;      // packed rows
  int16_t *d_hv;
  int32_t *d_norm;
};

static bool dist_stream_eligible(int path, bool same, int symmetric, uint32_t n_ref, uint32_t n_qry, uint32_t hv_d,
                                 const void *d_a, const void *d_b) {
  if (path != 0 && path != 3) return false;
  if ((uint64_t)n_ref * n_qry < 128ull * 128ull && path == 0) return false;  // SIMT territory
  if (same && !symmetric) return false;  // full n x n of one matrix: both (i, j) and (j, i) - not worth a special walk
  return hg_narrow_shape_ok(hv_d, d_a, d_b) == HG_OK;
}

// Dist with the rows arriving in chunks.  same: ONE matrix, symmetric all-vs-all - chunk c is compared against
// every row up to its own end (j in the chunk, i < j).  Otherwise the queries go first, whole, and every ref chunk
// is compared against all of them.
//   narrow = true: single-plane kernel; returns HG_E_UNSUPPORTED (with max |hv|) if the rows turn out not to be
//     narrow: by then both matrices are complete in HBM and the caller runs another kernel on them.
//   narrow = false: two-limb kernel (the caller knows every |hv| fits 13 bits: packed rows of <= 13 bits), the
//     rows of a chunk split into the matrix-wide limb planes when they arrive.
static int dist_stream(hg_ctx *c, HostRows &R, HostRows &Q, bool same, uint32_t hv_d, uint32_t ksize, float ani_th, int symmetric,
                       bool narrow, hg_hit *d_hits, uint64_t cap, unsigned long long *d_cnt, int32_t *absmax, uint64_t *outliers,
                       uint32_t *n_chunks_out) {
  int rc;
  if (!c->copy_stream) {
    HG_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      HG_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
      HG_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
  }
  if (!c->ev_chunk[0])
    for (int i = 0; i < 10; i++) HG_CUDA(cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming));
  hg_narrow_mat Qm, Rm;
  hg_tc_mat Qt, Rt;
  if (narrow) {
    const size_t mq = hg_narrow_meta_bytes(Q.n), mr = same ? 0 : hg_narrow_meta_bytes(R.n);
    void *p_meta;
    if ((rc = hg_scratch(c, HG_S_NARROW_META, mq + mr, &p_meta))) return rc;
    if ((rc = hg_narrow_setup(c, Q.d_hv, Q.n, hv_d, HG_S_QRY_LIMBS, p_meta, &Qm))) return rc;
    if (!same && (rc = hg_narrow_setup(c, R.d_hv, R.n, hv_d, HG_S_REF_LIMBS, (uint8_t *)p_meta + mq, &Rm))) return rc;
  } else {
    if ((rc = hg_tc_setup(c, Q.d_hv, Q.n, hv_d, HG_S_QRY_LIMBS, &Qt))) return rc;
    if (!same && (rc = hg_tc_setup(c, R.d_hv, R.n, hv_d, HG_S_REF_LIMBS, &Rt))) return rc;
  }
  HG_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c->stream));
  auto small = [&](HostRows &S) -> int {
    HG_CUDA(cudaMemcpyAsync(S.d_norm, S.norm, (size_t)S.n * 4, cudaMemcpyHostToDevice, c->stream));
    if (S.bits) HG_CUDA(cudaMemcpyAsync(S.d_bits, S.bits, S.n, cudaMemcpyHostToDevice, c->stream));
    return HG_OK;
  };
  if ((rc = small(R))) return rc;
  if (!same && (rc = small(Q))) return rc;
  // the copy stream starts once the compute stream has got here (the buffers may still be in use by an earlier call)
  HG_CUDA(cudaEventRecord(c->ev_chunk[0], c->stream));
  HG_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_chunk[0], 0));
  auto stage = [&](HostRows &S, uint32_t a, uint32_t rows, cudaEvent_t ev) -> int {  // copy stream
    HG_CUDA(cudaMemcpy2DAsync(S.d_stage + (size_t)a * S.width, S.width, S.base + (size_t)a * S.stride, S.stride, S.width, rows,
                              cudaMemcpyHostToDevice, c->copy_stream));
    HG_CUDA(cudaEventRecord(ev, c->copy_stream));
    return HG_OK;
  };
  auto ready = [&](HostRows &S, const hg_narrow_mat *M, const hg_tc_mat *T, uint32_t a, uint32_t rows, cudaEvent_t ev) -> int {  // compute stream
    HG_CUDA(cudaStreamWaitEvent(c->stream, ev, 0));
    if (S.bits) {
      int r2 = hg_launch_unpack(c, S.d_stage + (size_t)a * S.width, S.width, S.d_bits + a, rows, hv_d, S.d_hv + (size_t)a * hv_d);
      if (r2) return r2;
    }
    return M ? hg_narrow_prep_rows(c, M, a, rows) : hg_tc_split_rows(c, T, a, rows);
  };
  uint32_t chunk_rows = 0;
  if (const char *e = getenv("HG_DIST_CHUNK_ROWS")) chunk_rows = (uint32_t)std::max(0, atoi(e));
  if (chunk_rows == 0) {
    const uint32_t want = std::min<uint32_t>(8, std::max<uint32_t>(1, R.n / 1280));  // measured on config 3: 6-8 chunks beat 2-4
    chunk_rows = (R.n + want - 1) / want;
  }
  chunk_rows = std::max<uint32_t>(256, (chunk_rows + 255) & ~255u);
  if ((uint64_t)chunk_rows * 8 < R.n) chunk_rows = (uint32_t)((((uint64_t)R.n + 7) / 8 + 255) & ~255ull);  // at most 8 chunks
  // all copies are queued first so that PCIe never waits for the host
  std::vector<std::pair<uint32_t, uint32_t>> chunks;
  for (uint32_t a = 0; a < R.n; a += chunk_rows) chunks.push_back({a, std::min(chunk_rows, R.n - a)});
  if (!same && (rc = stage(Q, 0, Q.n, c->ev_chunk[1]))) return rc;
  for (size_t k = 0; k < chunks.size(); k++)
    if ((rc = stage(R, chunks[k].first, chunks[k].second, c->ev_chunk[2 + k]))) return rc;
  if (!same && (rc = ready(Q, narrow ? &Qm : nullptr, &Qt, 0, Q.n, c->ev_chunk[1]))) return rc;
  for (size_t k = 0; k < chunks.size(); k++) {
    const uint32_t a = chunks[k].first, rows = chunks[k].second;
    if ((rc = ready(R, narrow ? (same ? &Qm : &Rm) : nullptr, same ? &Qt : &Rt, a, rows, c->ev_chunk[2 + k]))) return rc;
    if (narrow) {
      if (same)  // pairs (i, j): j in this chunk, i < j
        rc = hg_narrow_launch(c, &Qm, 0, a + rows, 0, R.d_norm, &Qm, a, rows, a, R.d_norm + a, ksize, ani_th, 1, d_hits, cap, d_cnt);
      else
        rc = hg_narrow_launch(c, &Rm, a, rows, a, R.d_norm + a, &Qm, 0, Q.n, 0, Q.d_norm, ksize, ani_th, symmetric, d_hits, cap, d_cnt);
    } else if (same) {
      rc = hg_tc_launch(c, &Qt, 0, a + rows, 0, R.d_norm, &Qt, a, rows, a, R.d_norm + a, ksize, ani_th, 1, d_hits, cap, d_cnt);
    } else {
      rc = hg_tc_launch(c, &Rt, a, rows, a, R.d_norm + a, &Qt, 0, Q.n, 0, Q.d_norm, ksize, ani_th, symmetric, d_hits, cap, d_cnt);
    }
    if (rc) return rc;
  }
  if (n_chunks_out) *n_chunks_out = (uint32_t)chunks.size();
  if (!narrow) return HG_OK;
  return hg_narrow_verdict(c, &Qm, same ? nullptr : &Rm, absmax, outliers);
}

// common driver of the host entries: streamed single-plane path if eligible, else (or if it declines) whole
// copies + the kernel the rows call for
static int dist_from_host(hg_ctx *c, HostRows &R, HostRows &Q, bool same, uint32_t hv_d, uint32_t ksize, float ani_th, int symmetric,
                          int path, int stream_mode /*0 no, 1 single-plane, 2 two-limb*/, hg_hit *hits, uint64_t cap,
                          uint64_t *n_hits, bool sorted, uint32_t *ani_milli) {
  int rc;
  bool allow_narrow = true;
  int32_t absmax = -1;
  const bool shape_ok = !(same && !symmetric) && hg_narrow_shape_ok(hv_d, R.d_hv, Q.d_hv) == HG_OK;
  bool stream_on = true;
  if (const char *e = getenv("HG_DIST_STREAM")) stream_on = atoi(e) != 0;
  if (stream_on && stream_mode == 1 && dist_stream_eligible(path, same, symmetric, R.n, Q.n, hv_d, R.d_hv, Q.d_hv)) {
    void *d_hits, *d_cnt;
    if ((rc = dist_out_buffers(c, cap, &d_hits, &d_cnt))) return rc;
    uint64_t outliers = 0;
    uint32_t n_chunks = 0;
    rc = dist_stream(c, R, Q, same, hv_d, ksize, ani_th, symmetric, true, (hg_hit *)d_hits, cap, (unsigned long long *)d_cnt, &absmax,
                     &outliers, &n_chunks);
    if (rc == HG_OK) {
      c->dist_path = 3;
      c->ev_used &= ~(3 << 4);
      snprintf(c->dist_reason, sizeof(c->dist_reason),
               "tensor-narrow%s: rows fit one s8 plane as x = 2a + s (max |hv| = %d, %llu outlier elements corrected per candidate); "
               "tcgen05 kind::i8, one MMA per K step; %u row chunk(s) overlapping their H2D",
               path == 3 ? " (forced)" : "", absmax, (unsigned long long)outliers, n_chunks);
      return dist_tail(c, d_hits, d_cnt, hits, cap, n_hits, sorted, ani_milli);
    }
    if (rc != HG_E_UNSUPPORTED || path == 3) return rc;
    allow_narrow = false;  // both matrices are complete in HBM now: another kernel on the same data
  } else if (stream_on && stream_mode == 2 && shape_ok && hv_d % 512 == 0 && (uint64_t)R.n * Q.n >= 128ull * 128ull) {
    void *d_hits, *d_cnt;
    if ((rc = dist_out_buffers(c, cap, &d_hits, &d_cnt))) return rc;
    uint32_t n_chunks = 0;
    rc = dist_stream(c, R, Q, same, hv_d, ksize, ani_th, symmetric, false, (hg_hit *)d_hits, cap, (unsigned long long *)d_cnt, nullptr,
                     nullptr, &n_chunks);
    if (rc) return rc;
    c->dist_path = 2;
    c->ev_used &= ~(3 << 4);
    snprintf(c->dist_reason, sizeof(c->dist_reason),
             "tensor: hv_quant_bits fit two s8 limbs; tcgen05 kind::i8; %u row chunk(s) overlapping their H2D", n_chunks);
    return dist_tail(c, d_hits, d_cnt, hits, cap, n_hits, sorted, ani_milli);
  } else {
    auto whole = [&](HostRows &S) -> int {
      HG_CUDA(cudaMemcpy2DAsync(S.d_stage, S.width, S.base, S.stride, S.width, S.n, cudaMemcpyHostToDevice, c->stream));
      HG_CUDA(cudaMemcpyAsync(S.d_norm, S.norm, (size_t)S.n * 4, cudaMemcpyHostToDevice, c->stream));
      if (S.bits) {
        HG_CUDA(cudaMemcpyAsync(S.d_bits, S.bits, S.n, cudaMemcpyHostToDevice, c->stream));
        return hg_launch_unpack(c, S.d_stage, S.width, S.d_bits, S.n, hv_d, S.d_hv);
      }
      return HG_OK;
    };
    if ((rc = whole(R))) return rc;
    if (!same && (rc = whole(Q))) return rc;
  }
  return dist_finish(c, R.d_hv, R.d_norm, R.n, Q.d_hv, Q.d_norm, Q.n, hv_d, ksize, ani_th, symmetric, path, hits, cap, n_hits, sorted,
                     ani_milli, allow_narrow, absmax);
}

static int dist_host(hg_ctx *c, const int16_t *ref, const int32_t *ref_norm, uint32_t n_ref, const int16_t *qry,
                     const int32_t *qry_norm, uint32_t n_qry, uint32_t hv_d, uint32_t ksize, float ani_th, int symmetric,
                     int path, hg_hit *hits, uint64_t cap, uint64_t *n_hits, bool sorted, uint32_t *ani_milli) {
  if (!c || !n_hits) { hg_set_error("hg_dist: NULL argument"); return HG_E_INVALID; }
  *n_hits = 0;
  if ((n_ref && (!ref || !ref_norm)) || (n_qry && (!qry || !qry_norm)) || (cap && !hits)) {
    hg_set_error("hg_dist: NULL argument"); return HG_E_INVALID;
  }
  if (n_ref == 0 || n_qry == 0) return HG_OK;
  HG_CUDA(cudaSetDevice(c->device));
  int rc;
  const bool same = (ref == qry && ref_norm == qry_norm && n_ref == n_qry);
  const size_t rb = (size_t)n_ref * hv_d * 2, qb = same ? 0 : (size_t)n_qry * hv_d * 2;
  void *d_mat, *d_norm;
  if ((rc = hg_scratch(c, HG_S_HV, rb + qb + 512, &d_mat))) return rc;
  if ((rc = hg_scratch(c, HG_S_SMALL, ((size_t)n_ref + n_qry) * 4 + 256, &d_norm))) return rc;
  int16_t *d_ref = (int16_t *)d_mat;
  const size_t rb_al = (rb + 255) & ~(size_t)255;
  int16_t *d_qry = same ? d_ref : (int16_t *)((uint8_t *)d_mat + rb_al);
  int32_t *d_rn = (int32_t *)d_norm, *d_qn = same ? d_rn : d_rn + n_ref;
  HostRows R = {(const uint8_t *)ref, (uint64_t)hv_d * 2, (uint64_t)hv_d * 2, nullptr, ref_norm, n_ref, (uint8_t *)d_ref, nullptr, d_ref, d_rn};
  HostRows Q = {(const uint8_t *)qry, (uint64_t)hv_d * 2, (uint64_t)hv_d * 2, nullptr, qry_norm, n_qry, (uint8_t *)d_qry, nullptr, d_qry, d_qn};
  return dist_from_host(c, R, same ? R : Q, same, hv_d, ksize, ani_th, symmetric, path, 1, hits, cap, n_hits, sorted, ani_milli);
}

// `dist` straight from what the sketch file holds (FileSketch.hv bit-packed + hv_quant_bits + hv_norm_2,
// src/types.rs:224-235): only the packed bytes cross PCIe (b/16 of the i16 matrix), decompress_file_sketch
// (src/hd.rs:171-232) runs on the device, then the dist kernel and, if asked, the output-stage sort.
extern "C" int hg_dist_packed(hg_ctx *c, const uint8_t *ref_packed, uint64_t ref_stride, const uint8_t *ref_bits,
                              const int32_t *ref_norm, uint32_t n_ref, const uint8_t *qry_packed, uint64_t qry_stride,
                              const uint8_t *qry_bits, const int32_t *qry_norm, uint32_t n_qry, uint32_t hv_d, uint32_t ksize,
                              float ani_th, int symmetric, int path, int sorted, hg_hit *hits, uint32_t *ani_milli, uint64_t cap,
                              uint64_t *n_hits) {
  if (!c || !n_hits) { hg_set_error("hg_dist_packed: NULL argument"); return HG_E_INVALID; }
  *n_hits = 0;
  if ((n_ref && (!ref_packed || !ref_bits || !ref_norm)) || (n_qry && (!qry_packed || !qry_bits || !qry_norm)) || (cap && !hits)) {
    hg_set_error("hg_dist_packed: NULL argument"); return HG_E_INVALID;
  }
  if (hv_d == 0 || hv_d % 256 != 0 || ref_stride % 4 != 0 || qry_stride % 4 != 0) {
    hg_set_error("hg_dist_packed: hv_d %% 256 or row stride %% 4"); return HG_E_INVALID;
  }
  if (n_ref == 0 || n_qry == 0) return HG_OK;
  const bool same = (ref_packed == qry_packed && ref_bits == qry_bits && ref_norm == qry_norm && n_ref == n_qry);
  uint32_t rmax = 0, qmax = 0;
  for (uint32_t g = 0; g < n_ref; g++) {
    if (ref_bits[g] < 1 || ref_bits[g] > 16) { hg_set_error("hg_dist_packed: ref sketch %u has hv_quant_bits %u", g, (unsigned)ref_bits[g]); return HG_E_INVALID; }
    rmax = std::max<uint32_t>(rmax, ref_bits[g]);
  }
  for (uint32_t g = 0; g < n_qry && !same; g++) {
    if (qry_bits[g] < 1 || qry_bits[g] > 16) { hg_set_error("hg_dist_packed: query sketch %u has hv_quant_bits %u", g, (unsigned)qry_bits[g]); return HG_E_INVALID; }
    qmax = std::max<uint32_t>(qmax, qry_bits[g]);
  }
  const size_t rw = (size_t)rmax * hv_d / 8, qw = (size_t)qmax * hv_d / 8;  // bytes of a row that can be live
  if (rw > ref_stride || (!same && qw > qry_stride)) { hg_set_error("hg_dist_packed: row stride below hv_quant_bits * hv_d / 8"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  int rc;
  const size_t rb = (size_t)n_ref * hv_d * 2, qb = same ? 0 : (size_t)n_qry * hv_d * 2;
  const size_t rb_al = (rb + 255) & ~(size_t)255;
  const size_t rp = ((size_t)n_ref * rw + 255) & ~(size_t)255, qp = same ? 0 : (size_t)n_qry * qw;
  void *d_mat, *d_small, *d_pk;
  if ((rc = hg_scratch(c, HG_S_HV, rb + qb + 512, &d_mat))) return rc;
  if ((rc = hg_scratch(c, HG_S_SMALL, ((size_t)n_ref + n_qry) * 5 + 512, &d_small))) return rc;
  if ((rc = hg_scratch(c, HG_S_TABLES, rp + qp + 256, &d_pk))) return rc;
  int16_t *d_ref = (int16_t *)d_mat, *d_qry = same ? d_ref : (int16_t *)((uint8_t *)d_mat + rb_al);
  int32_t *d_rn = (int32_t *)d_small, *d_qn = same ? d_rn : d_rn + n_ref;
  uint8_t *d_rbits = (uint8_t *)d_small + ((size_t)n_ref + n_qry) * 4, *d_qbits = same ? d_rbits : d_rbits + n_ref;
  uint8_t *d_rp = (uint8_t *)d_pk, *d_qp = d_rp + rp;
  HostRows R = {ref_packed, ref_stride, rw, ref_bits, ref_norm, n_ref, d_rp, d_rbits, d_ref, d_rn};
  HostRows Q = {qry_packed, qry_stride, qw, qry_bits, qry_norm, n_qry, d_qp, d_qbits, d_qry, d_qn};
  // b-bit two's-complement values are below 2^(b-1): with b > 10 no row can fit the single s8 plane
  // (x = 2a + s spans 510), and with b <= 13 every element fits the int8 limb split, so that case needs no scan
  const uint32_t bmax = std::max(rmax, qmax);
  const bool wide = bmax > 10;
  if (path == 0 && wide && bmax <= 13 && (uint64_t)n_ref * n_qry >= 128ull * 128ull) {
    c->tc_is_hint = 1;  // the caller asked for auto: if the tensor kernel declines the shape (hv_d > 32768 ...) SIMT takes over
    rc = dist_from_host(c, R, same ? R : Q, same, hv_d, ksize, ani_th, symmetric, 2, 2, hits, cap, n_hits, sorted != 0, ani_milli);
    c->tc_is_hint = 0;
    if (c->dist_path == 2 && !strstr(c->dist_reason, "chunk"))
      snprintf(c->dist_reason, sizeof(c->dist_reason), "tensor: hv_quant_bits <= %u fits two s8 limbs; tcgen05 kind::i8", bmax);
    return rc;
  }
  return dist_from_host(c, R, same ? R : Q, same, hv_d, ksize, ani_th, symmetric, path, wide ? 0 : 1, hits, cap, n_hits, sorted != 0,
                        ani_milli);
}

extern "C" int hg_dist(hg_ctx *c, const int16_t *ref, const int32_t *ref_norm, uint32_t n_ref, const int16_t *qry,
                       const int32_t *qry_norm, uint32_t n_qry, uint32_t hv_d, uint32_t ksize, float ani_th,
                       int symmetric, int path, hg_hit *hits, uint64_t cap, uint64_t *n_hits) {
  return dist_host(c, ref, ref_norm, n_ref, qry, qry_norm, n_qry, hv_d, ksize, ani_th, symmetric, path, hits, cap, n_hits,
                   false, nullptr);
}

extern "C" int hg_dist_sorted(hg_ctx *c, const int16_t *ref, const int32_t *ref_norm, uint32_t n_ref, const int16_t *qry,
                              const int32_t *qry_norm, uint32_t n_qry, uint32_t hv_d, uint32_t ksize, float ani_th,
                              int symmetric, int path, hg_hit *hits, uint32_t *ani_milli, uint64_t cap, uint64_t *n_hits) {
  return dist_host(c, ref, ref_norm, n_ref, qry, qry_norm, n_qry, hv_d, ksize, ani_th, symmetric, path, hits, cap, n_hits,
                   true, ani_milli);
}

extern "C" int hg_sort_hits_dev(hg_ctx *c, hg_hit *d_hits, uint64_t n, uint32_t *d_ani_milli) {
  if (!c || (n && !d_hits)) { hg_set_error("hg_sort_hits_dev: NULL argument"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(c->device));
  return hg_launch_sort_hits(c, d_hits, n, d_ani_milli);
}
