// peer.cu — the GPUs of one box working on ONE dist call (ref rows sharded, query HVs
// broadcast, hit lists gathered).  The reference has nothing to replace here: dist is a single rayon
// loop over all pairs on the host (src/dist.rs:231-294).
//
// Every member GPU owns a "window" of device memory that all other members can address over NVLink
// (one process per GPU: cudaIpc handles; one process driving several GPUs: peer access).  A sharded
// dist then needs no collective library on its data path:
//   1. each member turns ITS rows into the operand form of the tensor kernel (one s8 plane + per-row
//      constants, or two s8 limb planes) inside its own window and PUSHES that slice, with plain
//      16-byte stores over NVLink, to the same offsets of every other window - the "all-gather" (or,
//      when one member holds all queries, the "broadcast") of what the kernel actually reads: 1 byte
//      per element instead of the 2 of the int16 rows, and no pre-pass replicated on every GPU;
//   2. a flag barrier (st.release.sys / ld.acquire.sys on words in the windows, bounded spin);
//   3. every member runs the dist kernel on its share of the output tiles - the non-empty tiles of the
//      all-vs-all are dealt round-robin (tile t to member t mod N), which balances the triangle to
//      within one tile; ref x query keeps the member's own ref rows - and appends its hits straight
//      into the ROOT's hit list over NVLink (warp-aggregated atomic on the root's counter);
//   4. a second flag barrier; the root's stream then owns the complete hit list.
// Measured on 2 B200s (tools/ipc_probe.cu): push of a 21 MB slice + both barriers 49 us (NCCL
// all_gather of the same plane 69 us), flag barrier 6.3 us.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "hg_common.cuh"

namespace {

constexpr size_t OFF_FLAGS = 0;      // u32[HG_MAX_PEERS]: flags[r] = last barrier epoch member r has reached
constexpr size_t OFF_STATUS = 64;    // u32: a barrier of this member timed out
constexpr size_t OFF_COUNT = 128;    // u64: hit counter (the root's is the one in use)
constexpr size_t OFF_STATS_Q = 256;  // u32[HG_MAX_PEERS][4]: pre-pass statistics of the gathered (query) matrix, one set per member
constexpr size_t OFF_STATS_R = 384;  // u32[HG_MAX_PEERS][4]: ... of the members' own ref rows (ref x query)
constexpr size_t OFF_HITS = 4096;    // hg_hit[cap]
constexpr int32_t TC_MAX_ABS = 8127; // |x| <= 8127 splits into two s8 limbs

struct PeerPtrs { uint8_t *p[HG_MAX_PEERS]; };

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One warp: lane t tells member t that this member has reached `epoch`, then waits until member t has.
// Everything this stream wrote before (its pushes) is complete when the kernel starts; the fence + release
// order it before the flag.  The spin is bounded: a member that never arrives (a failed launch, a dead
// process) turns into an error status instead of a hung GPU.
__global__ void peer_barrier_kernel(PeerPtrs w, int rank, int world, uint32_t epoch, unsigned long long timeout_ns) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();
  uint32_t *theirs = reinterpret_cast<uint32_t *>(w.p[t] + OFF_FLAGS) + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const uint32_t *mine = reinterpret_cast<const uint32_t *>(w.p[rank] + OFF_FLAGS) + t;
  const unsigned long long t0 = global_ns();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    if (global_ns() - t0 > timeout_ns) {
      atomicExch(reinterpret_cast<uint32_t *>(w.p[rank] + OFF_STATUS), 1u);
      break;
    }
  }
}

// Copies byte ranges of this member's window to the same offsets of every other window.
struct PushRange { uint64_t off, bytes; };  // 4-byte granular; ranges that are 16-byte aligned go as uint4
struct PushArgs { PushRange r[10]; int n; };

__global__ void peer_push_kernel(PeerPtrs w, int rank, int world, PushArgs a) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (int k = 0; k < a.n; ++k) {
    const uint64_t off = a.r[k].off, bytes = a.r[k].bytes;
    if (((off | bytes) & 15) == 0) {
      const uint4 *src = reinterpret_cast<const uint4 *>(w.p[rank] + off);
      for (size_t i = tid; i < bytes / 16; i += nth) {
        const uint4 v = src[i];
        for (int m = 0; m < world; ++m)
          if (m != rank) reinterpret_cast<uint4 *>(w.p[m] + off)[i] = v;
      }
    } else {
      const uint32_t *src = reinterpret_cast<const uint32_t *>(w.p[rank] + off);
      for (size_t i = tid; i < bytes / 4; i += nth) {
        const uint32_t v = src[i];
        for (int m = 0; m < world; ++m)
          if (m != rank) reinterpret_cast<uint32_t *>(w.p[m] + off)[i] = v;
      }
    }
  }
}

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// where the pieces of one sharded dist live inside every window (same call arguments -> same layout everywhere)
struct ShardLayout {
  uint64_t norm, arrays, entries, plane, end;
  uint32_t set_cap;
};
ShardLayout make_layout(int world, uint32_t n_total, uint32_t hv_d, uint64_t cap, int n_planes) {
  ShardLayout l;
  l.norm = align_up(OFF_HITS + cap * sizeof(hg_hit), 1024);
  l.arrays = align_up(l.norm + (uint64_t)n_total * 4, 256);
  l.set_cap = n_planes == 1 ? hg_narrow_set_cap(n_total) : 0;
  l.entries = align_up(l.arrays + (n_planes == 1 ? hg_narrow_arrays_bytes(n_total) : 0), 256);
  l.plane = align_up(l.entries + (uint64_t)world * l.set_cap * 4, 1024);
  l.end = l.plane + (uint64_t)n_planes * n_total * hv_d + 1024;
  return l;
}

}  // namespace

struct hg_peer {
  hg_ctx *ctx;
  int rank, world;
  int local;                     // members driven by one process (peer access) rather than one process per GPU (cudaIpc)
  uint64_t window_bytes;
  uint8_t *win[HG_MAX_PEERS];    // every member's window as THIS member addresses it (win[rank] = its own allocation)
  bool opened[HG_MAX_PEERS];
  bool connected;
  uint32_t epoch;
  unsigned long long timeout_ns;
  // the last sharded dist
  int root;
  uint64_t cap;
};

static PeerPtrs peer_ptrs(const hg_peer *p) {
  PeerPtrs w;
  for (int r = 0; r < HG_MAX_PEERS; ++r) w.p[r] = r < p->world ? p->win[r] : nullptr;
  return w;
}

extern "C" uint64_t hg_peer_window_need(uint32_t gathered_rows, uint32_t hv_d, uint64_t hit_cap) {
  const ShardLayout a = make_layout(HG_MAX_PEERS, gathered_rows, hv_d, hit_cap, 1), b = make_layout(HG_MAX_PEERS, gathered_rows, hv_d, hit_cap, 2);
  return std::max(a.end, b.end) + 4096;
}

static int peer_alloc(hg_ctx *ctx, int rank, int world, uint64_t window_bytes, hg_peer **out) {
  if (!ctx || !out) { hg_set_error("hg_peer_create: NULL argument"); return HG_E_INVALID; }
  if (world < 1 || world > HG_MAX_PEERS || rank < 0 || rank >= world) { hg_set_error("hg_peer_create: rank %d of %d (at most %d members)", rank, world, HG_MAX_PEERS); return HG_E_INVALID; }
  if (window_bytes < OFF_HITS + 4096) window_bytes = OFF_HITS + 4096;
  HG_CUDA(cudaSetDevice(ctx->device));
  hg_peer *p = new hg_peer();
  memset(p, 0, sizeof(*p));
  p->ctx = ctx;
  p->rank = rank;
  p->world = world;
  p->window_bytes = window_bytes;
  p->timeout_ns = 10ull * 1000000000ull;
  if (const char *e = getenv("HG_PEER_TIMEOUT_MS")) p->timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
  void *w = nullptr;
  cudaError_t e = cudaMalloc(&w, window_bytes);
  if (e != cudaSuccess) { delete p; return hg_cuda_fail(e, "cudaMalloc(window)", __FILE__, __LINE__); }
  e = cudaMemset(w, 0, OFF_HITS);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(w); delete p; return hg_cuda_fail(e, "cudaMemset(window)", __FILE__, __LINE__); }
  p->win[rank] = (uint8_t *)w;
  *out = p;
  return HG_OK;
}

// One process per GPU: allocate this member's window and return the handle the other members open.
extern "C" int hg_peer_create(hg_ctx *ctx, int rank, int world, uint64_t window_bytes, uint8_t handle_out[HG_IPC_HANDLE_BYTES],
                              hg_peer **out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == HG_IPC_HANDLE_BYTES, "IPC handle size");
  if (!handle_out) { hg_set_error("hg_peer_create: handle_out is NULL"); return HG_E_INVALID; }
  int rc = peer_alloc(ctx, rank, world, window_bytes, out);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, (*out)->win[rank]);
  if (e != cudaSuccess) { hg_peer_destroy(*out); *out = nullptr; return hg_cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__); }
  memcpy(handle_out, &h, HG_IPC_HANDLE_BYTES);
  if (world == 1) (*out)->connected = true;
  return HG_OK;
}

// `handles`: world x HG_IPC_HANDLE_BYTES in rank order (this member's own entry is ignored).  Every member must have
// created its window before anyone connects (the caller's bootstrap - an all-gather of the handles - implies that).
extern "C" int hg_peer_connect(hg_peer *p, const uint8_t *handles) {
  if (!p || !handles) { hg_set_error("hg_peer_connect: NULL argument"); return HG_E_INVALID; }
  if (p->local) { hg_set_error("hg_peer_connect: a single-process group is connected at creation"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(p->ctx->device));
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank || p->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * HG_IPC_HANDLE_BYTES, HG_IPC_HANDLE_BYTES);
    void *w = nullptr;
    HG_CUDA(cudaIpcOpenMemHandle(&w, h, cudaIpcMemLazyEnablePeerAccess));
    p->win[r] = (uint8_t *)w;
    p->opened[r] = true;
  }
  p->connected = true;
  return HG_OK;
}

// One process driving `world` GPUs: windows with mutual peer access.  out[0..world) receives the members.
extern "C" int hg_peer_create_local(hg_ctx *const *ctxs, int world, uint64_t window_bytes, hg_peer **out) {
  if (!ctxs || !out) { hg_set_error("hg_peer_create_local: NULL argument"); return HG_E_INVALID; }
  if (world < 1 || world > HG_MAX_PEERS) { hg_set_error("hg_peer_create_local: %d members (at most %d)", world, HG_MAX_PEERS); return HG_E_INVALID; }
  for (int r = 0; r < world; ++r) out[r] = nullptr;
  int rc = HG_OK;
  for (int r = 0; r < world && !rc; ++r) rc = peer_alloc(ctxs[r], r, world, window_bytes, &out[r]);
  for (int r = 0; r < world && !rc; ++r) {
    cudaSetDevice(ctxs[r]->device);
    for (int m = 0; m < world && !rc; ++m) {
      if (m == r || ctxs[m]->device == ctxs[r]->device) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[m]->device);
      if (!can) { hg_set_error("device %d cannot address device %d (no peer access)", ctxs[r]->device, ctxs[m]->device); rc = HG_E_UNSUPPORTED; break; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[m]->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) rc = hg_cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
    }
  }
  if (rc) {
    for (int r = 0; r < world; ++r) if (out[r]) { hg_peer_destroy(out[r]); out[r] = nullptr; }
    return rc;
  }
  for (int r = 0; r < world; ++r) {
    out[r]->local = 1;
    out[r]->connected = true;
    for (int m = 0; m < world; ++m) out[r]->win[m] = out[m]->win[m];
  }
  return HG_OK;
}

extern "C" void hg_peer_destroy(hg_peer *p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  for (int r = 0; r < p->world; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->win[r]);
  if (p->win[p->rank]) cudaFree(p->win[p->rank]);
  delete p;
}

extern "C" int hg_peer_rank(const hg_peer *p) { return p ? p->rank : -1; }
extern "C" int hg_peer_world(const hg_peer *p) { return p ? p->world : 0; }

static int peer_barrier(hg_peer *p) {
  peer_barrier_kernel<<<1, 32, 0, p->ctx->stream>>>(peer_ptrs(p), p->rank, p->world, ++p->epoch, p->timeout_ns);
  p->ctx->launches++;
  HG_CUDA(cudaGetLastError());
  return HG_OK;
}

extern "C" int hg_peer_barrier(hg_peer *p) {
  if (!p || !p->connected) { hg_set_error("hg_peer_barrier: not connected"); return HG_E_INVALID; }
  HG_CUDA(cudaSetDevice(p->ctx->device));
  return peer_barrier(p);
}

static int peer_push(hg_peer *p, const PushArgs &a) {
  if (p->world == 1 || a.n == 0) return HG_OK;
  uint64_t bytes = 0;
  for (int k = 0; k < a.n; ++k) bytes += a.r[k].bytes;
  const unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint64_t>(bytes / (16 * 256 * 4), 1), (uint64_t)p->ctx->sm_count * 4);
  peer_push_kernel<<<blocks, 256, 0, p->ctx->stream>>>(peer_ptrs(p), p->rank, p->world, a);
  p->ctx->launches++;
  HG_CUDA(cudaGetLastError());
  return HG_OK;
}

// ---------------------------------------------------------------------------------------------------
// sharded dist
// ---------------------------------------------------------------------------------------------------
namespace {
struct ShardCall {
  const int16_t *d_ref_hv; const int32_t *d_ref_norm; uint32_t n_ref_local, ref_row0;
  const int16_t *d_qry_hv; const int32_t *d_qry_norm; uint32_t n_qry_local, qry_row0, n_qry_total;
  uint32_t hv_d, ksize; float ani_th; int symmetric, root; uint64_t cap;
};
}  // namespace

static int shard_check(hg_peer *p, const ShardCall &a) {
  if (!p || !p->connected) { hg_set_error("hg_dist_sharded: the group is not connected"); return HG_E_INVALID; }
  if (a.hv_d == 0 || a.hv_d % 256 != 0) { hg_set_error("hg_dist_sharded: hv_d %u must be a multiple of 256", a.hv_d); return HG_E_INVALID; }
  if (a.root < 0 || a.root >= p->world) { hg_set_error("hg_dist_sharded: root %d of %d", a.root, p->world); return HG_E_INVALID; }
  if ((uint64_t)a.qry_row0 + a.n_qry_local > a.n_qry_total) { hg_set_error("hg_dist_sharded: query rows [%u, +%u) outside %u", a.qry_row0, a.n_qry_local, a.n_qry_total); return HG_E_INVALID; }
  if ((a.n_qry_local && (!a.d_qry_hv || !a.d_qry_norm)) || (!a.symmetric && a.n_ref_local && (!a.d_ref_hv || !a.d_ref_norm))) {
    hg_set_error("hg_dist_sharded: NULL argument"); return HG_E_INVALID;
  }
  if (((uintptr_t)a.d_qry_hv | (uintptr_t)a.d_ref_hv) & 15) { hg_set_error("hg_dist_sharded: HV rows must be 16-byte aligned"); return HG_E_INVALID; }
  const uint64_t need = std::max(make_layout(p->world, a.n_qry_total, a.hv_d, a.cap, 1).end, make_layout(p->world, a.n_qry_total, a.hv_d, a.cap, 2).end);
  if (need > p->window_bytes) {
    hg_set_error("hg_dist_sharded: the windows hold %llu bytes, this call needs %llu (hg_peer_window_need)", (unsigned long long)p->window_bytes,
                 (unsigned long long)need);
    return HG_E_CAPACITY;
  }
  return HG_OK;
}

// every allocation of the call, before anything that waits for another member is enqueued (growing a scratch
// buffer synchronises the device)
static int shard_reserve(hg_peer *p, const ShardCall &a) {
  hg_ctx *c = p->ctx;
  HG_CUDA(cudaSetDevice(c->device));
  if (!a.symmetric && a.n_ref_local) {
    void *x;
    int rc;
    if ((rc = hg_scratch(c, HG_S_REF_LIMBS, 2 * (uint64_t)a.n_ref_local * a.hv_d + 1024, &x))) return rc;
    if ((rc = hg_scratch(c, HG_S_NARROW_META, hg_narrow_arrays_bytes(a.n_ref_local) + (uint64_t)hg_narrow_set_cap(a.n_ref_local) * 4 + 256, &x))) return rc;
  }
  return HG_OK;
}

// rows whose first element is `d_rows` are rows [row0, ...) of a matrix: the matrix's (virtual) first element
static const int16_t *matrix_base(const int16_t *d_rows, uint32_t row0, uint32_t hv_d) {
  return (const int16_t *)((uintptr_t)d_rows - (uintptr_t)row0 * hv_d * 2);
}

// use_path 3: single s8 plane; 2: two s8 limb planes.  Enqueues everything on the member's stream; returns without waiting.
static int shard_enqueue(hg_peer *p, const ShardCall &a, int use_path, hg_narrow_mat *Qn_out, hg_narrow_mat *Rn_out) {
  hg_ctx *c = p->ctx;
  HG_CUDA(cudaSetDevice(c->device));
  int rc;
  const int rank = p->rank, world = p->world;
  uint8_t *W = p->win[rank];
  const ShardLayout lay = make_layout(world, a.n_qry_total, a.hv_d, a.cap, use_path == 3 ? 1 : 2);
  int32_t *normW = (int32_t *)(W + lay.norm);
  hg_hit *hits = (hg_hit *)(p->win[a.root] + OFF_HITS);
  unsigned long long *counter = (unsigned long long *)(p->win[a.root] + OFF_COUNT);
  p->root = a.root;
  p->cap = a.cap;
  c->ev_used &= ~(3 << 4);
  HG_PROF(c, 4);
  if (rank == a.root) HG_CUDA(cudaMemsetAsync(W + OFF_COUNT, 0, 8, c->stream));
  if (a.n_qry_local)
    HG_CUDA(cudaMemcpyAsync(normW + a.qry_row0, a.d_qry_norm, (size_t)a.n_qry_local * 4, cudaMemcpyDeviceToDevice, c->stream));
  PushArgs pa;
  pa.n = 0;
  auto add = [&](uint64_t off, uint64_t bytes) { if (bytes) { pa.r[pa.n].off = off; pa.r[pa.n].bytes = bytes; pa.n++; } };
  const uint64_t r0 = a.qry_row0, nl = a.n_qry_local;
  add(lay.norm + r0 * 4, nl * 4);
  if (use_path == 3) {
    hg_narrow_mat Q, R;
    if ((rc = hg_narrow_attach(c, matrix_base(a.d_qry_hv, a.qry_row0, a.hv_d), a.n_qry_total, a.hv_d, (int8_t *)(W + lay.plane), W + lay.arrays,
                               (uint32_t *)(W + lay.entries), (uint32_t *)(W + OFF_STATS_Q), (uint32_t)world, (uint32_t)rank,
                               (uint32_t)rank * lay.set_cap, lay.set_cap, &Q)))
      return rc;
    if ((rc = hg_narrow_prep_rows(c, &Q, a.qry_row0, a.n_qry_local))) return rc;
    if (!a.symmetric) {
      void *plane = c->d_scratch[HG_S_REF_LIMBS], *meta = c->d_scratch[HG_S_NARROW_META];
      const uint32_t nr = a.n_ref_local;
      if ((rc = hg_narrow_attach(c, a.d_ref_hv, nr, a.hv_d, (int8_t *)plane, meta, (uint32_t *)((uint8_t *)meta + hg_narrow_arrays_bytes(nr)),
                                 (uint32_t *)(W + OFF_STATS_R), (uint32_t)world, (uint32_t)rank, 0, hg_narrow_set_cap(nr), &R)))
        return rc;
      if ((rc = hg_narrow_prep_rows(c, &R, 0, nr))) return rc;
    }
    add(lay.plane + r0 * a.hv_d, nl * a.hv_d);
    for (int k = 0; k < 5; ++k) add(lay.arrays + ((uint64_t)k * a.n_qry_total + r0) * 4, nl * 4);
    if (nl) add(lay.entries + (uint64_t)rank * lay.set_cap * 4, (uint64_t)lay.set_cap * 4);
    add(OFF_STATS_Q + 16 * (uint64_t)rank, 16);
    if (!a.symmetric) add(OFF_STATS_R + 16 * (uint64_t)rank, 16);
    if ((rc = peer_push(p, pa))) return rc;
    if ((rc = peer_barrier(p))) return rc;
    if (a.symmetric)
      rc = hg_narrow_launch_ex(c, &Q, 0, a.n_qry_total, 0, normW, &Q, 0, a.n_qry_total, 0, normW, a.ksize, a.ani_th, 1, hits, a.cap, counter,
                               (uint32_t)world, (uint32_t)rank);
    else
      rc = hg_narrow_launch_ex(c, &R, 0, a.n_ref_local, a.ref_row0, a.d_ref_norm, &Q, 0, a.n_qry_total, 0, normW, a.ksize, a.ani_th, 0, hits,
                               a.cap, counter, 1, 0);
    if (rc) return rc;
    if (Qn_out) *Qn_out = Q;
    if (Rn_out && !a.symmetric) *Rn_out = R;
  } else {
    hg_tc_mat Q, R;
    if ((rc = hg_tc_shape_ok(a.hv_d, a.d_qry_hv, a.d_ref_hv))) return rc;
    if ((rc = hg_tc_attach(c, matrix_base(a.d_qry_hv, a.qry_row0, a.hv_d), a.n_qry_total, a.hv_d, (int8_t *)(W + lay.plane), &Q))) return rc;
    if ((rc = hg_tc_split_rows(c, &Q, a.qry_row0, a.n_qry_local))) return rc;
    if (!a.symmetric && a.n_ref_local) {
      if ((rc = hg_tc_attach(c, a.d_ref_hv, a.n_ref_local, a.hv_d, (int8_t *)c->d_scratch[HG_S_REF_LIMBS], &R))) return rc;
      if ((rc = hg_tc_split_rows(c, &R, 0, a.n_ref_local))) return rc;
    }
    add(lay.plane + r0 * a.hv_d, nl * a.hv_d);
    add(lay.plane + ((uint64_t)a.n_qry_total + r0) * a.hv_d, nl * a.hv_d);
    if ((rc = peer_push(p, pa))) return rc;
    if ((rc = peer_barrier(p))) return rc;
    if (a.symmetric)
      rc = hg_tc_launch_ex(c, &Q, 0, a.n_qry_total, 0, normW, &Q, 0, a.n_qry_total, 0, normW, a.ksize, a.ani_th, 1, hits, a.cap, counter,
                           (uint32_t)world, (uint32_t)rank);
    else
      rc = hg_tc_launch_ex(c, &R, 0, a.n_ref_local, a.ref_row0, a.d_ref_norm, &Q, 0, a.n_qry_total, 0, normW, a.ksize, a.ani_th, 0, hits, a.cap,
                           counter, 1, 0);
    if (rc) return rc;
  }
  HG_PROF(c, 5);
  return peer_barrier(p);
}

static int peer_status(hg_peer *p) {  // after a stream synchronisation
  uint32_t st = 0;
  HG_CUDA(cudaMemcpy(&st, p->win[p->rank] + OFF_STATUS, 4, cudaMemcpyDeviceToHost));
  if (st) {
    cudaMemset(p->win[p->rank] + OFF_STATUS, 0, 4);
    hg_set_error("a member of the GPU group did not reach a barrier within %llu ms", p->timeout_ns / 1000000ull);
    return HG_E_CUDA;
  }
  return HG_OK;
}

// verdict of the single-plane pre-passes of ALL members (their statistics were pushed into every window, so every
// member reads the same answer): HG_OK narrow, HG_E_UNSUPPORTED not narrow (*absmax = max |hv| over all rows)
static int shard_verdict(hg_peer *p, const ShardCall &a, int32_t *absmax) {
  hg_ctx *c = p->ctx;
  uint32_t sq[4 * HG_MAX_PEERS], sr[4 * HG_MAX_PEERS];
  HG_CUDA(cudaMemcpyAsync(sq, p->win[p->rank] + OFF_STATS_Q, sizeof(sq), cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaMemcpyAsync(sr, p->win[p->rank] + OFF_STATS_R, sizeof(sr), cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  int rc = peer_status(p);
  if (rc) return rc;
  uint32_t declined = 0, amax = 0;
  for (int r = 0; r < p->world; ++r) {
    declined |= sq[4 * r + 3];
    amax = std::max(amax, sq[4 * r]);
    if (!a.symmetric) { declined |= sr[4 * r + 3]; amax = std::max(amax, sr[4 * r]); }
  }
  *absmax = (int32_t)amax;
  return declined ? HG_E_UNSUPPORTED : HG_OK;
}

static void shard_reason(hg_peer *p, int path, int32_t absmax, bool forced) {
  hg_ctx *c = p->ctx;
  c->dist_path = path;
  if (path == 3)
    snprintf(c->dist_reason, sizeof(c->dist_reason), "tensor-narrow%s: member %d of %d GPUs, operand planes pushed over NVLink windows, tiles dealt round-robin%s",
             forced ? " (forced)" : "", p->rank, p->world, forced ? "" : "; rows fit one s8 plane as x = 2a + s");
  else
    snprintf(c->dist_reason, sizeof(c->dist_reason), "tensor%s: member %d of %d GPUs, two s8 limb planes pushed over NVLink windows, tiles dealt round-robin (max |hv| = %d)",
             forced ? " (forced)" : "", p->rank, p->world, absmax);
}

// Collective over the group (every member calls it with the same scalars and its own rows):
//   symmetric != 0: all-vs-all (j > i) over the n_qry_total-row matrix whose rows [qry_row0, +n_qry_local) this member
//                   holds (d_qry_hv, d_qry_norm2); the ref arguments are ignored;
//   symmetric == 0: this member's n_ref_local ref rows (global index ref_row0 + local row) against ALL n_qry_total query
//                   rows, of which this member contributes [qry_row0, +n_qry_local) - possibly none, possibly all.
//   path: 0 auto (one host read of the members' pre-pass verdict, then the two-limb kernel if the rows are not narrow),
//         3 / 2 force the single-plane / two-limb tensor kernel: nothing is waited for, the call only enqueues.
// The hits (global indices) accumulate in member `root`'s window: hg_dist_sharded_hits reads them.
extern "C" int hg_dist_sharded_dev(hg_peer *p, const int16_t *d_ref_hv, const int32_t *d_ref_norm2, uint32_t n_ref_local,
                                   uint32_t ref_row0, const int16_t *d_qry_hv, const int32_t *d_qry_norm2, uint32_t n_qry_local,
                                   uint32_t qry_row0, uint32_t n_qry_total, uint32_t hv_d, uint32_t ksize, float ani_th,
                                   int symmetric, int path, int root, uint64_t cap) {
  ShardCall a = {d_ref_hv, d_ref_norm2, symmetric ? 0u : n_ref_local, ref_row0, d_qry_hv, d_qry_norm2, n_qry_local, qry_row0, n_qry_total,
                 hv_d, ksize, ani_th, symmetric, root, cap};
  int rc;
  if ((rc = shard_check(p, a))) return rc;
  if (path != 0 && path != 2 && path != 3) { hg_set_error("hg_dist_sharded_dev: path %d (0 auto, 2 two-limb, 3 single plane)", path); return HG_E_INVALID; }
  if (n_qry_total == 0) return HG_OK;
  if ((rc = shard_reserve(p, a))) return rc;
  if (path == 2 || path == 3) {
    if (path == 3 && (rc = hg_narrow_shape_ok(hv_d, d_qry_hv, d_ref_hv))) return rc;
    if ((rc = shard_enqueue(p, a, path, nullptr, nullptr))) return rc;
    shard_reason(p, path, -1, true);
    return HG_OK;
  }
  int32_t absmax = -1;
  rc = hg_narrow_shape_ok(hv_d, d_qry_hv, d_ref_hv);
  if (rc == HG_OK) {
    if ((rc = shard_enqueue(p, a, 3, nullptr, nullptr))) return rc;
    rc = shard_verdict(p, a, &absmax);
    if (rc == HG_OK) { shard_reason(p, 3, absmax, false); return HG_OK; }
    if (rc != HG_E_UNSUPPORTED) return rc;
  }
  if (absmax > TC_MAX_ABS) {
    hg_set_error("hg_dist_sharded_dev: max |hv| = %d exceeds the 13-bit budget of the int8 limb split; the sharded path has no SIMT kernel", absmax);
    return HG_E_UNSUPPORTED;
  }
  if ((rc = shard_enqueue(p, a, 2, nullptr, nullptr))) return rc;
  shard_reason(p, 2, absmax, false);
  return HG_OK;
}

// After hg_dist_sharded_dev: waits for this member's stream.  On the root: the number of hits (HG_E_CAPACITY with the
// need if it exceeds cap), the hits - in dump_ani_file's order if `sorted` (ani_milli as hg_dist_sorted) - copied to
// HOST memory.  On the other members: n_hits = 0 (hits may be NULL).
extern "C" int hg_dist_sharded_hits(hg_peer *p, int sorted, hg_hit *hits, uint32_t *ani_milli, uint64_t cap, uint64_t *n_hits) {
  if (!p || !n_hits) { hg_set_error("hg_dist_sharded_hits: NULL argument"); return HG_E_INVALID; }
  hg_ctx *c = p->ctx;
  *n_hits = 0;
  HG_CUDA(cudaSetDevice(c->device));
  int rc;
  unsigned long long cnt = 0;
  if (p->rank == p->root) HG_CUDA(cudaMemcpyAsync(&cnt, p->win[p->rank] + OFF_COUNT, 8, cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  if ((rc = peer_status(p))) return rc;
  if (p->rank != p->root) return HG_OK;
  *n_hits = cnt;
  if (cnt > cap || cnt > p->cap) {
    hg_set_error("hg_dist_sharded: %llu pairs pass the threshold, capacity is %llu", cnt, (unsigned long long)std::min<uint64_t>(cap, p->cap));
    return HG_E_CAPACITY;
  }
  if (cnt == 0) return HG_OK;
  if (!hits) { hg_set_error("hg_dist_sharded_hits: hits is NULL"); return HG_E_INVALID; }
  hg_hit *d_hits = (hg_hit *)(p->win[p->rank] + OFF_HITS);
  void *d_milli = nullptr;
  if (sorted) {
    if (ani_milli && (rc = hg_scratch(c, HG_S_MISC, cnt * 4 + 256, &d_milli))) return rc;
    if ((rc = hg_launch_sort_hits(c, d_hits, cnt, (uint32_t *)d_milli))) return rc;
  }
  HG_CUDA(cudaMemcpyAsync(hits, d_hits, cnt * sizeof(hg_hit), cudaMemcpyDeviceToHost, c->stream));
  if (d_milli) HG_CUDA(cudaMemcpyAsync(ani_milli, d_milli, cnt * 4, cudaMemcpyDeviceToHost, c->stream));
  HG_CUDA(cudaStreamSynchronize(c->stream));
  return HG_OK;
}

// device pointers of the root's hit list and counter as this member addresses them (benchmarks, callers with their own tail)
extern "C" int hg_peer_hit_buffers(hg_peer *p, int root, hg_hit **d_hits, unsigned long long **d_count) {
  if (!p || root < 0 || root >= p->world || !p->connected) { hg_set_error("hg_peer_hit_buffers: bad argument"); return HG_E_INVALID; }
  if (d_hits) *d_hits = (hg_hit *)(p->win[root] + OFF_HITS);
  if (d_count) *d_count = (unsigned long long *)(p->win[root] + OFF_COUNT);
  return HG_OK;
}

// ---------------------------------------------------------------------------------------------------
// one process, all the GPUs of the box: what `hyper-gen sketch -D gpu` / `hyper-gen dist` call
// (the reference is single-GPU: CudaDevice::new(0), src/sketch_cuda.rs:52)
// ---------------------------------------------------------------------------------------------------
struct hg_group {
  int n;
  hg_ctx *ctx[HG_MAX_PEERS];
  hg_peer *peer[HG_MAX_PEERS];
  uint64_t window_bytes;
};

extern "C" int hg_group_create(int n_devices, const int *ordinals, hg_group **out) {
  if (!out) { hg_set_error("hg_group_create: out is NULL"); return HG_E_INVALID; }
  *out = nullptr;
  int avail = 0;
  cudaError_t e = cudaGetDeviceCount(&avail);
  if (e != cudaSuccess || avail == 0) {
    hg_set_error("hg_group_create: no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return HG_E_CUDA;
  }
  if (n_devices <= 0) n_devices = std::min(avail, HG_MAX_PEERS);  // all visible GPUs
  if (n_devices > HG_MAX_PEERS) { hg_set_error("hg_group_create: at most %d devices", HG_MAX_PEERS); return HG_E_INVALID; }
  hg_group *g = new hg_group();
  memset(g, 0, sizeof(*g));
  g->n = n_devices;
  for (int i = 0; i < n_devices; ++i) {
    const int rc = hg_init(ordinals ? ordinals[i] : i, &g->ctx[i]);
    if (rc) { hg_group_destroy(g); return rc; }
  }
  *out = g;
  return HG_OK;
}

extern "C" void hg_group_destroy(hg_group *g) {
  if (!g) return;
  for (int i = 0; i < g->n; ++i) if (g->peer[i]) hg_peer_destroy(g->peer[i]);
  for (int i = 0; i < g->n; ++i) if (g->ctx[i]) hg_destroy(g->ctx[i]);
  delete g;
}

extern "C" int hg_group_size(const hg_group *g) { return g ? g->n : 0; }
extern "C" hg_ctx *hg_group_ctx(hg_group *g, int i) { return g && i >= 0 && i < g->n ? g->ctx[i] : nullptr; }

static int group_windows(hg_group *g, uint64_t need) {
  if (g->peer[0] && g->window_bytes >= need) return HG_OK;
  for (int i = 0; i < g->n; ++i) if (g->peer[i]) { hg_peer_destroy(g->peer[i]); g->peer[i] = nullptr; }
  need += need / 8;
  int rc = hg_peer_create_local(g->ctx, g->n, need, g->peer);
  if (rc) return rc;
  g->window_bytes = need;
  return HG_OK;
}

// contiguous split of n items with weights w[i] = off[i + 1] - off[i] into `parts` runs of about equal weight
static std::vector<uint32_t> split_by_weight(const uint64_t *off, uint32_t n, int parts) {
  std::vector<uint32_t> b(parts + 1, n);
  b[0] = 0;
  const uint64_t total = off[n] - off[0];
  uint32_t g = 0;
  for (int k = 1; k < parts; ++k) {
    const uint64_t target = off[0] + total / parts * k + (total % parts) * k / parts;
    while (g < n && off[g + 1] - (off[g + 1] - off[g]) / 2 <= target) ++g;  // an item goes where its midpoint falls
    b[k] = g;
  }
  return b;
}

// hg_sketch_fasta_batch over all GPUs of the group: the files are split into one contiguous run per GPU of about
// equal bytes (genomes are independent - no collective); one host thread per GPU drives its pipeline.
extern "C" int hg_group_sketch_fasta_batch(hg_group *g, const uint8_t *raw, const uint64_t *file_off, uint32_t n_files,
                                           const hg_sketch_params *p, int16_t *hv, uint8_t *packed, uint8_t *quant_bits,
                                           int32_t *norm2, uint32_t *n_hashes) {
  if (!g || !file_off || !p) { hg_set_error("hg_group_sketch_fasta_batch: NULL argument"); return HG_E_INVALID; }
  if (n_files == 0) return HG_OK;
  for (uint32_t f = 0; f < n_files; ++f)
    if (file_off[f + 1] < file_off[f]) { hg_set_error("file_off not monotone at %u", f); return HG_E_INVALID; }
  const int parts = (int)std::min<uint32_t>((uint32_t)g->n, n_files);
  const std::vector<uint32_t> b = split_by_weight(file_off, n_files, parts);
  std::vector<int> rcs(parts, HG_OK);
  std::vector<std::string> msgs(parts);
  std::vector<std::thread> th;
  const uint32_t D = p->hv_d;
  for (int k = 0; k < parts; ++k) {
    th.emplace_back([&, k]() {
      const uint32_t f0 = b[k], m = b[k + 1] - b[k];
      if (m == 0) return;
      rcs[k] = hg_sketch_fasta_batch(g->ctx[k], raw, file_off + f0, m, p, hv ? hv + (size_t)f0 * D : nullptr,
                                     packed ? packed + (size_t)f0 * 2 * D : nullptr, quant_bits ? quant_bits + f0 : nullptr,
                                     norm2 ? norm2 + f0 : nullptr, n_hashes ? n_hashes + f0 : nullptr);
      if (rcs[k]) msgs[k] = hg_last_error();
    });
  }
  for (auto &t : th) t.join();
  for (int k = 0; k < parts; ++k)
    if (rcs[k]) { hg_set_error("GPU %d (files %u..%u): %s", g->ctx[k]->device, b[k], b[k + 1], msgs[k].c_str()); return rcs[k]; }
  return HG_OK;
}

// hg_dist_packed over all GPUs of the group.  Every GPU receives a contiguous block of the packed ref rows and of the
// packed query rows over its own PCIe link, unpacks them, and the sharded dist above does the rest; the hits come back
// from GPU 0 (sorted there if asked).  Same arguments and capacity contract as hg_dist_packed.
extern "C" int hg_group_dist_packed(hg_group *g, const uint8_t *ref_packed, uint64_t ref_stride, const uint8_t *ref_bits,
                                    const int32_t *ref_norm, uint32_t n_ref, const uint8_t *qry_packed, uint64_t qry_stride,
                                    const uint8_t *qry_bits, const int32_t *qry_norm, uint32_t n_qry, uint32_t hv_d,
                                    uint32_t ksize, float ani_th, int symmetric, int sorted, hg_hit *hits, uint32_t *ani_milli,
                                    uint64_t cap, uint64_t *n_hits) {
  if (!g || !n_hits) { hg_set_error("hg_group_dist_packed: NULL argument"); return HG_E_INVALID; }
  *n_hits = 0;
  if ((n_ref && (!ref_packed || !ref_bits || !ref_norm)) || (n_qry && (!qry_packed || !qry_bits || !qry_norm)) || (cap && !hits)) {
    hg_set_error("hg_group_dist_packed: NULL argument"); return HG_E_INVALID;
  }
  if (hv_d == 0 || hv_d % 256 != 0 || ref_stride % 4 != 0 || qry_stride % 4 != 0) {
    hg_set_error("hg_group_dist_packed: hv_d %% 256 or row stride %% 4"); return HG_E_INVALID;
  }
  if (n_ref == 0 || n_qry == 0) return HG_OK;
  const bool same = (ref_packed == qry_packed && ref_bits == qry_bits && ref_norm == qry_norm && n_ref == n_qry);
  // a single GPU, too little work to share, or a shape only the single-GPU entry handles (one matrix against itself
  // with both orders, the j > i filter over two different matrices): the single-GPU entry
  if (g->n == 1 || (uint64_t)n_ref * n_qry < 1024ull * 1024ull || hv_d % 512 != 0 || (same && !symmetric) || (!same && symmetric))
    return hg_dist_packed(g->ctx[0], ref_packed, ref_stride, ref_bits, ref_norm, n_ref, qry_packed, qry_stride, qry_bits, qry_norm, n_qry,
                          hv_d, ksize, ani_th, symmetric, 0, sorted, hits, ani_milli, cap, n_hits);
  uint32_t rmax = 0, qmax = 0;
  for (uint32_t i = 0; i < n_ref; ++i) {
    if (ref_bits[i] < 1 || ref_bits[i] > 16) { hg_set_error("hg_group_dist_packed: ref sketch %u has hv_quant_bits %u", i, (unsigned)ref_bits[i]); return HG_E_INVALID; }
    rmax = std::max<uint32_t>(rmax, ref_bits[i]);
  }
  for (uint32_t i = 0; i < n_qry && !same; ++i) {
    if (qry_bits[i] < 1 || qry_bits[i] > 16) { hg_set_error("hg_group_dist_packed: query sketch %u has hv_quant_bits %u", i, (unsigned)qry_bits[i]); return HG_E_INVALID; }
    qmax = std::max<uint32_t>(qmax, qry_bits[i]);
  }
  if (same) qmax = rmax;
  const uint32_t bmax = std::max(rmax, qmax);
  if (bmax > 13) {  // beyond the limb split: the exact SIMT kernel of the single-GPU entry
    return hg_dist_packed(g->ctx[0], ref_packed, ref_stride, ref_bits, ref_norm, n_ref, qry_packed, qry_stride, qry_bits, qry_norm, n_qry,
                          hv_d, ksize, ani_th, symmetric, 0, sorted, hits, ani_milli, cap, n_hits);
  }
  const size_t rw = (size_t)rmax * hv_d / 8, qw = (size_t)qmax * hv_d / 8;
  if (rw > ref_stride || qw > qry_stride) { hg_set_error("hg_group_dist_packed: row stride below hv_quant_bits * hv_d / 8"); return HG_E_INVALID; }
  int rc;
  const int N = g->n;
  if ((rc = group_windows(g, hg_peer_window_need(n_qry, hv_d, cap)))) return rc;
  auto bound = [&](uint32_t n, int k) -> uint32_t { return k >= N ? n : (uint32_t)(((uint64_t)n * k / N) & ~3ull); };
  struct Member { uint32_t r0, rn, q0, qn; int16_t *d_ref, *d_qry; int32_t *d_rn, *d_qn; };
  std::vector<Member> M(N);
  // phase 1: every allocation, on every GPU
  for (int k = 0; k < N; ++k) {
    hg_ctx *c = g->ctx[k];
    Member &m = M[k];
    m.q0 = bound(n_qry, k); m.qn = bound(n_qry, k + 1) - m.q0;
    m.r0 = same ? 0 : bound(n_ref, k); m.rn = same ? 0 : bound(n_ref, k + 1) - m.r0;
    HG_CUDA(cudaSetDevice(c->device));
    void *d_mat, *d_small, *d_pk;
    const size_t qb = ((size_t)m.qn * hv_d * 2 + 255) & ~(size_t)255, rb = (size_t)m.rn * hv_d * 2;
    if ((rc = hg_scratch(c, HG_S_HV, qb + rb + 512, &d_mat))) return rc;
    if ((rc = hg_scratch(c, HG_S_SMALL, ((size_t)m.qn + m.rn) * 5 + 512, &d_small))) return rc;
    if ((rc = hg_scratch(c, HG_S_TABLES, (size_t)m.qn * qw + (size_t)m.rn * rw + 512, &d_pk))) return rc;
    m.d_qry = (int16_t *)d_mat;
    m.d_ref = (int16_t *)((uint8_t *)d_mat + qb);
    m.d_qn = (int32_t *)d_small;
    m.d_rn = m.d_qn + m.qn;
    ShardCall a = {m.d_ref, m.d_rn, m.rn, m.r0, m.d_qry, m.d_qn, m.qn, m.q0, n_qry, hv_d, ksize, ani_th, symmetric, 0, cap};
    if ((rc = shard_check(g->peer[k], a))) return rc;
    if ((rc = shard_reserve(g->peer[k], a))) return rc;
  }
  // phase 2: H2D + unpack + sharded dist enqueued on every GPU; nothing here waits for another GPU
  auto enqueue_all = [&](int use_path) -> int {
    for (int k = 0; k < N; ++k) {
      hg_ctx *c = g->ctx[k];
      Member &m = M[k];
      HG_CUDA(cudaSetDevice(c->device));
      ShardCall a = {m.d_ref, m.d_rn, m.rn, m.r0, m.d_qry, m.d_qn, m.qn, m.q0, n_qry, hv_d, ksize, ani_th, symmetric, 0, cap};
      int r2 = shard_enqueue(g->peer[k], a, use_path, nullptr, nullptr);
      if (r2) return r2;
    }
    return HG_OK;
  };
  for (int k = 0; k < N; ++k) {
    hg_ctx *c = g->ctx[k];
    Member &m = M[k];
    HG_CUDA(cudaSetDevice(c->device));
    uint8_t *d_pk = (uint8_t *)c->d_scratch[HG_S_TABLES];
    uint8_t *d_bits = (uint8_t *)c->d_scratch[HG_S_SMALL] + ((size_t)m.qn + m.rn) * 4;
    auto load = [&](const uint8_t *packed, uint64_t stride, size_t w, const uint8_t *bits, const int32_t *norm, uint32_t r0, uint32_t n,
                    uint8_t *d_stage, uint8_t *d_b, int32_t *d_n, int16_t *d_hv) -> int {
      if (n == 0) return HG_OK;
      HG_CUDA(cudaMemcpy2DAsync(d_stage, w, packed + (size_t)r0 * stride, stride, w, n, cudaMemcpyHostToDevice, c->stream));
      HG_CUDA(cudaMemcpyAsync(d_b, bits + r0, n, cudaMemcpyHostToDevice, c->stream));
      HG_CUDA(cudaMemcpyAsync(d_n, norm + r0, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
      return hg_launch_unpack(c, d_stage, w, d_b, n, hv_d, d_hv);
    };
    if ((rc = load(qry_packed, qry_stride, qw, qry_bits, qry_norm, m.q0, m.qn, d_pk, d_bits, m.d_qn, m.d_qry))) return rc;
    if ((rc = load(ref_packed, ref_stride, rw, ref_bits, ref_norm, m.r0, m.rn, d_pk + (size_t)m.qn * qw, d_bits + m.qn, m.d_rn, m.d_ref))) return rc;
  }
  // b-bit values are below 2^(b-1): above 10 bits no row fits the single plane, and up to 13 bits every element fits two limbs
  int used = bmax > 10 ? 2 : 3;
  int32_t absmax = -1;
  if ((rc = enqueue_all(used))) return rc;
  if (used == 3) {
    int verdict = HG_OK;
    for (int k = 0; k < N; ++k) {
      ShardCall a = {M[k].d_ref, M[k].d_rn, M[k].rn, M[k].r0, M[k].d_qry, M[k].d_qn, M[k].qn, M[k].q0, n_qry, hv_d, ksize, ani_th, symmetric, 0, cap};
      cudaSetDevice(g->ctx[k]->device);
      const int v = shard_verdict(g->peer[k], a, &absmax);
      if (v != HG_OK && v != HG_E_UNSUPPORTED) return v;
      if (v) verdict = v;
    }
    if (verdict == HG_E_UNSUPPORTED) {  // not narrow (the same answer on every GPU): the two-limb kernel on the rows already in HBM
      used = 2;
      if ((rc = enqueue_all(2))) return rc;
    }
  }
  for (int k = 0; k < N; ++k) shard_reason(g->peer[k], used, absmax, false);
  for (int k = N - 1; k >= 0; --k) {
    uint64_t nh = 0;
    rc = hg_dist_sharded_hits(g->peer[k], sorted, k == 0 ? hits : nullptr, k == 0 ? ani_milli : nullptr, cap, &nh);
    if (k == 0) *n_hits = nh;
    if (rc) return rc;
  }
  return HG_OK;
}
