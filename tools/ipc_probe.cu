// ipc_probe.cu — what the peer-window design of the multi-GPU dist relies on, measured on the box:
//   one process per GPU (fork), cudaIpc windows, pushes over NVLink (16-byte stores), flag barrier
//   kernels (st.release.sys / ld.acquire.sys, bounded spin), remote warp-aggregated atomic append,
//   and how many CTAs a cluster of 2 / 4 keeps active.
//   nvcc -arch=sm_100a -O3 -o build/ipc_probe tools/ipc_probe.cu && build/ipc_probe <n_ranks>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      fprintf(stderr, "[rank %d] %s failed: %s (line %d)\n", g_rank, #x, cudaGetErrorString(e_), __LINE__); \
      _exit(2);                                                                          \
    }                                                                                    \
  } while (0)

static int g_rank = -1;
constexpr int MAXR = 8;
struct Shared {
  std::atomic<int> arrive[64];
  cudaIpcMemHandle_t handle[MAXR];
  double result[MAXR][8];
};
static Shared *g_sh;
static int g_n;
static int g_bar_idx = 0;
static void host_barrier() {
  const int i = g_bar_idx++;
  g_sh->arrive[i].fetch_add(1);
  while (g_sh->arrive[i].load() < g_n) usleep(50);
}

struct Peers { uint8_t *p[MAXR]; };

__global__ void push_kernel(const uint4 *src, Peers peers, int rank, int n, size_t off16, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    for (int r = 0; r < n; ++r)
      if (r != rank) reinterpret_cast<uint4 *>(peers.p[r])[off16 + i] = v;
  }
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// flags live at the start of every window: flags[r] = last epoch rank r has reached
__global__ void barrier_kernel(Peers peers, int rank, int n, uint32_t epoch, uint32_t *status) {
  const int t = threadIdx.x;
  if (t < n) {
    __threadfence_system();
    uint32_t *dst = reinterpret_cast<uint32_t *>(peers.p[t]) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
    const uint32_t *mine = reinterpret_cast<const uint32_t *>(peers.p[rank]) + t;
    const unsigned long long t0 = gtime();
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      if (gtime() - t0 > 2000000000ull) { atomicExch(status, 1u); break; }  // 2 s
    } while (true);
  }
}

struct Hit { uint32_t i, j; int32_t dot; float ani; };
__global__ void append_kernel(unsigned long long *counter, Hit *hits, int rank, uint32_t rounds, uint32_t keep_mod) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (uint32_t r = 0; r < rounds; ++r) {
    const bool keep = ((gw * 131u + r * 17u + lane) % keep_mod) == 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (!bal) continue;
    unsigned long long base = 0;
    if (lane == (uint32_t)(__ffs(bal) - 1)) base = atomicAdd(counter, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
    if (keep) {
      Hit h{(uint32_t)rank, gw, (int32_t)r, 1.0f};
      hits[base + __popc(bal & ((1u << lane) - 1u))] = h;
    }
  }
}

__global__ void __launch_bounds__(320, 1) dummy_cluster_kernel(int *x) {
  extern __shared__ uint8_t sm[];
  if (x && threadIdx.x == 9999) x[0] = sm[0];
}

static void run(int rank, int n) {
  g_rank = rank;
  CK(cudaSetDevice(rank));
  const size_t WIN = 96ull << 20;
  uint8_t *win;
  CK(cudaMalloc(&win, WIN));
  CK(cudaMemset(win, 0, WIN));
  CK(cudaDeviceSynchronize());
  CK(cudaIpcGetMemHandle(&g_sh->handle[rank], win));
  host_barrier();
  Peers peers;
  for (int r = 0; r < n; ++r) {
    if (r == rank) { peers.p[r] = win; continue; }
    void *p;
    CK(cudaIpcOpenMemHandle(&p, g_sh->handle[r], cudaIpcMemLazyEnablePeerAccess));
    peers.p[r] = (uint8_t *)p;
  }
  host_barrier();
  if (rank == 0) printf("ipc: %d ranks opened each other's windows\n", n);

  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  uint32_t *d_status;
  CK(cudaMalloc(&d_status, 4));
  CK(cudaMemset(d_status, 0, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  uint32_t epoch = 0;
  auto barrier = [&]() { barrier_kernel<<<1, 32, 0, st>>>(peers, rank, n, ++epoch, d_status); };

  // ---- 1. push: my 40 MB / n slice to every peer, pattern-checked ----
  const size_t total = 40ull << 20, slice = total / n, n16 = slice / 16;
  uint8_t *src;
  CK(cudaMalloc(&src, slice));
  CK(cudaMemset(src, 0x10 + rank, slice));
  const size_t off16 = ((1 << 20) + rank * slice) / 16;  // data after a 1 MB flag / counter area
  for (int rep = 0; rep < 3; ++rep) {
    barrier();
    if (rep == 2) CK(cudaEventRecord(e0, st));
    push_kernel<<<148 * 4, 256, 0, st>>>((const uint4 *)src, peers, rank, n, off16, n16);
    barrier();
    if (rep == 2) CK(cudaEventRecord(e1, st));
  }
  CK(cudaStreamSynchronize(st));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  g_sh->result[rank][0] = ms;
  // check what the peers pushed to me
  {
    uint8_t *h = (uint8_t *)malloc(total);
    CK(cudaMemcpy(h, win + (1 << 20), total, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (int r = 0; r < n; ++r)
      if (r != rank)
        for (size_t i = 0; i < slice; i += 4097) bad += h[r * slice + i] != (uint8_t)(0x10 + r);
    g_sh->result[rank][1] = (double)bad;
    free(h);
  }
  // ---- 2. barrier latency ----
  barrier();
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < 200; ++i) barrier();
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  g_sh->result[rank][2] = ms / 200 * 1e3;
  // ---- 3. remote append into rank 0's window ----
  unsigned long long *counter = (unsigned long long *)(peers.p[0] + 4096);
  Hit *hits = (Hit *)(peers.p[0] + (48ull << 20));
  host_barrier();
  if (rank == 0) CK(cudaMemsetAsync(win + 4096, 0, 8, st));
  barrier();
  CK(cudaEventRecord(e0, st));
  append_kernel<<<148, 256, 0, st>>>(counter, hits, rank, 2000, 97);  // ~ 148*8*2000*32/97 = 780 k hits per rank
  CK(cudaEventRecord(e1, st));
  barrier();
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  g_sh->result[rank][3] = ms;
  if (rank == 0) {
    unsigned long long cnt;
    CK(cudaMemcpy(&cnt, win + 4096, 8, cudaMemcpyDeviceToHost));
    g_sh->result[0][4] = (double)cnt;
    Hit *h = (Hit *)malloc(cnt * 16);
    CK(cudaMemcpy(h, win + (48ull << 20), cnt * 16, cudaMemcpyDeviceToHost));
    unsigned long long per[MAXR] = {0}, badrec = 0;
    for (unsigned long long i = 0; i < cnt; ++i) { if (h[i].i < (uint32_t)n && h[i].ani == 1.0f) per[h[i].i]++; else badrec++; }
    printf("append: %llu records in rank 0's window, malformed %llu, per rank:", cnt, badrec);
    for (int r = 0; r < n; ++r) printf(" %llu", per[r]);
    printf("\n");
    free(h);
  }
  uint32_t stt;
  CK(cudaMemcpy(&stt, d_status, 4, cudaMemcpyDeviceToHost));
  g_sh->result[rank][5] = stt;
  // ---- 4. cluster occupancy of a 1-CTA-per-SM kernel ----
  if (rank == 0) {
    CK(cudaFuncSetAttribute(dummy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    CK(cudaFuncSetAttribute(dummy_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int cs : {1, 2, 4, 8}) {
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cfg.blockDim = dim3(320); cfg.gridDim = dim3(148 / cs * cs); cfg.dynamicSmemBytes = 210 * 1024;
      int nc = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, dummy_cluster_kernel, &cfg);
      printf("cluster size %d: max active clusters %d (%d CTAs) %s\n", cs, nc, nc * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  host_barrier();
  if (rank == 0) {
    for (int r = 0; r < n; ++r)
      printf("rank %d: push of %.1f MB to each of %d peers + 2 barriers %.3f ms (%.1f GB/s out), bad bytes %.0f, barrier %.1f us, "
             "append kernel %.3f ms, timeouts %.0f\n",
             r, slice / 1e6, n - 1, g_sh->result[r][0], slice * (n - 1) / g_sh->result[r][0] / 1e6, g_sh->result[r][1],
             g_sh->result[r][2], g_sh->result[r][3], g_sh->result[r][5]);
  }
  for (int r = 0; r < n; ++r)
    if (r != rank) cudaIpcCloseMemHandle(peers.p[r]);
  cudaFree(win);
}

int main(int argc, char **argv) {
  g_n = argc > 1 ? atoi(argv[1]) : 2;
  if (g_n < 1 || g_n > MAXR) return 1;
  g_sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  memset((void *)g_sh, 0, sizeof(Shared));
  pid_t pids[MAXR];
  for (int r = 1; r < g_n; ++r) {
    pids[r] = fork();
    if (pids[r] == 0) { run(r, g_n); _exit(0); }
  }
  run(0, g_n);
  int rc = 0;
  for (int r = 1; r < g_n; ++r) { int st; waitpid(pids[r], &st, 0); rc |= st; }
  return rc != 0;
}
