/*
 * ref_driver.cpp — TEST INFRASTRUCTURE.  Host-side launcher for the reference kernel
 * compiled as CPU code (see ref_shim.h).  Emulates the reference's launch geometry and
 * result collection, src/sketch_cuda.rs:119-166: 512 k-mer starts per thread (:130),
 * n_threads = ceil(n_kmers / 512) (:131), n_hash_per_thread = max(512 / scaled * 4, 8)
 * (:136), a zero-filled u64 slot array (:138), and a host scan that keeps every non-zero
 * slot in a set (:158-163).
 */
#include "ref_shim.h"
#include <cstdlib>
#include <cstring>
#include <vector>

thread_local hgref_dim3 blockIdx, blockDim, threadIdx;

extern "C" uint64_t t1ha2_atonce(uint8_t *data, size_t length, uint64_t seed);
extern "C" void cuda_kmer_t1ha2(uint8_t *seq, const size_t n_bps, const size_t n_kmer_per_thread,
                                const size_t n_hash_per_thread, const size_t ksize,
                                const uint64_t threshold, const uint64_t seed, const bool canonical,
                                uint64_t *kmer_scaled_hash);

extern "C" __attribute__((visibility("default"))) uint64_t
hgref_t1ha2_atonce(const uint8_t *data, uint64_t length, uint64_t seed) {
  uint8_t buf[64] = {0};
  memcpy(buf, data, length > 32 ? 32 : length);
  return t1ha2_atonce(buf, (size_t)length, seed);
}

/* returns the number of distinct hashes, writes min(n, cap) of them sorted */
extern "C" __attribute__((visibility("default"))) uint64_t
hgref_extract_kmer_t1ha2(const uint8_t *seq, uint64_t n_bps, uint32_t k, uint64_t scaled,
                         uint64_t seed, int canonical, uint64_t *out, uint64_t cap) {
  if (n_bps < k) return 0;
  const size_t n_kmers = n_bps - k + 1, kmer_per_thread = 512;
  const size_t n_threads = (n_kmers + kmer_per_thread - 1) / kmer_per_thread;
  const size_t n_hash_per_thread = std::max<size_t>(kmer_per_thread / scaled * 4, 8);
  std::vector<uint64_t> slots(n_hash_per_thread * n_threads, 0);
  std::vector<uint8_t> s(seq, seq + n_bps);
  blockDim = {1024, 1, 1}; /* LaunchConfig::for_num_elems */
  for (size_t t = 0; t < n_threads; t++) {
    blockIdx = {(unsigned)(t / 1024), 0, 0};
    threadIdx = {(unsigned)(t % 1024), 0, 0};
    cuda_kmer_t1ha2(s.data(), n_bps, kmer_per_thread, n_hash_per_thread, k, UINT64_MAX / scaled,
                    seed, canonical != 0, slots.data());
  }
  std::vector<uint64_t> v;
  for (uint64_t h : slots)
    if (h != 0) v.push_back(h);
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
  for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
  return v.size();
}
